#!/bin/bash
# round-2 call T: all GPU tests, the full bench line, reference arm, configs rxr / r4r
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --diag > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?"; cut -c1-3500 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err | cut -c1-300
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== ref exit $?"; cut -c1-2500 gpurun_out/bench_ref.json; tail -n 3 gpurun_out/bench_ref.err
timeout 900 python bench.py --config rxr --no-cpu-baseline --no-store-leg > gpurun_out/bench_rxr.json 2> gpurun_out/bench_rxr.err; echo "== rxr exit $?"; cut -c1-3000 gpurun_out/bench_rxr.json; tail -n 3 gpurun_out/bench_rxr.err | cut -c1-300
timeout 900 python bench.py --config r4r --no-cpu-baseline --no-store-leg > gpurun_out/bench_r4r.json 2> gpurun_out/bench_r4r.err; echo "== r4r exit $?"; cut -c1-3000 gpurun_out/bench_r4r.json; tail -n 3 gpurun_out/bench_r4r.err | cut -c1-300
