#!/bin/bash
# round-2 call G: all GPU tests, the full bench line (store leg, HBM rooflines), step profile, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --diag > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?"; cut -c1-3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err | cut -c1-400
timeout 300 python tools/step_profile.py > gpurun_out/step_profile.txt 2>&1; echo "== step_profile exit $?"; head -n 32 gpurun_out/step_profile.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== ref exit $?"; cut -c1-2500 gpurun_out/bench_ref.json; tail -n 3 gpurun_out/bench_ref.err
