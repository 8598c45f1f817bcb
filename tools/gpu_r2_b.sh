#!/bin/bash
# round-2 call B: full GPU suite after the parity / ADVICE / LayerNorm changes, row-kernel timings, quick bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 40 gpurun_out/pytest_gpu.log
timeout 300 python tools/kbench.py --no-gemm > gpurun_out/kbench_rows.log 2>&1; echo "== kbench exit $?"; tail -n 12 gpurun_out/kbench_rows.log
timeout 400 python bench.py --quick --steps 24 --warmup 12 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== bench exit $?"; cat gpurun_out/bench_quick.json; tail -n 3 gpurun_out/bench_quick.err
