#!/bin/bash
# round-2 call H: wgrad GEMMs on a side stream -- tests, A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
HAMT_WGRAD_STREAM=0 timeout 600 python bench.py --quick > gpurun_out/bench_quick_off.json 2> gpurun_out/bench_quick_off.err; echo "== off exit $?"; cat gpurun_out/bench_quick_off.json; tail -n 3 gpurun_out/bench_quick_off.err
HAMT_WGRAD_STREAM=1 timeout 600 python bench.py --quick > gpurun_out/bench_quick_on.json 2> gpurun_out/bench_quick_on.err; echo "== on exit $?"; cat gpurun_out/bench_quick_on.json; tail -n 3 gpurun_out/bench_quick_on.err
