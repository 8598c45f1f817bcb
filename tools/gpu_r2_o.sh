#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_vit.log 2>&1; echo "== pytest vit exit $?"; tail -n 5 gpurun_out/pytest_vit.log | cut -c1-300
timeout 900 python bench.py --config e2e --steps 12 --warmup 3 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "== e2e exit $?"; cat gpurun_out/bench_e2e.json; tail -n 5 gpurun_out/bench_e2e.err | cut -c1-300
