#!/bin/bash
# round-2 final validation: all GPU tests, smoke(), the full bench line (+ per-signature table)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 1 gpurun_out/smoke.log | cut -c1-400
timeout 900 python bench.py --diag > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?"; cut -c1-700 gpurun_out/bench.json
