#!/bin/bash
# one-minute sanity check of the in-tree build: GEMM + attention parity, device-resident bench value
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gemm_gpu.py tests/test_attn_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_qc.log 2>&1; echo "== pytest exit $?"; tail -n 2 gpurun_out/pytest_qc.log | cut -c1-200
timeout 200 python bench.py --quick > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== quick exit $?"; cat gpurun_out/bench_quick.json
