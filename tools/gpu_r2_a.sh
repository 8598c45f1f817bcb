#!/bin/bash
# round-2 call A: opt-in experimental parity tests, then default-vs-experimental kernel timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
HAMT_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_rowops_gpu.py -m gpu -q -k "wide or experimental or variant" --timeout 300 > gpurun_out/pytest_exp.log 2>&1; echo "== pytest exp exit $?"; tail -n 8 gpurun_out/pytest_exp.log
timeout 500 python tools/kbench_wide.py > gpurun_out/kbench_wide.log 2>&1; echo "== kbench_wide exit $?"; cat gpurun_out/kbench_wide.log | tail -n 12
