#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_attn.log 2>&1; echo "== pytest attn exit $?"; tail -n 3 gpurun_out/pytest_attn.log | cut -c1-300
timeout 300 python tools/kbench.py --no-gemm --attn-only --attn-ab > gpurun_out/kbench_attn.log 2>&1; echo "== kbench exit $?"; grep '"attn"' gpurun_out/kbench_attn.log | cut -c1-230
