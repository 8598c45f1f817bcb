#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/kbench.py --no-gemm --attn-only --attn-ab > gpurun_out/kbench_attn.log 2>&1; echo "== kbench exit $?"; grep '"attn"' gpurun_out/kbench_attn.log | cut -c1-230
