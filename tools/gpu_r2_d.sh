#!/bin/bash
# round-2 call D: tcgen05 attention forward v2 (polling MMA issuer, Pd over Q/K, 4 stages) -- parity, A/B timings, ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_attn.log 2>&1; echo "== pytest attn exit $?"; tail -n 12 gpurun_out/pytest_attn.log
timeout 300 python tools/kbench.py --no-gemm --attn-only --attn-ab --attn-drop > gpurun_out/kbench_attn.log 2>&1; echo "== kbench exit $?"; grep tcgen05 gpurun_out/kbench_attn.log | tail -n 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -c 8 -f -o gpurun_out/prof_attn_fwd python tools/ncu_attn.py > gpurun_out/prof_attn_fwd.log 2>&1; echo "== ncu exit $?"; tail -n 3 gpurun_out/prof_attn_fwd.log
