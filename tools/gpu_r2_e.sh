#!/bin/bash
# round-2 call E: tcgen05 attention backward -- parity (both implementations), A/B timings, ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_attn.log 2>&1; echo "== pytest attn exit $?"; tail -n 25 gpurun_out/pytest_attn.log | cut -c1-220
timeout 300 python tools/kbench.py --no-gemm --attn-only --attn-ab > gpurun_out/kbench_attn.log 2>&1; echo "== kbench exit $?"; grep '"attn"' gpurun_out/kbench_attn.log | cut -c1-230
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -c 4 -f -o gpurun_out/prof_attn_bwd python tools/ncu_attn.py > gpurun_out/prof_attn_bwd.log 2>&1; echo "== ncu exit $?"; tail -n 3 gpurun_out/prof_attn_bwd.log
