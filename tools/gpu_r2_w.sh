#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/graph_idle.py sap > gpurun_out/graph_idle_sap.txt 2>&1; echo "== idle sap exit $?"; grep -v Warning gpurun_out/graph_idle_sap.txt | tail -n 16 | cut -c1-200
timeout 300 python tools/graph_idle.py mlm > gpurun_out/graph_idle_mlm.txt 2>&1; echo "== idle mlm exit $?"; grep -v Warning gpurun_out/graph_idle_mlm.txt | tail -n 16 | cut -c1-200
timeout 300 python -m pytest tests/test_attn_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_attn.log 2>&1; echo "== pytest attn exit $?"; tail -n 2 gpurun_out/pytest_attn.log | cut -c1-200
timeout 300 python tools/kbench.py --no-gemm --attn-only --attn-long > gpurun_out/kbench_attn_long.log 2>&1; echo "== kbench long exit $?"; grep '"attn"' gpurun_out/kbench_attn_long.log | cut -c1-230
