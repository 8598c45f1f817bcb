#!/bin/bash
# one gpurun call: unit tests per file (separate processes so a trap in one cannot poison the rest) + micro-bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in tests/test_gemm_gpu.py tests/test_rowops_gpu.py tests/test_attn_gpu.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q -x --timeout 300 > gpurun_out/$n.log 2>&1
  echo "== $n exit $?"; tail -n 25 gpurun_out/$n.log
done
timeout 600 python tools/kbench.py > gpurun_out/kbench.log 2>&1; echo "== kbench exit $?"; tail -n 30 gpurun_out/kbench.log
