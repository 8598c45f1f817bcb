#!/bin/bash
# round-2 call C: tcgen05 attention forward -- parity (both implementations), A/B timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/pytest_attn.log 2>&1; echo "== pytest attn exit $?"; tail -n 25 gpurun_out/pytest_attn.log
timeout 300 python -m pytest tests/test_finetune_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -k "later_forwards or (gradients_vs_oracle and itm)" > gpurun_out/pytest_fix.log 2>&1; echo "== pytest fixes exit $?"; tail -n 5 gpurun_out/pytest_fix.log
timeout 300 python tools/kbench.py --no-gemm --attn-only --attn-ab --attn-drop > gpurun_out/kbench_attn.log 2>&1; echo "== kbench exit $?"; tail -n 40 gpurun_out/kbench_attn.log
