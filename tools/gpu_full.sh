#!/bin/bash
# one gpurun call: all GPU tests, the bench line, per-kernel step profile, kernel micro-bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?"; cat gpurun_out/bench.json
timeout 300 python tools/step_profile.py > gpurun_out/step_profile.txt 2>&1; echo "== step_profile exit $?"; head -n 32 gpurun_out/step_profile.txt
timeout 400 python tools/kbench.py > gpurun_out/kbench.log 2>&1; echo "== kbench exit $?"; tail -n 30 gpurun_out/kbench.log
