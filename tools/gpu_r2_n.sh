#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_vit.py > gpurun_out/debug_vit.log 2>&1; echo "== debug exit $?"; grep -E "depth|qkv.weight|patch_embed.proj.weight" gpurun_out/debug_vit.log | cut -c1-160
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 30 gpurun_out/pytest_gpu.log | grep -v Warning | cut -c1-400
cat gpurun_out/parity_vit_*.txt
timeout 600 python bench.py --quick > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== quick exit $?"; cat gpurun_out/bench_quick.json
timeout 900 python bench.py --config e2e --steps 12 --warmup 3 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "== e2e exit $?"; cat gpurun_out/bench_e2e.json; tail -n 5 gpurun_out/bench_e2e.err | cut -c1-300
