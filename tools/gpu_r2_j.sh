#!/bin/bash
# round-2 call J: 16-warp softmax backward in the tcgen05 attention backward + refitted GEMM cost model
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_gemm_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_attn.log 2>&1; echo "== pytest attn+gemm exit $?"; tail -n 5 gpurun_out/pytest_attn.log | cut -c1-300
timeout 300 python tools/kbench.py --no-gemm --attn-only --attn-ab > gpurun_out/kbench_attn.log 2>&1; echo "== kbench exit $?"; grep '"attn"' gpurun_out/kbench_attn.log | cut -c1-230
timeout 600 python bench.py --quick > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== quick exit $?"; cat gpurun_out/bench_quick.json; tail -n 2 gpurun_out/bench_quick.err | cut -c1-300
