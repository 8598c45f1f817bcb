#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_graph_gpu.py tests/test_model_gpu.py tests/test_vit.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python tools/kbench_wide.py > gpurun_out/kbench_wide.log 2>&1; echo "== kbench_wide exit $?"; cat gpurun_out/kbench_wide.log | cut -c1-900
HAMT_BRANCH_STREAMS=0 timeout 600 python bench.py --quick > gpurun_out/bench_quick_b0.json 2> gpurun_out/bench_quick_b0.err; echo "== b0 exit $?"; cat gpurun_out/bench_quick_b0.json
HAMT_BRANCH_STREAMS=1 timeout 600 python bench.py --quick > gpurun_out/bench_quick_b1.json 2> gpurun_out/bench_quick_b1.err; echo "== b1 exit $?"; cat gpurun_out/bench_quick_b1.json; tail -n 3 gpurun_out/bench_quick_b1.err | cut -c1-300
HAMT_BRANCH_STREAMS=1 timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_branch.log 2>&1; echo "== pytest branch exit $?"; tail -n 4 gpurun_out/pytest_branch.log | cut -c1-300
timeout 900 python bench.py --config e2e --steps 12 --warmup 3 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "== e2e exit $?"; cat gpurun_out/bench_e2e.json | cut -c1-1500; tail -n 3 gpurun_out/bench_e2e.err | cut -c1-300
