#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_graph_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_gg.log 2>&1; echo "== pytest gemm+graph exit $?"; grep -E "passed|failed|^E  .*assert|AssertionError" gpurun_out/pytest_gg.log | cut -c1-700 | head -8
timeout 300 python tools/kbench_wide.py > gpurun_out/kbench_wide.log 2>&1; echo "== kbench_wide exit $?"; cat gpurun_out/kbench_wide.log | cut -c1-1000
timeout 600 python bench.py --quick > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== quick exit $?"; cat gpurun_out/bench_quick.json
