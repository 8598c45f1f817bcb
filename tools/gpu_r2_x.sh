#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_graph_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_x.log 2>&1; echo "== pytest exit $?"; tail -n 3 gpurun_out/pytest_x.log | cut -c1-300
timeout 600 python bench.py --quick > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== quick exit $?"; cat gpurun_out/bench_quick.json
timeout 600 python bench.py --quick --tasks mlm > gpurun_out/bench_quick_mlm.json 2> gpurun_out/bench_quick_mlm.err; echo "== quick mlm exit $?"; cat gpurun_out/bench_quick_mlm.json
