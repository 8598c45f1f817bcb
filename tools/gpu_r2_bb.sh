#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_graph_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_bb.log 2>&1; echo "== pytest exit $?"; tail -n 3 gpurun_out/pytest_bb.log | cut -c1-300
cd tools && timeout 300 python kbench_decoder.py > ../gpurun_out/kbench_decoder2.log 2>&1; echo "== decoder exit $?"; cat ../gpurun_out/kbench_decoder2.log | cut -c1-1500; cd ..
timeout 600 python bench.py --quick > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== quick exit $?"; cat gpurun_out/bench_quick.json
timeout 600 python bench.py --quick > gpurun_out/bench_quick2.json 2> gpurun_out/bench_quick2.err; echo "== quick2 exit $?"; cat gpurun_out/bench_quick2.json
