#!/bin/bash
mkdir -p gpurun_out
for W in 1 0; do
  HAMT_WGRAD_STREAM=$W timeout 600 python -m pytest tests/test_graph_gpu.py -m gpu -q --timeout 600 -k "drift or sprel or sap" > gpurun_out/pytest_drift_w$W.log 2>&1; echo "== pytest drift (wgrad stream $W) exit $?"; grep -E "passed|failed|^E  .*assert" gpurun_out/pytest_drift_w$W.log | cut -c1-300 | head -8
done
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gemm.log 2>&1; echo "== pytest gemm exit $?"; tail -n 3 gpurun_out/pytest_gemm.log | cut -c1-300
timeout 300 python tools/kbench_wide.py > gpurun_out/kbench_wide.log 2>&1; echo "== kbench_wide exit $?"; cat gpurun_out/kbench_wide.log | cut -c1-900
HAMT_GELU_DER=0 timeout 600 python bench.py --quick > gpurun_out/bench_quick_g0.json 2> gpurun_out/bench_quick_g0.err; echo "== gelu_der 0 exit $?"; cat gpurun_out/bench_quick_g0.json
HAMT_GELU_DER=1 timeout 600 python bench.py --quick > gpurun_out/bench_quick_g1.json 2> gpurun_out/bench_quick_g1.err; echo "== gelu_der 1 exit $?"; cat gpurun_out/bench_quick_g1.json
HAMT_BRANCH_STREAMS=1 timeout 600 python bench.py --quick > gpurun_out/bench_quick_b1.json 2> gpurun_out/bench_quick_b1.err; echo "== branch 1 exit $?"; cat gpurun_out/bench_quick_b1.json; tail -n 3 gpurun_out/bench_quick_b1.err | cut -c1-300
HAMT_BRANCH_STREAMS=1 timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_branch.log 2>&1; echo "== pytest branch exit $?"; tail -n 4 gpurun_out/pytest_branch.log | cut -c1-300
