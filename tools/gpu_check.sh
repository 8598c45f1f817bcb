#!/bin/bash
# one gpurun call: full GPU test suite, reference arm (CPU port + informational GPU-eager port), ncu capture of the big FFN GEMMs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 12 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== ref exit $?"; cat gpurun_out/bench_ref.json; tail -n 3 gpurun_out/bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 3 -f -o gpurun_out/prof_gelu python tools/ncu_gelu.py > gpurun_out/prof_gelu.log 2>&1; echo "== ncu gelu exit $?"
