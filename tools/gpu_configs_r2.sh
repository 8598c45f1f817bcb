#!/bin/bash
# BASELINE configs 3 / 4 / 5 bench lines (final round-2 tree)
mkdir -p gpurun_out
timeout 900 python bench.py --config rxr --no-cpu-baseline --no-store-leg --diag > gpurun_out/bench_rxr.json 2> gpurun_out/bench_rxr.err; echo "== rxr exit $?"; cut -c1-300 gpurun_out/bench_rxr.json
timeout 900 python bench.py --config e2e --steps 12 --warmup 3 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "== e2e exit $?"; cut -c1-300 gpurun_out/bench_e2e.json
