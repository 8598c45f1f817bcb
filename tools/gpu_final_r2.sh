#!/bin/bash
# round-2 final validation: all GPU tests, smoke(), the full bench line (+ per-signature table), reference arm, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 2 gpurun_out/smoke.log | cut -c1-400
timeout 900 python bench.py --diag > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?"; cut -c1-1200 gpurun_out/bench.json; grep -E "^\[gemm\] \((804|30522|768, 768, 804)" gpurun_out/bench.err | cut -c1-200
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== ref exit $?"; cut -c1-400 gpurun_out/bench_ref.json
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --profile-range > gpurun_out/launches_bench.log 2>&1; echo "== launch list exit $?"
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -n 36 gpurun_out/launches_summary.txt; rm -f gpurun_out/launches.csv
