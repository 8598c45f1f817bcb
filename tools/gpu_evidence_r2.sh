#!/bin/bash
# round-2 evidence call: e2e-stage line (graphed), launch list of the bench command, --set full of attention and the FFN GEMMs
mkdir -p gpurun_out
timeout 900 python bench.py --config e2e --steps 12 --warmup 3 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "== e2e exit $?"; cat gpurun_out/bench_e2e.json | cut -c1-400; tail -n 3 gpurun_out/bench_e2e.err | cut -c1-300
NCU="ncu --clock-control none --profile-from-start off"
BENCH="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --profile-range"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1; echo "== launch list exit $?"
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -n 40 gpurun_out/launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 16 -f -o gpurun_out/prof_attn python tools/ncu_attn.py > gpurun_out/prof_attn.log 2>&1; echo "== ncu attn exit $?"
python tools/ncu_summary.py gpurun_out/prof_attn.ncu-rep > gpurun_out/ncu_attn_summary.txt 2>&1; cat gpurun_out/ncu_attn_summary.txt | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 8 -f -o gpurun_out/prof_gelu python tools/ncu_gelu.py > gpurun_out/prof_gelu.log 2>&1; echo "== ncu gelu exit $?"
python tools/ncu_summary.py gpurun_out/prof_gelu.ncu-rep > gpurun_out/ncu_gelu_summary.txt 2>&1; cat gpurun_out/ncu_gelu_summary.txt | cut -c1-400
