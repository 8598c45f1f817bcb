#!/bin/bash
# ncu captures for profiles/ (numbers printed by bench.py under ncu are NOT bench values):
#  (1) launch list: gpu__time_duration.sum of every kernel of the timed device-resident region of bench.py (12 steps = one
#      pass over the 6-task schedule), bracketed by cudaProfilerStart/Stop (--profile-range);
#  (2) --set full of the dominant kernels inside the same region: tcgen05 GEMM, attention bwd/fwd, LayerNorm bwd/fwd.
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
BENCH="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --profile-range"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -n 30 gpurun_out/launches_summary.txt
# SAP step = step index 1 of the schedule: skip the ~470 launches of step 0 (MLM), then take the first launches of each kind
timeout 600 $NCU --set full --import-source on -k regex:gemm_tcgen05 -s 150 -c 14 -f -o gpurun_out/prof_gemm $BENCH > gpurun_out/prof_gemm.log 2>&1
echo "gemm full exit $?"
timeout 600 $NCU --set full --import-source on -k regex:'attn_|ln_bwd|ln_fwd' -s 60 -c 12 -f -o gpurun_out/prof_rows $BENCH > gpurun_out/prof_rows.log 2>&1
echo "rows full exit $?"
ls -la gpurun_out/*.ncu-rep
