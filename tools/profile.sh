#!/bin/bash
# ncu captures for profiles/: (1) launch list with per-launch device time for one SAP + one MLM step, (2) --set full on the GEMM,
# attention and LN kernels.  Numbers printed by bench.py under ncu are NOT bench values.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -s 2600 -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 5 --tasks sap,mlm --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
$NCU --set full --import-source on -k regex:gemm_tcgen05 -s 400 -c 6 -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 3 --tasks sap --no-cpu-baseline > gpurun_out/prof_gemm.log 2>&1
echo "gemm full exit $?"
$NCU --set full --import-source on -k regex:attn_ -s 60 -c 6 -o gpurun_out/prof_attn \
    python bench.py --steps 1 --warmup 3 --tasks sap --no-cpu-baseline > gpurun_out/prof_attn.log 2>&1
echo "attn full exit $?"
ls -la gpurun_out/*.ncu-rep
