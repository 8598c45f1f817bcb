#!/bin/bash
mkdir -p gpurun_out
cd tools && timeout 300 python kbench_decoder.py > ../gpurun_out/kbench_decoder.log 2>&1; echo "== decoder exit $?"; cat ../gpurun_out/kbench_decoder.log | cut -c1-1500; cd ..
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 10 -f -o gpurun_out/prof_gelu python tools/ncu_gelu.py > gpurun_out/prof_gelu.log 2>&1; echo "== ncu gelu exit $?"
python tools/ncu_summary.py gpurun_out/prof_gelu.ncu-rep > gpurun_out/ncu_gelu_summary.txt 2>&1; cat gpurun_out/ncu_gelu_summary.txt | cut -c1-400; rm -f gpurun_out/prof_gelu.ncu-rep
