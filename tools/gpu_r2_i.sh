#!/bin/bash
mkdir -p gpurun_out
cd tools && timeout 600 python gemm_sweep.py > ../gpurun_out/gemm_sweep.log 2>&1; echo "== sweep exit $?"; cat ../gpurun_out/gemm_sweep.log | cut -c1-900
