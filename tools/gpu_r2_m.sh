#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_vit.py > gpurun_out/debug_vit.log 2>&1; echo "== exit $?"; cat gpurun_out/debug_vit.log | cut -c1-200 | tail -80
