#!/bin/bash
# round-2 N = 2 data-parallel experiment (gpurun --gpus 2): wire dtype of the gradient exchange, quick mode (device-resident value leg)
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 2 --quick "$@" 2>gpurun_out/scale2.err | tail -n 1; }
: > gpurun_out/scale_n2.jsonl
timeout 300 python bench.py --quick 2>/dev/null | tail -n 1 | tee gpurun_out/scale_n1.json
run --grad-wire fp32 | tee -a gpurun_out/scale_n2.jsonl
run --grad-wire bf16 | tee -a gpurun_out/scale_n2.jsonl
run --grad-wire fp32 --dp-mode after | tee -a gpurun_out/scale_n2.jsonl
tail -n 5 gpurun_out/scale2.err | cut -c1-300
