#!/bin/bash
# round-2 N = 8 check (gpurun --gpus 8): weak scaling of the overlapped gradient exchange, fp32 and bf16 wire, quick mode
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 8 --quick "$@" 2>gpurun_out/scale8.err | tail -n 1; }
: > gpurun_out/scale_n8.jsonl
timeout 200 python bench.py --quick 2>/dev/null | tail -n 1 | tee -a gpurun_out/scale_n8.jsonl
run --grad-wire fp32 | tee -a gpurun_out/scale_n8.jsonl
run --grad-wire bf16 | tee -a gpurun_out/scale_n8.jsonl
grep -v "Warning\|run_backward\|\*\*\*" gpurun_out/scale8.err | tail -n 4 | cut -c1-300
