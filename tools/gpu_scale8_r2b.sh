#!/bin/bash
# round-2 N = 8 experiment B (gpurun --gpus 8): SMs reserved for the overlapped NCCL all-reduce during the backward, and the un-overlapped exchange
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 8 --quick "$@" 2>gpurun_out/scale8.err | tail -n 1; }
: > gpurun_out/scale_n8b.jsonl
run --nccl-sms 16 | tee -a gpurun_out/scale_n8b.jsonl
run --nccl-sms 24 | tee -a gpurun_out/scale_n8b.jsonl
run --dp-mode after | tee -a gpurun_out/scale_n8b.jsonl
grep -v "Warning\|run_backward\|\*\*\*" gpurun_out/scale8.err | tail -n 3 | cut -c1-300
