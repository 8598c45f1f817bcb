#!/bin/bash
# round-2 N = 2 experiment B (gpurun --gpus 2): NCCL stream priority / channel cap for the overlapped gradient exchange, quick mode
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 2 --quick "$@" 2>gpurun_out/scale2.err | tail -n 1; }
: > gpurun_out/scale_n2b.jsonl
run --nccl-low-prio | tee -a gpurun_out/scale_n2b.jsonl
run | tee -a gpurun_out/scale_n2b.jsonl
run --nccl-channels 8 | tee -a gpurun_out/scale_n2b.jsonl
run --nccl-channels 4 | tee -a gpurun_out/scale_n2b.jsonl
grep -v "Warning\|run_backward\|\*\*\*" gpurun_out/scale2.err | tail -n 4 | cut -c1-300
