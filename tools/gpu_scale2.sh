#!/bin/bash
# N = 2 data-parallel experiment: SM reservation for the overlapped NCCL all-reduce (bench.py --nccl-sms R), quick mode
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 2 --quick "$@" 2>gpurun_out/scale2.err | tail -n 1; }
timeout 300 python bench.py --quick 2>/dev/null | tail -n 1 | tee gpurun_out/scale_n1.json
for R in 0 8 16; do run --nccl-sms $R | tee -a gpurun_out/scale_n2.jsonl; done
run --nccl-sms 8 --dp-mode after | tee -a gpurun_out/scale_n2.jsonl
tail -n 5 gpurun_out/scale2.err
