#!/bin/bash
# the driver's N = 2 form of the FULL bench line (all legs), to make sure the multi-rank path of every leg works on the final tree
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 24 --warmup 12 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "== n2 exit $?"; tail -n 1 gpurun_out/bench_n2.json | cut -c1-1500; grep -v "Warning\|run_backward\|\*\*\*" gpurun_out/bench_n2.err | tail -n 5 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --no-gpu-eager > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "== ref n2 exit $?"; tail -n 1 gpurun_out/bench_ref_n2.json | cut -c1-300
