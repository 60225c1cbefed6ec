#!/bin/bash
# 2-GPU bench (NCCL): weak-scaling line + config 3 sharded + config 4 dealt over the ranks
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/t_bench8.json 2> gpurun_out/t_bench8.err; echo "rc=$?" >> gpurun_out/t_bench8.err
