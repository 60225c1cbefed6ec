#!/bin/bash
# N-GPU bench (NCCL): weak-scaling line + config 3 sharded + config 4 dealt over the ranks.  usage: gpu_t.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/t_bench$N.json 2> gpurun_out/t_bench$N.err; echo "rc=$?" >> gpurun_out/t_bench$N.err
