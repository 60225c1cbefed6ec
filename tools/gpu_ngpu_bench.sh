#!/bin/bash
# N-GPU bench (NCCL).  usage: gpu_ngpu_bench.sh N [steps] [extra bench flags]
set -u
N=${1:-2}; S=${2:-10}; shift; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps $S --warmup 3 "$@" > gpurun_out/t_bench$N.json 2> gpurun_out/t_bench$N.err; echo "rc=$?" >> gpurun_out/t_bench$N.err
