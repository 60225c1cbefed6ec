#!/bin/bash
# ncu launch list of the bench command itself (no extras): the kernels' shares of a step
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/y_bench_plain.json 2> gpurun_out/y_bench_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/y_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/y_ncu.log 2>&1
