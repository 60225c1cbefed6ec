#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --no-extras --steps 5 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "rc=$?" >> gpurun_out/s_bench.err
