#!/bin/bash
set -u
mkdir -p gpurun_out
SOAK_VERBOSE=1 timeout 60 python -u tools/soak_join_vs_lattice.py 30 0 > gpurun_out/s_soak.log 2>&1; echo "rc=$?" >> gpurun_out/s_soak.log
