#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_join.py -m gpu -x -q -k "dense_and_clustered or intra or edge" > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
