#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_triangle_pipeline.py tests/test_slidegraph.py -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hypothes or triangle or list" >> gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
