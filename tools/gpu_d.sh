#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slidegraph.py tests/test_cpp_adapter.py -m gpu -q --durations=10 > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
