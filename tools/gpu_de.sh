#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slidegraph.py tests/test_cpp_adapter.py -m gpu -q --durations=10 > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
SLIDE_PR_TRACE=1 timeout 120 python tools/trace_e2e.py > gpurun_out/e_trace.log 2>&1
bash tools/gpu_e.sh
