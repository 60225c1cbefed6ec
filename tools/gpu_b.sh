#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
