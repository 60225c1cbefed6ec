#!/bin/bash
# full GPU suite + default bench after the windowed kernel was removed
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest.log
timeout 600 python bench.py > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "rc=$?" >> gpurun_out/k_bench.err
