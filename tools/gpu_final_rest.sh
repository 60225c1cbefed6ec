#!/bin/bash
# the rest of the GPU suite (everything but the pair-join / cache files run by gpu_final_join_bench.sh and the two config-3 full-size tests)
mkdir -p gpurun_out
timeout 235 python -m pytest tests -m gpu -q -x --durations=6 -k "not config3" --ignore=tests/test_gpu_join.py --ignore=tests/test_gpu_cache.py > gpurun_out/x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/x_pytest.log
