#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/s_ab.log
timeout 900 python -m pytest tests/test_gpu_join.py tests/test_gpu_cache.py -m gpu -x -q -k "not config3 and not config4" > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
timeout 200 python tools/ab_search.py 2 10 4 >> gpurun_out/s_ab.log 2>&1
for c in 4 5; do timeout 200 python tools/ncu_cfg_target.py $c 2>&1 | tail -1 >> gpurun_out/s_ab.log; done
