#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_join.py -m gpu -x -q -k "not config3 and not config4 and not config5" > gpurun_out/q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/q_pytest.log
timeout 200 python tools/ab_search.py 2 10 4 > gpurun_out/q_ab2.log 2>&1
timeout 300 python tools/ab_search.py 3 2 4 > gpurun_out/q_ab3.log 2>&1
for c in 4 5; do timeout 200 python tools/ncu_cfg_target.py $c > gpurun_out/q_c$c.log 2>&1; done
SLIDE_PR_TRACE=1 timeout 100 python tools/trace_e2e.py > gpurun_out/q_trace.log 2>&1
