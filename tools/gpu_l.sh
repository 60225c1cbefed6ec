#!/bin/bash
# first GPU run of the pair-join scorer: its parity suite, timings against the lattice kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_join.py -m gpu -x -q --durations=8 > gpurun_out/l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/l_pytest.log
timeout 200 python tools/ab_search.py 2 10 > gpurun_out/l_ab2.log 2>&1
timeout 300 python tools/ab_search.py 3 2 4,3 > gpurun_out/l_ab3.log 2>&1
for c in 4 5; do SLIDE_PR_TRACE=1 timeout 200 python tools/ncu_cfg_target.py $c > gpurun_out/l_c$c.log 2>&1; done
