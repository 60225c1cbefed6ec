#!/bin/bash
# pair-join scorer: parity suite, timings against the lattice kernels, ncu capture on config 2
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_join.py -m gpu -x -q --durations=5 > gpurun_out/n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/n_pytest.log
timeout 200 python tools/ab_search.py 2 10 4 > gpurun_out/n_ab2.log 2>&1
timeout 300 python tools/ab_search.py 3 2 4 > gpurun_out/n_ab3.log 2>&1
for c in 4 5; do timeout 200 python tools/ncu_cfg_target.py $c > gpurun_out/n_c$c.log 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spr_join_score -c 1 -o gpurun_out/n_join_c2 python tools/ncu_step_target.py default 2 > gpurun_out/n_ncu.log 2>&1
timeout 300 ncu -i gpurun_out/n_join_c2.ncu-rep --page raw --csv > gpurun_out/n_join_c2_raw.csv 2> gpurun_out/n_err1.log
timeout 300 ncu -i gpurun_out/n_join_c2.ncu-rep --page source --csv > gpurun_out/n_join_c2_src.csv 2> gpurun_out/n_err2.log
