#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_join.py -m gpu -x -q -k "not config3 and not config4 and not config5" > gpurun_out/o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/o_pytest.log
timeout 200 python tools/ab_search.py 2 10 4 > gpurun_out/o_ab2.log 2>&1
timeout 300 python tools/ab_search.py 3 2 4 > gpurun_out/o_ab3.log 2>&1
for c in 4 5; do timeout 200 python tools/ncu_cfg_target.py $c > gpurun_out/o_c$c.log 2>&1; done
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:spr_join_score -c 1 --csv --log-file gpurun_out/o_join_c2.csv python tools/ncu_step_target.py default 2 > gpurun_out/o_ncu.log 2>&1
