#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/s_ab.log
timeout 600 python -m pytest tests/test_gpu_join.py -m gpu -x -q -k "not config3 and not config4 and not config5" > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
timeout 200 python tools/ab_search.py 2 10 4 >> gpurun_out/s_ab.log 2>&1
timeout 300 python tools/ab_search.py 3 2 4 >> gpurun_out/s_ab.log 2>&1
for c in 4 5; do timeout 200 python tools/ncu_cfg_target.py $c 2>&1 | tail -1 >> gpurun_out/s_ab.log; done
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:spr_join_score -c 1 --csv --log-file gpurun_out/s_join_c2.csv python tools/ncu_step_target.py join 2 > gpurun_out/s_ncu.log 2>&1
