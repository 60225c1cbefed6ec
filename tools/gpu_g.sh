#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/shard_emul.py 3 8 > gpurun_out/g_shard.log 2>&1
timeout 120 python tools/ncu_cfg_target.py 4 > gpurun_out/g_c4.log 2>&1
timeout 120 python tools/ncu_cfg_target.py 5 > gpurun_out/g_c5.log 2>&1
SLIDE_PR_TRACE=1 timeout 120 python tools/ncu_cfg_target.py 5 > gpurun_out/g_c5_trace.log 2>&1
SLIDE_PR_TRACE=1 timeout 120 python tools/ncu_cfg_target.py 4 > gpurun_out/g_c4_trace.log 2>&1
M=smsp__inst_executed.sum,gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/g_c4_launches.csv python tools/ncu_cfg_target.py 4 > gpurun_out/g_ncu4.log 2>&1
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/g_c5_launches.csv python tools/ncu_cfg_target.py 5 > gpurun_out/g_ncu5.log 2>&1
