#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spr_join_score -c 1 -o gpurun_out/p_join_c2 python tools/ncu_step_target.py default 2 > gpurun_out/p_ncu.log 2>&1
timeout 300 ncu -i gpurun_out/p_join_c2.ncu-rep --page raw --csv > gpurun_out/p_join_c2_raw.csv 2> gpurun_out/p_err1.log
timeout 300 ncu -i gpurun_out/p_join_c2.ncu-rep --page source --csv > gpurun_out/p_join_c2_src.csv 2> gpurun_out/p_err2.log
