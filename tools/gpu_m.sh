#!/bin/bash
# ncu capture of the pair-join kernel on config 2
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spr_join_score -c 1 -o gpurun_out/m_join_c2 python tools/ncu_step_target.py default 2 > gpurun_out/m_ncu.log 2>&1
timeout 300 ncu -i gpurun_out/m_join_c2.ncu-rep --page raw --csv > gpurun_out/m_join_c2_raw.csv 2> gpurun_out/m_err1.log
timeout 300 ncu -i gpurun_out/m_join_c2.ncu-rep --page source --csv > gpurun_out/m_join_c2_src.csv 2> gpurun_out/m_err2.log
