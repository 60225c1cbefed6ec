#!/bin/bash
set -u
mkdir -p gpurun_out
for lib in variants/libslide_pr_r1.so variants/libslide_pr_oldprobe.so slide_slam_b200/libslide_pr.so; do
  SLIDE_PR_LIB=$lib timeout 120 python tools/ab_search.py 2 20 >> gpurun_out/f_ab.log 2>&1
done
SLIDE_PR_LIB=variants/libslide_pr_oldprobe.so timeout 200 python tools/ab_search.py 3 2 >> gpurun_out/f_ab.log 2>&1
timeout 200 python tools/ab_search.py 3 2 >> gpurun_out/f_ab.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest.log
M=smsp__inst_executed.sum,gpu__time_duration.sum
for lib in variants/libslide_pr_r1.so slide_slam_b200/libslide_pr.so; do
  SLIDE_PR_LIB=$lib timeout 200 ncu --metrics $M --clock-control none -k regex:spr_bound --csv --log-file gpurun_out/f_launches_$(basename $lib .so).csv python tools/ab_search.py 2 0 > gpurun_out/f_ncu_$(basename $lib .so).log 2>&1
done
