#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "windowed or config2_full or every_hypothesis or bound" > gpurun_out/j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/j_pytest.log
for lib in slide_slam_b200/libslide_pr.so variants/libslide_pr_win2.so; do
  echo "== $lib" >> gpurun_out/j_c5.log
  SLIDE_PR_LIB=$lib timeout 120 python tools/ncu_cfg_target.py 5 >> gpurun_out/j_c5.log 2>&1
done
echo "== window off" >> gpurun_out/j_c5.log; SLIDE_PR_WINDOW=0 timeout 120 python tools/ncu_cfg_target.py 5 >> gpurun_out/j_c5.log 2>&1
timeout 120 python tools/ab_search.py 2 10 > gpurun_out/j_ab.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "config5" >> gpurun_out/j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/j_pytest.log
M=smsp__inst_executed.sum,gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/j_c5_launches.csv python tools/ncu_cfg_target.py 5 > gpurun_out/j_ncu5.log 2>&1
timeout 200 ncu --metrics $M --clock-control none -k regex:spr_score --csv --log-file gpurun_out/j_x_launches.csv python tools/ncu_step_target.py exhaustive > gpurun_out/j_ncux.log 2>&1
