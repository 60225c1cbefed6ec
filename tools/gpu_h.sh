#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cache.py -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/h_pytest.log
for lib in slide_slam_b200/libslide_pr.so variants/libslide_pr_mb4.so variants/libslide_pr_mb5.so; do
  echo "== $lib" >> gpurun_out/h_c5.log
  SLIDE_PR_LIB=$lib timeout 120 python tools/ncu_cfg_target.py 5 >> gpurun_out/h_c5.log 2>&1
done
echo "== default" >> gpurun_out/h_c4.log; timeout 120 python tools/ncu_cfg_target.py 4 >> gpurun_out/h_c4.log 2>&1
echo "== no refine" >> gpurun_out/h_c4.log; SLIDE_PR_REFINE_MIN=-1 timeout 120 python tools/ncu_cfg_target.py 4 >> gpurun_out/h_c4.log 2>&1
timeout 300 python tools/shard_emul.py 3 8 > gpurun_out/h_shard.log 2>&1
timeout 300 python tools/shard_emul.py 3 2 >> gpurun_out/h_shard.log 2>&1
