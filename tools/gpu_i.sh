#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_cache.py -m gpu -q -x --durations=5 > gpurun_out/i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/i_pytest.log
timeout 120 python tools/ab_search.py 2 10 > gpurun_out/i_ab.log 2>&1
timeout 300 python tools/ab_search.py 3 2 >> gpurun_out/i_ab.log 2>&1
echo "== refine off" >> gpurun_out/i_ab.log; SLIDE_PR_REFINE_MIN=-1 timeout 300 python tools/ab_search.py 3 2 >> gpurun_out/i_ab.log 2>&1
echo "== default" >> gpurun_out/i_c4.log; timeout 120 python tools/ncu_cfg_target.py 4 >> gpurun_out/i_c4.log 2>&1
echo "== no refine" >> gpurun_out/i_c4.log; SLIDE_PR_REFINE_MIN=-1 timeout 120 python tools/ncu_cfg_target.py 4 >> gpurun_out/i_c4.log 2>&1
timeout 300 python tools/shard_emul.py 3 8 > gpurun_out/i_shard.log 2>&1
