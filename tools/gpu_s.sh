#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cache.py -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
cat > /tmp/c4.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from slide_slam_b200 import synth
from slide_slam_b200.place_recognition import PlaceRecognition
pr = PlaceRecognition(bench.ROS)
maps = synth.config_robots(8, 5000)
pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
for rep in range(4):
    t0 = time.perf_counter()
    outs = pr.findTransformationBatch(maps, pairs)
    dt = time.perf_counter() - t0
    print("batch", rep, "s", dt, "pairs/s", 28 / dt, "kernel_ms_sum", sum(o.match.kernel_ms for o in outs), "prepare_ms_sum", sum(o.match.prepare_ms for o in outs), flush=True)
PY
timeout 300 python /tmp/c4.py > gpurun_out/s_c4.log 2>&1
