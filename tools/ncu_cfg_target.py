"""ncu target: one findTransformation of BASELINE config 4 (pair 0-1 of 5000 landmarks) or config 5 (one 300-landmark
query against 50000, index prebuilt).  usage: ncu_cfg_target.py 4|5"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from slide_slam_b200 import synth  # noqa: E402
from slide_slam_b200.place_recognition import PlaceRecognition  # noqa: E402
cfg = int(sys.argv[1])
pr = PlaceRecognition(bench.ROS)
if cfg == 4:
    maps = synth.config_robots(8, 5000)
    a, b = maps[0], maps[1]
else:
    a, qs = synth.config_stream(50000, n_queries=2, n_sub=300)
    b = qs[1]
    pr.findTransformation(a, qs[0])
for i in range(2):
    t0 = time.perf_counter()
    f, xyz, tf, info, ri, qi = pr.findTransformation(a, b)
    print(f"cfg{cfg} wall_ms={(time.perf_counter()-t0)*1e3:.2f} kernel_ms={info.match.kernel_ms:.2f} prepare_ms={info.match.prepare_ms:.2f} best={info.best_num_inliers} "
          f"hyp={info.match.hypotheses_scored} launches={info.match.gpu_launches} reuse={info.match.reuse}", flush=True)
