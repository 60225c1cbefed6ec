"""Developer aid: kernel time of the default (bound-and-verify) and the exhaustive search on
config-2-sized maps with and without overlap."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slide_slam_b200 import synth
from slide_slam_b200.place_recognition import PlaceRecognition
ROS = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 5.0, "match_threshold_position": 0.5,
       "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 15}
pr = PlaceRecognition(ROS)
ref = synth.make_pair(2000, seed=1002, classes="five", outlier_frac=0.1)[0]
qry = synth.make_pair(2000, seed=555, classes="five", outlier_frac=0.1)[1]
for name, (r, q) in {"unrelated": (ref, qry), "overlap": synth.make_pair(2000, seed=1002, classes="five", outlier_frac=0.1)[:2]}.items():
    r = r.copy(); q = q.copy()
    r[:, 1:3] -= r[:, 1:3].mean(0); q[:, 1:3] -= q[:, 1:3].mean(0)
    half = 1.2 * max(np.abs(r[:, 1:3]).max(), np.abs(q[:, 1:3]).max())
    pr.prepare(r, q, half, half)
    for mode in (False, True):
        ms = []
        for _ in range(3):
            res, _ = pr.search(exhaustive=mode)
            ms.append(res.kernel_ms)
        print(name, "exhaustive" if mode else "bound-and-verify", "best", res.best_num_inliers, "idx", res.best_hyp_index, "kernel_ms %.3f" % min(ms), flush=True)
