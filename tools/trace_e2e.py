"""Developer aid: per-phase host timings of findTransformation on two alternating config-2 pairs
(run with SLIDE_PR_TRACE=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slide_slam_b200 import synth
from slide_slam_b200.place_recognition import PlaceRecognition
ROS = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 5.0, "match_threshold_position": 0.5,
       "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 15}
pr = PlaceRecognition(ROS)
pairs = [synth.make_pair(2000, seed=1002, classes="five", outlier_frac=0.1)[:2], synth.make_pair(2000, seed=1002 + 100000, classes="five", outlier_frac=0.1)[:2]]
for i in range(12):
    ref, qry = pairs[i & 1]
    t0 = time.perf_counter()
    out = pr.findTransformation(ref, qry)
    print("wall ms", (time.perf_counter() - t0) * 1e3, "kernel", out[3].match.kernel_ms, "prepare", out[3].match.prepare_ms, flush=True)
