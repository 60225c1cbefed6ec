"""Small ncu target: prepares one BASELINE config and runs the lattice search a few times."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slide_slam_b200 import synth  # noqa: E402
from slide_slam_b200.place_recognition import PlaceRecognition  # noqa: E402

ROS = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 5.0, "match_threshold_position": 0.5,
       "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 15}

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "default"   # default (bound-and-verify) | exhaustive
pr = PlaceRecognition(ROS)
ref, qry, truth = synth.config_pair(cfg, n)
found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(ref, qry)
sref, sqry = ref.copy(), qry.copy()
sref[:, 1:3] -= np.array(info.centroid_ref[:])
sqry[:, 1:3] -= np.array(info.centroid_qry[:])
pr.prepare(sref, sqry, info.half_x, info.half_y)
for i in range(reps):
    res, _ = pr.search(collect_stats=(mode == "stats" and i == reps - 1), exhaustive=(mode == "exhaustive"))
    print(f"cfg{cfg} n={len(ref)} kernel_ms={res.kernel_ms:.3f} hyp={res.hypotheses_scored} "
          f"hyp/s={res.hypotheses_scored / res.kernel_ms * 1e3:.3e} hits={res.filter_hits} "
          f"groups probed={res.groups_probed} skipped={res.groups_skipped}", flush=True)
print("best", info.best_num_inliers, "found", found)
