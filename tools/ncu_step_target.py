"""ncu target: exactly ONE search step of bench.py's config-2 workload (no findTransformation before it),
so that the launch list of the process is the launch list of a step.
usage: ncu_step_target.py join|bound|exhaustive [config]   (default = join: the library's default search)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from slide_slam_b200.place_recognition import PlaceRecognition  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "join"
if mode == "default":
    mode = "join"
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
(ref, qry, _), _ = bench.workload(cfg, 0)
r = bench._ranges(ref, qry)
sref, sqry = ref.copy(), qry.copy()
sref[:, 1:3] -= r["centroid_ref"]; sqry[:, 1:3] -= r["centroid_qry"]
pr = PlaceRecognition(bench.ROS)
pr.prepare(sref, sqry, r["half_x"], r["half_y"])
res, _ = pr.search(exhaustive=(mode == "exhaustive"), engine={"join": "join", "bound": "lattice"}.get(mode))
print(f"mode={mode} cfg={cfg} kernel_ms={res.kernel_ms:.3f} hyp={res.hypotheses_scored} best={res.best_num_inliers} launches={res.gpu_launches}")
