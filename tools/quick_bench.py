"""Developer timing probe (not the contract bench): kernel time of the lattice search on the
BASELINE configs for both kernel variants."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slide_slam_b200 import synth  # noqa: E402
from slide_slam_b200.place_recognition import PlaceRecognition  # noqa: E402

ROS = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 5.0, "match_threshold_position": 0.5,
       "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 15}


def run(cfg, variant, n=None, reps=3, stats=False, exhaustive=False):
    os.environ["SLIDE_PR_VARIANT"] = str(variant)
    pr = PlaceRecognition(ROS)
    ref, qry, truth = synth.config_pair(cfg, n)
    t0 = time.time()
    found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(ref, qry)
    e2e = time.time() - t0
    sref, sqry = ref.copy(), qry.copy()
    sref[:, 1:3] -= np.array(info.centroid_ref[:])
    sqry[:, 1:3] -= np.array(info.centroid_qry[:])
    pr.prepare(sref, sqry, info.half_x, info.half_y)
    ms = []
    for _ in range(reps):
        res, _ = pr.search(collect_stats=stats, exhaustive=exhaustive)
        ms.append(res.kernel_ms)
    hyp = info.match.hypotheses_scored
    print(f"cfg{cfg} n={len(ref)} variant={variant} mode={res.search_mode} best_search={res.best_num_inliers} idx={res.best_hyp_index} found={found} best={info.best_num_inliers} hyp={hyp} "
          f"kernel_ms={min(ms):.3f} ({hyp / min(ms) * 1e3:.3e} hyp/s) prepare_ms={info.match.prepare_ms:.1f} "
          f"e2e_s={e2e:.3f} hits={res.filter_hits} hit_rate={res.filter_hits / max(hyp * len(qry), 1):.5f} "
          f"yaw_err={xyz_yaw[3] - truth['yaw']:.2e}", flush=True)
    pr.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2"]
    for cfg in which:
        for variant in (1, 0):
            run(int(cfg), variant)
            run(int(cfg), variant, exhaustive=True)
        run(int(cfg), 1, stats=True, reps=1)
