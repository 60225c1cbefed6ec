"""ncu target for the SlideGraph half: generate_and_score + run_semantic_clipper on a synthetic pair.
usage: ncu_sg_target.py [n_landmarks]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slide_slam_b200 import synth  # noqa: E402
from slide_slam_b200.place_recognition import PlaceRecognition, delaunay  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ref, qry, truth = synth.make_pair(n, seed=900 + n, classes="five", overlap=0.5, sigma=0.005)
pr = PlaceRecognition({"search_xy_step_size": 0.5, "match_threshold_position": 0.5}, slidegraph={"descriptor_matching_threshold": 0.05, "seed": 1})
ir, iq = delaunay(np.ascontiguousarray(ref[:, 1:3])), delaunay(np.ascontiguousarray(qry[:, 1:3]))
tr = np.ascontiguousarray(ref[:, 1:3][ir].reshape(-1, 6)); tq = np.ascontiguousarray(qry[:, 1:3][iq].reshape(-1, 6))
lr, lq = np.ascontiguousarray(ref[ir, 0]), np.ascontiguousarray(qry[iq, 0])
pr.prepare(ref, qry, 200.0, 200.0)
for i in range(2):
    t0 = time.perf_counter()
    res, gi, _ = pr.generate_and_score(tr, tq, 0.05, lr, lq, want_lists=False)
    print(f"generate_and_score n={n} T={len(tr)},{len(tq)} matches={gi.n_matches} best={res.best_num_inliers} call_ms={(time.perf_counter()-t0)*1e3:.3f} "
          f"match_ms={gi.match_ms:.3f} kabsch_ms={gi.kabsch_ms:.3f} score_ms={gi.score_ms:.3f}", flush=True)
for i in range(2):
    t0 = time.perf_counter()
    found, tf = pr.findInterLoopClosureWithClipper(ref, qry)
    sc = pr.last_sc
    print(f"with_clipper found={found} call_ms={(time.perf_counter()-t0)*1e3:.3f} tri_matches={sc.n_triangle_matches} assoc={sc.n_associations} nnz={sc.nnz_upper} "
          f"inliers={sc.n_inliers} delaunay_ms={sc.delaunay_ms:.3f} match_ms={sc.match_ms:.3f} affinity_ms={sc.affinity_ms:.3f} solve_ms={sc.solve_ms:.3f} "
          f"yaw_err={abs(np.angle(np.exp(1j*(np.arctan2(tf[1,0],tf[0,0])-truth['yaw'])))):.2e}", flush=True)
