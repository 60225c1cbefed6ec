"""Soak test (developer aid, GPU): the pair-join scorer against the lattice kernels of the same library on many random
map pairs -- winner of the whole lattice, and every per-hypothesis count of a random slice -- without the cost of the CPU
oracle.  usage: soak_join_vs_lattice.py [seconds] [seed0]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slide_slam_b200 import synth  # noqa: E402
from slide_slam_b200.place_recognition import PlaceRecognition  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t_end = time.time() + budget
n_ok = n_hyp = 0
while time.time() < t_end:
    rng = np.random.default_rng(90000 + seed)
    n = int(rng.integers(30, 800))
    n_b = int(rng.integers(10, 800)) if seed % 3 == 0 else n
    ref, qry, _ = synth.make_pair(n, seed=50000 + seed, classes=str(rng.choice(["five", "forest_urban"])), overlap=float(rng.uniform(0.0, 0.9)),
                                  outlier_frac=float(rng.choice([0.0, 0.1, 0.4])), sigma=float(rng.choice([0.0, 0.05, 0.3])), n_b=n_b)
    if seed % 5 == 0:   # clusters of near-duplicates
        k = min(40, len(ref))
        ref[:k, 1:3] = ref[rng.integers(0, len(ref), k), 1:3] + rng.normal(0, 0.2, (k, 2))
    step = float(rng.choice([0.5, 0.5, 1.0]))
    ros = {"search_xy_step_size": step, "search_yaw_step_size_degrees": float(rng.choice([5.0, 10.0, 30.0])),
           "match_threshold_position": float(step * rng.choice([0.5, 1.0, 1.0, 1.5, 3.0])), "match_threshold_dimension": float(rng.choice([1.0, 0.3])),
           "ignore_dimension": int(seed % 4 == 1), "min_num_inliers": 5, "disable_yaw_search": int(seed % 7 == 6)}
    if os.environ.get("SOAK_VERBOSE"):
        print("seed", seed, "n", len(ref), len(qry), ros, flush=True)
    pj, pl = PlaceRecognition(ros), PlaceRecognition(ros, engine="lattice")
    fj = pj.findTransformation(ref, qry)
    if os.environ.get("SOAK_VERBOSE"):
        print("  join ok", fj[3].match.kernel_ms, fj[3].match.hypotheses_scored, flush=True)
    fl = pl.findTransformation(ref, qry)
    if os.environ.get("SOAK_VERBOSE"):
        print("  lattice ok", fl[3].match.kernel_ms, flush=True)
    ij, il = fj[3], fl[3]
    assert ij.match.search_mode == 2 and il.match.search_mode in (0, 1), (seed, ij.match.search_mode, il.match.search_mode)
    assert (fj[0], ij.best_num_inliers, ij.match.best_hyp_index) == (fl[0], il.best_num_inliers, il.match.best_hyp_index), \
        (seed, ij.best_num_inliers, ij.match.best_hyp_index, il.best_num_inliers, il.match.best_hyp_index)
    assert fj[4].tolist() == fl[4].tolist() and fj[5].tolist() == fl[5].tolist() and fj[1].tolist() == fl[1].tolist(), seed
    sref, sqry = ref.copy(), qry.copy()
    sref[:, 1:3] -= np.array(ij.centroid_ref[:]); sqry[:, 1:3] -= np.array(ij.centroid_qry[:])
    pj.prepare(sref, sqry, ij.half_x, ij.half_y); pl.prepare(sref, sqry, il.half_x, il.half_y)
    nt = ij.match.n_translations
    if nt > 0:
        tb = int(rng.integers(0, nt)); te = min(nt, tb + int(rng.integers(1, 4000)))
        _, cj = pj.search(tb, te, want_counts=True)
        _, cl = pl.search(tb, te, want_counts=True, exhaustive=True)
        bad = np.nonzero(cj != cl)[0]
        assert bad.size == 0, (seed, tb, te, bad[:5], cj[bad[:5]], cl[bad[:5]])
        n_hyp += len(cj)
    n_ok += 1
    if n_ok % 10 == 0:
        print(f"{n_ok} pairs, seed {seed}, {n_hyp} counts compared", flush=True)
    seed += 1
    pj.close(); pl.close()
print(f"soak ok: {n_ok} map pairs (seeds up to {seed - 1}), {n_hyp} per-hypothesis counts compared, both engines agree", flush=True)
