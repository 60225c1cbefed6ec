"""CPU suite for the product's host logic (slide_slam_b200/csrc/spr_host.cpp + spr_core.h):
the lattice/chunk builder and the occupancy-bitmap + candidate-list index are driven through a
test-only single-thread emulation of the kernel's per-thread code (tests/emu) and compared with
the oracle hypothesis by hypothesis.  The emulation is a test harness, not a product path."""
import numpy as np
import pytest

import spr_helpers as H
from oracle import pyoracle as O


def _check_counts(op, ref, qry, hx, hy):
    p = H.to_capi_params(op)
    lat = O.enumerate_lattice(op, hx, hy)
    if lat is None:
        r = H.emu_match_maps(p, ref, qry, hx, hy)
        assert r["rc"] == 2  # SLIDE_PR_SANITY_RETURN
        return None
    nt, ny = len(lat[0]), len(lat[3])
    n, status, tx, ty, yaw = H.emu_lattice(p, hx, hy, nt)
    assert status == 0 and n == nt
    assert np.array_equal(tx, lat[0]) and np.array_equal(ty, lat[1]) and np.array_equal(yaw, lat[3])  # bit-exact
    o = O.match_maps(op, ref, qry, hx, hy, want_counts=True)
    for engine in ("lattice", "join"):   # the chunked bitmap kernels' code and the pair-join scorer's
        e = H.emu_match_maps(p, ref, qry, hx, hy, n_counts=nt * ny, engine=engine)
        assert e["rc"] == 0, e["err"]
        assert e["hypotheses_scored"] == o["hypotheses_scored"] == nt * ny
        bad = np.nonzero(e["counts"] != o["counts"])[0]
        assert bad.size == 0, f"{engine}: first mismatching hypotheses {bad[:5]}: {e['counts'][bad[:5]]} != {o['counts'][bad[:5]]}"
        assert e["best_num_inliers"] == o["best_num_inliers"] and e["best_hyp_index"] == o["best_hyp_index"], engine
    return e


def test_indoor_fixture_all_counts():
    maps, cases = H.golden_maps(), H.golden_cases()
    for name in ("indoor01_forest_yaml_nodim", "indoor12_forest_yaml_nodim"):
        c = cases[name]
        ref, qry = H.shifted_maps(maps, c)
        e = _check_counts(O.make_params(**c["params"]), ref, qry, c["half_x"], c["half_y"])
        assert e["best_hyp_index"] == c["best_hyp_index"]


def test_golden_slices_of_large_cases():
    maps, cases, counts = H.golden_maps(), H.golden_cases(), H.golden_counts()
    for name in ("parking01_forest_yaml", "c1_forest_yaml", "prtest_inter_lsq1"):
        c = cases[name]
        ref, qry = H.shifted_maps(maps, c)
        op = O.make_params(**c["params"])
        lo, hi = counts[name + "__slice"]
        ny = len(O.enumerate_lattice(op, 6.0, 6.0)[3])  # yaw candidates do not depend on the range
        # translation-aligned sub-slice of the golden hypothesis slice
        tb, te = -(-int(lo) // ny), int(hi) // ny
        want = counts[name][tb * ny - int(lo): te * ny - int(lo)]
        for engine in ("lattice", "join"):
            e = H.emu_match_maps(H.to_capi_params(op), ref, qry, c["half_x"], c["half_y"], tb, te, n_counts=(te - tb) * ny, engine=engine)
            assert e["rc"] == 0, e["err"]
            assert np.array_equal(e["counts"], want), engine


@pytest.mark.parametrize("seed", range(12))
def test_random_maps_every_hypothesis(seed):
    rng = np.random.default_rng(100 + seed)
    n_ref, n_qry = int(rng.integers(1, 60)), int(rng.integers(1, 50))
    ref, qry = H.random_maps(rng, n_ref, n_qry, extent=float(rng.uniform(3, 15)), n_labels=int(rng.integers(1, 5)),
                             grid=(0.25 if seed % 3 == 0 else None))
    step = float(rng.choice([0.25, 0.5, 0.5, 1.0, 0.3]))
    thr = float(rng.choice([0.5, 0.75, 0.3, 1.1]))
    op = O.make_params(match_xy_step_size=step, yaw_step_deg=float(rng.choice([30.0, 45.0, 17.0])),
                       match_threshold=thr, match_threshold_dimension=float(rng.choice([1.0, 0.3])),
                       ignore_dimension=int(seed % 4 == 1), disable_yaw_search=int(seed % 5 == 4))
    hx = float(rng.uniform(4, 14))
    hy = hx if seed % 2 else float(rng.uniform(4, 14))
    _check_counts(op, ref, qry, hx, hy)


@pytest.mark.parametrize("seed", range(24))
def test_pair_join_scorer_dense_and_clustered_maps(seed):
    """The pair-join scorer's own hard cases, every hypothesis against the oracle: clusters of near-duplicate reference
    landmarks (first-match attribution through the lower-index neighbour lists), many query landmarks per block (several
    quads, full pair lists), thresholds from a third of the step to four steps, one label or many, rectangular ranges."""
    rng = np.random.default_rng(7000 + seed)
    n_ref, n_qry = int(rng.integers(20, 160)), int(rng.integers(20, 120))
    n_labels = int(rng.choice([1, 1, 2, 6]))
    ref, qry = H.random_maps(rng, n_ref, n_qry, extent=float(rng.uniform(4, 12)), n_labels=n_labels,
                             dup_frac=float(rng.choice([0.0, 0.3, 0.6])), grid=(0.125 if seed % 4 == 0 else None))
    if seed % 3 == 0:   # a tight cluster: up to 12 reference landmarks within one threshold of each other
        k = min(12, n_ref)
        ref[:k, 1:3] = ref[0, 1:3] + rng.normal(0, 0.1, (k, 2))
        ref[:k, 0] = ref[0, 0]
    step = float(rng.choice([0.5, 0.25, 0.4]))
    thr = float(step * rng.choice([0.34, 1.0, 1.0, 1.7, 4.0]))
    op = O.make_params(match_xy_step_size=step, yaw_step_deg=float(rng.choice([45.0, 60.0, 36.0])), match_threshold=thr,
                       match_threshold_dimension=float(rng.choice([1.0, 0.4])), ignore_dimension=int(seed % 5 == 2),
                       disable_yaw_search=int(seed % 6 == 5))
    hx = float(rng.uniform(3, 9))
    hy = hx if seed % 2 else float(rng.uniform(3, 9))
    p = H.to_capi_params(op)
    lat = O.enumerate_lattice(op, hx, hy)
    if lat is None:
        pytest.skip("range below one lattice step")
    n = len(lat[0]) * len(lat[3])
    o = O.match_maps(op, ref, qry, hx, hy, want_counts=True)
    e = H.emu_match_maps(p, ref, qry, hx, hy, n_counts=n, engine="join")
    assert e["rc"] == 0, e["err"]
    bad = np.nonzero(e["counts"] != o["counts"])[0]
    assert bad.size == 0, f"first mismatching hypotheses {bad[:5]}: {e['counts'][bad[:5]]} != {o['counts'][bad[:5]]}"
    assert (e["best_num_inliers"], e["best_hyp_index"]) == (o["best_num_inliers"], o["best_hyp_index"])


def test_edge_cases():
    op = O.make_params(match_xy_step_size=0.5, yaw_step_deg=45.0)
    rng = np.random.default_rng(1)
    ref, qry = H.random_maps(rng, 10, 8, extent=4.0)
    # empty maps: every hypothesis scores 0; the first one wins (PR.cpp:125,361)
    for r, q in ((np.zeros((0, 7)), qry), (ref, np.zeros((0, 7))), (np.zeros((0, 7)), np.zeros((0, 7)))):
        e = _check_counts(op, r, q, 6.0, 6.0)
        assert e["best_num_inliers"] == 0 and e["best_hyp_index"] == 0
    # query labels that never occur in the reference, NaN labels, -0.0 vs 0.0 labels
    q2 = qry.copy()
    q2[:, 0] = 77.0
    assert _check_counts(op, ref, q2, 6.0, 6.0)["best_num_inliers"] == 0
    r3, q3 = ref.copy(), qry.copy()
    r3[0, 0] = np.nan
    q3[0, 0] = np.nan
    r3[1, 0] = -0.0
    q3[1, 0] = 0.0
    _check_counts(op, r3, q3, 6.0, 6.0)
    # more distinct labels than the reference-side builder collects by linear search (16): its sorted-vector path,
    # with NaN and -0.0 labels among them, in descending and in shuffled order of appearance
    r6, q6 = H.random_maps(rng, 60, 50, extent=4.0)
    r6[:, 0] = np.arange(60)[::-1] % 23 - 3.0
    q6[:, 0] = rng.integers(-3, 20, 50)
    r6[5, 0] = np.nan; r6[7, 0] = -0.0; q6[3, 0] = np.nan
    assert _check_counts(op, r6, q6, 6.0, 6.0)["best_num_inliers"] > 0
    r6[:, 0] = rng.permutation(60) % 16 * 0.5      # exactly 16 labels: still the linear-search path
    q6[:, 0] = rng.integers(0, 16, 50) * 0.5
    assert _check_counts(op, r6, q6, 6.0, 6.0)["best_num_inliers"] > 0
    # all landmarks identical (every reference object matches every query object)
    r4 = np.tile(np.array([[2, 1.0, 1.0, 0, 0.5, 0, 0]], float), (6, 1))
    q4 = np.tile(np.array([[2, 0.0, 0.0, 0, 0.5, 0, 0]], float), (5, 1))
    assert _check_counts(op, r4, q4, 6.0, 6.0)["best_num_inliers"] == 5
    # more identical query landmarks than one round of the pair-join kernel holds: its u8 counters must not wrap
    q5 = np.tile(np.array([[2, 0.0, 0.0, 0, 0.5, 0, 0]], float), (700, 1))
    assert _check_counts(O.make_params(match_xy_step_size=0.5, yaw_step_deg=90.0), r4, q5, 3.0, 3.0)["best_num_inliers"] == 700
    # maps far from the origin (intra mode searches in the map frame: no centroid shift): coordinates of 3e5 m, where
    # one ulp is 6e-11 m -- the enumeration margins of the pair-join scorer scale with the coordinate magnitude
    far_r, far_q = ref.copy(), qry.copy()
    far_r[:, 1] += 3.0e5; far_r[:, 2] -= 2.0e5; far_q[:, 1] += 3.0e5; far_q[:, 2] -= 2.0e5
    _check_counts(O.make_params(match_xy_step_size=0.5, yaw_step_deg=90.0, disable_yaw_search=1), far_r, far_q, 6.0, 6.0)
    # threshold far above the step: 15 x 15 lattice samples per pair (general path of the pair-join kernel)
    _check_counts(O.make_params(match_xy_step_size=0.1, yaw_step_deg=60.0, match_threshold=0.75), ref, qry, 4.0, 4.0)
    # sanity-check early return and zero rings
    assert _check_counts(op, ref, qry, 0.3, 0.3) is None
    for engine in ("lattice", "join"):
        e = H.emu_match_maps(H.to_capi_params(op), ref, qry, 0.0, 0.0, engine=engine)
        assert e["rc"] == 0 and e["hypotheses_scored"] == 0 and e["best_num_inliers"] == -10000
    # rectangular half ranges (disable_yaw_search keeps x and y ranges apart, PR.cpp:777-782)
    op2 = O.make_params(match_xy_step_size=0.5, disable_yaw_search=1)
    _check_counts(op2, ref, qry, 12.0, 5.5)
    # threshold <= 0 can never match (strict '<' on a non-negative distance)
    _check_counts(O.make_params(match_xy_step_size=0.5, yaw_step_deg=90.0, match_threshold=0.0), ref, qry, 5.0, 5.0)
    # intra-mode defaults: the (0, 0) translation is skipped
    op3 = O.make_params(inter_loop_closure=0, yaw_half_range_intra_deg=10.0, yaw_step_deg=2.0)
    _check_counts(op3, ref, qry, 5.0, 5.0)


def test_thresholds_are_exact():
    import ctypes as C
    import math
    L = H.emu_lib()
    # (sqrt(d2) < thr) == (d2 < Tstar) around the boundary -- checked through the emulation's
    # distance test on a two-landmark scene placed exactly at / next to the threshold
    for thr in (0.5, 0.75, 0.3, 1.0 / 3.0, 2.0 ** 0.5):
        op = O.make_params(match_xy_step_size=1.0, disable_yaw_search=1, match_threshold=thr, ignore_dimension=1)
        for eps in (-2, -1, 0, 1, 2):
            d = thr
            for _ in range(abs(eps)):
                d = math.nextafter(d, math.inf if eps > 0 else -math.inf)
            ref = np.array([[1, d, 0, 0, 0, 0, 0]], float)
            qry = np.array([[1, 0, 0, 0, 0, 0, 0]], float)
            _check_counts(op, ref, qry, 10.0, 10.0)
