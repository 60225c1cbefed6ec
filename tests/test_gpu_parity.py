"""GPU parity suite (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle and
the committed golden vectors.  Integer results (inlier counts, correspondences, best hypothesis,
the winner's R_t entries) are compared bit-exactly; refined transforms within 1e-5 relative."""
import ctypes as C
import os

import numpy as np
import pytest

import spr_helpers as H
from oracle import pyoracle as O
from slide_slam_b200 import capi, synth
from slide_slam_b200.place_recognition import PlaceRecognition

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # BASELINE.json north_star: transform parameters within 1e-5 relative


@pytest.fixture(autouse=True, scope="module")
def _lattice_engine():
    """This module pins the lattice kernels (bound-and-verify, exhaustive verification, bounds): handles are
    created with SLIDE_PR_ENGINE=lattice.  The default engine (pair-join scorer) is covered by tests/test_gpu_join.py."""
    old = os.environ.get("SLIDE_PR_ENGINE")
    os.environ["SLIDE_PR_ENGINE"] = "lattice"
    yield
    if old is None:
        os.environ.pop("SLIDE_PR_ENGINE", None)
    else:
        os.environ["SLIDE_PR_ENGINE"] = old


@pytest.fixture(scope="module")
def gold():
    return H.golden_maps(), H.golden_cases(), H.golden_counts()


def make_pr(params: dict, variant: int | None = None, **members):
    if variant is not None:
        os.environ["SLIDE_PR_VARIANT"] = str(variant)
    else:
        os.environ.pop("SLIDE_PR_VARIANT", None)
    pr = PlaceRecognition(H.rosparams_from_golden(params))
    os.environ.pop("SLIDE_PR_VARIANT", None)
    if "use_lsq" in params:
        pr.use_lsq = bool(params["use_lsq"])
    if "inter_loop_closure" in params:
        pr.inter_loop_closure = bool(params["inter_loop_closure"])
    for k, v in members.items():
        setattr(pr, k, v)
    return pr


ALL_CASES = ["indoor01_sloam_yaml", "indoor01_forest_yaml_nodim", "indoor02_sloam_yaml", "indoor02_forest_yaml_nodim",
             "indoor12_sloam_yaml", "indoor12_forest_yaml_nodim", "indoor10_sloam_yaml", "indoor10_forest_yaml_nodim",
             "indoor01_sloam_yaml_nolsq", "indoor01_defaults_2deg", "indoor01_noyaw", "parking01_forest_yaml",
             "parking02_forest_yaml", "forest01_forest_yaml", "prtest_inter_lsq1", "prtest_intra_lsq1",
             "prtest_inter_lsq0", "prtest_intra_lsq0", "c1_forest_yaml"]


@pytest.mark.parametrize("name", ALL_CASES)
def test_find_transformation_matches_golden(gold, name):
    maps, cases, _ = gold
    c = cases[name]
    pr = make_pr(c["params"])
    found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(maps[c["ref"]], maps[c["qry"]])
    assert found == c["found"]
    assert info.best_num_inliers == c["best_num_inliers"]
    assert info.match.best_hyp_index == c["best_hyp_index"]
    assert info.match.hypotheses_scored == c["hypotheses_scored"]
    assert ri.tolist() == c["ref_idx"] and qi.tolist() == c["qry_idx"]
    assert list(info.R_t) == c["R_t"]                      # lattice winner: bit-exact
    assert info.half_x == c["half_x"] and info.half_y == c["half_y"]
    assert list(info.centroid_ref) == c["centroid_ref"] and list(info.centroid_qry) == c["centroid_qry"]
    if found:
        np.testing.assert_allclose(xyz_yaw, c["xyz_yaw"], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(tf.ravel(), c["transform"], rtol=RTOL, atol=1e-9)
    pr.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_every_hypothesis_count_indoor(gold, variant):
    maps, cases, counts = gold
    c = cases["indoor01_forest_yaml_nodim"]
    ref, qry = H.shifted_maps(maps, c)
    pr = make_pr(c["params"], variant=variant)
    pr.prepare(ref, qry, c["half_x"], c["half_y"])
    nt, ny, _ = pr.lattice_info()
    res, got = pr.search(0, nt, want_counts=True, collect_stats=True)
    assert np.array_equal(got, counts["indoor01_forest_yaml_nodim"])
    assert res.best_hyp_index == c["best_hyp_index"] and res.best_num_inliers == c["best_num_inliers"]
    assert res.hypotheses_scored == nt * ny and res.filter_hits > 0
    pr.close()


@pytest.mark.parametrize("name", ["parking01_forest_yaml", "c1_forest_yaml", "prtest_inter_lsq1"])
def test_count_slices_of_large_cases(gold, name):
    maps, cases, counts = gold
    c = cases[name]
    ref, qry = H.shifted_maps(maps, c)
    pr = make_pr(c["params"])
    pr.prepare(ref, qry, c["half_x"], c["half_y"])
    nt, ny, _ = pr.lattice_info()
    lo, hi = (int(v) for v in counts[name + "__slice"])
    tb, te = -(-lo // ny), hi // ny
    res, got = pr.search(tb, te, want_counts=True)
    assert np.array_equal(got, counts[name][tb * ny - lo: te * ny - lo])
    assert res.hypotheses_scored == (te - tb) * ny
    # the slice's own winner must be the first maximum of the slice
    sl = counts[name][tb * ny - lo: te * ny - lo]
    assert res.best_num_inliers == int(sl.max()) and res.best_hyp_index == tb * ny + int(np.argmax(sl))
    pr.close()


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("seed", range(10))
def test_random_maps_every_hypothesis(seed, variant):
    rng = np.random.default_rng(100 + seed)
    n_ref, n_qry = int(rng.integers(1, 60)), int(rng.integers(1, 50))
    ref, qry = H.random_maps(rng, n_ref, n_qry, extent=float(rng.uniform(3, 15)), n_labels=int(rng.integers(1, 5)),
                             grid=(0.25 if seed % 3 == 0 else None))
    step = float(rng.choice([0.25, 0.5, 0.5, 1.0, 0.3]))
    thr = float(rng.choice([0.5, 0.75, 0.3, 1.1]))
    kw = dict(match_xy_step_size=step, yaw_step_deg=float(rng.choice([30.0, 45.0, 17.0])), match_threshold=thr,
              match_threshold_dimension=float(rng.choice([1.0, 0.3])), ignore_dimension=int(seed % 4 == 1),
              disable_yaw_search=int(seed % 5 == 4))
    hx = float(rng.uniform(4, 14))
    hy = hx if seed % 2 else float(rng.uniform(4, 14))
    op = O.make_params(**kw)
    want = O.match_maps(op, ref, qry, hx, hy, want_counts=True)
    pr = make_pr(kw, variant=variant)
    pr.prepare(ref, qry, hx, hy)
    nt, ny, _ = pr.lattice_info()
    res, got = pr.search(0, nt, want_counts=True)
    assert np.array_equal(got, want["counts"])
    assert (res.best_num_inliers, res.best_hyp_index) == (want["best_num_inliers"], want["best_hyp_index"])
    m = pr.MatchMaps(ref, qry, hx, hy)
    assert m.best_num_inliers == want["best_num_inliers"]
    assert m.ref_idx.tolist() == want["ref_idx"].tolist() and m.qry_idx.tolist() == want["qry_idx"].tolist()
    assert m.R_t.ravel().tolist() == want["R_t"].ravel().tolist()
    # MatchMaps returns the rows themselves (PR.cpp:344-350): reference rows and ORIGINAL query rows
    assert np.array_equal(m.map_objects_matched, ref[want["ref_idx"]][:, :4].reshape(-1, 4))
    assert np.array_equal(m.detection_objects_matched, qry[want["qry_idx"]][:, :4].reshape(-1, 4))
    pr.close()


@pytest.fixture(params=["default", "always_refine", "never_refine"])
def refine_mode(request):
    """The bounds of the candidate double groups are refined against half-cell bitmap variants when
    many of them remain (SLIDE_PR_REFINE_MIN, default 32): run with the default, always and never."""
    val = {"default": None, "always_refine": "0", "never_refine": "-1"}[request.param]
    if val is None:
        os.environ.pop("SLIDE_PR_REFINE_MIN", None)
    else:
        os.environ["SLIDE_PR_REFINE_MIN"] = val
    yield request.param
    os.environ.pop("SLIDE_PR_REFINE_MIN", None)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("seed", range(6))
def test_upper_bounds_dominate_exact_counts(seed, variant, refine_mode):
    """Bound phase of the bound-and-verify search: the bound of EVERY hypothesis is >= its exact
    inlier count (and <= the number of query landmarks), so pruning can never drop the winner."""
    rng = np.random.default_rng(700 + seed)
    n_ref, n_qry = int(rng.integers(5, 200)), int(rng.integers(5, 120))
    ref, qry = H.random_maps(rng, n_ref, n_qry, extent=float(rng.uniform(5, 20)), n_labels=int(rng.integers(1, 7)),
                             grid=(0.25 if seed % 3 == 0 else None))
    kw = dict(match_xy_step_size=float(rng.choice([0.25, 0.5, 1.0])), yaw_step_deg=float(rng.choice([30.0, 45.0])),
              match_threshold=float(rng.choice([0.5, 0.75, 0.3, 1.1])), match_threshold_dimension=1.0,
              ignore_dimension=int(seed % 2))
    hx = float(rng.uniform(4, 12))
    pr = make_pr(kw, variant=variant)
    pr.prepare(ref, qry, hx, hx)
    nt, ny, _ = pr.lattice_info()
    res_e, exact = pr.search(0, nt, want_counts=True)
    res_b, bound = pr.search(0, nt, want_counts=True, bounds_only=True)
    assert res_e.search_mode == 0 and res_b.search_mode == 1
    assert exact.shape == bound.shape and (exact >= 0).all()
    assert (bound >= exact).all() and (bound <= n_qry).all()
    # a slice of the lattice (re-chunked) gives the same bounds
    lo, hi = nt // 3, nt // 3 + max(nt // 5, 1)
    _, part = pr.search(lo, hi, want_counts=True, bounds_only=True)
    if refine_mode == "never_refine":
        assert np.array_equal(part, bound[lo * ny:hi * ny])
    else:  # which groups get the refined (tighter) bound depends on the running best of the searched slice
        assert (part >= exact[lo * ny:hi * ny]).all() and (part <= n_qry).all()
    res_p, _ = pr.search()            # default: bound-and-verify
    res_x, _ = pr.search(exhaustive=True)
    assert res_p.search_mode == 1 and res_x.search_mode == 0
    assert (res_p.best_num_inliers, res_p.best_hyp_index) == (res_x.best_num_inliers, res_x.best_hyp_index)
    assert (res_x.best_num_inliers, res_x.best_hyp_index) == (res_e.best_num_inliers, res_e.best_hyp_index)
    pr.close()


@pytest.mark.parametrize("kind", ["overlap", "unrelated", "one_label", "tiny_query"])
def test_bound_and_verify_equals_exhaustive(kind, refine_mode):
    """Same winner (count, canonical index, correspondences) with and without pruning, with a
    strong peak, with no peak at all (unrelated maps: pruning barely helps) and sharded."""
    if kind == "overlap":
        ref, qry, _ = synth.make_pair(400, seed=31, classes="five", outlier_frac=0.1)
    elif kind == "unrelated":
        ref = synth.make_pair(400, seed=32, classes="five")[0]
        qry = synth.make_pair(400, seed=33, classes="five")[1]
    elif kind == "one_label":
        ref, qry, _ = synth.make_pair(300, seed=34, classes="five")
        ref[:, 0] = 1.0; qry[:, 0] = 1.0
    else:
        ref, qry, _ = synth.make_pair(400, seed=35, classes="forest_urban", n_b=12)
    ros = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 10.0, "match_threshold_position": 0.5,
           "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 5}
    pr = PlaceRecognition(ros)
    found, _, _, info, ri, qi = pr.findTransformation(ref, qry)      # default mode
    assert info.match.search_mode == 1
    sref, sqry = ref.copy(), qry.copy()
    sref[:, 1:3] -= np.array(info.centroid_ref[:]); sqry[:, 1:3] -= np.array(info.centroid_qry[:])
    pr.prepare(sref, sqry, info.half_x, info.half_y)
    res_x, _ = pr.search(exhaustive=True)
    assert res_x.search_mode == 0
    assert (info.match.best_num_inliers, info.match.best_hyp_index) == (res_x.best_num_inliers, res_x.best_hyp_index)
    _, ri_x, qi_x = pr.extract(res_x.best_hyp_index)
    assert ri.tolist() == ri_x.tolist() and qi.tolist() == qi_x.tolist()
    lib = capi.lib()
    for n_shards in (2, 5):
        recs = (capi.TopkRecord * n_shards)()
        for r in range(n_shards):
            res, _ = pr.search(shard_index=r, shard_count=n_shards)
            assert res.search_mode == 1
            lib.slide_pr_pack_record(C.byref(res), r, C.byref(recs[r]))
        w = lib.slide_pr_merge_records(recs, n_shards)
        assert (recs[w].inliers, recs[w].hyp_index) == (res_x.best_num_inliers, res_x.best_hyp_index)
    pr.close()


@pytest.mark.parametrize("smem", [60000, 9000, 1200, 300])
def test_bound_phase_planner_paths(smem):
    """The bound phase stages the bitmap planes in label batches, in row bands, or reads them in
    place, depending on the shared-memory budget (SLIDE_PR_BOUND_SMEM shrinks it for the test).
    Every path must give the same bounds and the same winner."""
    ref, qry, _ = synth.make_pair(600, seed=41, classes="five", outlier_frac=0.1)
    ros = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 15.0, "match_threshold_position": 0.5,
           "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 5}
    ref = ref.copy(); qry = qry.copy()
    ref[:, 1:3] -= ref[:, 1:3].mean(0); qry[:, 1:3] -= qry[:, 1:3].mean(0)
    half = 1.2 * max(np.abs(ref[:, 1:3]).max(), np.abs(qry[:, 1:3]).max())

    def run(budget):
        if budget:
            os.environ["SLIDE_PR_BOUND_SMEM"] = str(budget)
        else:
            os.environ.pop("SLIDE_PR_BOUND_SMEM", None)
        try:
            pr = PlaceRecognition(ros)
            pr.prepare(ref, qry, half, half)
            nt = pr.lattice_info()[0]
            lo, hi = nt // 2, nt // 2 + 4000
            res, _ = pr.search()
            _, bound = pr.search(lo, hi, want_counts=True, bounds_only=True)
            res_x, _ = pr.search(exhaustive=True)
            pr.close()
        finally:
            os.environ.pop("SLIDE_PR_BOUND_SMEM", None)
        assert res.search_mode == 1 and res_x.search_mode == 0
        assert (res.best_num_inliers, res.best_hyp_index) == (res_x.best_num_inliers, res_x.best_hyp_index)
        return res.best_hyp_index, bound

    want_idx, want_bound = run(0)
    got_idx, got_bound = run(smem)
    assert got_idx == want_idx and np.array_equal(got_bound, want_bound)


@pytest.mark.parametrize("smem", [0, 100000, 20000, 3000])
def test_refinement_paths(smem):
    """Refined bounds (half-cell bitmap variants) forced on: variants read in place (default for small
    query maps) or staged in row bands of shrinking size.  Bounds stay upper bounds, the winner is
    the exhaustive one."""
    ref, qry, _ = synth.make_pair(500, seed=47, classes="five", outlier_frac=0.2)
    ros = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 20.0, "match_threshold_position": 0.5,
           "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 5}
    ref = ref.copy(); qry = qry.copy()
    ref[:, 1:3] -= ref[:, 1:3].mean(0); qry[:, 1:3] -= qry[:, 1:3].mean(0)
    half = 1.2 * max(np.abs(ref[:, 1:3]).max(), np.abs(qry[:, 1:3]).max())
    os.environ["SLIDE_PR_REFINE_MIN"] = "0"
    if smem:
        os.environ["SLIDE_PR_BOUND_SMEM"] = str(smem)
    try:
        pr = PlaceRecognition(ros)
        pr.prepare(ref, qry, half, half)
        nt, ny, _ = pr.lattice_info()
        lo, hi = nt // 2, nt // 2 + 3000
        res, _ = pr.search()
        res_x, _ = pr.search(exhaustive=True)
        _, exact = pr.search(lo, hi, want_counts=True)
        _, bound = pr.search(lo, hi, want_counts=True, bounds_only=True)
        pr.close()
    finally:
        os.environ.pop("SLIDE_PR_REFINE_MIN", None)
        os.environ.pop("SLIDE_PR_BOUND_SMEM", None)
    assert (res.best_num_inliers, res.best_hyp_index) == (res_x.best_num_inliers, res_x.best_hyp_index)
    assert (bound >= exact).all() and (bound <= len(qry)).all()
    # the refinement really tightened something: fewer filter-level false positives than the plain bound
    os.environ["SLIDE_PR_REFINE_MIN"] = "-1"
    try:
        pr = PlaceRecognition(ros)
        pr.prepare(ref, qry, half, half)
        _, plain = pr.search(lo, hi, want_counts=True, bounds_only=True)
        pr.close()
    finally:
        os.environ.pop("SLIDE_PR_REFINE_MIN", None)
    assert (bound <= plain).all() and bound.sum() < plain.sum()


def test_bound_and_verify_with_many_query_landmarks():
    """More than 4095 query landmarks: 16 bit planes per bound; more than 2047: row bands allowed."""
    ref, qry, _ = synth.make_pair(4500, seed=43, classes="forest_urban")
    ros = {"search_xy_step_size": 2.0, "search_yaw_step_size_degrees": 30.0, "match_threshold_position": 0.5,
           "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 5}
    pr = PlaceRecognition(ros)
    found, _, _, info, ri, qi = pr.findTransformation(ref, qry)
    assert info.match.search_mode == 1
    sref, sqry = ref.copy(), qry.copy()
    sref[:, 1:3] -= np.array(info.centroid_ref[:]); sqry[:, 1:3] -= np.array(info.centroid_qry[:])
    pr.prepare(sref, sqry, info.half_x, info.half_y)
    res_x, _ = pr.search(exhaustive=True)
    assert (info.match.best_num_inliers, info.match.best_hyp_index) == (res_x.best_num_inliers, res_x.best_hyp_index)
    nt = pr.lattice_info()[0]
    _, exact = pr.search(nt // 2, nt // 2 + 300, want_counts=True)
    _, bound = pr.search(nt // 2, nt // 2 + 300, want_counts=True, bounds_only=True)
    assert (bound >= exact).all() and (bound <= len(qry)).all()
    pr.close()


def test_two_phase_sharded_search_with_shared_incumbent(gold):
    """Sharded search as slide_slam_b200.parallel.sharded_search runs it: bound phase of every shard,
    max of the seeds' inlier counts (the all-reduce), verification against that incumbent with the
    bounds reused.  Shards without anything >= the incumbent report -1; the merge is the winner."""
    maps, cases, _ = gold
    for name in ("parking01_forest_yaml", "c1_forest_yaml"):
        c = cases[name]
        ref, qry = H.shifted_maps(maps, c)
        lib = capi.lib()
        for n_shards in (2, 3, 8):
            prs = [make_pr(c["params"]) for _ in range(n_shards)]   # one handle per "rank"
            seeds = []
            for r, pr in enumerate(prs):
                pr.prepare(ref, qry, c["half_x"], c["half_y"])
                seeds.append(pr.search(shard_index=r, shard_count=n_shards, bounds_only=True)[0])
            inc = max(max(int(s.best_num_inliers), 0) for s in seeds)
            recs = (capi.TopkRecord * n_shards)()
            total, n_none = 0, 0
            for r, pr in enumerate(prs):
                res, _ = pr.search(shard_index=r, shard_count=n_shards, incumbent_inliers=inc, reuse_bounds=True)
                assert res.search_mode == 1
                lib.slide_pr_pack_record(C.byref(res), r, C.byref(recs[r]))
                total += res.hypotheses_scored
                n_none += int(res.best_hyp_index < 0)
                if res.best_hyp_index >= 0:
                    assert res.best_num_inliers >= inc
            w = lib.slide_pr_merge_records(recs, n_shards)
            assert total == c["hypotheses_scored"]
            assert recs[w].hyp_index == c["best_hyp_index"] and recs[w].inliers == c["best_num_inliers"]
            assert n_none <= n_shards - 1
            # an incumbent above the true best: nothing anywhere reaches it
            res, _ = prs[0].search(incumbent_inliers=c["best_num_inliers"] + 1)
            assert res.best_hyp_index == -1 and res.best_num_inliers == -10000
            # an incumbent equal to the true best still finds the winner (ties go to the real index)
            res, _ = prs[0].search(incumbent_inliers=c["best_num_inliers"])
            assert (res.best_hyp_index, res.best_num_inliers) == (c["best_hyp_index"], c["best_num_inliers"])
            for pr in prs:
                pr.close()


def test_edge_cases_through_the_abi():
    kw = dict(match_xy_step_size=0.5, yaw_step_deg=45.0)
    rng = np.random.default_rng(1)
    ref, qry = H.random_maps(rng, 10, 8, extent=4.0)
    pr = make_pr(kw)
    # empty maps: every hypothesis scores 0 and the first one wins (PR.cpp:125,361)
    for r, q in ((np.zeros((0, 7)), qry), (ref, np.zeros((0, 7))), (np.zeros((0, 7)), np.zeros((0, 7)))):
        m = pr.MatchMaps(r, q, 6.0, 6.0)
        want = O.match_maps(O.make_params(**kw), r, q, 6.0, 6.0)
        assert m.best_num_inliers == 0 and m.info.best_hyp_index == 0 and len(m.ref_idx) == 0
        assert m.R_t.ravel().tolist() == want["R_t"].ravel().tolist()
        assert m.info.hypotheses_scored == want["hypotheses_scored"]
    # sanity-check early return (PR.cpp:169-175) and zero rings (PR.cpp:125)
    m = pr.MatchMaps(ref, qry, 0.3, 0.3)
    assert m.status == capi.SANITY_RETURN
    m = pr.MatchMaps(ref, qry, 0.0, 0.0)
    assert m.status == 0 and m.best_num_inliers == -10000 and m.info.hypotheses_scored == 0
    # NaN labels never match; -0.0 == 0.0
    r3, q3 = ref.copy(), qry.copy()
    r3[0, 0] = np.nan; q3[0, 0] = np.nan; r3[1, 0] = -0.0; q3[1, 0] = 0.0
    m = pr.MatchMaps(r3, q3, 6.0, 6.0)
    want = O.match_maps(O.make_params(**kw), r3, q3, 6.0, 6.0)
    assert (m.best_num_inliers, m.info.best_hyp_index) == (want["best_num_inliers"], want["best_hyp_index"])
    assert m.ref_idx.tolist() == want["ref_idx"].tolist()
    # non-finite coordinates are rejected with an error, not silently scored
    bad = ref.copy(); bad[0, 1] = np.inf
    with pytest.raises(capi.SlidePrError):
        pr.MatchMaps(bad, qry, 6.0, 6.0)
    # size gate of findInterLoopClosure (PR.cpp:508-510)
    pr2 = make_pr(dict(kw, min_num_map_objects_to_start=20))
    assert pr2.findInterLoopClosure(ref, qry)[0] is False
    pr.close(); pr2.close()


def test_inter_and_intra_loop_closure(gold):
    maps, cases, _ = gold
    c = cases["prtest_inter_lsq1"]
    pr = make_pr(c["params"])
    found, tf = pr.findInterLoopClosure(maps[c["ref"]], maps[c["qry"]])
    op = O.make_params(**c["params"])
    wfound, wtf, _ = O.find_inter_loop_closure(op, maps[c["ref"]], maps[c["qry"]], n_threads=-1)
    assert found and wfound
    np.testing.assert_allclose(tf, wtf, rtol=RTOL, atol=1e-9)
    # findIntraLoopClosure(measurements, submap, I, I) (place_recognition_test.cpp:219-226)
    ci = cases["prtest_intra_lsq1"]
    pri = make_pr(ci["params"])
    found, tf = pri.findIntraLoopClosure(maps[ci["qry"]], maps[ci["ref"]], np.eye(4), np.eye(4))
    assert found
    x, y, _, yaw = ci["xyz_yaw"]
    want = np.eye(4)
    want[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    want[0, 3], want[1, 3] = x, y
    np.testing.assert_allclose(tf, want, rtol=RTOL, atol=1e-9)
    # with non-trivial poses: tf = candidate^-1 * query * loop_closure (PR.cpp:478-494)
    def pose(yaw, t):
        m = np.eye(4); m[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]; m[:3, 3] = t
        return m
    qp, cp = pose(0.3, [1.0, -2.0, 0.1]), pose(-0.2, [0.5, 0.5, 0.0])
    meas_local = maps[ci["qry"]].copy()
    inv = np.linalg.inv(qp)
    meas_local[:, 1:4] = (meas_local[:, 1:4] - qp[:3, 3]) @ inv[:3, :3].T
    found, tf2 = pri.findIntraLoopClosure(meas_local, maps[ci["ref"]], qp, cp)
    ofound, otf2, oinfo = O.find_intra_loop_closure(O.make_params(**ci["params"]), meas_local, maps[ci["ref"]], qp, cp)
    assert found and ofound
    np.testing.assert_allclose(tf2, otf2, rtol=RTOL, atol=1e-9)  # the oracle's findIntraLoopClosure (PR.cpp:389-496)
    np.testing.assert_allclose(tf2, np.linalg.inv(cp) @ qp @ want, rtol=1e-4, atol=1e-6)  # and the scene's own truth
    # identity poses through the oracle as well
    ofound, otf, _ = O.find_intra_loop_closure(O.make_params(**ci["params"]), maps[ci["qry"]], maps[ci["ref"]], np.eye(4), np.eye(4))
    assert ofound
    np.testing.assert_allclose(tf, otf, rtol=RTOL, atol=1e-9)
    pr.close(); pri.close()


def test_budget_path_equals_unlimited(gold):
    maps, cases, _ = gold
    c = cases["indoor01_forest_yaml_nodim"]
    ref, qry = H.shifted_maps(maps, c)
    pr = PlaceRecognition(dict(H.rosparams_from_golden(c["params"]), compute_budget_sec=3600.0))
    m = pr.MatchMaps(ref, qry, c["half_x"], c["half_y"])
    assert m.info.best_hyp_index == c["best_hyp_index"] and m.best_num_inliers == c["best_num_inliers"]
    assert m.info.rings_scored == m.info.n_rings and m.info.hypotheses_scored == c["hypotheses_scored"]
    pr.close()


def test_sharded_search_merges_to_the_unsharded_result(gold):
    maps, cases, _ = gold
    c = cases["parking01_forest_yaml"]
    ref, qry = H.shifted_maps(maps, c)
    pr = make_pr(c["params"])
    pr.prepare(ref, qry, c["half_x"], c["half_y"])
    lib = capi.lib()
    for n_shards in (2, 3, 8):
        recs = (capi.TopkRecord * n_shards)()
        total = 0
        for r in range(n_shards):
            res, _ = pr.search(shard_index=r, shard_count=n_shards)
            lib.slide_pr_pack_record(C.byref(res), r, C.byref(recs[r]))
            total += res.hypotheses_scored
        w = lib.slide_pr_merge_records(recs, n_shards)
        assert total == c["hypotheses_scored"]
        assert recs[w].hyp_index == c["best_hyp_index"] and recs[w].inliers == c["best_num_inliers"]
    res, ri, qi = pr.extract(c["best_hyp_index"])
    assert ri.tolist() == c["ref_idx"] and qi.tolist() == c["qry_idx"] and list(res.R_t) == c["R_t"]
    pr.close()


def test_score_hypotheses_list():
    """General (c, s, x, y) hypothesis lists -- the scorer behind pair/triplet-generated
    candidates -- against the oracle's single-hypothesis scorer (PR.cpp:246-357)."""
    rng = np.random.default_rng(7)
    ref, qry = H.random_maps(rng, 80, 60, extent=15.0)
    kw = dict(match_xy_step_size=0.5, match_threshold=0.6, match_threshold_dimension=0.8)
    op = O.make_params(**kw)
    pr = make_pr(kw)
    pr.prepare(ref, qry, 20.0, 20.0)
    n = 3000
    yaw = rng.uniform(-np.pi, np.pi, n)
    hyps = np.column_stack([np.cos(yaw), np.sin(yaw), rng.uniform(-8, 8, n), rng.uniform(-8, 8, n)])
    res, got = pr.score_hypotheses(hyps)
    want = np.array([O.score_one(op, ref, qry, *h)[0] for h in hyps])
    assert np.array_equal(got, want)
    assert res.best_num_inliers == want.max() and res.best_hyp_index == int(np.argmax(want))
    pr.close()


def test_triangle_matching_matches_oracle():
    rng = np.random.default_rng(11)
    tm = rng.uniform(-30, 30, (700, 6))
    td = np.vstack([tm[rng.integers(0, 700, 150)] + 0.01 * rng.normal(size=(150, 6)), rng.uniform(-30, 30, (900, 6))])
    mi, di, _ = O.match_triangles(tm, td, 0.1)
    pr = make_pr({})
    lib = capi.lib()
    cap = len(mi) + 10
    gm, gd = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    pm, pd = np.zeros(3 * cap, np.int32), np.zeros(3 * cap, np.int32)
    n = C.c_int64(0)
    tmc, tdc = np.ascontiguousarray(tm), np.ascontiguousarray(td)
    rc = lib.slide_pr_match_triangles(pr._h, capi.dptr(tmc), len(tmc), capi.dptr(tdc), len(tdc), 0.1, capi.iptr(gm),
                                      capi.iptr(gd), capi.iptr(pm), capi.iptr(pd), cap, C.byref(n))
    assert rc == 0 and n.value == len(mi)
    assert gm[:n.value].tolist() == mi.tolist() and gd[:n.value].tolist() == di.tolist()  # reference order
    for k in range(0, n.value, 37):
        assert pm[3 * k:3 * k + 3].tolist() == O.triangle_descriptor(tm[mi[k]])[1].tolist()
        assert pd[3 * k:3 * k + 3].tolist() == O.triangle_descriptor(td[di[k]])[1].tolist()
    pr.close()


def test_config2_full_size_properties():
    """BASELINE config 2 (2000 x 2000, 5 classes, 10 % outliers) at full size, where the oracle is
    too slow for a full run: size-independent properties + an oracle-checked slice."""
    ref, qry, truth = synth.config_pair(2)
    kw = dict(match_xy_step_size=0.5, yaw_step_deg=5.0, match_threshold=0.5, match_threshold_dimension=1.0,
              ignore_dimension=0, min_num_inliers=15)
    pr = make_pr(kw)
    found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(ref, qry)
    assert found and info.best_num_inliers >= 50
    # (1) the planted SE(2) offset is recovered (refined transform)
    # (5 deg yaw lattice over a 283 m map: only the landmarks near the pivot align, so the refined
    #  transform is good to a few cm / mrad at the pivot and ~1 m at the map centre's lever arm)
    assert abs(np.angle(np.exp(1j * (xyz_yaw[3] - truth["yaw"])))) < 0.02
    assert abs(xyz_yaw[0] - truth["t"][0]) < 3.0 and abs(xyz_yaw[1] - truth["t"][1]) < 3.0
    # (2) the winner's count equals the reference-order brute-force recount (self-check inside
    #     slide_pr_match_maps) and the oracle's single-hypothesis scorer
    op = O.make_params(**kw)
    sref, sqry = ref.copy(), qry.copy()
    sref[:, 1:3] -= np.array(info.centroid_ref[:]); sqry[:, 1:3] -= np.array(info.centroid_qry[:])
    R = np.array(info.R_t[:]).reshape(3, 3)
    n, ori, oqi = O.score_one(op, sref, sqry, R[0, 0], R[1, 0], R[0, 2], R[1, 2])
    assert n == info.best_num_inliers and ori.tolist() == ri.tolist() and oqi.tolist() == qi.tolist()
    # (3) oracle on a slice of translations around the winner: every count identical
    ny = info.match.n_yaw
    t_win = info.match.best_hyp_index // ny
    tb, te = max(t_win - 20, 0), t_win + 20
    want = O.match_maps(op, sref, sqry, info.half_x, info.half_y, tb * ny, te * ny, want_counts=True)
    pr.prepare(sref, sqry, info.half_x, info.half_y)
    res, got = pr.search(tb, te, want_counts=True)
    assert np.array_equal(got, want["counts"])
    # (4) sharded == unsharded, direct variant == queued variant
    best = (info.best_num_inliers, info.match.best_hyp_index)
    lib = capi.lib()
    recs = (capi.TopkRecord * 4)()
    for r in range(4):
        res, _ = pr.search(shard_index=r, shard_count=4)
        lib.slide_pr_pack_record(C.byref(res), r, C.byref(recs[r]))
    w = lib.slide_pr_merge_records(recs, 4)
    assert (recs[w].inliers, recs[w].hyp_index) == best
    pr.close()
    pr0 = make_pr(kw, variant=0)
    m = pr0.MatchMaps(sref, sqry, info.half_x, info.half_y)
    assert (m.best_num_inliers, m.info.best_hyp_index) == best
    pr0.close()


def test_streaming_reuse_of_reference_index_and_lattice():
    """Many queries against one accumulated map (BASELINE config 5): the reference-map index and
    the lattice are rebuilt only when their inputs change; reused structures give the same
    results as a fresh handle."""
    big, queries = synth.config_stream(n_map=300, n_queries=3, n_sub=40, seed=77)
    kw = dict(match_xy_step_size=0.5, yaw_step_deg=30.0, match_threshold=0.5, match_threshold_dimension=1.0,
              ignore_dimension=0, min_num_inliers=10)
    op = O.make_params(**kw)
    pr = make_pr(kw)
    flags = []
    for q in queries:
        found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(big, q)
        flags.append(info.match.reuse)
        fresh = make_pr(kw)
        f2, x2, t2, i2, r2, q2 = fresh.findTransformation(big, q)
        fresh.close()
        assert (found, info.best_num_inliers, info.match.best_hyp_index) == (f2, i2.best_num_inliers, i2.match.best_hyp_index)
        assert ri.tolist() == r2.tolist() and qi.tolist() == q2.tolist() and list(info.R_t) == list(i2.R_t)
        want = O.find_transformation(op, big, q, n_threads=-1)
        assert (found, info.best_num_inliers, info.match.best_hyp_index) == (want["found"], want["best_num_inliers"], want["best_hyp_index"])
        assert ri.tolist() == want["ref_idx"].tolist()
    assert flags[0] == 0 and all(f & 2 for f in flags[1:])  # the map's index is built once
    # a different reference map must not reuse the index
    other = big.copy(); other[0, 1] += 0.25
    found, _, _, info, _, _ = pr.findTransformation(other, queries[0])
    assert not (info.match.reuse & 2)
    want = O.find_transformation(op, other, queries[0], n_threads=-1)
    assert (info.best_num_inliers, info.match.best_hyp_index) == (want["best_num_inliers"], want["best_hyp_index"])
    pr.close()

