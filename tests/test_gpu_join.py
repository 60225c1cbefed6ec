"""GPU parity suite (-m gpu) of the DEFAULT search: the pair-join scorer (slide_slam_b200/csrc/spr_join.cu),
which gives every hypothesis of the lattice its exact inlier count.  Called through the C-ABI and compared with
  * the committed golden vectors and the CPU oracle: every per-hypothesis count (bit-exact), the winner, its
    correspondences and R_t; refined transforms within 1e-5 relative (BASELINE.json north_star);
  * the lattice kernels of the same library (exhaustive verification and bound-and-verify), at the full sizes of
    BASELINE.json configs 2-5 where the oracle can only check slices.
"""
import os

import numpy as np
import pytest

import spr_helpers as H
from oracle import pyoracle as O
from slide_slam_b200 import capi, synth
from slide_slam_b200.place_recognition import PlaceRecognition

pytestmark = pytest.mark.gpu

RTOL = 1e-5
JOIN = 2   # slide_pr_match_result.search_mode of the pair-join scorer
KW = dict(match_xy_step_size=0.5, yaw_step_deg=5.0, match_threshold=0.5, match_threshold_dimension=1.0,
          ignore_dimension=0, min_num_inliers=15)
ROS = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 5.0, "match_threshold_position": 0.5,
       "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 15}


@pytest.fixture(autouse=True, scope="module")
def _default_engine():
    old = os.environ.pop("SLIDE_PR_ENGINE", None)
    yield
    if old is not None:
        os.environ["SLIDE_PR_ENGINE"] = old


@pytest.fixture(scope="module")
def gold():
    return H.golden_maps(), H.golden_cases(), H.golden_counts()


def make_pr(params: dict, **kw):
    pr = PlaceRecognition(H.rosparams_from_golden(params), **kw)
    if "use_lsq" in params:
        pr.use_lsq = bool(params["use_lsq"])
    if "inter_loop_closure" in params:
        pr.inter_loop_closure = bool(params["inter_loop_closure"])
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
    assert info.match.search_mode == JOIN or info.match.status == capi.SANITY_RETURN
    assert found == c["found"]
    assert info.best_num_inliers == c["best_num_inliers"]
    assert info.match.best_hyp_index == c["best_hyp_index"]
    assert info.match.hypotheses_scored == c["hypotheses_scored"]
    assert ri.tolist() == c["ref_idx"] and qi.tolist() == c["qry_idx"]
    assert list(info.R_t) == c["R_t"]                      # lattice winner: bit-exact
    assert info.half_x == c["half_x"] and info.half_y == c["half_y"]
    if found:
        np.testing.assert_allclose(xyz_yaw, c["xyz_yaw"], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(tf.ravel(), c["transform"], rtol=RTOL, atol=1e-9)
    pr.close()


def test_every_hypothesis_count_indoor(gold):
    maps, cases, counts = gold
    c = cases["indoor01_forest_yaml_nodim"]
    ref, qry = H.shifted_maps(maps, c)
    pr = make_pr(c["params"])
    pr.prepare(ref, qry, c["half_x"], c["half_y"])
    nt, ny, _ = pr.lattice_info()
    res, got = pr.search(0, nt, want_counts=True)
    assert res.search_mode == JOIN
    assert np.array_equal(got, counts["indoor01_forest_yaml_nodim"])
    assert res.best_hyp_index == c["best_hyp_index"] and res.best_num_inliers == c["best_num_inliers"]
    assert res.hypotheses_scored == nt * ny
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
    assert res.search_mode == JOIN
    sl = counts[name][tb * ny - lo: te * ny - lo]
    assert np.array_equal(got, sl)
    assert res.hypotheses_scored == (te - tb) * ny
    assert res.best_num_inliers == int(sl.max()) and res.best_hyp_index == tb * ny + int(np.argmax(sl))
    pr.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_maps_every_hypothesis(seed):
    """thresholds above and below the step (one micro-tile / general path), duplicate and near-duplicate
    landmarks (first-match attribution), cylinders and cuboids, coordinates snapped onto the threshold"""
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
    pr = make_pr(kw)
    pr.prepare(ref, qry, hx, hy)
    nt, ny, _ = pr.lattice_info()
    res, got = pr.search(0, nt, want_counts=True)
    assert res.search_mode == JOIN
    assert np.array_equal(got, want["counts"])
    assert (res.best_num_inliers, res.best_hyp_index) == (want["best_num_inliers"], want["best_hyp_index"])
    m = pr.MatchMaps(ref, qry, hx, hy)
    assert m.info.search_mode == JOIN and m.best_num_inliers == want["best_num_inliers"]
    assert m.ref_idx.tolist() == want["ref_idx"].tolist() and m.qry_idx.tolist() == want["qry_idx"].tolist()
    assert m.R_t.ravel().tolist() == want["R_t"].ravel().tolist()
    # slices and shards partition the lattice: their winners merge into the winner
    cut = sorted(int(v) for v in rng.integers(0, nt + 1, 2))
    parts = [pr.search(0, cut[0])[0], pr.search(cut[0], cut[1])[0], pr.search(cut[1], nt)[0]]
    assert sum(p.hypotheses_scored for p in parts) == nt * ny
    keys = [(p.best_num_inliers, -p.best_hyp_index) for p in parts if p.best_hyp_index >= 0]
    assert max(keys) == (want["best_num_inliers"], -want["best_hyp_index"])
    shards = [pr.search(shard_index=r, shard_count=3)[0] for r in range(3)]
    assert sum(s.hypotheses_scored for s in shards) == nt * ny
    keys = [(s.best_num_inliers, -s.best_hyp_index) for s in shards if s.best_hyp_index >= 0]
    assert max(keys) == (want["best_num_inliers"], -want["best_hyp_index"])
    pr.close()


@pytest.mark.parametrize("seed", range(24))
def test_dense_and_clustered_maps_every_hypothesis(seed):
    """The kernel's own hard cases (same generator as tests/test_host_index.py runs through the CPU emulation): clusters of
    near-duplicate reference landmarks (first-match attribution), several quads and full pair lists per block (the list is
    flushed mid-pass), more than four cell bands per landmark (several passes), thresholds from a third of the step to four
    steps, one label or many, rectangular ranges.  Every per-hypothesis count against the oracle."""
    rng = np.random.default_rng(7000 + seed)
    n_ref, n_qry = int(rng.integers(20, 160)), int(rng.integers(20, 120))
    n_labels = int(rng.choice([1, 1, 2, 6]))
    ref, qry = H.random_maps(rng, n_ref, n_qry, extent=float(rng.uniform(4, 12)), n_labels=n_labels,
                             dup_frac=float(rng.choice([0.0, 0.3, 0.6])), grid=(0.125 if seed % 4 == 0 else None))
    if seed % 3 == 0:
        k = min(12, n_ref)
        ref[:k, 1:3] = ref[0, 1:3] + rng.normal(0, 0.1, (k, 2))
        ref[:k, 0] = ref[0, 0]
    step = float(rng.choice([0.5, 0.25, 0.4]))
    thr = float(step * rng.choice([0.34, 1.0, 1.0, 1.7, 4.0]))
    kw = dict(match_xy_step_size=step, yaw_step_deg=float(rng.choice([45.0, 60.0, 36.0])), match_threshold=thr,
              match_threshold_dimension=float(rng.choice([1.0, 0.4])), ignore_dimension=int(seed % 5 == 2),
              disable_yaw_search=int(seed % 6 == 5))
    hx = float(rng.uniform(3, 9))
    hy = hx if seed % 2 else float(rng.uniform(3, 9))
    op = O.make_params(**kw)
    if O.enumerate_lattice(op, hx, hy) is None:
        pytest.skip("range below one lattice step")
    want = O.match_maps(op, ref, qry, hx, hy, want_counts=True)
    pr = make_pr(kw)
    pr.prepare(ref, qry, hx, hy)
    nt, ny, _ = pr.lattice_info()
    res, got = pr.search(0, nt, want_counts=True)
    assert res.search_mode == JOIN
    bad = np.nonzero(got != want["counts"])[0]
    assert bad.size == 0, f"first mismatching hypotheses {bad[:5]}: {got[bad[:5]]} != {want['counts'][bad[:5]]}"
    res2, _ = pr.search()          # the whole-block scan path (no per-hypothesis output)
    assert (res2.best_num_inliers, res2.best_hyp_index) == (want["best_num_inliers"], want["best_hyp_index"])
    pr.close()


def test_edge_cases():
    kw = dict(match_xy_step_size=0.5, yaw_step_deg=45.0)
    rng = np.random.default_rng(1)
    ref, qry = H.random_maps(rng, 10, 8, extent=4.0)

    def check(kw, r, q, hx, hy):
        op = O.make_params(**kw)
        want = O.match_maps(op, r, q, hx, hy, want_counts=True)
        pr = make_pr(kw)
        pr.prepare(r, q, hx, hy)
        nt, ny, _ = pr.lattice_info()
        res, got = pr.search(0, nt, want_counts=True)
        assert np.array_equal(got, want["counts"])
        assert (res.best_num_inliers, res.best_hyp_index) == (want["best_num_inliers"], want["best_hyp_index"])
        pr.close()
        return res

    # empty maps: every hypothesis scores 0; the first one wins (PR.cpp:125,361)
    for r, q in ((np.zeros((0, 7)), qry), (ref, np.zeros((0, 7))), (np.zeros((0, 7)), np.zeros((0, 7)))):
        res = check(kw, r, q, 6.0, 6.0)
        assert res.best_num_inliers == 0 and res.best_hyp_index == 0
    q2 = qry.copy(); q2[:, 0] = 77.0                        # labels that never occur in the reference map
    assert check(kw, ref, q2, 6.0, 6.0).best_num_inliers == 0
    r3, q3 = ref.copy(), qry.copy()
    r3[0, 0] = np.nan; q3[0, 0] = np.nan; r3[1, 0] = -0.0; q3[1, 0] = 0.0
    check(kw, r3, q3, 6.0, 6.0)
    # all landmarks identical: every reference object matches every query object, each query counts once
    r4 = np.tile(np.array([[2, 1.0, 1.0, 0, 0.5, 0, 0]], float), (6, 1))
    q4 = np.tile(np.array([[2, 0.0, 0.0, 0, 0.5, 0, 0]], float), (5, 1))
    assert check(kw, r4, q4, 6.0, 6.0).best_num_inliers == 5
    # more identical query landmarks than a round of the kernel holds (224): the u8 counters must not wrap
    q5 = np.tile(np.array([[2, 0.0, 0.0, 0, 0.5, 0, 0]], float), (700, 1))
    assert check(dict(kw, yaw_step_deg=90.0), r4, q5, 3.0, 3.0).best_num_inliers == 700
    check(dict(match_xy_step_size=0.5, disable_yaw_search=1), ref, qry, 12.0, 5.5)       # rectangular ranges
    check(dict(match_xy_step_size=0.5, yaw_step_deg=90.0, match_threshold=0.0), ref, qry, 5.0, 5.0)   # nothing can match
    check(dict(inter_loop_closure=0, yaw_half_range_intra_deg=10.0, yaw_step_deg=2.0), ref, qry, 5.0, 5.0)   # (0, 0) skipped
    check(dict(match_xy_step_size=0.1, yaw_step_deg=60.0, match_threshold=0.75), ref, qry, 4.0, 4.0)  # 15 x 15 samples per pair


def _shifted(ref, qry, info):
    sref, sqry = ref.copy(), qry.copy()
    sref[:, 1:3] -= np.array(info.centroid_ref[:])
    sqry[:, 1:3] -= np.array(info.centroid_qry[:])
    return sref, sqry


def _check_against_oracle_and_lattice(ref, qry, half_width, shards=0, lattice_exhaustive=True):
    """findTransformation through the default engine; oracle: winner's count and correspondences, every count of
    slices around the winner / first ring / last ring; lattice kernels: the same winner from the bound-and-verify
    search (and from the exhaustive verification when asked)."""
    op = O.make_params(**KW)
    pr = PlaceRecognition(ROS)
    found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(ref, qry)
    assert info.match.search_mode == JOIN
    sref, sqry = _shifted(ref, qry, info)
    R = np.array(info.R_t[:]).reshape(3, 3)
    n, ori, oqi = O.score_one(op, sref, sqry, R[0, 0], R[1, 0], R[0, 2], R[1, 2])
    assert n == info.best_num_inliers == len(ri) and ori.tolist() == ri.tolist() and oqi.tolist() == qi.tolist()
    best = (info.best_num_inliers, info.match.best_hyp_index)
    ny, nt = info.match.n_yaw, info.match.n_translations
    assert info.match.hypotheses_scored == ny * nt
    t_win = best[1] // ny
    edge = max(2, half_width // 2)
    pr.prepare(sref, sqry, info.half_x, info.half_y)
    for tb, te in [(max(t_win - half_width, 0), min(t_win + half_width + 1, nt)), (0, edge), (nt - edge, nt), (nt // 3, nt // 3 + edge)]:
        want = O.match_maps(op, sref, sqry, info.half_x, info.half_y, tb * ny, te * ny, want_counts=True, n_threads=-1)
        res, got = pr.search(tb, te, want_counts=True)
        assert res.search_mode == JOIN and np.array_equal(got, want["counts"]), f"slice [{tb}, {te})"
        assert (res.best_num_inliers, res.best_hyp_index) == (want["best_num_inliers"], want["best_hyp_index"])
    res_l, _ = pr.search(engine="lattice")
    assert res_l.search_mode == 1 and (res_l.best_num_inliers, res_l.best_hyp_index) == best
    if lattice_exhaustive:
        res_x, _ = pr.search(exhaustive=True)
        assert res_x.search_mode == 0 and (res_x.best_num_inliers, res_x.best_hyp_index) == best
    if shards:
        got = [pr.search(shard_index=r, shard_count=shards)[0] for r in range(shards)]
        assert sum(g.hypotheses_scored for g in got) == ny * nt
        assert max((g.best_num_inliers, -g.best_hyp_index) for g in got) == (best[0], -best[1])
    pr.close()
    return found, info


def test_config2_full_size():
    ref, qry, truth = synth.config_pair(2)
    found, info = _check_against_oracle_and_lattice(ref, qry, 20, shards=8)
    assert found and info.match.n_yaw == 73 and info.match.hypotheses_scored > 6e7


@pytest.mark.parametrize("kind", ["unrelated", "one_label", "tiny_query"])
def test_other_map_shapes(kind):
    """no peak at all (unrelated maps), a single label (every pair is a candidate), a tiny query map"""
    if kind == "unrelated":
        ref = synth.make_pair(2000, seed=32, classes="five")[0]
        qry = synth.make_pair(2000, seed=33, classes="five")[1]
    elif kind == "one_label":
        ref, qry, _ = synth.make_pair(600, seed=34, classes="five")
        ref[:, 0] = 1.0; qry[:, 0] = 1.0
    else:
        ref, qry, _ = synth.make_pair(1500, seed=35, classes="forest_urban", n_b=12)
    _check_against_oracle_and_lattice(ref, qry, 6, shards=5)


def test_config3_20000_landmarks_full_size():
    ref, qry, truth = synth.config_pair(3)
    found, info = _check_against_oracle_and_lattice(ref, qry, 2, shards=8, lattice_exhaustive=False)
    assert found and info.best_num_inliers > 200 and info.match.hypotheses_scored > 4e8


def test_config4_5000_landmark_pairs_full_size():
    maps = synth.config_robots(8, 5000)
    n_found = 0
    for r, q in ((0, 1), (3, 4), (0, 4)):
        found, info = _check_against_oracle_and_lattice(maps[r], maps[q], 6 if (r, q) == (0, 1) else 2)
        n_found += int(found)
    assert n_found >= 2


def test_config5_streaming_queries_against_50000_landmarks_full_size():
    big, queries = synth.config_stream(50000, n_queries=3, n_sub=300)
    pr = PlaceRecognition(ROS)
    for k, q in enumerate(queries):
        found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(big, q)
        assert info.match.search_mode == JOIN and found and info.best_num_inliers >= 25
        assert bool(info.match.reuse & 2) == (k > 0)          # the reference map's join index is built once
    pr.close()
    _check_against_oracle_and_lattice(big, queries[0], 3, lattice_exhaustive=False)


def test_intra_loop_closure_batch_equals_single_calls(gold):
    """slide_pr_find_intra_loop_closure_batch (candidates enqueued back to back, one wait): per candidate exactly what
    findIntraLoopClosure returns -- found flag, inlier count, winner, transform -- including a candidate without a
    closure, an empty submap and a repeated candidate; and the lattice-engine fallback of the same entry."""
    maps, cases, _ = gold
    ci = cases["prtest_intra_lsq1"]

    def pose(yaw, t):
        m = np.eye(4); m[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]; m[:3, 3] = t
        return m
    qp = pose(0.3, [1.0, -2.0, 0.1])
    meas_local = maps[ci["qry"]].copy()
    inv = np.linalg.inv(qp)
    meas_local[:, 1:4] = (meas_local[:, 1:4] - qp[:3, 3]) @ inv[:3, :3].T
    sub = maps[ci["ref"]]
    rng = np.random.default_rng(5)
    far = sub.copy(); far[:, 1:3] += 500.0                      # nothing within reach: no closure
    jitter = sub.copy(); jitter[:, 1:3] += rng.normal(0, 0.05, (len(sub), 2))
    fewer = sub[: len(sub) // 2].copy()
    submaps = [sub, far, np.zeros((0, 7)), jitter, fewer, sub]
    cposes = [pose(-0.2, [0.5, 0.5, 0.0]), np.eye(4), np.eye(4), pose(0.1, [0.0, 1.0, 0.0]), pose(0.0, [2.0, 0.0, 0.0]), pose(-0.2, [0.5, 0.5, 0.0])]
    for engine in (None, "lattice"):
        pri = make_pr(ci["params"], engine=engine) if engine else make_pr(ci["params"])
        single = []
        for s_, c_ in zip(submaps, cposes):
            f, tf = pri.findIntraLoopClosure(meas_local, s_, qp, c_)
            single.append((f, tf, pri.last.best_num_inliers, pri.last.match.best_hyp_index))
        batch = pri.findIntraLoopClosureBatch(meas_local, submaps, qp, cposes)
        assert len(batch) == len(submaps)
        assert [b[0] for b in batch] == [s[0] for s in single]
        assert batch[0][0] and batch[5][0] and not batch[1][0] and not batch[2][0]
        for (bf, btf, bo), (sf, stf, sn, si) in zip(batch, single):
            assert bf == sf
            if sf:
                assert bo.best_num_inliers == sn and bo.match.best_hyp_index == si
                assert btf.tolist() == stf.tolist()          # same code path downstream of the search: bit-identical
        # the handle keeps working for ordinary calls afterwards
        f, tf = pri.findIntraLoopClosure(meas_local, sub, qp, cposes[0])
        assert f and tf.tolist() == single[0][1].tolist()
        assert pri.findIntraLoopClosureBatch(meas_local, [], qp, []) == []
        pri.close()
