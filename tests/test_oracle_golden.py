"""CPU suite: the oracle against the committed golden vectors, against an independent numpy
restatement, and against the reference's own pass criterion for the PRtest scene
(place_recognition_test.cpp:242-275).  No GPU needed."""
import math

import numpy as np
import pytest

from oracle import oracle_np
from oracle import pyoracle as O
import spr_helpers as H

FAST = ["indoor01_sloam_yaml", "indoor01_forest_yaml_nodim", "indoor02_forest_yaml_nodim", "indoor12_forest_yaml_nodim",
        "indoor10_forest_yaml_nodim", "indoor01_sloam_yaml_nolsq", "indoor01_defaults_2deg", "indoor01_noyaw",
        "prtest_intra_lsq1", "prtest_intra_lsq0"]


@pytest.fixture(scope="module")
def gold():
    return H.golden_maps(), H.golden_cases()


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_golden(gold, name):
    maps, cases = gold
    c = cases[name]
    p = O.make_params(**c["params"])
    r = O.find_transformation(p, maps[c["ref"]], maps[c["qry"]], n_threads=1)
    assert r["found"] == c["found"]
    assert r["best_num_inliers"] == c["best_num_inliers"]
    assert r["best_hyp_index"] == c["best_hyp_index"]
    assert r["hypotheses_scored"] == c["hypotheses_scored"]
    assert r["ref_idx"].tolist() == c["ref_idx"] and r["qry_idx"].tolist() == c["qry_idx"]
    assert r["R_t"].ravel().tolist() == c["R_t"]  # bit-exact
    if c["found"]:
        np.testing.assert_allclose(r["xyz_yaw"], c["xyz_yaw"], rtol=1e-12, atol=1e-12)


def test_threaded_oracle_equals_single_thread(gold):
    maps, cases = gold
    c = cases["indoor01_forest_yaml_nodim"]
    p = O.make_params(**c["params"])
    a = O.find_transformation(p, maps[c["ref"]], maps[c["qry"]], n_threads=1)
    b = O.find_transformation(p, maps[c["ref"]], maps[c["qry"]], n_threads=4)
    assert a["best_hyp_index"] == b["best_hyp_index"] and a["best_num_inliers"] == b["best_num_inliers"]
    assert a["ref_idx"].tolist() == b["ref_idx"].tolist()
    assert a["xyz_yaw"].tolist() == b["xyz_yaw"].tolist()


def test_golden_counts_slice(gold):
    maps, cases = gold
    counts = H.golden_counts()
    c = cases["indoor01_forest_yaml_nodim"]
    ref, qry = H.shifted_maps(maps, c)
    p = O.make_params(**c["params"])
    r = O.match_maps(p, ref, qry, c["half_x"], c["half_y"], want_counts=True)
    assert np.array_equal(r["counts"], counts["indoor01_forest_yaml_nodim"])
    assert int(r["counts"].max()) == c["best_num_inliers"]
    assert int(np.argmax(r["counts"])) == c["best_hyp_index"]  # first maximum == strict '>' rule


def test_numpy_twin_agrees_with_c_oracle(gold):
    """Two independently written restatements of MatchMaps must give identical counts."""
    maps, cases = gold
    c = cases["indoor01_forest_yaml_nodim"]
    ref, qry = H.shifted_maps(maps, c)
    p = O.make_params(**c["params"])
    r = oracle_np.match_maps(ref, qry, c["half_x"], c["half_y"], step=p.match_xy_step_size,
                             yaw_half=p.match_yaw_half_range, yaw_step=p.match_yaw_angle_step_size,
                             thr=p.match_threshold, thr_dim=p.match_threshold_dimension,
                             ignore_dimension=bool(p.ignore_dimension), want_counts=True)
    assert np.array_equal(r["counts"], H.golden_counts()["indoor01_forest_yaml_nodim"])
    assert r["best_hyp_index"] == c["best_hyp_index"]
    assert r["ref_idx"].tolist() == c["ref_idx"] and r["qry_idx"].tolist() == c["qry_idx"]
    assert r["R_t"].ravel().tolist() == c["R_t"]


def test_numpy_twin_dimension_rule():
    rng = np.random.default_rng(5)
    ref, qry = H.random_maps(rng, 25, 20, extent=6.0)
    p = O.make_params(match_xy_step_size=0.5, yaw_step_deg=30.0, match_threshold=0.6, match_threshold_dimension=0.4)
    a = O.match_maps(p, ref, qry, 7.0, 6.0, want_counts=True)
    b = oracle_np.match_maps(ref, qry, 7.0, 6.0, step=0.5, yaw_half=p.match_yaw_half_range,
                             yaw_step=p.match_yaw_angle_step_size, thr=0.6, thr_dim=0.4, want_counts=True)
    assert np.array_equal(a["counts"], b["counts"])
    assert a["best_hyp_index"] == b["best_hyp_index"]
    assert a["ref_idx"].tolist() == b["ref_idx"].tolist()


@pytest.mark.parametrize("mode", ["inter", "intra"])
@pytest.mark.parametrize("lsq", [1, 0])
def test_prtest_pass_criterion(gold, mode, lsq):
    """The reference's own acceptance rule for its synthetic scene: |dx|, |dy| <= 0.5 m and
    |dyaw| <= 3 deg against xyzYawExpected (place_recognition_test.cpp:242-275)."""
    maps, cases = gold
    c = cases[f"prtest_{mode}_lsq{lsq}"]
    assert c["found"]
    exp = maps[f"prtest_expected_{mode}"]
    got = np.array(c["xyz_yaw"])
    dx = round((exp[0] - got[0]) * 10000) / 10000
    dy = round((exp[1] - got[1]) * 10000) / 10000
    dyaw = (round((exp[3] - got[3]) * 10000) / 10000) * 180.0 / math.pi
    if dyaw > 180:
        dyaw -= 360
    elif dyaw < -180:
        dyaw += 360
    # with use_lsq the refined transform meets the reference's bound; without it the raw lattice
    # winner is the FIRST translation (scan order) that reaches the best count, which may sit up to
    # match_threshold_position (0.75 m in sloam.yaml) from the truth -- the reference logs that case
    # as an error too (it asserts nothing), so only the looser bound is pinned.
    tol = 0.5 if lsq else 0.75
    assert abs(dx) <= tol and abs(dy) <= tol and abs(dyaw) <= 3.0


def test_lattice_quirks():
    # 5 deg accumulates to 73 yaw candidates, 15 deg to 24, 2 deg to 180 (PR.cpp:140-145)
    for deg, n in ((5.0, 73), (15.0, 24), (2.0, 180)):
        p = O.make_params(yaw_step_deg=deg)
        assert len(O.enumerate_lattice(p, 5.0, 5.0)[3]) == n
    # intra defaults: half 5, step 0.5 -> one ring whose sample x = y = 0.0 exactly is skipped
    p = O.make_params(inter_loop_closure=0)
    tx, ty, ring, yaw = O.enumerate_lattice(p, 5.0, 5.0)
    assert len(tx) == 21 * 21 - 1 and not np.any((tx == 0.0) & (ty == 0.0))
    # sanity-check early return: outer step smaller than the lattice step (PR.cpp:169-175)
    p = O.make_params(match_xy_step_size=0.5)
    assert O.enumerate_lattice(p, 0.3, 0.3) is None
    r = O.match_maps(p, np.zeros((1, 7)), np.zeros((1, 7)), 0.3, 0.3)
    assert r["status"] == 1
    # zero half range -> zero rings, best stays -10000 (PR.cpp:125)
    r = O.match_maps(p, np.zeros((1, 7)), np.zeros((1, 7)), 0.0, 0.0)
    assert r["status"] == 0 and r["best_num_inliers"] == -10000 and r["hypotheses_scored"] == 0


def test_first_match_not_nearest():
    """PR.cpp:299-355: the recorded correspondence is the FIRST reference object within the
    threshold in input order, not the nearest."""
    ref = np.array([[1, 0.4, 0.0, 0, 1, 0, 0], [1, 0.05, 0.0, 0, 1, 0, 0]], float)
    qry = np.array([[1, 0.0, 0.0, 0, 1, 0, 0]], float)
    p = O.make_params(ignore_dimension=1)
    n, ri, qi = O.score_one(p, ref, qry, 1.0, 0.0, 0.0, 0.0)
    assert n == 1 and ri.tolist() == [0]
    # dimension rule is keyed on the REFERENCE object's d2 == d3 == 0 (PR.cpp:318)
    ref = np.array([[1, 0.0, 0.0, 0, 1.0, 0, 0]], float)        # cylinder: |1.0 - 1.9| = 0.9 < 1
    qry = np.array([[1, 0.0, 0.0, 0, 1.9, 5.0, 5.0]], float)
    assert O.score_one(O.make_params(), ref, qry, 1.0, 0.0, 0.0, 0.0)[0] == 1
    ref = np.array([[1, 0.0, 0.0, 0, 1.0, 0.1, 0]], float)      # cuboid: (0.9 + 4.9 + 5) / 3 > 1
    assert O.score_one(O.make_params(), ref, qry, 1.0, 0.0, 0.0, 0.0)[0] == 0
    # strict '<' on the distance (PR.cpp:333)
    ref = np.array([[1, 0.5, 0.0, 0, 0, 0, 0]], float)
    qry = np.array([[1, 0.0, 0.0, 0, 0, 0, 0]], float)
    assert O.score_one(O.make_params(ignore_dimension=1), ref, qry, 1.0, 0.0, 0.0, 0.0)[0] == 0


def test_svd_and_lsq():
    rng = np.random.default_rng(0)
    for _ in range(20):
        A = rng.normal(size=(3, 3))
        U, S, V = O.svd3(A)
        np.testing.assert_allclose(U @ np.diag(S) @ V.T, A, atol=1e-12)
        np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(S, np.linalg.svd(A, compute_uv=False), atol=1e-12)
    # Kabsch recovers a planted yaw + translation; planar (z = 0) input keeps det(R) = +1
    for planar in (False, True):
        src = rng.uniform(-10, 10, (12, 3))
        if planar:
            src[:, 2] = 0
        yaw, t = 0.7, np.array([3.0, -2.0, 0.0 if planar else 0.5])
        R = np.array([[math.cos(yaw), -math.sin(yaw), 0], [math.sin(yaw), math.cos(yaw), 0], [0, 0, 1]])
        tgt = src @ R.T + t
        xyzyaw, tf = O.solve_lsq(tgt, src)
        np.testing.assert_allclose(xyzyaw, [t[0], t[1], t[2], yaw], atol=1e-9)
        np.testing.assert_allclose(tf[:3, :3], R, atol=1e-9)


def test_triangle_descriptor_and_matching():
    rng = np.random.default_rng(3)
    tm = rng.uniform(-5, 5, (40, 6))
    td = np.vstack([tm[:10] + 0.001 * rng.normal(size=(10, 6)), rng.uniform(-5, 5, (30, 6))])
    mi, di, df = O.match_triangles(tm, td, 0.1)
    # brute-force numpy check of SC.cpp:66-99
    def desc(t):
        p = t.reshape(3, 2)
        return np.sort(np.linalg.norm(p - p.mean(0), axis=1))
    want = [(i, j) for i in range(40) for j in range(40) if np.linalg.norm(desc(tm[i]) - desc(td[j])) < 0.1]
    assert list(zip(mi.tolist(), di.tolist())) == want
    assert all((i, i) in want for i in range(10))
    # estimate_tf recovers a planted 2-D rigid transform (SC.cpp:122-138)
    a = rng.uniform(-5, 5, (9, 2))
    th = -1.1
    Rm = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
    b = a @ Rm.T + np.array([0.5, 4.0])
    tf = O.estimate_tf(a, b)
    np.testing.assert_allclose(tf[:2, :2], Rm, atol=1e-10)
    np.testing.assert_allclose(tf[:2, 2], [0.5, 4.0], atol=1e-10)
