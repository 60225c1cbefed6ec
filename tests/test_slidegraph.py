"""SlideGraph entry points: the host Delaunay triangulation (CPU suite, against qhull through scipy --
the same qhull 8.0.2 the reference links) and, on the GPU, run_semantic_clipper /
findInterLoopClosureWithClipper against the oracle's restatement of semantic_clipper.cpp:140-275 and
against the planted transform of a synthetic map pair."""
import math

import numpy as np
import pytest

from oracle import pyoracle as O
from slide_slam_b200 import capi, synth
from slide_slam_b200.place_recognition import PlaceRecognition, delaunay


def qhull_triangles(xy):
    from scipy.spatial import Delaunay
    return Delaunay(xy, qhull_options="Qt Qbb Qc Qz Q12").simplices   # observation.cpp:25-26


def as_set(tri):
    return {tuple(sorted(t)) for t in np.asarray(tri).tolist()}


@pytest.mark.parametrize("n,seed", [(3, 0), (4, 1), (7, 2), (50, 3), (365, 4), (2000, 5), (20000, 6)])
def test_delaunay_equals_qhull_on_points_in_general_position(n, seed):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-300, 300, (n, 2))
    tri = delaunay(xy)
    assert as_set(tri) == as_set(qhull_triangles(xy))
    p = xy[tri]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    assert (cross > 0).all()                                            # counter-clockwise


def test_delaunay_on_the_reference_fixture_maps():
    import spr_helpers as H
    maps = H.golden_maps()
    for name in ("forest0", "forest1", "parking0", "indoor0"):
        xy = np.ascontiguousarray(maps[name][:, 1:3])
        assert as_set(delaunay(xy)) == as_set(qhull_triangles(xy)), name


def test_delaunay_degenerate_inputs():
    grid = np.array([(i, j) for i in range(12) for j in range(9)], float)   # cocircular quadruples everywhere
    tri = delaunay(grid)
    p = grid[tri]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    assert len(tri) == 2 * 11 * 8 and (cross > 0).all() and abs(cross.sum() / 2 - 88.0) < 1e-9   # a triangulation of the 11 x 8 rectangle
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 10, (50, 2))
    assert len(delaunay(np.vstack([pts, pts]))) == len(qhull_triangles(pts))     # coincident points add no vertex
    assert len(delaunay(np.column_stack([np.arange(10.0), 2 * np.arange(10.0)]))) == 0   # collinear: no Delaunay facet
    assert len(delaunay(np.zeros((0, 2)))) == 0 and len(delaunay(np.ones((2, 2)))) == 0
    with pytest.raises(capi.SlidePrError):
        delaunay(np.array([[0.0, 0.0], [1.0, np.nan], [2.0, 1.0]]))


def _tris6(xy, tri):
    return np.ascontiguousarray(xy[tri].reshape(-1, 6))


@pytest.mark.gpu
@pytest.mark.parametrize("n,source", [(300, "qhull"), (300, "own"), (1200, "own")])
def test_run_semantic_clipper_against_the_oracle(n, source):
    ref, qry, truth = synth.make_pair(n, seed=77 + n, classes="five", overlap=0.5, sigma=0.005)
    rx, qx = np.ascontiguousarray(ref[:, 1:3]), np.ascontiguousarray(qry[:, 1:3])
    tr = _tris6(rx, qhull_triangles(rx) if source == "qhull" else delaunay(rx))
    tq = _tris6(qx, qhull_triangles(qx) if source == "qhull" else delaunay(qx))
    sigma, eps, min_pairs, thr = 0.1, 0.1, 10, 0.05
    mi, di, _ = O.match_triangles(tr, tq, thr)
    u0 = np.random.default_rng(5).uniform(0, 1, 3 * len(mi))
    ofound, otf, oinfo = O.run_semantic_clipper(tr, tq, sigma, eps, min_pairs, thr, u0)
    pr = PlaceRecognition({})
    if source == "qhull":    # triangles handed over in qhull's facet order, like the reference sees them
        found, tf, info = pr.run_semantic_clipper(ref, qry, sigma, eps, min_pairs, thr, tr, tq, u0)
    else:                    # internal triangulation == delaunay(): same triangles, same order
        found, tf, info = pr.run_semantic_clipper(ref, qry, sigma, eps, min_pairs, thr, u0=u0)
    assert found == ofound and found
    assert info.n_triangle_matches == oinfo["n_triangle_matches"] == len(mi)
    assert info.n_associations == oinfo["n_associations"] and info.nnz_upper == oinfo["nnz"]
    assert info.n_inliers == oinfo["n_inliers"]
    assert abs(info.score - oinfo["score"]) < 1e-5 * abs(oinfo["score"])
    np.testing.assert_allclose(tf, otf, rtol=1e-6, atol=1e-6)   # floating point: 2-D Kabsch of the same selected pairs
    # the reference's tfFromQuery2Ref maps model (reference-map) points onto data (query-map) points: p_B = R^T (p_A - t)
    yaw = math.atan2(tf[1, 0], tf[0, 0])
    assert abs(np.angle(np.exp(1j * (yaw + truth["yaw"])))) < 2e-3
    pr.close()


@pytest.mark.gpu
def test_find_inter_loop_closure_with_clipper_recovers_the_planted_transform():
    ref, qry, truth = synth.make_pair(800, seed=4100, classes="five", overlap=0.5, sigma=0.005)
    # objects at exactly (0, 0) are dropped as invalid (PR.cpp:576-612)
    ref2 = np.vstack([ref, [[1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0]]])
    pr = PlaceRecognition({}, slidegraph={"descriptor_matching_threshold": 0.05, "sigma": 0.1, "epsilon": 0.1, "seed": 3})
    found, tf = pr.findInterLoopClosureWithClipper(ref2, qry)
    assert found and pr.last_sc.n_inliers >= 10
    yaw = math.atan2(tf[1, 0], tf[0, 0])
    assert abs(np.angle(np.exp(1j * (yaw - truth["yaw"])))) < 2e-3       # tfFromQueryToRef after the inversion
    assert abs(tf[0, 3] - truth["t"][0]) < 0.2 and abs(tf[1, 3] - truth["t"][1]) < 0.2
    # the class signature only removes candidates; the closure is still found
    pr2 = PlaceRecognition({}, slidegraph={"descriptor_matching_threshold": 0.05, "use_class_signature": 1, "seed": 3})
    found2, tf2 = pr2.findInterLoopClosureWithClipper(ref, qry)
    assert found2 and pr2.last_sc.n_triangle_matches < pr.last_sc.n_triangle_matches
    np.testing.assert_allclose(tf2, tf, atol=5e-2)
    # size gate (PR.cpp:618) and unrelated maps
    pr3 = PlaceRecognition({}, slidegraph={"min_num_map_objects_to_start": 5000})
    assert pr3.findInterLoopClosureWithClipper(ref, qry)[0] is False
    other, _, _ = synth.make_pair(800, seed=999, classes="five")
    f4, _ = pr.findInterLoopClosureWithClipper(other, qry)
    assert not f4 or pr.last_sc.n_inliers < 40
    pr.close(); pr2.close(); pr3.close()
