"""Descriptor / hypothesis-generation half (SlideGraph front, semantic_clipper.cpp:49-138):
triangle descriptors + matching (with the optional class signature), 2-D Kabsch hypotheses from
matched triangles, and their scoring with the MatchMaps predicate."""
import ctypes as C
import math

import numpy as np
import pytest

import spr_helpers as H
from oracle import pyoracle as O
from slide_slam_b200 import capi, synth
from slide_slam_b200.place_recognition import PlaceRecognition


def delaunay_triangles(xy):
    """The reference's triangle source (observation.cpp:13-88): qhull 'Qt Qbb Qc Qz Q12 d'
    (scipy bundles the same qhull 8.0.2).  Returns (t x 6 coordinates, t x 3 vertex indices)."""
    from scipy.spatial import Delaunay
    tri = Delaunay(xy, qhull_options="Qt Qbb Qc Qz Q12")
    idx = tri.simplices.astype(np.int64)
    return np.ascontiguousarray(xy[idx].reshape(-1, 6)), idx


def test_estimate_tf_and_hypotheses_host():
    """Host-only entry points (no GPU needed): slide_pr_estimate_tf / slide_pr_triangle_hypotheses
    against the oracle's estimate_tf (SC.cpp:122-138)."""
    lib = capi.lib()
    rng = np.random.default_rng(2)
    for _ in range(50):
        k = int(rng.integers(3, 12))
        a = np.ascontiguousarray(rng.uniform(-20, 20, (k, 2)))
        th = rng.uniform(-math.pi, math.pi)
        Rm = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        b = np.ascontiguousarray(a @ Rm.T + rng.uniform(-5, 5, 2) + rng.normal(0, 0.01, (k, 2)))
        tf = np.zeros(9)
        assert lib.slide_pr_estimate_tf(capi.dptr(a), capi.dptr(b), k, capi.dptr(tf)) == 0
        np.testing.assert_allclose(tf.reshape(3, 3), O.estimate_tf(a, b), rtol=1e-9, atol=1e-9)
    # hypotheses from matched triangles == estimate_tf on the permuted vertices
    tm = np.ascontiguousarray(rng.uniform(-10, 10, (30, 6)))
    td = np.ascontiguousarray(rng.uniform(-10, 10, (40, 6)))
    mi = rng.integers(0, 30, 25).astype(np.int32)
    di = rng.integers(0, 40, 25).astype(np.int32)
    pm = np.ascontiguousarray(np.stack([O.triangle_descriptor(tm[i])[1] for i in mi]).astype(np.int32))
    pd = np.ascontiguousarray(np.stack([O.triangle_descriptor(td[j])[1] for j in di]).astype(np.int32))
    hyps = np.zeros((25, 4))
    assert lib.slide_pr_triangle_hypotheses(capi.dptr(tm), capi.dptr(td), capi.iptr(mi), capi.iptr(di), capi.iptr(pm),
                                            capi.iptr(pd), 25, capi.dptr(hyps)) == 0
    for k in range(25):
        a = td[di[k]].reshape(3, 2)[pd[k]]
        b = tm[mi[k]].reshape(3, 2)[pm[k]]
        tf = O.estimate_tf(a, b)
        np.testing.assert_allclose(hyps[k], [tf[0, 0], tf[1, 0], tf[0, 2], tf[1, 2]], rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_labeled_triangle_matching():
    rng = np.random.default_rng(5)
    tm = np.ascontiguousarray(rng.uniform(-30, 30, (500, 6)))
    td = np.ascontiguousarray(np.vstack([tm[rng.integers(0, 500, 200)] + 0.01 * rng.normal(size=(200, 6)),
                                         rng.uniform(-30, 30, (400, 6))]))
    lm = np.ascontiguousarray(rng.integers(0, 3, (500, 3)).astype(np.float64))
    ld = np.ascontiguousarray(rng.integers(0, 3, (600, 3)).astype(np.float64))
    mi, di, _ = O.match_triangles(tm, td, 0.1)
    keep = []
    for i, j in zip(mi, di):  # class signature: labels of the vertices paired by the sorted order must agree
        pmv, pdv = O.triangle_descriptor(tm[i])[1], O.triangle_descriptor(td[j])[1]
        if all(lm[i][pmv[k]] == ld[j][pdv[k]] for k in range(3)):
            keep.append((int(i), int(j)))
    pr = PlaceRecognition({})
    lib = capi.lib()
    cap = len(mi) + 8
    gm, gd = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    n = C.c_int64(0)
    rc = lib.slide_pr_match_triangles_labeled(pr._h, capi.dptr(tm), capi.dptr(lm), len(tm), capi.dptr(td), capi.dptr(ld), len(td),
                                              0.1, capi.iptr(gm), capi.iptr(gd), None, None, cap, C.byref(n))
    assert rc == 0 and n.value == len(keep) and 0 < len(keep) < len(mi)
    assert list(zip(gm[:n.value].tolist(), gd[:n.value].tolist())) == keep
    pr.close()


@pytest.mark.gpu
def test_triangle_hypotheses_recover_the_planted_transform():
    """Delaunay triangles -> descriptor matching -> one 2-D Kabsch hypothesis per match -> scoring
    with the MatchMaps predicate: the winner recovers the planted offset, and every count equals
    the oracle's single-hypothesis scorer on the same hypothesis list."""
    ref, qry, truth = synth.make_pair(300, seed=4242, classes="five", overlap=0.5, sigma=0.005)
    tr, ir = delaunay_triangles(np.ascontiguousarray(ref[:, 1:3]))
    tq, iq = delaunay_triangles(np.ascontiguousarray(qry[:, 1:3]))
    lr = np.ascontiguousarray(ref[ir, 0])
    lq = np.ascontiguousarray(qry[iq, 0])
    kw = dict(match_xy_step_size=0.5, match_threshold=0.5, match_threshold_dimension=1.0)
    pr = H_make_pr(kw)
    lib = capi.lib()
    cap = 200000
    gm, gd = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    pm, pd = np.zeros(3 * cap, np.int32), np.zeros(3 * cap, np.int32)
    n = C.c_int64(0)
    assert lib.slide_pr_match_triangles_labeled(pr._h, capi.dptr(tr), capi.dptr(lr), len(tr), capi.dptr(tq), capi.dptr(lq),
                                                len(tq), 0.05, capi.iptr(gm), capi.iptr(gd), capi.iptr(pm), capi.iptr(pd),
                                                cap, C.byref(n)) == 0
    m = n.value
    assert 0 < m <= cap
    hyps = np.zeros((m, 4))
    assert lib.slide_pr_triangle_hypotheses(capi.dptr(tr), capi.dptr(tq), capi.iptr(gm), capi.iptr(gd), capi.iptr(pm),
                                            capi.iptr(pd), m, capi.dptr(hyps)) == 0
    pr.prepare(ref, qry, 200.0, 200.0)
    res, counts = pr.score_hypotheses(hyps)
    op = O.make_params(**kw)
    sample = np.unique(np.concatenate([np.arange(0, m, max(m // 300, 1)), [res.best_hyp_index]]))
    for k in sample:
        assert counts[k] == O.score_one(op, ref, qry, *hyps[k])[0]
    assert res.best_num_inliers == counts.max() and res.best_hyp_index == int(np.argmax(counts))
    c, s, x, y = hyps[res.best_hyp_index]
    assert res.best_num_inliers >= 100
    assert abs(math.atan2(s, c) - truth["yaw"]) < 5e-3
    assert abs(x - truth["t"][0]) < 0.3 and abs(y - truth["t"][1]) < 0.3
    pr.close()


def H_make_pr(kw):
    return PlaceRecognition(H.rosparams_from_golden(kw))


@pytest.mark.gpu
@pytest.mark.parametrize("n_landmarks,thr", [(2000, 0.05), (20000, 0.02)])
def test_generator_half_on_the_device(n_landmarks, thr):
    """BASELINE config 2 / config 3 sized maps (T ~ 4 000 / 40 000 Delaunay triangles per map): the
    binned + windowed + radix-sorted device matching gives the oracle's all-pairs match list in the
    reference's order; the fused device pipeline (Kabsch per match, list scoring) gives the host
    pipeline's hypotheses (1e-9), the oracle's counts and the same winner."""
    ref, qry, truth = synth.make_pair(n_landmarks, seed=900 + n_landmarks, classes="five", overlap=0.5, sigma=0.005)
    tr, ir = delaunay_triangles(np.ascontiguousarray(ref[:, 1:3]))
    tq, iq = delaunay_triangles(np.ascontiguousarray(qry[:, 1:3]))
    lr, lq = np.ascontiguousarray(ref[ir, 0]), np.ascontiguousarray(qry[iq, 0])
    kw = dict(match_xy_step_size=0.5, match_threshold=0.5, match_threshold_dimension=1.0)
    pr = H_make_pr(kw)
    # (1) unlabeled match list == the oracle's, order included
    omi, odi, _ = O.match_triangles(tr, tq, thr)
    mi, di, pm, pd = pr.match_triangles(tr, tq, thr)
    assert len(omi) > 100 and mi.tolist() == omi.tolist() and di.tolist() == odi.tolist()
    for k in range(0, len(mi), max(len(mi) // 200, 1)):
        assert pm[k].tolist() == O.triangle_descriptor(tr[mi[k]])[1].tolist()
        assert pd[k].tolist() == O.triangle_descriptor(tq[di[k]])[1].tolist()
    # (2) fused pipeline with the class signature
    pr.prepare(ref, qry, 200.0, 200.0)
    res, info, L = pr.generate_and_score(tr, tq, thr, lr, lq)
    lmi, ldi, lpm, lpd = pr.match_triangles(tr, tq, thr, lr, lq)
    assert info.n_matches == len(lmi) and L["model_idx"].tolist() == lmi.tolist() and L["data_idx"].tolist() == ldi.tolist()
    keep = set(zip(omi.tolist(), odi.tolist()))
    assert all((a, b) in keep for a, b in zip(lmi.tolist(), ldi.tolist())) and 0 < len(lmi) < len(omi)
    lib = capi.lib()
    m = len(lmi)
    hyps = np.zeros((m, 4))
    lpm_c, lpd_c = np.ascontiguousarray(lpm.reshape(-1)), np.ascontiguousarray(lpd.reshape(-1))
    assert lib.slide_pr_triangle_hypotheses(capi.dptr(tr), capi.dptr(tq), capi.iptr(np.ascontiguousarray(lmi)), capi.iptr(np.ascontiguousarray(ldi)),
                                            capi.iptr(lpm_c), capi.iptr(lpd_c), m, capi.dptr(hyps)) == 0
    np.testing.assert_allclose(L["hyps"], hyps, rtol=1e-9, atol=1e-9)     # closed-form 2 x 2 polar factor vs Jacobi SVD
    op = O.make_params(**kw)
    for k in np.unique(np.concatenate([np.arange(0, m, max(m // (150 if n_landmarks <= 2000 else 24), 1)), [res.best_hyp_index]])):
        a, b = tq[ldi[k]].reshape(3, 2)[lpd[k]], tr[lmi[k]].reshape(3, 2)[lpm[k]]
        tf = O.estimate_tf(a, b)                                         # the oracle's estimate_tf (SC.cpp:122-138)
        np.testing.assert_allclose(L["hyps"][k], [tf[0, 0], tf[1, 0], tf[0, 2], tf[1, 2]], rtol=1e-9, atol=1e-9)
        assert L["counts"][k] == O.score_one(op, ref, qry, *L["hyps"][k])[0]
    assert res.best_num_inliers == L["counts"].max() and res.best_hyp_index == int(np.argmax(L["counts"]))
    c, s, x, y = L["hyps"][res.best_hyp_index]
    assert res.best_num_inliers >= 0.3 * n_landmarks                        # half the landmarks are shared
    assert abs(math.atan2(s, c) - truth["yaw"]) < 2e-3
    assert abs(x - truth["t"][0]) < 0.3 and abs(y - truth["t"][1]) < 0.3
    # (3) edge cases: no matches, empty lists
    r0, i0, _ = pr.generate_and_score(tr, tq, 1e-12, lr, lq, want_lists=False)
    assert i0.n_matches == 0 and r0.best_hyp_index == -1
    r1, i1, _ = pr.generate_and_score(tr[:0], tq, thr, want_lists=False)
    assert i1.n_matches == 0
    pr.close()
