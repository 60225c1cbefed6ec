"""CPU suite: the CLIPPER oracle (oracle/clipper_oracle.c) against the known answers the
reference's own tests hold (tests/clipper_kats.py) -- the first parity evidence of this repo that is
pinned by reference-held vectors rather than by oracle-generated goldens."""
import numpy as np
import pytest

from oracle import pyoracle as O
import clipper_kats as K


def test_affinity_matrix_equals_the_reference_mtrue():
    """CSO/test/affinity_test.cpp:15-107"""
    model, data = K.kat_clouds()
    p = O.clipper_params()
    A, Mu = O.clipper_score_pairwise(p, model, data)
    n = model.shape[1] * data.shape[1]
    assert A.shape == (n, 2)
    for i in range(model.shape[1]):                      # all-to-all hypothesis, affinity_test.cpp:67-73
        for j in range(data.shape[1]):
            assert A[i * data.shape[1] + j].tolist() == [i, j]
    M = O.clipper_affinity_matrix(Mu)
    assert np.array_equal(np.diag(M), np.ones(n))        # :84
    assert np.array_equal(M, M.T)                        # :87
    assert np.array_equal(M, K.MTRUE)                    # :106  exact, as EXPECT_EQ(M, Mtrue)
    assert np.array_equal((M != 0).astype(float), M)     # M == C for perfect data (:91)


def test_dense_clique_selects_the_three_true_associations():
    """CSO/test/clipper_test.cpp:15-68.  The reference draws u0 from std::random_device; the
    gradient ascent is local, so the outcome depends on u0: of 200 uniform u0 (numpy seeds 0..199)
    187 reach the 3-clique {(0,0), (1,1), (2,2)} the test expects and 13 stop at one of the graph's
    2-cliques (the reference's own test is therefore probabilistic, ~93 %).  Checked here: u0 = ones
    (the natural deterministic choice) and the first seeds reach it; the rate over 200 seeds is the
    one an independent numpy restatement of clipper.cpp:172-323 gave (tools/clipper_np.py)."""
    model, data = K.kat_clouds()
    p = O.clipper_params()
    A, Mu = O.clipper_score_pairwise(p, model, data)
    n_ok = 0
    for seed in [None] + list(range(200)):
        u0 = np.ones(len(A)) if seed is None else np.random.default_rng(seed).uniform(0, 1, len(A))
        sol = O.clipper_find_dense_clique(p, Mu, u0)
        inl = A[sol["nodes"]]
        ok = len(inl) == 3 and (inl[:, 0] == inl[:, 1]).all() and sorted(inl[:, 0].tolist()) == [0, 1, 2]
        if seed is None or seed in (0, 2, 3, 4):
            assert ok, seed
            assert abs(sol["score"] - 3.0) < 1e-6       # spectral radius of a 3-clique with unit diagonal
        else:
            assert len(inl) in (2, 3)                   # always a clique of the consistency graph
            assert all(K.MTRUE[a, b] == 1 for a in sol["nodes"] for b in sol["nodes"])
        n_ok += int(ok and seed is not None)
    assert n_ok == 187


def test_dsd_known_answers():
    """CSO/test/dsd_test.cpp:15-44 and :48-80"""
    M = K.dsd_matrix()
    Mu = np.triu(M, 1)
    assert O.clipper_dsd(Mu).tolist() == K.DSD_NODES
    assert O.clipper_dsd(Mu, K.DSD_S).tolist() == K.DSD_NODES


def test_k2ij_enumerates_the_strict_upper_triangle():
    import ctypes as C
    L = O.lib()
    for n in (2, 3, 7, 12, 40):
        seen = []
        for k in range(n * (n - 1) // 2):
            i, j = C.c_longlong(), C.c_longlong()
            L.clipper_oracle_k2ij(C.c_longlong(k), C.c_longlong(n), C.byref(i), C.byref(j))
            seen.append((i.value, j.value))
        assert seen == [(i, j) for i in range(n) for j in range(i + 1, n)]


def test_roundings_and_planted_clique():
    """40 planted consistent associations among 160: the clique is found from several u0."""
    rng = np.random.default_rng(4)
    n1 = 40
    model = rng.uniform(-20, 20, (2, n1))
    a = 0.7
    R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    data = R @ model + np.array([[3.0], [-2.0]]) + rng.normal(0, 0.002, (2, n1))
    A = np.array([(i, i) for i in range(n1)] + [(int(rng.integers(n1)), int(rng.integers(n1))) for _ in range(120)], np.int32)
    p = O.clipper_params(sigma=0.01, epsilon=0.06)
    _, Mu = O.clipper_score_pairwise(p, model, data, A)
    for s in range(3):
        sol = O.clipper_find_dense_clique(p, Mu, np.random.default_rng(s).uniform(0, 1, len(A)))
        good = set(sol["nodes"].tolist())
        true_nodes = {k for k in range(len(A)) if A[k, 0] == A[k, 1]}
        assert len(good & true_nodes) >= 30 and not (good - true_nodes)  # DSD_HEU keeps round(F) nodes: fewer than 40 with noisy weights
    pn = O.clipper_params(sigma=0.01, epsilon=0.06, rounding=O.ROUND_NONZERO)
    soln = O.clipper_find_dense_clique(pn, Mu, np.ones(len(A)))
    assert len(set(soln["nodes"].tolist()) & set(range(40))) >= 38
