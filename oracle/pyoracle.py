"""ctypes binding of the CPU oracle (oracle/slide_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (slide_slam_b200) never imports it.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libslide_oracle.so")
_ALT_PATH = os.path.join(_HERE, "libslide_oracle_alt.so")  # -DSLIDE_ORACLE_ALT_SUM_ORDER (see slide_oracle.c)


class Params(C.Structure):
    """slide_oracle_params -- rosparams of place_recognition.cpp:24-75."""

    _fields_ = [
        ("compute_budget_sec", C.c_double),
        ("dilation_factor", C.c_double),
        ("match_xy_step_size", C.c_double),
        ("match_yaw_half_range", C.c_double),
        ("disable_yaw_search", C.c_int),
        ("match_yaw_angle_step_size", C.c_double),
        ("match_threshold", C.c_double),
        ("match_threshold_dimension", C.c_double),
        ("ignore_dimension", C.c_int),
        ("min_num_inliers", C.c_int),
        ("use_lsq", C.c_int),
        ("min_num_map_objects_to_start", C.c_int),
        ("match_x_half_range_intra", C.c_double),
        ("match_y_half_range_intra", C.c_double),
        ("match_yaw_half_range_intra", C.c_double),
        ("inter_loop_closure", C.c_int),
    ]


class MatchResult(C.Structure):
    _fields_ = [
        ("status", C.c_int),
        ("best_num_inliers", C.c_int),
        ("R_t", C.c_double * 9),
        ("n_matched", C.c_int),
        ("hypotheses_scored", C.c_longlong),
        ("best_hyp_index", C.c_longlong),
        ("n_rings", C.c_int),
        ("n_yaw", C.c_int),
    ]


class TfResult(C.Structure):
    _fields_ = [
        ("found", C.c_int),
        ("match_status", C.c_int),
        ("best_num_inliers", C.c_int),
        ("hypotheses_scored", C.c_longlong),
        ("best_hyp_index", C.c_longlong),
        ("R_t", C.c_double * 9),
        ("xyz_yaw", C.c_double * 4),
        ("transform", C.c_double * 16),
        ("centroid_ref", C.c_double * 2),
        ("centroid_qry", C.c_double * 2),
        ("half_x", C.c_double),
        ("half_y", C.c_double),
        ("yaw_half", C.c_double),
        ("n_matched", C.c_int),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "slide_oracle.c")
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("slide_oracle.c", "slide_oracle.h", "clipper_oracle.c", "clipper_oracle.h"))
    stale = any((not os.path.exists(q)) or os.path.getmtime(q) < newest for q in (_LIB_PATH, _ALT_PATH))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libslide_oracle.so", "libslide_oracle_alt.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_libs = {}
_use_alt = False
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


@contextlib.contextmanager
def alt_sum_order():
    """Inside the block every call goes to the oracle built with the OTHER candidate for Eigen's
    summation order in PR.cpp:257-258: x' = c*qx + ((-s)*qy + x)."""
    global _use_alt
    old, _use_alt = _use_alt, True
    try:
        yield
    finally:
        _use_alt = old


def lib():
    if _use_alt not in _libs:
        build()
        L = C.CDLL(_ALT_PATH if _use_alt else _LIB_PATH)
        L.slide_oracle_deg2rad.restype = C.c_double
        L.slide_oracle_deg2rad.argtypes = [C.c_double]
        L.slide_oracle_enumerate_lattice.restype = C.c_longlong
        L.slide_oracle_enumerate_lattice.argtypes = [
            C.POINTER(Params), C.c_double, C.c_double, _dp, _dp, _ip, C.c_longlong, _dp, C.c_int, _ip]
        L.slide_oracle_match_maps.argtypes = [
            C.POINTER(Params), _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double,
            C.c_longlong, C.c_longlong, _ip, _ip, _ip, C.c_longlong, C.POINTER(MatchResult)]
        L.slide_oracle_match_maps_mt.argtypes = [
            C.POINTER(Params), _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double,
            C.c_longlong, C.c_longlong, _ip, _ip, C.c_int, C.POINTER(MatchResult)]
        L.slide_oracle_match_maps_mt_counts.argtypes = [
            C.POINTER(Params), _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double,
            C.c_longlong, C.c_longlong, _ip, _ip, C.c_int, _ip, C.c_longlong, C.POINTER(MatchResult)]
        L.slide_oracle_score_one.argtypes = [
            C.POINTER(Params), _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_double,
            C.c_double, _ip, _ip]
        L.slide_oracle_find_transformation.argtypes = [
            C.POINTER(Params), _dp, C.c_int, _dp, C.c_int, _ip, _ip, C.c_int, C.POINTER(TfResult)]
        L.slide_oracle_find_inter_loop_closure.argtypes = [
            C.POINTER(Params), _dp, C.c_int, _dp, C.c_int, C.c_int, _dp, C.POINTER(TfResult)]
        L.slide_oracle_solve_lsq.argtypes = [_dp, _dp, C.c_int, _dp, _dp]
        L.slide_oracle_svd3.argtypes = [_dp, _dp, _dp, _dp]
        L.slide_oracle_triangle_descriptor.argtypes = [_dp, _dp, _ip]
        L.slide_oracle_match_triangles.restype = C.c_longlong
        L.slide_oracle_match_triangles.argtypes = [
            _dp, C.c_int, _dp, C.c_int, C.c_double, _ip, _ip, _dp, C.c_longlong]
        L.slide_oracle_estimate_tf.argtypes = [_dp, _dp, C.c_int, _dp]
        L.slide_oracle_find_intra_loop_closure.argtypes = [
            C.POINTER(Params), _dp, C.c_int, _dp, C.c_int, _dp, _dp, C.c_int, _dp, C.POINTER(TfResult)]
        _libs[_use_alt] = L
    return _libs[_use_alt]


def deg2rad(deg: float) -> float:
    return lib().slide_oracle_deg2rad(deg)


def make_params(**kw) -> Params:
    """Defaults of place_recognition.cpp:24-75 (budget disabled); angles given in degrees
    through the *_deg keywords are converted with the reference's expression."""
    p = Params()
    lib().slide_oracle_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "yaw_step_deg":
            p.match_yaw_angle_step_size = deg2rad(v)
        elif k == "yaw_half_range_deg":
            p.match_yaw_half_range = deg2rad(v)
        elif k == "yaw_half_range_intra_deg":
            p.match_yaw_half_range_intra = deg2rad(v)
        else:
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
    return p


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def enumerate_lattice(p: Params, half_x: float, half_y: float):
    n_yaw = C.c_int(0)
    n = lib().slide_oracle_enumerate_lattice(C.byref(p), half_x, half_y, None, None, None, 0,
                                             None, 0, C.byref(n_yaw))
    if n < 0:
        return None
    tx = np.zeros(n, np.float64)
    ty = np.zeros(n, np.float64)
    ring = np.zeros(n, np.int32)
    yaw = np.zeros(n_yaw.value, np.float64)
    lib().slide_oracle_enumerate_lattice(
        C.byref(p), half_x, half_y, tx.ctypes.data_as(_dp), ty.ctypes.data_as(_dp),
        ring.ctypes.data_as(_ip), n, yaw.ctypes.data_as(_dp), n_yaw.value, C.byref(n_yaw))
    return tx, ty, ring, yaw


def match_maps(p: Params, ref7, qry7, half_x, half_y, hyp_begin=0, hyp_end=-1,
               want_counts=False, n_threads=1):
    ref7, rp = _d(ref7)
    qry7, qp = _d(qry7)
    n_ref, n_qry = ref7.shape[0] if ref7.size else 0, qry7.shape[0] if qry7.size else 0
    ri = np.full(max(n_qry, 1), -1, np.int32)
    qi = np.full(max(n_qry, 1), -1, np.int32)
    res = MatchResult()
    counts = None
    cp, cap = None, 0
    if want_counts:
        lat = enumerate_lattice(p, half_x, half_y)
        total = 0 if lat is None else len(lat[0]) * len(lat[3])
        end = total if hyp_end < 0 else min(hyp_end, total)
        cap = max(end - hyp_begin, 0)
        counts = np.full(max(cap, 1), -1, np.int32)
        cp = counts.ctypes.data_as(_ip)
    if n_threads == 1:
        lib().slide_oracle_match_maps(C.byref(p), rp, n_ref, qp, n_qry, half_x, half_y, hyp_begin,
                                      hyp_end, ri.ctypes.data_as(_ip), qi.ctypes.data_as(_ip),
                                      cp, cap, C.byref(res))
        if counts is not None:
            counts = counts[:cap]
    else:
        lib().slide_oracle_match_maps_mt_counts(C.byref(p), rp, n_ref, qp, n_qry, half_x, half_y,
                                                hyp_begin, hyp_end, ri.ctypes.data_as(_ip),
                                                qi.ctypes.data_as(_ip), n_threads, cp, cap, C.byref(res))
        if counts is not None:
            counts = counts[:cap]
    k = max(res.n_matched, 0)
    return {
        "status": res.status, "best_num_inliers": res.best_num_inliers,
        "R_t": np.array(res.R_t[:], np.float64).reshape(3, 3),
        "ref_idx": ri[:k].copy(), "qry_idx": qi[:k].copy(),
        "hypotheses_scored": res.hypotheses_scored, "best_hyp_index": res.best_hyp_index,
        "n_rings": res.n_rings, "n_yaw": res.n_yaw, "counts": counts,
    }


def score_one(p: Params, ref7, qry7, c, s, x, y):
    ref7, rp = _d(ref7)
    qry7, qp = _d(qry7)
    n_qry = qry7.shape[0]
    ri = np.full(max(n_qry, 1), -1, np.int32)
    qi = np.full(max(n_qry, 1), -1, np.int32)
    n = lib().slide_oracle_score_one(C.byref(p), rp, ref7.shape[0], qp, n_qry, c, s, x, y,
                                     ri.ctypes.data_as(_ip), qi.ctypes.data_as(_ip))
    return n, ri[:n].copy(), qi[:n].copy()


def _tf_dict(res: TfResult, ri=None, qi=None):
    k = max(res.n_matched, 0)
    d = {
        "found": bool(res.found), "match_status": res.match_status,
        "best_num_inliers": res.best_num_inliers, "hypotheses_scored": res.hypotheses_scored,
        "best_hyp_index": res.best_hyp_index,
        "R_t": np.array(res.R_t[:], np.float64).reshape(3, 3),
        "xyz_yaw": np.array(res.xyz_yaw[:], np.float64),
        "transform": np.array(res.transform[:], np.float64).reshape(4, 4),
        "centroid_ref": np.array(res.centroid_ref[:]), "centroid_qry": np.array(res.centroid_qry[:]),
        "half_x": res.half_x, "half_y": res.half_y, "yaw_half": res.yaw_half, "n_matched": k,
    }
    if ri is not None:
        d["ref_idx"], d["qry_idx"] = ri[:k].copy(), qi[:k].copy()
    return d


def find_transformation(p: Params, ref7, qry7, n_threads=1):
    ref7, rp = _d(ref7)
    qry7, qp = _d(qry7)
    n_qry = qry7.shape[0]
    ri = np.full(max(n_qry, 1), -1, np.int32)
    qi = np.full(max(n_qry, 1), -1, np.int32)
    res = TfResult()
    lib().slide_oracle_find_transformation(C.byref(p), rp, ref7.shape[0], qp, n_qry,
                                           ri.ctypes.data_as(_ip), qi.ctypes.data_as(_ip),
                                           n_threads, C.byref(res))
    return _tf_dict(res, ri, qi)


def find_inter_loop_closure(p: Params, ref7, qry7, n_threads=1):
    ref7, rp = _d(ref7)
    qry7, qp = _d(qry7)
    tf = np.zeros(16, np.float64)
    res = TfResult()
    found = lib().slide_oracle_find_inter_loop_closure(
        C.byref(p), rp, ref7.shape[0], qp, qry7.shape[0], n_threads, tf.ctypes.data_as(_dp),
        C.byref(res))
    return bool(found), tf.reshape(4, 4), _tf_dict(res)


def find_intra_loop_closure(p: Params, meas7, submap7, query_pose, candidate_pose, n_threads=1):
    meas7, mp = _d(meas7)
    submap7, sp = _d(submap7)
    qp_, qpp = _d(np.reshape(query_pose, 16))
    cp_, cpp = _d(np.reshape(candidate_pose, 16))
    tf = np.zeros(16, np.float64)
    res = TfResult()
    found = lib().slide_oracle_find_intra_loop_closure(
        C.byref(p), mp, meas7.shape[0] if meas7.size else 0, sp, submap7.shape[0] if submap7.size else 0,
        qpp, cpp, n_threads, tf.ctypes.data_as(_dp), C.byref(res))
    return bool(found), tf.reshape(4, 4), _tf_dict(res)


def solve_lsq(tgt3, src3):
    tgt3, tp = _d(tgt3)
    src3, sp = _d(src3)
    xyzyaw = np.zeros(4)
    tf = np.zeros(16)
    lib().slide_oracle_solve_lsq(tp, sp, tgt3.shape[0], xyzyaw.ctypes.data_as(_dp),
                                 tf.ctypes.data_as(_dp))
    return xyzyaw, tf.reshape(4, 4)


def svd3(A):
    A, ap = _d(A)
    U, S, V = np.zeros(9), np.zeros(3), np.zeros(9)
    lib().slide_oracle_svd3(ap, U.ctypes.data_as(_dp), S.ctypes.data_as(_dp), V.ctypes.data_as(_dp))
    return U.reshape(3, 3), S, V.reshape(3, 3)


def triangle_descriptor(tri6):
    tri6, tp = _d(tri6)
    d = np.zeros(3)
    perm = np.zeros(3, np.int32)
    lib().slide_oracle_triangle_descriptor(tp, d.ctypes.data_as(_dp), perm.ctypes.data_as(_ip))
    return d, perm


def match_triangles(tris_model, tris_data, threshold):
    tm, mp = _d(np.reshape(tris_model, (-1, 6)))
    td, dp = _d(np.reshape(tris_data, (-1, 6)))
    n = lib().slide_oracle_match_triangles(mp, tm.shape[0], dp, td.shape[0], threshold, None, None,
                                           None, 0)
    mi = np.zeros(max(n, 1), np.int32)
    di = np.zeros(max(n, 1), np.int32)
    df = np.zeros(max(n, 1), np.float64)
    lib().slide_oracle_match_triangles(mp, tm.shape[0], dp, td.shape[0], threshold,
                                       mi.ctypes.data_as(_ip), di.ctypes.data_as(_ip),
                                       df.ctypes.data_as(_dp), n)
    return mi[:n], di[:n], df[:n]


def estimate_tf(a2, b2):
    a2, ap = _d(a2)
    b2, bp = _d(b2)
    tf = np.zeros(9)
    lib().slide_oracle_estimate_tf(ap, bp, a2.shape[0], tf.ctypes.data_as(_dp))
    return tf.reshape(3, 3)


# ---------------------------------------------------------------------------------------------
# SlideGraph half: CLIPPER affinity + dense clique (oracle/clipper_oracle.c)
# ---------------------------------------------------------------------------------------------
ROUND_NONZERO, ROUND_DSD, ROUND_DSD_HEU = 0, 1, 2


class ClipperParams(C.Structure):
    """clipper_oracle_params: clipper::Params (clipper.h:28-60) + EuclideanDistance::Params."""
    _fields_ = [("sigma", C.c_double), ("epsilon", C.c_double), ("mindist", C.c_double),
                ("tol_u", C.c_double), ("tol_F", C.c_double), ("tol_Fop", C.c_double),
                ("maxiniters", C.c_int), ("maxoliters", C.c_int), ("beta", C.c_double), ("maxlsiters", C.c_int),
                ("eps", C.c_double), ("affinityeps", C.c_double), ("rescale_u0", C.c_int), ("rounding", C.c_int)]


class ClipperSolution(C.Structure):
    _fields_ = [("ifinal", C.c_int), ("n_nodes", C.c_int), ("score", C.c_double)]


class ScInfo(C.Structure):
    _fields_ = [("found", C.c_int), ("n_triangle_matches", C.c_longlong), ("n_associations", C.c_int),
                ("nnz", C.c_longlong), ("n_inliers", C.c_int), ("score", C.c_double)]


def clipper_params(**kw) -> ClipperParams:
    p = ClipperParams()
    lib().clipper_oracle_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _cols(D):
    """points as columns of a dim x n matrix (Eigen's invariants::Data), handed over column-major"""
    D = np.asarray(D, np.float64)
    flat = np.ascontiguousarray(D.T)  # n x dim row-major == dim x n column-major
    return flat, flat.ctypes.data_as(_dp), D.shape[0], D.shape[1]


def clipper_all_to_all(n1, n2):
    A = np.zeros((n1 * n2, 2), np.int32)
    lib().clipper_oracle_all_to_all(n1, n2, A.ctypes.data_as(_ip))
    return A


def clipper_score_pairwise(p: ClipperParams, D1, D2, A=None):
    """scorePairwiseConsistency (clipper.cpp:21-65).  Returns (A, M_upper)."""
    f1, p1, dim, n1 = _cols(D1)
    f2, p2, dim2, n2 = _cols(D2)
    assert dim == dim2
    A = clipper_all_to_all(n1, n2) if A is None or len(A) == 0 else np.ascontiguousarray(A, np.int32)
    m = len(A)
    M = np.zeros((m, m), np.float64)
    L = lib()
    L.clipper_oracle_score_pairwise.restype = C.c_longlong
    L.clipper_oracle_score_pairwise(C.byref(p), p1, n1, p2, n2, dim, A.ctypes.data_as(_ip), m, M.ctypes.data_as(_dp))
    return A, M


def clipper_affinity_matrix(M_upper):
    m = len(M_upper)
    out = np.zeros((m, m), np.float64)
    Mu = np.ascontiguousarray(M_upper, np.float64)
    lib().clipper_oracle_affinity_matrix(Mu.ctypes.data_as(_dp), m, out.ctypes.data_as(_dp))
    return out


def clipper_find_dense_clique(p: ClipperParams, M_upper, u0):
    Mu = np.ascontiguousarray(M_upper, np.float64)
    n = len(Mu)
    u0 = np.ascontiguousarray(u0, np.float64)
    nodes = np.zeros(max(n, 1), np.int32)
    u = np.zeros(max(n, 1), np.float64)
    sol = ClipperSolution()
    k = lib().clipper_oracle_find_dense_clique(C.byref(p), Mu.ctypes.data_as(_dp), n, u0.ctypes.data_as(_dp), C.byref(sol),
                                               nodes.ctypes.data_as(_ip), u.ctypes.data_as(_dp))
    return {"nodes": nodes[:k].copy(), "u": u[:n], "score": sol.score, "ifinal": sol.ifinal}


def clipper_dsd(M_upper, S=None):
    Mu = np.ascontiguousarray(M_upper, np.float64)
    n = len(Mu)
    nodes = np.zeros(max(n, 1), np.int32)
    if S is None or len(S) == 0:
        k = lib().clipper_oracle_dsd(Mu.ctypes.data_as(_dp), n, None, 0, nodes.ctypes.data_as(_ip))
    else:
        S = np.ascontiguousarray(S, np.int32)
        k = lib().clipper_oracle_dsd(Mu.ctypes.data_as(_dp), n, S.ctypes.data_as(_ip), len(S), nodes.ctypes.data_as(_ip))
    return nodes[:k].copy()


def run_semantic_clipper(tris_model, tris_data, sigma, epsilon, min_num_pairs, matching_threshold, u0):
    """semantic_clipper::run_semantic_clipper (SC.cpp:140-275) from the triangle lists on."""
    tm, mp = _d(np.reshape(tris_model, (-1, 6)))
    td, dp = _d(np.reshape(tris_data, (-1, 6)))
    u0 = np.ascontiguousarray(u0, np.float64)
    tf = np.zeros(16)
    info = ScInfo()
    L = lib()
    L.clipper_oracle_run_semantic_clipper.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                                      _dp, C.c_int, _dp, C.POINTER(ScInfo)]
    found = L.clipper_oracle_run_semantic_clipper(mp, tm.shape[0], dp, td.shape[0], sigma, epsilon, min_num_pairs,
                                                  matching_threshold, u0.ctypes.data_as(_dp), len(u0), tf.ctypes.data_as(_dp), C.byref(info))
    return bool(found), tf.reshape(4, 4), {"n_triangle_matches": info.n_triangle_matches, "n_associations": info.n_associations,
                                           "nnz": info.nnz, "n_inliers": info.n_inliers, "score": info.score}
