"""ctypes binding of the C-ABI in include/slide_pr.h (slide_slam_b200/libslide_pr.so).

The library holds the hand-written sm_100a kernels; there is no Python or CPU implementation
of the search behind this module -- if the library is missing or no CUDA device is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SLIDE_PR_LIB: developer override used to A/B-test kernel build variants (tools/build_variants.sh)
LIB_PATH = os.environ.get("SLIDE_PR_LIB") or os.path.join(_HERE, "libslide_pr.so")

OK, NOT_FOUND, SANITY_RETURN = 0, 1, 2
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NONFINITE, ERR_INTERNAL = -1, -2, -3, -4, -5

# every symbol include/slide_pr.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "slide_pr_abi_version", "slide_pr_deg2rad", "slide_pr_default_params", "slide_pr_create",
    "slide_pr_destroy", "slide_pr_set_params", "slide_pr_last_error", "slide_pr_match_maps",
    "slide_pr_prepare", "slide_pr_search", "slide_pr_lattice_info", "slide_pr_extract", "slide_pr_find_transformation",
    "slide_pr_find_inter_loop_closure", "slide_pr_find_intra_loop_closure", "slide_pr_find_intra_loop_closure_batch", "slide_pr_solve_lsq",
    "slide_pr_get_xyz_yaw_from_tf", "slide_pr_find_transformation_batch", "slide_pr_pack_record",
    "slide_pr_merge_records", "slide_pr_match_triangles", "slide_pr_score_hypotheses",
    "slide_pr_match_triangles_labeled", "slide_pr_estimate_tf", "slide_pr_triangle_hypotheses",
    "slide_pr_generate_and_score",
    "slide_pr_map_cache_put", "slide_pr_map_cache_drop", "slide_pr_map_cache_size", "slide_pr_find_transformation_cached",
    "slide_pr_measure_issue_peaks", "slide_pr_delaunay", "slide_pr_slidegraph_default_params", "slide_pr_run_semantic_clipper",
    "slide_pr_find_inter_loop_closure_with_clipper",
    "slide_clipper_default_params", "slide_pr_clipper_score_pairwise_consistency",
    "slide_pr_clipper_get_initial_associations", "slide_pr_clipper_get_affinity_matrix",
    "slide_pr_clipper_get_affinity_csr", "slide_pr_clipper_solve",
]


class Params(C.Structure):
    """slide_pr_params: the rosparams of place_recognition.cpp:24-75 as stored in the members."""

    _fields_ = [
        ("compute_budget_sec", C.c_double),
        ("dilation_factor", C.c_double),
        ("match_xy_step_size", C.c_double),
        ("match_yaw_half_range", C.c_double),
        ("match_yaw_angle_step_size", C.c_double),
        ("match_threshold", C.c_double),
        ("match_threshold_dimension", C.c_double),
        ("match_x_half_range_intra", C.c_double),
        ("match_y_half_range_intra", C.c_double),
        ("match_yaw_half_range_intra", C.c_double),
        ("disable_yaw_search", C.c_int32),
        ("ignore_dimension", C.c_int32),
        ("min_num_inliers", C.c_int32),
        ("use_lsq", C.c_int32),
        ("min_num_map_objects_to_start", C.c_int32),
        ("inter_loop_closure", C.c_int32),
        ("device", C.c_int32),
        ("exhaustive_search", C.c_int32),
    ]


class MatchResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("best_num_inliers", C.c_int32),
        ("R_t", C.c_double * 9),
        ("n_matched", C.c_int32),
        ("n_rings", C.c_int32),
        ("n_yaw", C.c_int32),
        ("rings_scored", C.c_int32),
        ("hypotheses_scored", C.c_int64),
        ("best_hyp_index", C.c_int64),
        ("n_translations", C.c_int64),
        ("kernel_ms", C.c_float),
        ("prepare_ms", C.c_float),
        ("gpu_launches", C.c_int64),
        ("filter_hits", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("groups_probed", C.c_int64),
        ("groups_skipped", C.c_int64),
        ("reuse", C.c_int32),
        ("search_mode", C.c_int32),
    ]


class SearchOpts(C.Structure):
    _fields_ = [
        ("trans_begin", C.c_int64),
        ("trans_end", C.c_int64),
        ("shard_index", C.c_int32),
        ("shard_count", C.c_int32),
        ("counts_out", C.POINTER(C.c_int32)),
        ("counts_cap", C.c_int64),
        ("stream", C.c_void_p),
        ("collect_stats", C.c_int32),
        ("exhaustive", C.c_int32),
        ("incumbent_inliers", C.c_int32),
        ("reuse_bounds", C.c_int32),
    ]


class TfResult(C.Structure):
    _fields_ = [
        ("found", C.c_int32),
        ("best_num_inliers", C.c_int32),
        ("n_matched", C.c_int32),
        ("reserved", C.c_int32),
        ("R_t", C.c_double * 9),
        ("xyz_yaw", C.c_double * 4),
        ("transform", C.c_double * 16),
        ("centroid_ref", C.c_double * 2),
        ("centroid_qry", C.c_double * 2),
        ("half_x", C.c_double),
        ("half_y", C.c_double),
        ("yaw_half", C.c_double),
        ("match", MatchResult),
    ]


class TopkRecord(C.Structure):
    _fields_ = [("hyp_index", C.c_int64), ("inliers", C.c_int32), ("rank", C.c_int32)]


class ClipperParams(C.Structure):
    """slide_clipper_params: clipper::Params (clipper.h:28-60) + EuclideanDistance::Params."""
    _fields_ = [("sigma", C.c_double), ("epsilon", C.c_double), ("mindist", C.c_double),
                ("tol_u", C.c_double), ("tol_F", C.c_double), ("tol_Fop", C.c_double),
                ("maxiniters", C.c_int32), ("maxoliters", C.c_int32), ("beta", C.c_double), ("maxlsiters", C.c_int32),
                ("eps", C.c_double), ("affinityeps", C.c_double), ("rescale_u0", C.c_int32), ("rounding", C.c_int32)]


class ClipperSolution(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("ifinal", C.c_int32), ("score", C.c_double), ("d", C.c_double),
                ("line_search_steps", C.c_int64), ("kernel_ms", C.c_float), ("reserved", C.c_int32)]


class SlidegraphParams(C.Structure):
    """slide_pr_slidegraph_params: rosparams sloam/place_recognition_slidegraph/* (PR.cpp:64-75)."""
    _fields_ = [("sigma", C.c_double), ("epsilon", C.c_double), ("matching_threshold", C.c_double),
                ("num_inliers_threshold", C.c_int32), ("min_num_map_objects_to_start", C.c_int32),
                ("use_class_signature", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64)]


class ScInfo(C.Structure):
    _fields_ = [("found", C.c_int32), ("n_inliers", C.c_int32), ("n_triangles_model", C.c_int32), ("n_triangles_data", C.c_int32),
                ("n_triangle_matches", C.c_int64), ("n_associations", C.c_int64), ("nnz_upper", C.c_int64), ("score", C.c_double),
                ("delaunay_ms", C.c_float), ("match_ms", C.c_float), ("affinity_ms", C.c_float), ("solve_ms", C.c_float)]


class GenerateInfo(C.Structure):
    _fields_ = [("n_matches", C.c_int64), ("n_triangles_model", C.c_int32), ("n_triangles_data", C.c_int32),
                ("match_ms", C.c_float), ("kabsch_ms", C.c_float), ("score_ms", C.c_float), ("reserved", C.c_int32)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_lib = None


def build(force: bool = False) -> str:
    """Compile the sm_100a library in-tree (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    cmd = ["make", "-C", csrc] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no Python/CPU fallback for the place-recognition search)")
    L = C.CDLL(LIB_PATH)
    L.slide_pr_abi_version.restype = C.c_int
    L.slide_pr_deg2rad.restype = C.c_double
    L.slide_pr_deg2rad.argtypes = [C.c_double]
    L.slide_pr_default_params.argtypes = [C.POINTER(Params)]
    L.slide_pr_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    L.slide_pr_destroy.argtypes = [C.c_void_p]
    L.slide_pr_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.slide_pr_last_error.restype = C.c_char_p
    L.slide_pr_last_error.argtypes = [C.c_void_p]
    L.slide_pr_match_maps.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, C.c_double, C.c_double,
                                      _ip, _ip, C.POINTER(MatchResult)]
    L.slide_pr_prepare.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, C.c_double, C.c_double]
    L.slide_pr_search.argtypes = [C.c_void_p, C.POINTER(SearchOpts), C.POINTER(MatchResult)]
    L.slide_pr_lattice_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64), _ip, _ip]
    L.slide_pr_extract.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, C.POINTER(MatchResult)]
    L.slide_pr_find_transformation.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, _ip, _ip,
                                               C.POINTER(TfResult)]
    L.slide_pr_find_inter_loop_closure.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, _dp,
                                                   C.POINTER(TfResult)]
    L.slide_pr_find_intra_loop_closure.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, _dp, _dp, _dp,
                                                   C.POINTER(TfResult)]
    L.slide_pr_find_intra_loop_closure_batch.argtypes = [C.c_void_p, _dp, C.c_int32, C.POINTER(_dp), _ip, _dp, _dp, C.c_int32, _dp,
                                                         C.POINTER(TfResult)]
    L.slide_pr_solve_lsq.argtypes = [_dp, _dp, C.c_int32, _dp, _dp]
    L.slide_pr_get_xyz_yaw_from_tf.argtypes = [_dp, _dp]
    L.slide_pr_find_transformation_batch.argtypes = [C.c_void_p, C.POINTER(_dp), _ip, C.c_int32, _ip, _ip,
                                                     C.c_int32, C.POINTER(TfResult)]
    L.slide_pr_pack_record.argtypes = [C.POINTER(MatchResult), C.c_int32, C.POINTER(TopkRecord)]
    L.slide_pr_merge_records.argtypes = [C.POINTER(TopkRecord), C.c_int32]
    L.slide_pr_match_triangles.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, C.c_double, _ip, _ip,
                                           _ip, _ip, C.c_int64, C.POINTER(C.c_int64)]
    L.slide_pr_match_triangles_labeled.argtypes = [C.c_void_p, _dp, _dp, C.c_int32, _dp, _dp, C.c_int32, C.c_double, _ip,
                                                   _ip, _ip, _ip, C.c_int64, C.POINTER(C.c_int64)]
    L.slide_pr_estimate_tf.argtypes = [_dp, _dp, C.c_int32, _dp]
    L.slide_pr_triangle_hypotheses.argtypes = [_dp, _dp, _ip, _ip, _ip, _ip, C.c_int64, _dp]
    L.slide_pr_score_hypotheses.argtypes = [C.c_void_p, _dp, C.c_int64, _ip, C.POINTER(MatchResult)]
    L.slide_pr_generate_and_score.argtypes = [C.c_void_p, _dp, _dp, C.c_int32, _dp, _dp, C.c_int32, C.c_double, C.POINTER(MatchResult),
                                              C.POINTER(GenerateInfo), _ip, _ip, _dp, _ip, C.c_int64]
    L.slide_pr_map_cache_put.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, _dp, C.c_int32]
    L.slide_pr_map_cache_drop.argtypes = [C.c_void_p, C.c_int64]
    L.slide_pr_map_cache_size.restype = C.c_int32
    L.slide_pr_map_cache_size.argtypes = [C.c_void_p]
    L.slide_pr_find_transformation_cached.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _ip, _ip, C.POINTER(TfResult)]
    L.slide_pr_measure_issue_peaks.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.slide_pr_delaunay.argtypes = [_dp, C.c_int32, _ip, C.c_int64, _lp]
    L.slide_pr_slidegraph_default_params.argtypes = [C.POINTER(SlidegraphParams)]
    L.slide_pr_run_semantic_clipper.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, C.POINTER(SlidegraphParams), _dp, C.c_int32,
                                                _dp, C.c_int32, _dp, C.c_int64, _dp, C.POINTER(ScInfo)]
    L.slide_pr_find_inter_loop_closure_with_clipper.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, C.POINTER(SlidegraphParams),
                                                                _dp, C.POINTER(ScInfo)]
    L.slide_clipper_default_params.argtypes = [C.POINTER(ClipperParams)]
    L.slide_pr_clipper_score_pairwise_consistency.argtypes = [C.c_void_p, C.POINTER(ClipperParams), _dp, C.c_int32, _dp, C.c_int32,
                                                              C.c_int32, _ip, C.c_int32, _lp]
    L.slide_pr_clipper_get_initial_associations.restype = C.c_int32
    L.slide_pr_clipper_get_initial_associations.argtypes = [C.c_void_p, _ip, C.c_int32]
    L.slide_pr_clipper_get_affinity_matrix.argtypes = [C.c_void_p, _dp, C.c_int64]
    L.slide_pr_clipper_get_affinity_csr.argtypes = [C.c_void_p, _lp, _ip, _dp, C.c_int64]
    L.slide_pr_clipper_solve.argtypes = [C.c_void_p, C.POINTER(ClipperParams), _dp, C.c_uint64, _ip, C.c_int32,
                                         C.POINTER(ClipperSolution), _dp]
    _lib = L
    return L


def default_params() -> Params:
    p = Params()
    lib().slide_pr_default_params(C.byref(p))
    return p


def as_rows7(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.size == 0:
        return a.reshape(0, 7)
    if a.ndim != 2 or a.shape[1] != 7:
        raise ValueError("object maps are n x 7 rows [label, x, y, z, d1, d2, d3]")
    return a


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(_ip)


class SlidePrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"slide_pr error {code}: {msg}")
        self.code = code
