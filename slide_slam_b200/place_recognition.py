"""Host-side mirror of the reference's `PlaceRecognition` class for the SlideMatch path
(backend/sloam/include/core/place_recognition.h:31-237), over the C-ABI in include/slide_pr.h.

Same member names, argument meaning and failure behaviour as the reference: maps are
`n x 7` arrays of `[label, x, y, z, d1, d2, d3]` rows (std::vector<Eigen::Vector7d>), results
are returned instead of written through out-references, "not found" is `False`, not an
exception.  All scoring happens in the CUDA library; nothing here computes a match on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi

# rosparam name under sloam/place_recognition/  ->  (Params field, converter)   PR.cpp:24-75
_ROSPARAMS = {
    "compute_budget_sec": ("compute_budget_sec", float),
    "dilation_factor": ("dilation_factor", float),
    "search_xy_step_size": ("match_xy_step_size", float),
    "match_yaw_half_range": ("match_yaw_half_range", "deg"),
    "disable_yaw_search": ("disable_yaw_search", int),
    "search_yaw_step_size_degrees": ("match_yaw_angle_step_size", "deg"),
    "match_threshold_position": ("match_threshold", float),
    "match_threshold_dimension": ("match_threshold_dimension", float),
    "ignore_dimension": ("ignore_dimension", int),
    "min_num_inliers": ("min_num_inliers", int),
    "use_nonlinear_least_squares": ("use_lsq", int),
    "min_num_map_objects_to_start": ("min_num_map_objects_to_start", int),
    "match_x_half_range_intra": ("match_x_half_range_intra", float),
    "match_y_half_range_intra": ("match_y_half_range_intra", float),
    "match_yaw_half_range_intra": ("match_yaw_half_range_intra", "deg"),
    # not a rosparam of the reference: which kernels search the lattice.  0 (default) = the pair-join scorer (exact
    # count of every hypothesis); 1 = the lattice kernels, every hypothesis verified; 2 = the lattice kernels,
    # bound-and-verify.  Same winner, count and correspondences in every case.
    "exhaustive_search": ("exhaustive_search", int),
}


# rosparam name under sloam/place_recognition_slidegraph/  ->  SlidegraphParams field   PR.cpp:64-75
_SLIDEGRAPH = {"num_inliners_threshold": "num_inliers_threshold", "descriptor_matching_threshold": "matching_threshold",
               "sigma": "sigma", "epsilon": "epsilon", "min_num_map_objects_to_start": "min_num_map_objects_to_start",
               "use_class_signature": "use_class_signature", "seed": "seed"}


def delaunay(xy) -> np.ndarray:
    """Observation::delaunayTriangulation (observation.cpp:13-88) on the host: t x 3 vertex ids."""
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    n = C.c_int64(0)
    lib = capi.lib()
    rc = lib.slide_pr_delaunay(capi.dptr(xy), len(xy), None, 0, C.byref(n))
    if rc < 0:
        raise capi.SlidePrError(rc, "non-finite coordinate")
    tri = np.zeros((max(n.value, 1), 3), np.int32)
    lib.slide_pr_delaunay(capi.dptr(xy), len(xy), capi.iptr(tri), n.value, C.byref(n))
    return tri[:n.value]


@dataclass
class MatchMapsResult:
    """Outputs of MatchMaps (PR.h:70-74)."""
    status: int
    R_t: np.ndarray                 # 3x3
    best_num_inliers: int
    map_objects_matched: np.ndarray        # k x 4 [label, x, y, z] rows of the reference map
    detection_objects_matched: np.ndarray  # k x 4 rows of the (un-transformed) query map
    ref_idx: np.ndarray
    qry_idx: np.ndarray
    info: capi.MatchResult


class PlaceRecognition:
    """Drop-in for the reference class on the SlideMatch path.  `params` uses the rosparam names
    of place_recognition.cpp:24-75 (angles in degrees, like the yaml files)."""

    ENGINES = {"join": 0, "lattice_exhaustive": 1, "lattice": 2}

    def __init__(self, params: dict | None = None, device: int = -1, slidegraph: dict | None = None, engine: str | None = None):
        """engine: "join" (default: the pair-join scorer), "lattice" (bound-and-verify lattice kernels) or
        "lattice_exhaustive"; sets the exhaustive_search parameter.  The environment variable
        SLIDE_PR_ENGINE=lattice makes the lattice kernels the default of every handle (A/B runs)."""
        self._lib = capi.lib()
        self._p = capi.default_params()
        self._p.device = device
        self._sg = capi.SlidegraphParams()
        self._lib.slide_pr_slidegraph_default_params(C.byref(self._sg))
        for k, v in (slidegraph or {}).items():
            if k not in _SLIDEGRAPH:
                raise KeyError(f"unknown place_recognition_slidegraph rosparam {k!r}")
            setattr(self._sg, _SLIDEGRAPH[k], v)
        # public members of the reference class (PR.h:34-43)
        self.visualize_matching_results = False
        self.min_loop_closure_overlap_percentage_ = 0.1
        for k, v in (params or {}).items():
            if k in ("visualize_matching_results", "min_loop_closure_overlap_percentage"):
                continue
            if k not in _ROSPARAMS:
                raise KeyError(f"unknown place_recognition rosparam {k!r}")
            field, conv = _ROSPARAMS[k]
            setattr(self._p, field, self._lib.slide_pr_deg2rad(float(v)) if conv == "deg" else conv(v))
        if engine is not None:
            self._p.exhaustive_search = self.ENGINES[engine]
        import os
        lattice = self._p.exhaustive_search != 0 or os.environ.get("SLIDE_PR_ENGINE") == "lattice" or os.environ.get("SLIDE_PR_EXHAUSTIVE", "0") not in ("", "0")
        self.engine = "lattice" if lattice else "join"   # what an unqualified search of this handle runs
        h = C.c_void_p()
        rc = self._lib.slide_pr_create(C.byref(self._p), C.byref(h))
        if rc != capi.OK:
            raise capi.SlidePrError(rc, self._lib.slide_pr_last_error(None).decode())
        self._h = h

    # -- public mutable members, forwarded to the handle ------------------------------------
    @property
    def use_lsq(self) -> bool:
        return bool(self._p.use_lsq)

    @use_lsq.setter
    def use_lsq(self, v: bool):
        self._p.use_lsq = int(bool(v))
        self._lib.slide_pr_set_params(self._h, C.byref(self._p))

    @property
    def inter_loop_closure(self) -> bool:
        return bool(self._p.inter_loop_closure)

    @inter_loop_closure.setter
    def inter_loop_closure(self, v: bool):
        self._p.inter_loop_closure = int(bool(v))
        self._lib.slide_pr_set_params(self._h, C.byref(self._p))

    @property
    def params(self) -> capi.Params:
        return self._p

    def close(self):
        if getattr(self, "_h", None):
            self._lib.slide_pr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            raise capi.SlidePrError(rc, self._lib.slide_pr_last_error(self._h).decode())
        return rc

    # -- PlaceRecognition::MatchMaps (PR.cpp:98-387) ----------------------------------------
    def MatchMaps(self, reference_objects, query_objects, match_x_half_range: float,
                  match_y_half_range: float) -> MatchMapsResult:
        """The half ranges are the members findTransformation sets before calling MatchMaps
        (PR.cpp:786-787 / 808-809)."""
        ref, qry = capi.as_rows7(reference_objects), capi.as_rows7(query_objects)
        ri = np.zeros(max(len(qry), 1), np.int32)
        qi = np.zeros(max(len(qry), 1), np.int32)
        res = capi.MatchResult()
        self._check(self._lib.slide_pr_match_maps(self._h, capi.dptr(ref), len(ref), capi.dptr(qry), len(qry),
                                                  float(match_x_half_range), float(match_y_half_range),
                                                  capi.iptr(ri), capi.iptr(qi), C.byref(res)))
        k = max(res.n_matched, 0)
        ri, qi = ri[:k].copy(), qi[:k].copy()
        return MatchMapsResult(
            status=res.status, R_t=np.array(res.R_t[:]).reshape(3, 3), best_num_inliers=res.best_num_inliers,
            map_objects_matched=ref[ri][:, :4].copy() if k else np.zeros((0, 4)),
            detection_objects_matched=qry[qi][:, :4].copy() if k else np.zeros((0, 4)),
            ref_idx=ri, qry_idx=qi, info=res)

    # -- staged form (sharded search) ---------------------------------------------------------
    def prepare(self, reference_objects, query_objects, half_x: float, half_y: float):
        ref, qry = capi.as_rows7(reference_objects), capi.as_rows7(query_objects)
        self._check(self._lib.slide_pr_prepare(self._h, capi.dptr(ref), len(ref), capi.dptr(qry), len(qry),
                                               float(half_x), float(half_y)))
        self._n_qry = len(qry)

    def lattice_info(self):
        """(n_translations, n_yaw, n_rings) of the prepared lattice (PR.cpp:136-241)."""
        nt, ny, nr = C.c_int64(0), C.c_int32(0), C.c_int32(0)
        self._check(self._lib.slide_pr_lattice_info(self._h, C.byref(nt), C.byref(ny), C.byref(nr)))
        return nt.value, ny.value, nr.value

    def search(self, trans_begin: int = 0, trans_end: int = -1, shard_index: int = 0, shard_count: int = 1,
               want_counts: bool = False, stream: int | None = None, collect_stats: bool = False,
               exhaustive: bool = False, bounds_only: bool = False, incumbent_inliers: int = 0,
               reuse_bounds: bool = False, engine: str | None = None):
        """Default: the handle's engine (the pair-join scorer unless it was created otherwise).
        exhaustive=True: the lattice kernels with every hypothesis verified; bounds_only=True: their bound phase
        only (counts = upper bounds); engine="lattice": their bound-and-verify search; engine="join": the pair-join
        scorer.  Same winner, count and correspondences in every mode."""
        o = capi.SearchOpts()
        o.trans_begin, o.trans_end, o.shard_index, o.shard_count = trans_begin, trans_end, shard_index, shard_count
        o.stream = stream
        o.collect_stats = int(collect_stats)
        o.exhaustive = 2 if bounds_only else int(exhaustive)  # 2: bound phase only (counts = upper bounds)
        if engine is not None and not bounds_only and not exhaustive:
            o.exhaustive = {"lattice": 3, "join": 4}[engine]
        o.incumbent_inliers = int(incumbent_inliers)
        o.reuse_bounds = int(reuse_bounds)
        counts = None
        res = capi.MatchResult()
        if want_counts:
            if trans_end < 0:
                raise ValueError("want_counts needs an explicit trans_end")
            n_trans, n_yaw = self.lattice_info()[:2]
            n_t = max(min(trans_end, n_trans) - trans_begin, 0)
            counts = np.full(max(n_t * max(n_yaw, 1), 1), -1, np.int32)
            o.counts_out = capi.iptr(counts)
            o.counts_cap = counts.size
        self._check(self._lib.slide_pr_search(self._h, C.byref(o), C.byref(res)))
        if counts is not None:
            counts = counts[: max(min(trans_end, res.n_translations) - trans_begin, 0) * res.n_yaw]
        return res, counts

    def extract(self, hyp_index: int, res: capi.MatchResult | None = None):
        res = res or capi.MatchResult()
        ri = np.zeros(max(self._n_qry, 1), np.int32)
        qi = np.zeros(max(self._n_qry, 1), np.int32)
        self._check(self._lib.slide_pr_extract(self._h, int(hyp_index), capi.iptr(ri), capi.iptr(qi), C.byref(res)))
        k = max(res.n_matched, 0)
        return res, ri[:k].copy(), qi[:k].copy()

    def score_hypotheses(self, hyps4, want_counts: bool = True):
        hyps = np.ascontiguousarray(hyps4, np.float64).reshape(-1, 4)
        counts = np.zeros(max(len(hyps), 1), np.int32) if want_counts else None
        res = capi.MatchResult()
        self._check(self._lib.slide_pr_score_hypotheses(self._h, capi.dptr(hyps), len(hyps),
                                                        capi.iptr(counts) if want_counts else None, C.byref(res)))
        return res, (counts[: len(hyps)] if want_counts else None)

    def match_triangles(self, tris_model6, tris_data6, threshold: float, labels_model3=None, labels_data3=None):
        """semantic_clipper::match_triangles (SC.cpp:111-118) on the device; returns (model_idx, data_idx,
        perm_model k x 3, perm_data k x 3) in the reference's order."""
        tm = np.ascontiguousarray(tris_model6, np.float64).reshape(-1, 6)
        td = np.ascontiguousarray(tris_data6, np.float64).reshape(-1, 6)
        lm = None if labels_model3 is None else np.ascontiguousarray(labels_model3, np.float64).reshape(-1, 3)
        ld = None if labels_data3 is None else np.ascontiguousarray(labels_data3, np.float64).reshape(-1, 3)
        n = C.c_int64(0)
        args = (self._h, capi.dptr(tm), None if lm is None else capi.dptr(lm), len(tm), capi.dptr(td),
                None if ld is None else capi.dptr(ld), len(td), float(threshold))
        self._check(self._lib.slide_pr_match_triangles_labeled(*args, None, None, None, None, 0, C.byref(n)))
        k = n.value
        mi, di = np.zeros(max(k, 1), np.int32), np.zeros(max(k, 1), np.int32)
        pm, pd = np.zeros((max(k, 1), 3), np.int32), np.zeros((max(k, 1), 3), np.int32)
        self._check(self._lib.slide_pr_match_triangles_labeled(*args, capi.iptr(mi), capi.iptr(di), capi.iptr(pm), capi.iptr(pd), k, C.byref(n)))
        return mi[:k], di[:k], pm[:k], pd[:k]

    def generate_and_score(self, tris_model6, tris_data6, threshold: float, labels_model3=None, labels_data3=None,
                           want_lists: bool = True, cap: int | None = None):
        """The generator half on the device (slide_pr_generate_and_score): matched triangles -> 2-D Kabsch
        hypotheses -> MatchMaps predicate -> best hypothesis, for the map pair given to prepare().
        Returns (MatchResult, GenerateInfo, lists) with lists = dict(model_idx, data_idx, hyps, counts) or None."""
        tm = np.ascontiguousarray(tris_model6, np.float64).reshape(-1, 6)
        td = np.ascontiguousarray(tris_data6, np.float64).reshape(-1, 6)
        lm = None if labels_model3 is None else np.ascontiguousarray(labels_model3, np.float64).reshape(-1, 3)
        ld = None if labels_data3 is None else np.ascontiguousarray(labels_data3, np.float64).reshape(-1, 3)
        res, info = capi.MatchResult(), capi.GenerateInfo()
        args = (self._h, capi.dptr(tm), None if lm is None else capi.dptr(lm), len(tm), capi.dptr(td),
                None if ld is None else capi.dptr(ld), len(td), float(threshold), C.byref(res), C.byref(info))
        if not want_lists:
            self._check(self._lib.slide_pr_generate_and_score(*args, None, None, None, None, 0))
            return res, info, None
        if cap is None:  # size the outputs with a first call
            self._check(self._lib.slide_pr_generate_and_score(*args, None, None, None, None, 0))
            cap = int(info.n_matches)
        mi, di = np.zeros(max(cap, 1), np.int32), np.zeros(max(cap, 1), np.int32)
        hyps, counts = np.zeros((max(cap, 1), 4)), np.zeros(max(cap, 1), np.int32)
        self._check(self._lib.slide_pr_generate_and_score(*args, capi.iptr(mi), capi.iptr(di), capi.dptr(hyps), capi.iptr(counts), cap))
        k = min(int(info.n_matches), cap)
        return res, info, {"model_idx": mi[:k], "data_idx": di[:k], "hyps": hyps[:k], "counts": counts[:k]}

    # -- PlaceRecognition::findTransformation (PR.cpp:736-945) ---------------------------------
    def findTransformation(self, reference_objects, query_objects):
        """Returns (found, xyzYaw[4], transform_out 4x4, TfResult, ref_idx, qry_idx)."""
        ref, qry = capi.as_rows7(reference_objects), capi.as_rows7(query_objects)
        ri = np.zeros(max(len(qry), 1), np.int32)
        qi = np.zeros(max(len(qry), 1), np.int32)
        out = capi.TfResult()
        rc = self._check(self._lib.slide_pr_find_transformation(self._h, capi.dptr(ref), len(ref), capi.dptr(qry),
                                                                len(qry), capi.iptr(ri), capi.iptr(qi), C.byref(out)))
        k = max(out.n_matched, 0)
        return (rc == capi.OK, np.array(out.xyz_yaw[:]), np.array(out.transform[:]).reshape(4, 4), out,
                ri[:k].copy(), qi[:k].copy())

    # -- PlaceRecognition::findInterLoopClosure (PR.cpp:498-538) --------------------------------
    def findInterLoopClosure(self, reference_objects, query_objects):
        """Returns (closure_found, tfFromQueryToRef 4x4)."""
        ref, qry = capi.as_rows7(reference_objects), capi.as_rows7(query_objects)
        tf = np.zeros(16)
        out = capi.TfResult()
        rc = self._check(self._lib.slide_pr_find_inter_loop_closure(self._h, capi.dptr(ref), len(ref), capi.dptr(qry),
                                                                    len(qry), capi.dptr(tf), C.byref(out)))
        self.last = out
        return rc == capi.OK, tf.reshape(4, 4)

    # -- device-resident map cache (databaseManager::robotMapDict_, databaseManager.h:99-102) -------
    def cache_put(self, robot_id: int, version: int, objects):
        rows = capi.as_rows7(objects)
        self._check(self._lib.slide_pr_map_cache_put(self._h, int(robot_id), int(version), capi.dptr(rows), len(rows)))
        self._cache_sizes = getattr(self, "_cache_sizes", {})
        self._cache_sizes[int(robot_id)] = len(rows)

    def cache_drop(self, robot_id: int) -> bool:
        return self._lib.slide_pr_map_cache_drop(self._h, int(robot_id)) == capi.OK

    def cache_size(self) -> int:
        return int(self._lib.slide_pr_map_cache_size(self._h))

    def findTransformationCached(self, ref_robot_id: int, qry_robot_id: int):
        """findTransformation on two cached maps; returns what findTransformation returns."""
        nq = max(getattr(self, "_cache_sizes", {}).get(int(qry_robot_id), 1), 1)
        ri, qi = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
        out = capi.TfResult()
        rc = self._check(self._lib.slide_pr_find_transformation_cached(self._h, int(ref_robot_id), int(qry_robot_id), capi.iptr(ri),
                                                                       capi.iptr(qi), C.byref(out)))
        k = max(out.n_matched, 0)
        return (rc == capi.OK, np.array(out.xyz_yaw[:]), np.array(out.transform[:]).reshape(4, 4), out, ri[:k].copy(), qi[:k].copy())

    def findTransformationBatch(self, maps, pairs):
        """slide_pr_find_transformation_batch: maps = list of n x 7 arrays, pairs = [(ref_index, qry_index), ...]."""
        rows = [capi.as_rows7(m) for m in maps]
        ptrs = (C.POINTER(C.c_double) * len(rows))(*[capi.dptr(r) for r in rows])
        sizes = np.array([len(r) for r in rows], np.int32)
        ro = np.array([p[0] for p in pairs], np.int32); qo = np.array([p[1] for p in pairs], np.int32)
        out = (capi.TfResult * max(len(pairs), 1))()
        self._check(self._lib.slide_pr_find_transformation_batch(self._h, ptrs, capi.iptr(sizes), len(rows), capi.iptr(ro), capi.iptr(qo),
                                                                 len(pairs), out))
        return [out[i] for i in range(len(pairs))]

    # -- PlaceRecognition::findInterLoopClosureWithClipper (PR.cpp:541-630) --------------------------
    def findInterLoopClosureWithClipper(self, reference_objects, query_objects):
        """SlideGraph: Delaunay triangles -> descriptor matching -> CLIPPER -> 2-D Kabsch.
        Returns (closure_found, tfFromQueryToRef 4x4); self.last_sc holds the stage statistics."""
        ref, qry = capi.as_rows7(reference_objects), capi.as_rows7(query_objects)
        tf = np.eye(4).reshape(16).copy()
        info = capi.ScInfo()
        rc = self._check(self._lib.slide_pr_find_inter_loop_closure_with_clipper(self._h, capi.dptr(ref), len(ref), capi.dptr(qry),
                                                                                 len(qry), C.byref(self._sg), capi.dptr(tf), C.byref(info)))
        self.last_sc = info
        return rc == capi.OK, tf.reshape(4, 4)

    def run_semantic_clipper(self, reference_map, query_map, sigma, epsilon, min_num_pairs, matching_threshold,
                             tris_model6=None, tris_data6=None, u0=None):
        """semantic_clipper::run_semantic_clipper (semantic_clipper.h:38).  Returns (found, tfFromQuery2Ref 4x4
        as the reference names it: model -> data, info)."""
        ref, qry = capi.as_rows7(reference_map), capi.as_rows7(query_map)
        sp = capi.SlidegraphParams()
        C.memmove(C.byref(sp), C.byref(self._sg), C.sizeof(sp))
        sp.sigma, sp.epsilon, sp.num_inliers_threshold, sp.matching_threshold = sigma, epsilon, int(min_num_pairs), matching_threshold
        tm = None if tris_model6 is None else np.ascontiguousarray(tris_model6, np.float64).reshape(-1, 6)
        td = None if tris_data6 is None else np.ascontiguousarray(tris_data6, np.float64).reshape(-1, 6)
        u = None if u0 is None else np.ascontiguousarray(u0, np.float64)
        tf = np.eye(4).reshape(16).copy()
        info = capi.ScInfo()
        rc = self._check(self._lib.slide_pr_run_semantic_clipper(
            self._h, capi.dptr(ref), len(ref), capi.dptr(qry), len(qry), C.byref(sp),
            None if tm is None else capi.dptr(tm), 0 if tm is None else len(tm), None if td is None else capi.dptr(td),
            0 if td is None else len(td), None if u is None else capi.dptr(u), 0 if u is None else len(u), capi.dptr(tf), C.byref(info)))
        return rc == capi.OK, tf.reshape(4, 4), info

    # -- PlaceRecognition::findIntraLoopClosure (PR.cpp:389-496) --------------------------------
    def findIntraLoopClosure(self, measurements, submap, query_pose, candidate_pose):
        """Poses are 4x4 matrices (SE3::matrix()).  Returns (closure_found, tfFromQuery2Candidate)."""
        meas, sub = capi.as_rows7(measurements), capi.as_rows7(submap)
        qp = np.ascontiguousarray(query_pose, np.float64).reshape(16)
        cp = np.ascontiguousarray(candidate_pose, np.float64).reshape(16)
        tf = np.zeros(16)
        out = capi.TfResult()
        rc = self._check(self._lib.slide_pr_find_intra_loop_closure(self._h, capi.dptr(meas), len(meas), capi.dptr(sub),
                                                                    len(sub), capi.dptr(qp), capi.dptr(cp), capi.dptr(tf),
                                                                    C.byref(out)))
        self.last = out
        return rc == capi.OK, tf.reshape(4, 4)

    def findIntraLoopClosureBatch(self, measurements, submaps, query_pose, candidate_poses):
        """Several candidate key poses for one set of measurements (slide_pr_find_intra_loop_closure_batch): submaps =
        list of n x 7 arrays, candidate_poses = list of 4x4 matrices.  Returns [(closure_found, tfFromQuery2Candidate,
        TfResult), ...] -- per candidate what findIntraLoopClosure returns."""
        meas = capi.as_rows7(measurements)
        rows = [capi.as_rows7(m) for m in submaps]
        k = len(rows)
        ptrs = (C.POINTER(C.c_double) * max(k, 1))(*[capi.dptr(r) for r in rows])
        sizes = np.array([len(r) for r in rows] or [0], np.int32)
        qp = np.ascontiguousarray(query_pose, np.float64).reshape(16)
        cps = np.ascontiguousarray(np.array(candidate_poses, np.float64).reshape(max(k, 1) if k else 0, 16) if k else np.zeros((1, 16)))
        tfs = np.zeros((max(k, 1), 16))
        out = (capi.TfResult * max(k, 1))()
        self._check(self._lib.slide_pr_find_intra_loop_closure_batch(self._h, capi.dptr(meas), len(meas), ptrs, capi.iptr(sizes), capi.dptr(qp),
                                                                     capi.dptr(cps), k, capi.dptr(tfs), out))
        return [(bool(out[i].found), tfs[i].reshape(4, 4).copy(), out[i]) for i in range(k)]

    # -- PlaceRecognition::solveLSQ / getxyzYawfromTF (PR.cpp:632-711) --------------------------
    def solveLSQ(self, map_objects_matched, detection_objects_matched):
        tgt = np.ascontiguousarray(map_objects_matched, np.float64).reshape(-1, 3)
        src = np.ascontiguousarray(detection_objects_matched, np.float64).reshape(-1, 3)
        xyzyaw, tf = np.zeros(4), np.zeros(16)
        self._check(self._lib.slide_pr_solve_lsq(capi.dptr(tgt), capi.dptr(src), len(tgt), capi.dptr(xyzyaw), capi.dptr(tf)))
        return xyzyaw, tf.reshape(4, 4)

    def getxyzYawfromTF(self, tf):
        tf = np.ascontiguousarray(tf, np.float64).reshape(16)
        out = np.zeros(4)
        self._lib.slide_pr_get_xyz_yaw_from_tf(capi.dptr(tf), capi.dptr(out))
        return out
