"""Host-side mirror of the reference's `clipper::CLIPPER` class (clipper_semantic_object/include/
clipper/clipper.h:78-182, its pybind module bindings/python/py_clipper.cpp exposes the same
methods) over the C-ABI in include/slide_pr.h.  The affinity matrix and the dense-clique solver live
on the GPU (slide_slam_b200/csrc/spr_clipper.cu); nothing here computes on the CPU.

Datasets are `dim x n` arrays with the points as columns, exactly like the reference's
Eigen `invariants::Data`; associations are `m x 2` integer arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

ROUND_NONZERO, ROUND_DSD, ROUND_DSD_HEU = 0, 1, 2


def default_params(**kw) -> capi.ClipperParams:
    """clipper::Params + invariants::EuclideanDistance::Params defaults, overridden by keywords."""
    p = capi.ClipperParams()
    capi.lib().slide_clipper_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class CLIPPER:
    """`clipper::CLIPPER(invariant, params)` with the EuclideanDistance invariant."""

    def __init__(self, params: capi.ClipperParams | None = None, device: int = -1, handle=None):
        self._lib = capi.lib()
        self.params = params or default_params()
        self._own = handle is None
        if handle is None:
            pp = capi.default_params()
            pp.device = device
            h = C.c_void_p()
            rc = self._lib.slide_pr_create(C.byref(pp), C.byref(h))
            if rc != capi.OK:
                raise capi.SlidePrError(rc, self._lib.slide_pr_last_error(None).decode())
            handle = h
        self._h = handle
        self.solution = None
        self.nnz_upper = 0

    def close(self):
        if self._own and self._h:
            self._lib.slide_pr_destroy(self._h)
        self._h = None

    def _check(self, rc):
        if rc < 0:
            raise capi.SlidePrError(rc, self._lib.slide_pr_last_error(self._h).decode())

    def score_pairwise_consistency(self, D1, D2, A=None):
        """CLIPPER::scorePairwiseConsistency (clipper.cpp:21-65); A empty -> all-to-all."""
        D1 = np.asarray(D1, np.float64); D2 = np.asarray(D2, np.float64)
        if D1.ndim != 2 or D2.ndim != 2 or D1.shape[0] != D2.shape[0]:
            raise ValueError("datasets are dim x n arrays (points as columns)")
        f1, f2 = np.ascontiguousarray(D1.T), np.ascontiguousarray(D2.T)   # == column-major dim x n
        nnz = C.c_int64(0)
        if A is None or len(A) == 0:
            ap, m = None, 0
        else:
            A = np.ascontiguousarray(A, np.int32)
            ap, m = capi.iptr(A), len(A)
        self._check(self._lib.slide_pr_clipper_score_pairwise_consistency(
            self._h, C.byref(self.params), capi.dptr(f1), D1.shape[1], capi.dptr(f2), D2.shape[1], D1.shape[0], ap, m, C.byref(nnz)))
        self.nnz_upper = nnz.value
        return nnz.value

    def get_initial_associations(self) -> np.ndarray:
        m = self._lib.slide_pr_clipper_get_initial_associations(self._h, None, 0)
        A = np.zeros((max(m, 1), 2), np.int32)
        self._lib.slide_pr_clipper_get_initial_associations(self._h, capi.iptr(A), m)
        return A[:m]

    def get_affinity_matrix(self) -> np.ndarray:
        """CLIPPER::getAffinityMatrix: dense, symmetric, ones on the diagonal (small problems only)."""
        m = self._lib.slide_pr_clipper_get_initial_associations(self._h, None, 0)
        M = np.zeros((m, m), np.float64)
        self._check(self._lib.slide_pr_clipper_get_affinity_matrix(self._h, capi.dptr(M), m * m))
        return M

    def get_constraint_matrix(self) -> np.ndarray:
        """CLIPPER::getConstraintMatrix: the pattern of the affinity matrix (clipper.cpp:62-64, 130-135)."""
        return (self.get_affinity_matrix() != 0).astype(np.float64)

    def get_affinity_csr(self):
        m = self._lib.slide_pr_clipper_get_initial_associations(self._h, None, 0)
        rp = np.zeros(m + 1, np.int64)
        self._check(self._lib.slide_pr_clipper_get_affinity_csr(self._h, rp.ctypes.data_as(C.POINTER(C.c_int64)), None, None, 0))
        nnz = int(rp[m])
        col, val = np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1), np.float64)
        self._check(self._lib.slide_pr_clipper_get_affinity_csr(self._h, rp.ctypes.data_as(C.POINTER(C.c_int64)), capi.iptr(col),
                                                                capi.dptr(val), nnz))
        return rp, col[:nnz], val[:nnz]

    def solve(self, u0=None, seed: int = 0):
        """CLIPPER::solve (clipper.cpp:69-78): u0 None -> a deterministic U[0,1) vector from `seed`
        (the reference draws it from std::random_device)."""
        m = self._lib.slide_pr_clipper_get_initial_associations(self._h, None, 0)
        nodes = np.zeros(max(m, 1), np.int32)
        u = np.zeros(max(m, 1), np.float64)
        sol = capi.ClipperSolution()
        up = None
        if u0 is not None:
            u0 = np.ascontiguousarray(u0, np.float64)
            if len(u0) != m:
                raise ValueError("u0 must have one entry per association")
            up = capi.dptr(u0)
        self._check(self._lib.slide_pr_clipper_solve(self._h, C.byref(self.params), up, seed, capi.iptr(nodes), m, C.byref(sol), capi.dptr(u)))
        self.solution = {"nodes": nodes[:sol.n_nodes].copy(), "u": u[:m].copy(), "score": sol.score, "ifinal": sol.ifinal,
                         "d": sol.d, "line_search_steps": sol.line_search_steps, "kernel_ms": sol.kernel_ms}
        return self.solution

    def get_selected_associations(self) -> np.ndarray:
        """CLIPPER::getSelectedAssociations (clipper.cpp:114-117, utils.cpp:96-104)."""
        return self.get_initial_associations()[self.solution["nodes"]]
