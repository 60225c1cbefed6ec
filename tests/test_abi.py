"""CPU suite: the C-ABI shared library loads, exports every symbol include/slide_pr.h declares,
agrees with the header on struct layout, and FAILS LOUDLY without a CUDA device (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from slide_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "slide_pr.h")).read()
    declared = sorted(set(re.findall(r"\b(slide_(?:pr|clipper)_[a-z0-9_]+)\s*\(", header)))
    assert sorted(capi.EXPORTS) == declared
    lib = capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.slide_pr_abi_version() == 2


def test_struct_layout_matches_header():
    # sizes implied by include/slide_pr.h on LP64
    assert C.sizeof(capi.Params) == 10 * 8 + 8 * 4
    assert C.sizeof(capi.MatchResult) == 8 + 72 + 16 + 24 + 8 + 16 + 16 + 16 + 8
    assert C.sizeof(capi.SearchOpts) == 16 + 8 + 8 + 8 + 8 + 8 + 8
    assert C.sizeof(capi.TopkRecord) == 16
    assert C.sizeof(capi.TfResult) == 16 + 72 + 32 + 128 + 32 + 24 + C.sizeof(capi.MatchResult)
    assert C.sizeof(capi.ClipperParams) == 6 * 8 + 8 + 8 + 8 + 16 + 8      # ints padded to the doubles' alignment
    assert C.sizeof(capi.ClipperSolution) == 8 + 8 + 8 + 8 + 8
    p = capi.ClipperParams()
    capi.lib().slide_clipper_default_params(C.byref(p))                      # clipper.h:28-60, euclidean_distance.h:27-30
    assert (p.sigma, p.epsilon, p.mindist, p.tol_u, p.tol_F, p.maxiniters, p.maxoliters, p.beta, p.maxlsiters, p.eps,
            p.affinityeps, p.rescale_u0, p.rounding) == (0.01, 0.06, 0.0, 1e-8, 1e-9, 200, 1000, 0.25, 99, 1e-9, 1e-4, 1, 2)


def test_default_params_are_the_reference_defaults():
    p = capi.default_params()  # PR.cpp:24-75
    assert p.match_xy_step_size == 0.5 and p.match_threshold == 0.5 and p.match_threshold_dimension == 1.0
    assert p.dilation_factor == 1.2 and p.min_num_inliers == 5 and p.use_lsq == 1 and p.inter_loop_closure == 1
    assert p.match_yaw_half_range == 180.0 * np.pi / 180.0
    assert p.match_yaw_angle_step_size == capi.lib().slide_pr_deg2rad(2.0)
    assert p.min_num_map_objects_to_start == 1 and p.ignore_dimension == 0 and p.disable_yaw_search == 0


def test_host_only_entry_points():
    lib = capi.lib()
    # solveLSQ / getxyzYawfromTF are host code and work without a device
    rng = np.random.default_rng(0)
    src = rng.uniform(-5, 5, (10, 3))
    yaw, t = -0.4, np.array([1.0, 2.0, 0.3])
    R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    tgt = src @ R.T + t
    xyzyaw, tf = np.zeros(4), np.zeros(16)
    assert lib.slide_pr_solve_lsq(capi.dptr(np.ascontiguousarray(tgt)), capi.dptr(np.ascontiguousarray(src)), 10,
                                  capi.dptr(xyzyaw), capi.dptr(tf)) == 0
    np.testing.assert_allclose(xyzyaw, [1.0, 2.0, 0.3, yaw], atol=1e-9)
    # deterministic top-k merge: max inliers, ties to the smallest canonical index (PR.cpp:361)
    recs = (capi.TopkRecord * 4)()
    for i, (h, n) in enumerate([(50, 7), (20, 9), (10, 9), (-1, -10000)]):
        recs[i].hyp_index, recs[i].inliers, recs[i].rank = h, n, i
    assert lib.slide_pr_merge_records(recs, 4) == 2


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = capi.lib()
    p = capi.default_params()
    h = C.c_void_p()
    rc = lib.slide_pr_create(C.byref(p), C.byref(h))
    assert rc == capi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.slide_pr_last_error(None)
    from slide_slam_b200.place_recognition import PlaceRecognition
    with pytest.raises(capi.SlidePrError):
        PlaceRecognition()
