"""A/B timing of the lattice search across library builds (SLIDE_PR_LIB=<.so> selects one): binds only
the entry points every build has.  usage: ab_search.py [config] [reps] [modes: 4 pair-join, 3 bound-and-verify, 1 lattice exhaustive]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from slide_slam_b200 import capi  # noqa: E402  (struct layouts only)

path = os.environ.get("SLIDE_PR_LIB") or os.path.join(ROOT, "slide_slam_b200", "libslide_pr.so")
L = C.CDLL(path)
_dp = C.POINTER(C.c_double)
L.slide_pr_default_params.argtypes = [C.POINTER(capi.Params)]
L.slide_pr_create.argtypes = [C.POINTER(capi.Params), C.POINTER(C.c_void_p)]
L.slide_pr_prepare.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, C.c_double, C.c_double]
L.slide_pr_search.argtypes = [C.c_void_p, C.POINTER(capi.SearchOpts), C.POINTER(capi.MatchResult)]
L.slide_pr_deg2rad.restype = C.c_double
L.slide_pr_deg2rad.argtypes = [C.c_double]
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
(ref, qry, _), _ = bench.workload(cfg, 0)
r = bench._ranges(ref, qry)
sref, sqry = np.ascontiguousarray(ref.copy()), np.ascontiguousarray(qry.copy())
sref[:, 1:3] -= r["centroid_ref"]; sqry[:, 1:3] -= r["centroid_qry"]
p = capi.Params()
L.slide_pr_default_params(C.byref(p))
p.match_yaw_angle_step_size = L.slide_pr_deg2rad(5.0)
p.min_num_inliers = 15
h = C.c_void_p()
assert L.slide_pr_create(C.byref(p), C.byref(h)) == 0
assert L.slide_pr_prepare(h, sref.ctypes.data_as(_dp), len(sref), sqry.ctypes.data_as(_dp), len(sqry), r["half_x"], r["half_y"]) == 0
import torch  # L2 flush between searches, like bench.py
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
names = {4: "pair-join", 3: "bound-and-verify", 1: "lattice exhaustive"}
modes = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4, 3, 1]
for mode in modes:
    ms = []
    for i in range(reps + 3):
        flush.fill_(1); torch.cuda.synchronize()
        o = capi.SearchOpts(); o.trans_end = -1; o.exhaustive = mode
        res = capi.MatchResult()
        assert L.slide_pr_search(h, C.byref(o), C.byref(res)) == 0
        if i >= 3:
            ms.append(res.kernel_ms)
    if ms:
        print(f"{os.path.basename(path)} cfg{cfg} {names.get(mode, mode)}: kernel_ms min {min(ms):.3f} median {np.median(ms):.3f} "
              f"best={res.best_num_inliers} idx={res.best_hyp_index} hyp={res.hypotheses_scored} launches={res.gpu_launches}", flush=True)
