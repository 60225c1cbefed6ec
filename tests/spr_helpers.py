"""Shared test helpers: golden fixtures, oracle <-> product parameter conversion, the CPU
emulation of the kernel logic (tests/emu), random map generators."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

from oracle import pyoracle as O
from slide_slam_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def golden_maps():
    return dict(np.load(os.path.join(GOLDEN, "maps.npz")))


def golden_cases():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return {c["name"]: c for c in json.load(f)["cases"]}


def golden_counts():
    return dict(np.load(os.path.join(GOLDEN, "counts.npz")))


def oracle_params(**kw):
    return O.make_params(**kw)


def to_capi_params(op: O.Params) -> capi.Params:
    """oracle params -> product params (same member names)."""
    p = capi.Params()
    for name, _ in capi.Params._fields_:
        if hasattr(op, name):
            setattr(p, name, getattr(op, name))
    p.device = -1
    return p


def rosparams_from_golden(params: dict) -> dict:
    """golden.json parameter dicts (oracle keyword names) -> rosparam names of PR.cpp:24-75."""
    m = {"match_xy_step_size": "search_xy_step_size", "yaw_step_deg": "search_yaw_step_size_degrees",
         "match_threshold": "match_threshold_position", "match_threshold_dimension": "match_threshold_dimension",
         "ignore_dimension": "ignore_dimension", "min_num_inliers": "min_num_inliers",
         "min_num_map_objects_to_start": "min_num_map_objects_to_start", "use_lsq": "use_nonlinear_least_squares",
         "disable_yaw_search": "disable_yaw_search", "yaw_half_range_deg": "match_yaw_half_range",
         "dilation_factor": "dilation_factor"}
    return {m[k]: v for k, v in params.items() if k in m}


def shifted_maps(maps, case):
    ref, qry = maps[case["ref"]].copy(), maps[case["qry"]].copy()
    ref[:, 1:3] -= np.array(case["centroid_ref"])
    qry[:, 1:3] -= np.array(case["centroid_qry"])
    return ref, qry


# ------------------------------------------------------------------ CPU emulation of the kernel logic
_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        so = os.path.join(HERE, "emu", "libspr_emu.so")
        srcs = [os.path.join(HERE, "emu", "spr_emu.cpp"), os.path.join(ROOT, "slide_slam_b200", "csrc", "spr_host.cpp")]
        deps = srcs + [os.path.join(ROOT, "slide_slam_b200", "csrc", f) for f in ("spr_core.h", "spr_types.h", "spr_host.h", "spr_join_core.h", "spr_join_types.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
            gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.run([gxx, "-O2", "-mpopcnt", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so] + srcs, check=True)
        L = C.CDLL(so)
        L.spr_emu_match_maps.argtypes = [C.POINTER(capi.Params), _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double,
                                         C.c_longlong, C.c_longlong, _ip, C.c_longlong, _ip, C.POINTER(C.c_longlong),
                                         C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_char_p, C.c_int]
        L.spr_emu_join_match_maps.argtypes = L.spr_emu_match_maps.argtypes
        L.spr_emu_lattice.restype = C.c_longlong
        L.spr_emu_lattice.argtypes = [C.POINTER(capi.Params), C.c_double, C.c_double, _dp, _dp, C.c_longlong, _ip, _dp,
                                      C.c_int, _ip]
        _emu = L
    return _emu


def emu_match_maps(p: capi.Params, ref7, qry7, half_x, half_y, trans_begin=0, trans_end=-1, n_counts=0, engine="lattice"):
    """engine: "lattice" = the chunked bitmap kernels' per-thread code, "join" = the pair-join scorer's."""
    ref7 = np.ascontiguousarray(ref7, np.float64).reshape(-1, 7)
    qry7 = np.ascontiguousarray(qry7, np.float64).reshape(-1, 7)
    counts = np.full(max(n_counts, 1), -1, np.int32)
    bc, bi, hs, fh = C.c_int(), C.c_longlong(), C.c_longlong(), C.c_longlong()
    eb = C.create_string_buffer(256)
    fn = emu_lib().spr_emu_join_match_maps if engine == "join" else emu_lib().spr_emu_match_maps
    rc = fn(C.byref(p), ref7.ctypes.data_as(_dp), len(ref7), qry7.ctypes.data_as(_dp), len(qry7),
                                      half_x, half_y, trans_begin, trans_end,
                                      counts.ctypes.data_as(_ip) if n_counts else None, n_counts, C.byref(bc), C.byref(bi),
                                      C.byref(hs), C.byref(fh), eb, 256)
    return {"rc": rc, "err": eb.value.decode(), "best_num_inliers": bc.value, "best_hyp_index": bi.value,
            "hypotheses_scored": hs.value, "filter_hits": fh.value, "counts": counts[:n_counts]}


def emu_lattice(p: capi.Params, half_x, half_y, cap):
    tx, ty = np.zeros(max(cap, 1)), np.zeros(max(cap, 1))
    yaw = np.zeros(1 << 16)
    n_yaw, status = C.c_int(), C.c_int()
    n = emu_lib().spr_emu_lattice(C.byref(p), half_x, half_y, tx.ctypes.data_as(_dp), ty.ctypes.data_as(_dp), cap,
                                  C.byref(n_yaw), yaw.ctypes.data_as(_dp), len(yaw), C.byref(status))
    return n, status.value, tx[:max(n, 0)], ty[:max(n, 0)], yaw[:n_yaw.value]


# ------------------------------------------------------------------ random maps for property tests
def random_maps(rng, n_ref, n_qry, extent=20.0, n_labels=3, planted=True, cyl_frac=0.5, dup_frac=0.1, grid=None):
    """Small random map pair with planted overlap, duplicate/near-duplicate landmarks, cylinders
    (d2 = d3 = 0) and cuboids.  grid: snap coordinates to a multiple of `grid` so that distances
    land exactly on the threshold (tie stress)."""
    def attrs(n):
        lab = rng.integers(0, n_labels, n).astype(np.float64)
        dims = rng.uniform(0.2, 2.5, (n, 3))
        cyl = rng.uniform(size=n) < cyl_frac
        dims[cyl, 1:] = 0.0
        return lab, dims
    ref_xy = rng.uniform(-extent, extent, (n_ref, 2))
    if n_ref > 3 and dup_frac > 0:
        k = max(1, int(dup_frac * n_ref))
        src = rng.integers(0, n_ref, k)
        dst = rng.integers(0, n_ref, k)
        ref_xy[dst] = ref_xy[src] + rng.normal(0, 0.2, (k, 2))
    lab, dims = attrs(n_ref)
    ref = np.column_stack([lab, ref_xy, rng.normal(0, 0.5, n_ref), dims])
    yaw = rng.uniform(-np.pi, np.pi)
    t = rng.uniform(-extent / 3, extent / 3, 2)
    qry = np.zeros((n_qry, 7))
    for j in range(n_qry):
        if planted and n_ref > 0 and rng.uniform() < 0.6:
            i = rng.integers(0, n_ref)
            p = ref[i, 1:3] - t
            c, s = np.cos(-yaw), np.sin(-yaw)
            q = np.array([c * p[0] - s * p[1], s * p[0] + c * p[1]]) + rng.normal(0, 0.05, 2)
            qry[j] = [ref[i, 0], q[0], q[1], ref[i, 3] + rng.normal(0, 0.05), *(ref[i, 4:7] + (ref[i, 4:7] != 0) * rng.normal(0, 0.1, 3))]
        else:
            l2, d2 = attrs(1)
            qry[j] = [l2[0] if rng.uniform() < 0.9 else 99.0, *rng.uniform(-extent, extent, 2), 0.0, *d2[0]]
    if grid:
        ref[:, 1:3] = np.round(ref[:, 1:3] / grid) * grid
        qry[:, 1:3] = np.round(qry[:, 1:3] / grid) * grid
    return ref, qry
