"""Deterministic synthetic object maps for the BASELINE.json configs.

Rows are the reference's landmark record [label, x, y, z, d1, d2, d3]
(place_recognition.h:51-57; built at databaseManager.cpp:64-96): a tree cylinder is
[8, root.xyz, radius, 0, 0], a cuboid is [label, centre.xyz, l, w, h].

World: density 0.025 landmarks/m^2, >= 1 m separation, z ~ N(0, 0.5).  Map A is a square
window of side L = sqrt(N / density); map B is a same-size window shifted so that the
shared area is `overlap`, expressed in its own frame through a known SE(3) offset
(roll = pitch = 0), with N(0, sigma) position noise and an optional outlier fraction.
numpy PCG64, seed = 1000 + config id.
"""
from __future__ import annotations

import math

import numpy as np

DENSITY = 0.025
MIN_SEP = 1.0


def _poisson_disc_uniform(rng, n, x0, x1, y0, y1, existing=None):
    """n uniform points in the box, rejection-sampled to MIN_SEP from each other and from
    `existing` (hash-grid accelerated)."""
    cell = MIN_SEP
    grid = {}

    def key(p):
        return (int(math.floor(p[0] / cell)), int(math.floor(p[1] / cell)))

    def ok(p):
        kx, ky = key(p)
        for ax in (kx - 1, kx, kx + 1):
            for ay in (ky - 1, ky, ky + 1):
                for q in grid.get((ax, ay), ()):
                    if (q[0] - p[0]) ** 2 + (q[1] - p[1]) ** 2 < MIN_SEP * MIN_SEP:
                        return False
        return True

    if existing is not None:
        for q in existing:
            grid.setdefault(key(q), []).append((float(q[0]), float(q[1])))
    out = []
    while len(out) < n:
        m = max(2 * (n - len(out)), 16)
        cand = np.column_stack([rng.uniform(x0, x1, m), rng.uniform(y0, y1, m)])
        for p in cand:
            if len(out) == n:
                break
            p = (float(p[0]), float(p[1]))
            if ok(p):
                grid.setdefault(key(p), []).append(p)
                out.append(p)
    return np.array(out, np.float64).reshape(-1, 2)


def _attributes(rng, n, classes):
    """labels + dims.  classes='forest_urban': 70 % tree cylinders (label 8), 30 % car cuboids
    (label 5); classes='five': labels 1..5 uniform, cuboid dims."""
    lab = np.zeros(n)
    dims = np.zeros((n, 3))
    if classes == "forest_urban":
        is_tree = rng.uniform(size=n) < 0.7
        lab[:] = np.where(is_tree, 8.0, 5.0)
        dims[:, 0] = np.where(is_tree, rng.uniform(0.1, 0.5, n), rng.uniform(2.0, 5.5, n))
        dims[:, 1] = np.where(is_tree, 0.0, rng.uniform(1.0, 2.5, n))
        dims[:, 2] = np.where(is_tree, 0.0, rng.uniform(0.5, 2.0, n))
    elif classes == "five":
        lab[:] = rng.integers(1, 6, n).astype(np.float64)
        dims[:, 0] = rng.uniform(2.0, 5.5, n)
        dims[:, 1] = rng.uniform(1.0, 2.5, n)
        dims[:, 2] = rng.uniform(0.5, 2.0, n)
    else:
        raise ValueError(classes)
    return lab, dims


def _to_frame(xyz, yaw, t):
    """world -> robot frame: p_r = Rz(yaw)^T (p_w - t)."""
    c, s = math.cos(yaw), math.sin(yaw)
    d = xyz - t[None, :]
    out = np.empty_like(d)
    out[:, 0] = c * d[:, 0] + s * d[:, 1]
    out[:, 1] = -s * d[:, 0] + c * d[:, 1]
    out[:, 2] = d[:, 2]
    return out


def make_pair(n: int, *, seed: int, classes: str = "forest_urban", overlap: float = 0.3,
              outlier_frac: float = 0.0, sigma: float = 0.05, n_b: int | None = None):
    """Two maps (A = reference, B = query) of n (n_b) landmarks.  Returns (A7, B7, truth) where
    truth = dict(yaw, t) is the pose of B's frame in A's (world) frame: p_A = Rz(yaw) p_B + t."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_b = n if n_b is None else n_b
    L = math.sqrt(n / DENSITY)
    n_shared = int(round(overlap * min(n, n_b)))
    shift = (1.0 - overlap) * L
    shared = _poisson_disc_uniform(rng, n_shared, shift, L, 0.0, L)
    only_a = _poisson_disc_uniform(rng, n - n_shared, 0.0, shift, 0.0, L, shared)
    only_b = _poisson_disc_uniform(rng, n_b - n_shared, L, L + shift, 0.0, L, shared)
    world_xy = np.vstack([shared, only_a, only_b])
    nw = world_xy.shape[0]
    world = np.column_stack([world_xy, rng.normal(0.0, 0.5, nw)])
    lab, dims = _attributes(rng, nw, classes)
    ia = np.arange(0, n)                                   # shared + only_a
    ib = np.concatenate([np.arange(0, n_shared), np.arange(n, nw)])
    A = np.column_stack([lab[ia], world[ia], dims[ia]])
    yaw = float(rng.uniform(-math.pi, math.pi))
    t = np.array([rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(-2, 2)])
    # choose the B-frame origin near B's window so coordinates stay local, then add the offset
    origin = np.array([L + 0.5 * shift - 0.5 * L, 0.5 * L, 0.0]) + t
    b_xyz = _to_frame(world[ib], yaw, origin)
    b_xyz[:, :2] += rng.normal(0.0, sigma, (len(ib), 2))
    B = np.column_stack([lab[ib], b_xyz, dims[ib]])
    if outlier_frac > 0:
        k = int(round(outlier_frac * n_b))
        idx = rng.choice(n_b, k, replace=False)
        fresh = np.column_stack([rng.uniform(shift, L + shift, k), rng.uniform(0, L, k),
                                 rng.normal(0.0, 0.5, k)])
        B[idx, 1:4] = _to_frame(fresh, yaw, origin)
    A = A[rng.permutation(n)]
    B = B[rng.permutation(n_b)]
    return np.ascontiguousarray(A), np.ascontiguousarray(B), {"yaw": yaw, "t": origin, "L": L}


def config_pair(config_id: int, n: int | None = None):
    """BASELINE.json configs 1..3 as (A7, B7, truth); n overrides the landmark count."""
    if config_id == 1:
        return make_pair(n or 200, seed=1001, classes="forest_urban")
    if config_id == 2:
        return make_pair(n or 2000, seed=1002, classes="five", outlier_frac=0.1)
    if config_id == 3:
        return make_pair(n or 20000, seed=1003, classes="forest_urban")
    raise ValueError(config_id)


def config_robots(n_robots: int = 8, n: int = 5000, seed: int = 1004):
    """Config 4: n_robots maps of n landmarks (5 classes) cut from one world; windows are laid
    on a ring so neighbours overlap by ~30 %.  Returns list of maps (each in its own frame)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = math.sqrt(n / DENSITY)
    side = 3.2 * L
    nw = int(DENSITY * side * side)
    world_xy = _poisson_disc_uniform(rng, nw, 0, side, 0, side)
    world = np.column_stack([world_xy, rng.normal(0, 0.5, nw)])
    lab, dims = _attributes(rng, nw, "five")
    maps = []
    for r in range(n_robots):
        ang = 2 * math.pi * r / n_robots
        cx = 0.5 * side + 0.9 * L * math.cos(ang)
        cy = 0.5 * side + 0.9 * L * math.sin(ang)
        inside = np.nonzero((np.abs(world[:, 0] - cx) <= L / 2) & (np.abs(world[:, 1] - cy) <= L / 2))[0]
        inside = inside[rng.permutation(len(inside))][:n]
        yaw = float(rng.uniform(-math.pi, math.pi))
        origin = np.array([cx + rng.uniform(-50, 50), cy + rng.uniform(-50, 50), rng.uniform(-2, 2)])
        xyz = _to_frame(world[inside], yaw, origin)
        xyz[:, :2] += rng.normal(0, 0.05, (len(inside), 2))
        maps.append(np.ascontiguousarray(np.column_stack([lab[inside], xyz, dims[inside]])))
    return maps


def config_stream(n_map: int = 50000, n_queries: int = 1000, n_sub: int = 300, seed: int = 1005):
    """Config 5: an accumulated map of n_map landmarks and n_queries submaps of n_sub landmarks,
    each a ~110 m window of the map seen from its own frame with noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = math.sqrt(n_map / DENSITY)
    world_xy = _poisson_disc_uniform(rng, n_map, 0, L, 0, L)
    world = np.column_stack([world_xy, rng.normal(0, 0.5, n_map)])
    lab, dims = _attributes(rng, n_map, "five")
    big = np.ascontiguousarray(np.column_stack([lab, world, dims]))
    w = math.sqrt(n_sub / DENSITY)
    queries = []
    for _ in range(n_queries):
        cx, cy = rng.uniform(w, L - w, 2)
        d2 = (world[:, 0] - cx) ** 2 + (world[:, 1] - cy) ** 2
        idx = np.argsort(d2)[:n_sub]
        idx = idx[rng.permutation(n_sub)]
        yaw = float(rng.uniform(-math.pi, math.pi))
        origin = np.array([cx + rng.uniform(-20, 20), cy + rng.uniform(-20, 20), rng.uniform(-2, 2)])
        xyz = _to_frame(world[idx], yaw, origin)
        xyz[:, :2] += rng.normal(0, 0.05, (n_sub, 2))
        queries.append(np.ascontiguousarray(np.column_stack([lab[idx], xyz, dims[idx]])))
    return big, queries
