"""Independent numpy restatement of PlaceRecognition::MatchMaps (place_recognition.cpp:98-387).

TEST INFRASTRUCTURE ONLY.  Written separately from oracle/slide_oracle.c (vectorised over
yaw x query x reference instead of nested scalar loops) so that the two restatements pin
each other; numpy ufuncs round every multiply and add separately (no FMA), like the
reference build.  Only practical for small maps.
"""
from __future__ import annotations

import math

import numpy as np


def yaw_candidates(yaw_half: float, yaw_step: float, disable: bool = False):
    if disable:                                    # PR.cpp:137-139
        return [0.0]
    out, y = [], -yaw_half
    while y < yaw_half:                            # PR.cpp:140-145
        out.append(y)
        y += yaw_step
    return out


def lattice(half_x: float, half_y: float, step: float):
    """Translations in canonical order (PR.cpp:151-241); None on the sanity-check return."""
    outer = 10 * step
    rings = math.ceil(min(half_x, half_y) / outer)
    if rings == 0:
        return []
    ox, oy = half_x / float(rings), half_y / float(rings)
    if ox < step or oy < step:
        return None
    out = []
    for k in range(rings):
        kd = float(k)
        xl, xr, yl, yr = -kd * ox, kd * ox, -kd * oy, kd * oy
        x_end, y_end = (kd + 1) * ox, (kd + 1) * oy
        x = -(kd + 1) * ox
        while x <= x_end:
            y = -(kd + 1) * oy
            while y <= y_end:
                if not ((xl <= x <= xr) and (yl <= y <= yr)):
                    out.append((x, y, k))
                y += step
            x += step
    return out


def match_maps(ref7, qry7, half_x, half_y, *, step=0.5, yaw_half=math.pi, yaw_step=None,
               thr=0.5, thr_dim=1.0, ignore_dimension=False, disable_yaw_search=False,
               want_counts=False):
    ref7 = np.asarray(ref7, np.float64).reshape(-1, 7)
    qry7 = np.asarray(qry7, np.float64).reshape(-1, 7)
    yaws = yaw_candidates(yaw_half, yaw_step, disable_yaw_search)
    lat = lattice(half_x, half_y, step)
    if lat is None:
        return {"status": 1}
    cs = np.array([[math.cos(a), math.sin(a)] for a in yaws], np.float64).reshape(-1, 2)
    c, s = cs[:, 0:1], cs[:, 1:2]
    qx, qy = qry7[None, :, 1], qry7[None, :, 2]
    rot_x = c * qx + (-s) * qy                     # (n_yaw, Nq)  PR.cpp:257-258 first two terms
    rot_y = s * qx + c * qy
    same = ref7[None, :, 0] == qry7[:, None, 0]    # (Nq, Nr)     PR.cpp:306
    if ignore_dimension:
        dim_ok = np.ones_like(same)
    else:
        dd = np.abs(ref7[None, :, 4:7] - qry7[:, None, 4:7])          # (Nq, Nr, 3)
        avg3 = (((0.0 + dd[..., 0]) + dd[..., 1]) + dd[..., 2]) / 3    # PR.cpp:324-329
        cyl = (ref7[:, 5] == 0) & (ref7[:, 6] == 0)                    # PR.cpp:318 (ref only)
        avg = np.where(cyl[None, :], dd[..., 0], avg3)
        dim_ok = avg < thr_dim
    compat = same & dim_ok
    best, best_h, best_rec = -10000, -1, None
    counts = []
    h = 0
    for (x, y, _k) in lat:
        xt = rot_x + x                              # (n_yaw, Nq)
        yt = rot_y + y
        dx = ref7[None, None, :, 1] - xt[:, :, None]
        dy = ref7[None, None, :, 2] - yt[:, :, None]
        ok = (np.sqrt(dx * dx + dy * dy) < thr) & compat[None]
        hit = ok.any(axis=2)                        # (n_yaw, Nq)
        cnt = hit.sum(axis=1)
        if want_counts:
            counts.append(cnt.astype(np.int32))
        a = int(np.argmax(cnt))                     # first max == strict '>' PR.cpp:361
        if cnt[a] > best:
            best, best_h = int(cnt[a]), h + a
            first = ok[a].argmax(axis=1)
            qi = np.nonzero(hit[a])[0]
            best_rec = (cs[a, 0], cs[a, 1], x, y, first[qi].astype(np.int32), qi.astype(np.int32))
        h += len(yaws)
    out = {"status": 0, "best_num_inliers": best, "best_hyp_index": best_h,
           "hypotheses_scored": h, "n_yaw": len(yaws)}
    if best_rec is not None:
        cc, ss, x, y, ri, qi = best_rec
        out["R_t"] = np.array([[cc, -ss, x], [ss, cc, y], [0, 0, 1.0]])
        out["ref_idx"], out["qry_idx"] = ri, qi
    if want_counts:
        out["counts"] = np.concatenate(counts) if counts else np.zeros(0, np.int32)
    return out
