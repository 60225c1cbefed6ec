#!/usr/bin/env python
"""Regenerates tests/golden/{maps.npz,golden.json,counts.npz}.

Run in the build container (needs /root/reference for the fixture maps):
    python tests/golden/make_golden.py

Inputs:
  * the reference's map fixtures backend/sloam/clipper_semantic_object/examples/data/
    robot{0,1,2}Map_{indoor,parking,forest}.txt (4 columns `label x y z`; dims are zero here --
    place_recognition_test.cpp:92-94 leaves them unset and relies on ignore_dimension),
  * the PRtest synthetic scene (place_recognition_test.cpp:13-28,160-207) regenerated with
    glibc's unseeded rand(),
  * BASELINE.json config 1 from slide_slam_b200.synth.
Outputs are produced by the CPU oracle (oracle/slide_oracle.c).  The reference binary cannot
be built in this image, so these goldens pin the ORACLE's contract.
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402
from slide_slam_b200 import synth  # noqa: E402

REF_DATA = "/root/reference/backend/sloam/clipper_semantic_object/examples/data"

# params/sloam.yaml:26-56 and params/sloam-forest-parking-lot.yaml
SLOAM_YAML = dict(match_xy_step_size=0.1, yaw_step_deg=15.0, match_threshold=0.75,
                  match_threshold_dimension=5.0, ignore_dimension=1, min_num_inliers=8,
                  min_num_map_objects_to_start=5)
FOREST_YAML = dict(match_xy_step_size=0.5, yaw_step_deg=5.0, match_threshold=0.5,
                   match_threshold_dimension=1.0, ignore_dimension=0, min_num_inliers=15,
                   min_num_map_objects_to_start=15)


def load_fixture(name):
    a = np.loadtxt(os.path.join(REF_DATA, name))
    m = np.zeros((a.shape[0], 7))
    m[:, :4] = a
    return m


def prtest_scene():
    """place_recognition_test.cpp:13-28 (generateObjects) and :160-199 (query construction)."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)  # an unseeded program starts from seed 1
    rand = libc.rand
    ref = np.zeros((50, 7))
    for i in range(50):
        ref[i, 0] = i % 10
        ref[i, 1] = 100 * (rand() % 100) / 100.0
        ref[i, 2] = 100 * (rand() % 100) / 100.0
        ref[i, 3] = 0
        ref[i, 4] = 0.5 + 2.0 * (rand() % 100) / 100.0
        ref[i, 5] = 0.5 + 2.0 * (rand() % 100) / 100.0
        ref[i, 6] = 0.5 + 2.0 * (rand() % 100) / 100.0
    out = {"prtest_ref": ref}
    for name, (xt, yt, yaw_deg) in (("inter", (5.25, 5.25, 90.0)), ("intra", (2.25, 1.5, 10.0))):
        yaw = yaw_deg * math.pi / 180.0
        q = []
        for o in ref:  # transformObjects :31-51
            t = o.copy()
            t[1] = o[1] * math.cos(yaw) - o[2] * math.sin(yaw) + xt
            t[2] = o[1] * math.sin(yaw) + o[2] * math.cos(yaw) + yt
            t[3] = o[3] + 0
            q.append(t)
        kept = []
        noise_mag = 0.0
        for t in q:  # :185-199 (erase-in-place loop == filter, rand() consumed in order)
            if rand() % 100 < 20:
                continue
            t[1] += noise_mag * (rand() % 100) / 100.0
            t[2] += noise_mag * (rand() % 100) / 100.0
            t[4] += noise_mag * (rand() % 100) / 100.0
            t[5] += noise_mag * (rand() % 100) / 100.0
            t[6] += noise_mag * (rand() % 100) / 100.0
            kept.append(t)
        out["prtest_qry_" + name] = np.array(kept)
        out["prtest_expected_" + name] = np.array([
            -xt * math.cos(yaw) - yt * math.sin(yaw), xt * math.sin(yaw) - yt * math.cos(yaw), 0.0,
            -yaw])  # :242-247
    return out


def run_case(name, maps, ref_key, qry_key, params, n_threads, **flags):
    p = O.make_params(**params, **flags)
    t0 = time.time()
    r = O.find_transformation(p, maps[ref_key], maps[qry_key], n_threads=n_threads)
    dt = time.time() - t0
    case = {
        "name": name, "ref": ref_key, "qry": qry_key, "params": dict(params, **flags),
        "found": r["found"], "match_status": r["match_status"],
        "best_num_inliers": r["best_num_inliers"], "hypotheses_scored": r["hypotheses_scored"],
        "best_hyp_index": r["best_hyp_index"], "R_t": r["R_t"].ravel().tolist(),
        "ref_idx": r["ref_idx"].tolist(), "qry_idx": r["qry_idx"].tolist(),
        "xyz_yaw": r["xyz_yaw"].tolist(), "transform": r["transform"].ravel().tolist(),
        "half_x": r["half_x"], "half_y": r["half_y"],
        "centroid_ref": r["centroid_ref"].tolist(), "centroid_qry": r["centroid_qry"].tolist(),
    }
    print(f"{name}: best={r['best_num_inliers']} hyp={r['hypotheses_scored']} "
          f"idx={r['best_hyp_index']} found={r['found']} ({dt:.1f}s)", flush=True)
    return case


def main():
    maps = {}
    for env in ("indoor", "parking", "forest"):
        for r in (0, 1, 2):
            maps[f"{env}{r}"] = load_fixture(f"robot{r}Map_{env}.txt")
    maps.update(prtest_scene())
    a, b, truth = synth.config_pair(1)
    maps["c1_ref"], maps["c1_qry"] = a, b
    maps["c1_truth"] = np.array([truth["yaw"], *truth["t"]])
    np.savez_compressed(os.path.join(HERE, "maps.npz"), **maps)

    nt = os.cpu_count() or 1
    cases = []
    # --- indoor fixture, both shipped parameter files, all pairs (PRtest mode 1 uses r0 vs r1)
    for (r, q) in ((0, 1), (0, 2), (1, 2), (1, 0)):
        cases.append(run_case(f"indoor{r}{q}_sloam_yaml", maps, f"indoor{r}", f"indoor{q}", SLOAM_YAML, nt))
        cases.append(run_case(f"indoor{r}{q}_forest_yaml_nodim", maps, f"indoor{r}", f"indoor{q}",
                              dict(FOREST_YAML, ignore_dimension=1, min_num_inliers=8), 1))
    cases.append(run_case("indoor01_sloam_yaml_nolsq", maps, "indoor0", "indoor1", SLOAM_YAML, nt, use_lsq=0))
    cases.append(run_case("indoor01_defaults_2deg", maps, "indoor0", "indoor1",
                          dict(match_xy_step_size=0.5, yaw_step_deg=2.0, ignore_dimension=1), nt))
    cases.append(run_case("indoor01_noyaw", maps, "indoor0", "indoor1",
                          dict(SLOAM_YAML, disable_yaw_search=1), nt))
    # --- parking + forest fixtures with their yaml
    cases.append(run_case("parking01_forest_yaml", maps, "parking0", "parking1", FOREST_YAML, nt))
    cases.append(run_case("parking02_forest_yaml", maps, "parking0", "parking2", FOREST_YAML, nt))
    cases.append(run_case("forest01_forest_yaml", maps, "forest0", "forest1", FOREST_YAML, nt))
    # --- PRtest mode 2 (launch file loads params/sloam.yaml): inter + intra, both use_lsq values
    for lsq in (1, 0):
        cases.append(run_case(f"prtest_inter_lsq{lsq}", maps, "prtest_ref", "prtest_qry_inter",
                              SLOAM_YAML, nt, use_lsq=lsq))
        # findIntraLoopClosure(query, reference, I, I): submap = reference, measurements = query
        cases.append(run_case(f"prtest_intra_lsq{lsq}", maps, "prtest_ref", "prtest_qry_intra",
                              SLOAM_YAML, nt, use_lsq=lsq, inter_loop_closure=0))
    # --- BASELINE config 1 (the reference's own CPU-runnable case)
    cases.append(run_case("c1_forest_yaml", maps, "c1_ref", "c1_qry", FOREST_YAML, nt))

    # --- per-hypothesis inlier counts on slices (for count-level GPU parity)
    counts = {}
    for (cname, rk, qk, prm, sl) in (
            ("indoor01_forest_yaml_nodim", "indoor0", "indoor1",
             dict(FOREST_YAML, ignore_dimension=1, min_num_inliers=8), (0, -1)),
            ("parking01_forest_yaml", "parking0", "parking1", FOREST_YAML, (4_000_000, 4_060_000)),
            ("c1_forest_yaml", "c1_ref", "c1_qry", FOREST_YAML, (4_750_000, 4_800_000)),
            ("prtest_inter_lsq1", "prtest_ref", "prtest_qry_inter", SLOAM_YAML, (12_000_000, 12_050_000))):
        case = next(c for c in cases if c["name"] == cname)
        p = O.make_params(**prm)
        ref, qry = maps[rk].copy(), maps[qk].copy()
        ref[:, 1:3] -= np.array(case["centroid_ref"])
        qry[:, 1:3] -= np.array(case["centroid_qry"])
        t0 = time.time()
        r = O.match_maps(p, ref, qry, case["half_x"], case["half_y"], sl[0], sl[1], want_counts=True)
        counts[cname] = r["counts"]
        counts[cname + "__slice"] = np.array([sl[0], sl[0] + len(r["counts"])], np.int64)
        print(f"counts {cname}: {len(r['counts'])} hyps, max {r['counts'].max()} ({time.time()-t0:.1f}s)")
    np.savez_compressed(os.path.join(HERE, "counts.npz"), **counts)

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": cases}, f, indent=1)


if __name__ == "__main__":
    main()
