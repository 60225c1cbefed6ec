#!/usr/bin/env python
"""bench.py -- hypotheses scored / s of the SlideMatch lattice search (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle)

A "step" is one pass of the hot path over one map pair: every hypothesis of the reference's
(x, y, yaw) lattice gets its exact inlier count and the best one is selected
(PlaceRecognition::MatchMaps, place_recognition.cpp:98-387).  Workload at N = 1: BASELINE.json
configs[1] -- two synthetic maps of 2,000 landmarks, 5 classes, 10 % outliers, lattice
0.5 m / 5 deg (params/sloam-forest-parking-lot.yaml).  At N > 1 every rank searches its own
config-2 map pair (BASELINE config 4's "one pair-set per GPU"), results are all-gathered (NCCL)
and merged deterministically: weak scaling.  `--workload shard` instead shards ONE pair's
hypothesis space over the ranks (BASELINE config 3's layout) and all-gathers the per-GPU top-1.

value  : hypotheses/s with both maps and their index structures already resident in HBM;
         CUDA events on the launching stream around each step, L2 flushed between steps.
e2e    : the same metric through the public findTransformation call with HOST buffers: index
         build on the host, H2D copies, kernels, D2H of the result, refinement -- wall clock.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROS = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 5.0, "match_threshold_position": 0.5,
       "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 15, "dilation_factor": 1.2}
ORACLE_KW = dict(match_xy_step_size=0.5, yaw_step_deg=5.0, match_threshold=0.5, match_threshold_dimension=1.0,
                 ignore_dimension=0, min_num_inliers=15)
ROOFLINE_TRAFFIC_BYTES = 6.1e6   # dram read 1.63 MB + write 4.51 MB per launch, profiles/r1_q_bound_lattice_summary.txt
ROOFLINE_TRAFFIC_SOURCE = "profiles/r1_q_bound_lattice_summary.txt (ncu --set full, per launch)"
ALG_BYTES_PER_HYP = 24.0  # 16 B hypothesis record + 8 B packed score
METRIC = "hypotheses_scored_per_s"
UNIT = "hypotheses/s"


def workload(config: int, rank: int = 0):
    from slide_slam_b200 import synth
    if config == 1:
        return synth.make_pair(200, seed=1001 + 100 * rank, classes="forest_urban"), "config1: 200 x 200 landmarks, trees + cars"
    if config == 2:
        return (synth.make_pair(2000, seed=1002 + 100 * rank, classes="five", outlier_frac=0.1),
                "config2: 2000 x 2000 landmarks, 5 classes, 10% outliers, exhaustive 0.5 m / 5 deg lattice")
    if config == 3:
        return synth.make_pair(20000, seed=1003 + 100 * rank, classes="forest_urban"), "config3: 20000 x 20000 landmarks"
    raise SystemExit(f"unknown config {config}")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_rate(ref, qry, seconds: float, threads: int):
    """The reference's CPU algorithm (oracle, OpenMP over hypotheses) on the first M hypotheses
    of the workload in canonical order; M sized for ~`seconds` of wall time."""
    from oracle import pyoracle as O
    op = O.make_params(**ORACLE_KW)
    sref, sqry = ref.copy(), qry.copy()
    half = _oracle_ranges(O, op, ref, qry)
    sref[:, 1:3] -= half["centroid_ref"]
    sqry[:, 1:3] -= half["centroid_qry"]
    m = 16 * threads
    t0 = time.perf_counter()
    O.match_maps(op, sref, sqry, half["half_x"], half["half_y"], 0, m, n_threads=threads)
    dt = max(time.perf_counter() - t0, 1e-3)
    m = int(max(m, min(m * seconds / dt, 5e7)))
    return op, sref, sqry, half, m


def _oracle_ranges(O, p, ref, qry):
    """centroids and half ranges with the reference's arithmetic (PR.cpp:713-734, 752-787)."""
    cr = [0.0, 0.0]
    for v in ref:
        cr[0] += v[1]; cr[1] += v[2]
    cr = [cr[0] / float(len(ref)), cr[1] / float(len(ref))]
    cq = [0.0, 0.0]
    for v in qry:
        cq[0] += v[1]; cq[1] += v[2]
    cq = [cq[0] / float(len(qry)), cq[1] / float(len(qry))]
    mx = max(np.abs(ref[:, 1] - cr[0]).max(), np.abs(qry[:, 1] - cq[0]).max())
    my = max(np.abs(ref[:, 2] - cr[1]).max(), np.abs(qry[:, 2] - cq[1]).max())
    m = max(mx, my)
    return {"centroid_ref": np.array(cr), "centroid_qry": np.array(cq), "half_x": float(m * p.dilation_factor),
            "half_y": float(m * p.dilation_factor)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import pyoracle as O
    (ref, qry, _), wname = workload(args.config)
    threads = host_threads()
    per_step = min(3.0, 150.0 / max(args.steps + args.warmup, 1))
    op, sref, sqry, half, m = cpu_reference_rate(ref, qry, per_step, threads)
    for _ in range(args.warmup):
        O.match_maps(op, sref, sqry, half["half_x"], half["half_y"], 0, m, n_threads=threads)
    t0 = time.perf_counter()
    scored = 0
    for _ in range(args.steps):
        r = O.match_maps(op, sref, sqry, half["half_x"], half["half_y"], 0, m, n_threads=threads)
        scored += r["hypotheses_scored"]
    dt = time.perf_counter() - t0
    value = scored / dt
    sample = f"first {m} hypotheses (canonical order) of the workload per step, {threads} OpenMP threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wname, "note": "reference CPU algorithm = oracle/slide_oracle.c (the reference needs ROS/Eigen and cannot be built here)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from slide_slam_b200 import capi
    from slide_slam_b200.place_recognition import PlaceRecognition

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the place-recognition search has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shard_mode = args.workload == "shard" and world > 1
    # weak scaling: every rank searches the SAME synthetic pair (fixed work per GPU), so that the
    # per-N values are comparable; the ranks do not share any data
    (ref, qry, truth), wname = workload(args.config, 0)
    pr = PlaceRecognition(ROS, device=local_rank)
    lib = capi.lib()

    # pinned host buffers for the end-to-end leg
    ref_pin = torch.from_numpy(ref).pin_memory()
    qry_pin = torch.from_numpy(qry).pin_memory()
    ref_h, qry_h = ref_pin.numpy(), qry_pin.numpy()

    found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(ref_h, qry_h)  # also the first warm-up
    sref, sqry = ref.copy(), qry.copy()
    sref[:, 1:3] -= np.array(info.centroid_ref[:])
    sqry[:, 1:3] -= np.array(info.centroid_qry[:])
    pr.prepare(sref, sqry, info.half_x, info.half_y)   # inputs + index structures now resident in HBM
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # results: one 16-byte (canonical index, inliers) record per searched pair.  Weak scaling (every
    # GPU matches its own pairs): the ranks run independently and the records of all timed steps
    # are all-gathered once at the end of the timed region.  Shard mode (one pair split over the
    # GPUs): the top-1 records are all-gathered and merged after every search.
    n_rec = 1 if shard_mode else max(args.steps, 1)
    rec_host = torch.zeros(2 * n_rec, dtype=torch.int64).pin_memory()
    rec_dev = torch.zeros(2 * n_rec, dtype=torch.int64, device=dev)
    rec_all = torch.zeros(2 * n_rec * max(world, 1), dtype=torch.int64, device=dev)
    inc_host = torch.zeros(1, dtype=torch.int64).pin_memory()
    inc_dev = torch.zeros(1, dtype=torch.int64, device=dev)

    def gather_and_merge(res):
        """shard mode: all-gather of the top-1 records + deterministic merge"""
        if world == 1 or not shard_mode:
            return
        rec_host[0], rec_host[1] = int(res.best_hyp_index), int(res.best_num_inliers)
        rec_dev.copy_(rec_host, non_blocking=True)
        dist.all_gather_into_tensor(rec_all, rec_dev)
        recs = (capi.TopkRecord * world)()
        for i, t in enumerate(rec_all.view(world, 2).cpu().tolist()):
            recs[i].hyp_index, recs[i].inliers, recs[i].rank = int(t[0]), int(t[1]), i
        return lib.slide_pr_merge_records(recs, world)

    def one_step():
        if shard_mode:
            # two-phase sharded search: bound phase, all-reduce(max) of the seeds' inlier counts (the
            # incumbent of the branch-and-bound), verification against the shared incumbent
            seed, _ = pr.search(shard_index=rank, shard_count=world, stream=stream.cuda_stream, bounds_only=True)
            inc_host[0] = max(int(seed.best_num_inliers), 0)
            inc_dev.copy_(inc_host, non_blocking=True)
            dist.all_reduce(inc_dev, op=dist.ReduceOp.MAX)
            if seed.search_mode == 0:       # no bound phase for this problem: already searched exhaustively
                res = seed
            else:
                res, _ = pr.search(shard_index=rank, shard_count=world, stream=stream.cuda_stream,
                                   incumbent_inliers=int(inc_dev.item()), reuse_bounds=True)
                res.gpu_launches += seed.gpu_launches
                res.kernel_ms += seed.kernel_ms
        else:
            res, _ = pr.search(stream=stream.cuda_stream)
        gather_and_merge(res)
        return res

    # clocks / throttle reasons are sampled from the warm-up to the end of the exhaustive leg: the
    # timed steps alone (milliseconds) are shorter than one nvidia-smi query
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        one_step()
    if world > 1 and not shard_mode:        # warm-up of the exchange too (NCCL sets its channels up lazily)
        for _ in range(2):
            rec_dev.copy_(rec_host, non_blocking=True)
            dist.all_gather_into_tensor(rec_all, rec_dev)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    step_ms, kern_ms, hyps, launches = [], [], 0, 0
    for i in range(args.steps):
        flush.fill_(1)                      # L2 flush between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = one_step()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        kern_ms.append(res.kernel_ms)
        hyps += res.hypotheses_scored
        launches += res.gpu_launches
        if not shard_mode:
            rec_host[2 * i], rec_host[2 * i + 1] = int(res.best_hyp_index), int(res.best_num_inliers)
    if world > 1 and not shard_mode:        # timed: the one exchange of the weak-scaling job
        torch.cuda.synchronize()
        dist.barrier()                      # untimed, like the L2 flushes: the ranks' untimed host work differs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rec_dev.copy_(rec_host, non_blocking=True)
        dist.all_gather_into_tensor(rec_all, rec_dev)
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        assert bool((rec_all.view(world, -1, 2)[:, :, 1] == int(res.best_num_inliers)).all())  # same pair on every rank
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = float(sum(step_ms))
    tot = torch.tensor([total_ms, float(hyps), float(sum(kern_ms)), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, hyps_all, launches_all = float(mx[0]), float(sm[1]), int(sm[3])
        kernel_total_ms = float(mx[2])
    else:
        hyps_all, launches_all, kernel_total_ms = float(hyps), int(launches), float(sum(kern_ms))
    # the same search with EVERY hypothesis verified exactly (no bound-and-verify pruning), rank 0's GPU
    exh_ms = []
    if rank == 0:
        for i in range(2 + min(args.steps, 5)):
            flush.fill_(1)
            r_x, _ = pr.search(stream=stream.cuda_stream, exhaustive=True)
            if i >= 2:
                exh_ms.append(r_x.kernel_ms)
        exh = {"value_per_gpu": float(r_x.hypotheses_scored) / (float(np.mean(exh_ms)) * 1e-3), "unit": UNIT,
               "kernel_ms_per_step": float(np.mean(exh_ms)), "best_num_inliers": int(r_x.best_num_inliers),
               "same_winner": bool(r_x.best_hyp_index == res.best_hyp_index and r_x.best_num_inliers == res.best_num_inliers),
               "note": "slide_pr_search_opts.exhaustive = 1: exact inlier count of every hypothesis (the mode used with counts_out)"}

    clocks = sampler.stop() if rank == 0 else None  # before the host-timed leg: nvidia-smi polling perturbs wall-clock timing

    # end-to-end through the public API with host buffers.  Two DISTINCT map pairs alternate so that
    # nothing (lattice, reference-map index) can be reused from the previous step: every step pays
    # the full host index build, the H2D copies, the kernels, the D2H of the result and the refinement.
    (ref_b, qry_b, _), _ = workload(args.config, 1000)
    pairs = [(ref_h, qry_h), (torch.from_numpy(ref_b).pin_memory().numpy(), torch.from_numpy(qry_b).pin_memory().numpy())]
    for _ in range(max(args.warmup, 3)):  # untimed warm-up of both pairs (the page-locked buffers grow to their final size)
        for pr_ref, pr_qry in pairs:
            pr.findTransformation(pr_ref, pr_qry)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_hyps = 0
    reused = 0
    for i in range(args.steps):
        f2, _, _, info2, _, _ = pr.findTransformation(*pairs[i & 1])
        e2e_hyps += info2.match.hypotheses_scored
        reused |= info2.match.reuse
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = torch.tensor([e2e_s, float(e2e_hyps)], dtype=torch.float64, device=dev)
    if world > 1:
        a = e2e.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = e2e.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_hyps = float(a[0]), float(b[1])

    if rank == 0:
        value = hyps_all / (total_ms * 1e-3)
        peak, peak_src = measured_peaks()
        # roofline of the dominant kernel (spr_score_lattice_kernel): algorithmic bytes per launch
        # = 24 B x hypotheses of the launch, over its CUDA-event duration (rank 0's launches)
        k_hyps = float(hyps) / max(args.steps, 1)
        k_ms = float(np.mean(kern_ms))
        achieved = ALG_BYTES_PER_HYP * k_hyps / (k_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong" if shard_mode else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "landmarks": [int(len(ref)), int(len(qry))],
                       "hypotheses_per_pair": int(info.match.hypotheses_scored),
                       "lattice": "0.5 m / 5 deg, dilation 1.2 (sloam-forest-parking-lot.yaml)",
                       "parallelism": ("hypothesis space of one pair sharded over %d GPUs: bound phase, NCCL all-reduce(max) of the incumbent, verification, NCCL all-gather of top-1" % world) if shard_mode
                       else ("one map pair per GPU and step (the same synthetic pair on every rank), ranks independent, one NCCL all-gather of all result records at the end of the timed region" if world > 1 else "single GPU"),
                       "l2": "flushed between timed steps (256 MiB write)", "decision_arithmetic": "fp64, non-fused (bit-exact vs reference)",
                       "search": "bound-and-verify (library default): bitmap-filter upper bound of every hypothesis, exact fp64 verification of "
                                 "those whose bound reaches the running best; winner, inlier count and correspondences identical to the exhaustive search"},
            "best_num_inliers": int(info.best_num_inliers), "closure_found": bool(found),
            "kernel_ms_per_step": kernel_total_ms / max(args.steps, 1),
            "exhaustive": exh,
            "gpu_launches": launches_all,
            "clocks": clocks,
            "e2e": {"value": e2e_hyps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(info.match.h2d_bytes),
                    "d2h_bytes_per_step": int(info.match.d2h_bytes), "ms_per_step": e2e_s / max(args.steps, 1) * 1e3,
                    "host_index_build_ms": float(info2.match.prepare_ms), "kernel_ms": float(info2.match.kernel_ms),
                    "index_reused_between_steps": bool(reused),
                    "note": "two distinct map pairs alternate, so every step rebuilds and re-uploads all index structures",
                    "api": "PlaceRecognition.findTransformation -> slide_pr_find_transformation (host buffers)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read + dram__bytes_write per launch of the dominant kernel
                         # (spr_bound_lattice_kernel, 4 launches per step) from the committed ncu capture:
                         # the bit planes of the bounds carried between the launches; maps and bitmaps
                         # stay in L2 / shared memory
                         "traffic": ROOFLINE_TRAFFIC_BYTES, "traffic_source": ROOFLINE_TRAFFIC_SOURCE,
                         "peak_source": peak_src, "kernel": "spr_bound_lattice_kernel",
                         "note": "algorithmic 24 B/hypothesis (SURVEY 8d) x hypotheses of a step / summed kernel time of the step; "
                                 "the kernel is ALU-pipe / shared-memory bound on bitmaps staged in shared memory, not HBM-bound: "
                                 "see DESIGN.md section 5 for the issue-slot and pipe utilisation from ncu"},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle as O
            threads = host_threads()
            op, oref, oqry, half, m = cpu_reference_rate(ref, qry, args.cpu_seconds, threads)
            t0 = time.perf_counter()
            r = O.match_maps(op, oref, oqry, half["half_x"], half["half_y"], 0, m, n_threads=threads)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": r["hypotheses_scored"] / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": f"first {m} hypotheses (canonical order) of the same workload, oracle with {threads} OpenMP threads, {dt:.1f} s"}
        print(json.dumps(out), flush=True)
    pr.close()
    if world > 1:
        dist.destroy_process_group()


def run_extra(args, rank, world, local_rank):
    """BASELINE configs 4 and 5 (not the driver's default line): map-pair matches / s for N-robot
    all-pairs matching, and per-query latency of streaming submap queries against one map."""
    import torch
    from slide_slam_b200 import synth
    from slide_slam_b200.place_recognition import PlaceRecognition
    torch.cuda.set_device(local_rank)
    pr = PlaceRecognition(ROS, device=local_rank)
    if args.config == 4:
        maps = synth.config_robots(8, args.landmarks or 5000)
        pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
        mine = pairs[rank::world]
        for r, q in mine[:1]:
            pr.findTransformation(maps[r], maps[q])  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hyps, found = 0, 0
        for r, q in mine:
            f, _, _, info, _, _ = pr.findTransformation(maps[r], maps[q])
            hyps += info.match.hypotheses_scored
            found += int(f)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"metric": "map_pair_matches_per_s", "value": len(mine) / dt, "unit": "pairs/s", "n_gpus": world,
                          "rank": rank, "pairs": len(mine), "closures_found": found, "hypotheses_per_s": hyps / dt,
                          "config": {"workload": f"config4: 8 robots, 28 pairs of {len(maps[0])} landmarks, pairs dealt round-robin"},
                          "data": "synthetic", "dtype": "f64"}), flush=True)
    else:
        big, queries = synth.config_stream(args.landmarks or 50000, n_queries=max(args.steps, 1), n_sub=300)
        pr.findTransformation(big, queries[0])  # builds the map's index (reused by the stream)
        lat, hyps, found, reuse = [], 0, 0, 0
        for q in queries:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f, _, _, info, _, _ = pr.findTransformation(big, q)
            lat.append((time.perf_counter() - t0) * 1e3)
            hyps += info.match.hypotheses_scored
            found += int(f)
            reuse += int(bool(info.match.reuse & 2))
        lat = np.array(lat)
        print(json.dumps({"metric": "streaming_query_latency_ms", "value": float(np.median(lat)), "unit": "ms", "higher_is_better": False,
                          "n_gpus": 1, "queries": len(queries), "p50": float(np.percentile(lat, 50)), "p95": float(np.percentile(lat, 95)),
                          "p99": float(np.percentile(lat, 99)), "closures_found": found, "index_reused": reuse,
                          "hypotheses_per_query": hyps // max(len(queries), 1), "hypotheses_per_s": hyps / (lat.sum() * 1e-3),
                          "config": {"workload": f"config5: {len(queries)} submap queries of 300 landmarks against a {len(big)}-landmark map"},
                          "data": "synthetic", "dtype": "f64"}), flush=True)
    pr.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--workload", default="pairs", choices=["pairs", "shard"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--landmarks", type=int, default=0, help="override the landmark count of configs 4 / 5")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config in (4, 5):
        run_extra(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
