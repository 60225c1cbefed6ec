#!/usr/bin/env python
"""bench.py -- hypotheses scored / s of the SlideMatch lattice search (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle)

A "step" is one pass of the hot path over one map pair: EVERY hypothesis of the reference's
(x, y, yaw) lattice gets its exact inlier count and the best one is selected
(PlaceRecognition::MatchMaps, place_recognition.cpp:98-387).  Workload at N = 1: BASELINE.json
configs[1] -- two synthetic maps of 2,000 landmarks, 5 classes, 10 % outliers, lattice
0.5 m / 5 deg (params/sloam-forest-parking-lot.yaml).  At N > 1 every rank matches a config-2
map pair per step (BASELINE config 4's "one pair-set per GPU"): weak scaling, no data-path
collective.

value  : hypotheses SCORED / s -- the library's default search (`slide_pr_search`, pair-join scorer: the exact
         inlier count of EVERY hypothesis, arg-max on the device) with both maps and their index structures
         resident in HBM; CUDA events on the launching stream around each step, L2 flushed between steps.
e2e    : the same metric through the public findTransformation call with HOST buffers: index build, H2D
         copies, kernels, D2H of the result, refinement -- wall clock.
lattice_kernels : the same pair through the library's other engine (occupancy-bitmap lattice kernels): every
         hypothesis verified exactly, and bound-and-verify (an upper bound for every hypothesis, exact
         verification where the bound reaches the running best).  Same winner; not the metric.
roofline: the dominant kernel against the instruction-issue peak MEASURED in the same run by a
         micro-kernel (the path is issue bound on shared-memory / L1 resident structures);
         the HBM figure of SURVEY.md section 8d is kept beside it.
extras : generator (triangle-hypothesis set (T) on the same pair), config3_shard (one 20 000 x 20 000
         pair, hypothesis space sharded over the ranks), config4 (28 pairs of 5 000 landmarks dealt
         over the ranks), config5 (streaming 300-landmark queries against 50 000 landmarks).
"""
from __future__ import annotations

import argparse
import gc
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
ALG_BYTES_PER_HYP = 24.0  # 16 B hypothesis record + 8 B packed score
METRIC = "hypotheses_scored_per_s"
UNIT = "hypotheses/s"
NCU_COUNTS = os.path.join(ROOT, "profiles", "r2_instr_counts.json")  # warp instructions / DRAM bytes per step, from ncu


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


def config_dict(wname, ref, qry, world, n_hyp):
    """the `config` object: identical in both arms (n_hyp: lattice size, counted by whichever arm runs)"""
    return {"workload": wname, "landmarks": [int(len(ref)), int(len(qry))], "hypotheses_per_pair": int(n_hyp),
            "lattice": "0.5 m / 5 deg, dilation 1.2 (sloam-forest-parking-lot.yaml)", "pairs_per_step": int(max(world, 1)),
            "decision_arithmetic": "fp64, non-fused", "l2": "flushed between timed steps (256 MiB write)"}


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
        busy = [v for v in sm if v > 0.5 * (max(smax) if smax else 1)]
        return {"sm_mhz": float(np.median(busy or sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
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


def _ranges(ref, qry, dilation=ROS["dilation_factor"]):
    """centroids and half ranges with the reference's arithmetic (PR.cpp:713-734, 752-787); plain numpy, no library"""
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
    return {"centroid_ref": np.array(cr), "centroid_qry": np.array(cq), "half_x": float(m * dilation), "half_y": float(m * dilation)}


def cpu_reference_setup(ref, qry):
    from oracle import pyoracle as O
    op = O.make_params(**ORACLE_KW)
    sref, sqry = ref.copy(), qry.copy()
    half = _ranges(ref, qry, op.dilation_factor)
    sref[:, 1:3] -= half["centroid_ref"]
    sqry[:, 1:3] -= half["centroid_qry"]
    return O, op, sref, sqry, half


def cpu_reference_rate(ref, qry, seconds: float, threads: int, gpu=None):
    """The reference's CPU algorithm (oracle) on the first M hypotheses of the workload in canonical
    order; M sized for ~`seconds` of wall time with `threads` threads (1 = the faithful single loop:
    the reference runs MatchMaps on one std::thread, sloamNode.cpp:109)."""
    O, op, sref, sqry, half = cpu_reference_setup(ref, qry)
    m = 16 * threads
    t0 = time.perf_counter()
    O.match_maps(op, sref, sqry, half["half_x"], half["half_y"], 0, m, n_threads=threads)
    dt = max(time.perf_counter() - t0, 1e-3)
    m = int(max(m, min(m * seconds / dt, 5e7)))
    t0 = time.perf_counter()
    r = O.match_maps(op, sref, sqry, half["half_x"], half["half_y"], 0, m, want_counts=gpu is not None, n_threads=threads)
    dt = time.perf_counter() - t0
    what = "single thread, the reference's own loop nest" if threads == 1 else f"{threads} OpenMP threads over hypotheses"
    out = {"value": r["hypotheses_scored"] / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"first {m} hypotheses (canonical order) of the same workload, oracle/slide_oracle.c, {what}, {dt:.1f} s"}
    if gpu is not None:   # the results on the CPU's slice must match the GPU's on the same slice
        n_yaw = gpu.lattice_info()[1]
        te = m // n_yaw                       # whole translations inside the sample
        if te > 0:
            _, got = gpu.search(0, te, want_counts=True)
            out["slice_equals_gpu"] = bool(np.array_equal(got, r["counts"][: te * n_yaw]))
            out["slice_hypotheses_compared"] = int(te * n_yaw)
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    (ref, qry, _), wname = workload(args.config)
    threads = host_threads()
    O, op, sref, sqry, half = cpu_reference_setup(ref, qry)
    lat = O.enumerate_lattice(op, half["half_x"], half["half_y"])
    n_hyp = 0 if lat is None else len(lat[0]) * len(lat[3])
    per_step = min(3.0, 150.0 / max(args.steps + args.warmup, 1))
    m = 16 * threads
    t0 = time.perf_counter()
    O.match_maps(op, sref, sqry, half["half_x"], half["half_y"], 0, m, n_threads=threads)
    m = int(max(m, min(m * per_step / max(time.perf_counter() - t0, 1e-3), 5e7)))
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
        "config": config_dict(wname, ref, qry, world, n_hyp),
        "note": "reference CPU algorithm = oracle/slide_oracle.c, a line-by-line restatement (the reference needs ROS/Eigen and cannot be built here); rank 0 only",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def load_ncu_counts():
    try:
        with open(NCU_COUNTS) as f:
            return json.load(f)
    except Exception:
        return {}


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
    # weak scaling: every rank searches the SAME synthetic pair (fixed work per GPU), so that the
    # per-N values are comparable; the ranks do not share any data
    (ref, qry, truth), wname = workload(args.config, 0)
    pr = PlaceRecognition(ROS, device=local_rank)                        # library defaults: the pair-join scorer
    prl = PlaceRecognition(ROS, device=local_rank, engine="lattice")     # the lattice kernels (bound-and-verify / exhaustive)
    lib = capi.lib()

    ref_h = torch.from_numpy(ref).pin_memory().numpy()
    qry_h = torch.from_numpy(qry).pin_memory().numpy()
    found, xyz_yaw, tf, info, ri, qi = pr.findTransformation(ref_h, qry_h)  # also the first warm-up
    sref, sqry = ref.copy(), qry.copy()
    sref[:, 1:3] -= np.array(info.centroid_ref[:])
    sqry[:, 1:3] -= np.array(info.centroid_qry[:])
    pr.prepare(sref, sqry, info.half_x, info.half_y)   # inputs + index structures now resident in HBM
    prl.prepare(sref, sqry, info.half_x, info.half_y)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # one 16-byte (canonical index, inliers) record per searched pair; the records of all timed steps
    # are all-gathered once at the end of the timed region (the ranks' searches are independent)
    n_rec = max(args.steps, 1)
    rec_host = torch.zeros(2 * n_rec, dtype=torch.int64).pin_memory()
    rec_dev = torch.zeros(2 * n_rec, dtype=torch.int64, device=dev)
    rec_all = torch.zeros(2 * n_rec * max(world, 1), dtype=torch.int64, device=dev)

    def timed_steps(handle, exhaustive, steps):
        step_ms, kern_ms, hyps, launches, last = [], [], 0, 0, None
        gc.collect(); gc.disable()              # no collector pause between a step's two events (as timeit does)
        try:
            return _timed_steps(handle, exhaustive, steps, step_ms, kern_ms, hyps, launches, last)
        finally:
            gc.enable()

    def _timed_steps(handle, exhaustive, steps, step_ms, kern_ms, hyps, launches, last):
        for i in range(steps):
            flush.fill_(1)                      # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            res, _ = handle.search(stream=stream.cuda_stream, exhaustive=exhaustive)
            e1.record(stream)
            e1.synchronize()
            # the step's events enclose the library's own pair of events around its kernels (same stream), so the step cannot be
            # shorter; with the L2 flush kernel right in front the outer pair has been seen to read up to 2 % LESS than the inner
            # one -- the longer of the two is what counts
            step_ms.append(max(e0.elapsed_time(e1), float(res.kernel_ms))); kern_ms.append(res.kernel_ms)
            hyps += res.hypotheses_scored; launches += res.gpu_launches
            if handle is pr:
                rec_host[2 * i], rec_host[2 * i + 1] = int(res.best_hyp_index), int(res.best_num_inliers)
            last = res
        return step_ms, kern_ms, hyps, launches, last

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        pr.search(stream=stream.cuda_stream)
        prl.search(stream=stream.cuda_stream, exhaustive=True)
        prl.search(stream=stream.cuda_stream)
    if world > 1:        # warm-up of the exchange too (NCCL sets its channels up lazily)
        for _ in range(2):
            rec_dev.copy_(rec_host, non_blocking=True)
            dist.all_gather_into_tensor(rec_all, rec_dev)
        torch.cuda.synchronize()
        dist.barrier()
    torch.cuda.synchronize()

    # ---- the metric: every hypothesis scored exactly
    step_ms, kern_ms, hyps, launches, res_x = timed_steps(pr, False, args.steps)
    exchange_ms = 0.0
    if world > 1:        # timed: the one exchange of the weak-scaling job
        torch.cuda.synchronize()
        dist.barrier()                      # untimed, like the L2 flushes: the ranks' untimed host work differs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rec_dev.copy_(rec_host, non_blocking=True)
        dist.all_gather_into_tensor(rec_all, rec_dev)
        e1.record(stream)
        e1.synchronize()
        exchange_ms = e0.elapsed_time(e1)
        step_ms.append(exchange_ms)
        assert bool((rec_all.view(world, -1, 2)[:, :, 1] == int(res_x.best_num_inliers)).all())  # same pair on every rank
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # ---- the lattice kernels on the same pair: bound-and-verify, and every hypothesis verified exactly
    n_l = max(min(args.steps, 5), 1)
    b_step_ms, b_kern_ms, b_hyps, b_launches, res_b = timed_steps(prl, False, n_l)
    lx_step_ms, lx_kern_ms, lx_hyps, lx_launches, res_lx = timed_steps(prl, True, n_l)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    def reduce_over_ranks(total_ms, count, kern, launches_):
        t = torch.tensor([total_ms, float(count), kern, float(launches_)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            return float(mx[0]), float(sm[1]), float(mx[2]), int(sm[3])
        return total_ms, float(count), kern, int(launches_)

    total_ms, hyps_all, kernel_total_ms, launches_all = reduce_over_ranks(float(sum(step_ms)), hyps, float(sum(kern_ms)), launches)
    b_total_ms, b_hyps_all, b_kernel_total_ms, b_launches_all = reduce_over_ranks(float(sum(b_step_ms)), b_hyps, float(sum(b_kern_ms)), b_launches)
    lx_total_ms, lx_hyps_all, lx_kernel_total_ms, lx_launches_all = reduce_over_ranks(float(sum(lx_step_ms)), lx_hyps, float(sum(lx_kern_ms)), lx_launches)

    peaks = None
    if rank == 0:   # issue-rate micro-benchmark, same run, same clocks
        import ctypes as C
        a, f, m = C.c_double(), C.c_double(), C.c_double()
        if lib.slide_pr_measure_issue_peaks(pr._h, C.byref(a), C.byref(f), C.byref(m)) == 0:
            peaks = {"alu_pipe_winst_per_s": a.value, "fma_pipe_winst_per_s": f.value, "issue_winst_per_s": m.value}
    clocks = sampler.stop() if rank == 0 else None  # before the host-timed legs: nvidia-smi polling perturbs wall-clock timing

    # ---- end to end through the public API with host buffers.  Two DISTINCT map pairs alternate so that
    # nothing (lattice, reference-map index) can be reused from the previous step: every step pays the full
    # index build, the H2D copies, the kernels, the D2H of the result and the refinement.
    (ref_b, qry_b, _), _ = workload(args.config, 1000)
    pairs = [(ref_h, qry_h), (torch.from_numpy(ref_b).pin_memory().numpy(), torch.from_numpy(qry_b).pin_memory().numpy())]

    def e2e_leg(handle):
        for _ in range(max(args.warmup, 3)):  # untimed warm-up of both pairs (the page-locked buffers grow to their final size)
            for a_, b_ in pairs:
                handle.findTransformation(a_, b_)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gc.collect(); gc.disable()
        t0 = time.perf_counter()
        n_h, reused, last = 0, 0, None
        for i in range(args.steps):
            _, _, _, last, _, _ = handle.findTransformation(*pairs[i & 1])
            n_h += last.match.hypotheses_scored
            reused |= last.match.reuse
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gc.enable()
        t = torch.tensor([dt, float(n_h)], dtype=torch.float64, device=dev)
        if world > 1:
            a_ = t.clone(); dist.all_reduce(a_, op=dist.ReduceOp.MAX)
            b_ = t.clone(); dist.all_reduce(b_, op=dist.ReduceOp.SUM)
            dt, n_h = float(a_[0]), float(b_[1])
        return dt, n_h, reused, last

    x_s, x_hyps, x_reused, x_info = e2e_leg(pr)
    d_s, d_hyps, d_reused, d_info = e2e_leg(prl)

    extras = run_extras(args, rank, world, local_rank, dev) if not args.no_extras else {}

    if rank == 0:
        n_steps = max(args.steps, 1)
        value = hyps_all / (total_ms * 1e-3)
        hbm_peak, hbm_src = measured_peaks()
        counts = load_ncu_counts()
        k_ms = float(np.mean(kern_ms))                          # rank 0's step: rotate + spr_join_score_kernel
        k_hyps = float(hyps) / n_steps
        winst = counts.get("join_c2_winst_per_step")
        roof = {"bound": "issue", "kernel": "spr_join_score_kernel (one launch per search; > 99 % of the step)",
                "unit": "Gwinst/s", "achieved": None, "peak": None, "frac": None,
                "traffic": counts.get("join_c2_dram_bytes_per_launch"),
                "traffic_source": counts.get("source", "no ncu capture committed"),
                "peak_source": "issue-rate micro-kernel (LOP3 + IMAD chains interleaved) in this run, slide_pr_measure_issue_peaks",
                "note": "achieved = warp instructions of the step's launches (ncu smsp__inst_executed.sum, committed) / their CUDA-event "
                        "duration in this run; the kernel joins L1 / shared-memory resident landmark bins and counts in shared memory, so "
                        "instruction issue is the binding roof, not HBM (SURVEY.md section 8d)"}
        if winst and peaks:
            ach = winst / (k_ms * 1e-3)
            roof.update({"achieved": ach / 1e9, "peak": peaks["issue_winst_per_s"] / 1e9, "frac": ach / peaks["issue_winst_per_s"],
                         "winst_per_hypothesis": winst / k_hyps,
                         "alu_pipe": {"share_of_instructions": counts.get("join_c2_alu_share"),
                                      "peak_Gwinst_per_s": peaks["alu_pipe_winst_per_s"] / 1e9,
                                      "frac": (ach * counts["join_c2_alu_share"] / peaks["alu_pipe_winst_per_s"])
                                      if counts.get("join_c2_alu_share") else None}})
        hbm_ach = ALG_BYTES_PER_HYP * k_hyps / (k_ms * 1e-3) / 1e9
        roof["hbm"] = {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "peak_source": hbm_src,
                       "note": "algorithmic 24 B/hypothesis (SURVEY 8d) x hypotheses of a step / kernel time: the looser roof"}
        if peaks:
            roof["measured_peaks"] = {k: v / 1e9 for k, v in peaks.items()}
            roof["measured_peaks"]["unit"] = "Gwinst/s"
        b_winst, lx_winst = counts.get("search_c2_winst_per_step"), counts.get("exhaustive_c2_winst_per_step")
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / n_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(wname, ref, qry, world, int(info.match.hypotheses_scored)),
            "parallelism": ("one map pair per GPU and step (the same synthetic pair on every rank), ranks independent, one NCCL all-gather of "
                            "all result records at the end of the timed region" if world > 1 else "single GPU"),
            "best_num_inliers": int(res_x.best_num_inliers), "closure_found": bool(found),
            "kernel_ms_per_step": kernel_total_ms / n_steps, "gpu_launches": launches_all,
            "exchange_ms": exchange_ms,   # rank 0: the one all-gather of all steps' records (inside the timed region, N > 1)
            "pairs_per_s": world * n_steps / (total_ms * 1e-3),
            "search_mode": int(res_x.search_mode),
            "lattice_kernels": {
                "note": "the library's other engine on the same pair (engine='lattice' / exhaustive_search = 2 | 1): occupancy-bitmap lattice "
                        "kernels; bound_and_verify = upper bound of every hypothesis + exact verification of those reaching the running best; "
                        "exhaustive = every hypothesis verified; same winner, count and correspondences as the pair-join scorer",
                "same_winner": bool(res_b.best_hyp_index == res_x.best_hyp_index and res_b.best_num_inliers == res_x.best_num_inliers and
                                    res_lx.best_hyp_index == res_x.best_hyp_index and res_lx.best_num_inliers == res_x.best_num_inliers),
                "bound_and_verify": {"pairs_per_s": world * n_l / (b_total_ms * 1e-3), "ms_per_pair": b_total_ms / n_l,
                                     "kernel_ms_per_pair": b_kernel_total_ms / n_l, "gpu_launches_per_pair": b_launches_all / (world * n_l),
                                     "issue_frac": (b_winst / (float(np.mean(b_kern_ms)) * 1e-3) / peaks["issue_winst_per_s"]) if (b_winst and peaks) else None,
                                     "e2e_pairs_per_s": world * n_steps / d_s, "e2e_ms_per_pair": d_s / n_steps * 1e3,
                                     "e2e_h2d_bytes_per_pair": int(d_info.match.h2d_bytes), "e2e_host_prepare_ms": float(d_info.match.prepare_ms)},
                "exhaustive": {"hypotheses_per_s": lx_hyps_all / (lx_total_ms * 1e-3), "ms_per_pair": lx_total_ms / n_l,
                               "kernel_ms_per_pair": lx_kernel_total_ms / n_l, "gpu_launches_per_pair": lx_launches_all / (world * n_l),
                               "issue_frac": (lx_winst / (float(np.mean(lx_kern_ms)) * 1e-3) / peaks["issue_winst_per_s"]) if (lx_winst and peaks) else None}},
            "clocks": clocks,
            "e2e": {"value": x_hyps / x_s, "unit": UNIT, "h2d_bytes_per_step": int(x_info.match.h2d_bytes),
                    "d2h_bytes_per_step": int(x_info.match.d2h_bytes), "ms_per_step": x_s / n_steps * 1e3,
                    "host_prepare_ms": float(x_info.match.prepare_ms), "kernel_ms": float(x_info.match.kernel_ms),
                    "index_reused_between_steps": bool(x_reused), "search_mode": int(x_info.match.search_mode),
                    "pairs_per_s": world * n_steps / x_s,
                    "note": "two distinct map pairs alternate, so every step rebuilds all index structures (landmark bins, query groups, lattice blocks)",
                    "api": "PlaceRecognition.findTransformation -> slide_pr_find_transformation (host buffers)"},
            "roofline": roof,
        }
        out.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            pr.prepare(sref, sqry, info.half_x, info.half_y)   # the e2e legs re-prepared the handle
            out["cpu_baseline"] = cpu_reference_rate(ref, qry, args.cpu_seconds, threads, gpu=pr)
            out["cpu_baseline_1thread"] = cpu_reference_rate(ref, qry, min(args.cpu_seconds, 8.0), 1)
        print(json.dumps(out), flush=True)
    pr.close(); prl.close()
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, rank, world, local_rank, dev):
    """Extra keys of the default line: the (T) hypothesis set on the config-2 pair, config 3 sharded over
    the ranks (strong scaling), config 4 (28 pairs dealt over the ranks), config 5 (streaming queries)."""
    import torch
    import torch.distributed as dist
    from slide_slam_b200 import parallel, synth
    from slide_slam_b200.place_recognition import PlaceRecognition, delaunay
    out = {}

    def ranks_max_sum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            return mx.tolist(), sm.tolist()
        return t.tolist(), t.tolist()

    pr = PlaceRecognition(ROS, device=local_rank)
    # ---- (T): triangle-generated hypothesis set on the config-2 pair (rank 0)
    if rank == 0:
        (ref, qry, truth), _ = workload(2, 0)
        t0 = time.perf_counter()
        ir, iq = delaunay(np.ascontiguousarray(ref[:, 1:3])), delaunay(np.ascontiguousarray(qry[:, 1:3]))
        dl_ms = (time.perf_counter() - t0) * 1e3
        tr = np.ascontiguousarray(ref[:, 1:3][ir].reshape(-1, 6)); tq = np.ascontiguousarray(qry[:, 1:3][iq].reshape(-1, 6))
        lr, lq = np.ascontiguousarray(ref[ir, 0]), np.ascontiguousarray(qry[iq, 0])
        pr.prepare(ref, qry, 400.0, 400.0)
        pr.generate_and_score(tr, tq, 0.1, lr, lq, want_lists=False)      # warm-up
        ms, res, gi = [], None, None
        for _ in range(5):
            t0 = time.perf_counter()
            res, gi, _ = pr.generate_and_score(tr, tq, 0.1, lr, lq, want_lists=False)
            ms.append((time.perf_counter() - t0) * 1e3)
        yaw = float(np.arctan2(res.R_t[3], res.R_t[0])) if res.best_hyp_index >= 0 else None
        out["generator"] = {
            "note": "hypothesis set (T) of SURVEY 8d on the config-2 pair: Delaunay triangles (host), descriptor binning + windowed matching "
                    "+ radix sort + per-match 2-D Kabsch + MatchMaps predicate on the device (slide_pr_generate_and_score), class signature on",
            "triangles": [int(len(tr)), int(len(tq))], "hypotheses": int(gi.n_matches), "best_num_inliers": int(res.best_num_inliers),
            "call_ms": float(np.median(ms)), "delaunay_host_ms": dl_ms, "match_ms": float(gi.match_ms), "kabsch_ms": float(gi.kabsch_ms),
            "score_ms": float(gi.score_ms), "hypotheses_per_s": float(gi.n_matches) / (float(np.median(ms)) * 1e-3),
            "yaw_error_rad": (abs(float(np.angle(np.exp(1j * (yaw - truth["yaw"]))))) if yaw is not None else None)}
    # ---- config 3: ONE 20 000 x 20 000 pair, hypothesis space sharded over the ranks
    (ref3, qry3, _), _ = workload(3, 0)
    rng_ = _ranges(ref3, qry3)   # every rank prepares the same shifted pair (replicated maps)
    s3r, s3q = ref3.copy(), qry3.copy()
    s3r[:, 1:3] -= rng_["centroid_ref"]; s3q[:, 1:3] -= rng_["centroid_qry"]
    hx = rng_["half_x"]
    pr.prepare(s3r, s3q, hx, hx)
    wins = []
    t_ms = []
    for it in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = parallel.sharded_search(pr, rank, world, device=dev)
        win = parallel.allgather_winner(res.best_hyp_index, res.best_num_inliers, device=dev) if world > 1 else \
            parallel.ShardWinner(res.best_num_inliers, res.best_hyp_index, 0)
        torch.cuda.synchronize()
        t_ms.append((time.perf_counter() - t0) * 1e3)
        wins.append((win.inliers, win.hyp_index, float(res.kernel_ms), int(res.hypotheses_scored)))
    (mx, sm) = ranks_max_sum([min(t_ms[1:]), wins[-1][2], float(wins[-1][3])])
    if rank == 0:
        out["config3_shard"] = {
            "note": "one 20000 x 20000 pair, hypothesis space (lattice blocks) sharded round-robin over the ranks, every hypothesis counted "
                    "exactly, all-gather of the 16-byte top-1 records (strong scaling); wall clock, max over ranks, best of 2 after a warm-up",
            "n_gpus": world, "ms_per_pair": mx[0], "hypotheses": int(sm[2]), "hypotheses_per_s": sm[2] / (mx[0] * 1e-3),
            "slowest_rank_kernel_ms": mx[1], "mean_rank_kernel_ms": sm[1] / world,
            "best_num_inliers": int(wins[-1][0]), "best_hyp_index": int(wins[-1][1])}
    # ---- config 4: 28 pairs of 5000 landmarks dealt round-robin over the ranks
    maps = synth.config_robots(8, 5000)
    pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
    mine = pairs[rank::world]
    pr.findTransformationBatch(maps, mine)  # warm-up: the page-locked / device buffers of every map slot exist afterwards (slots are pooled)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = pr.findTransformationBatch(maps, mine)   # every map through the device cache once: one index per reference map
    torch.cuda.synchronize()
    dt4 = time.perf_counter() - t0
    n_found = sum(int(o.found) for o in outs); n_h = sum(int(o.match.hypotheses_scored) for o in outs)
    n_built = sum(int(not (o.match.reuse & 2)) for o in outs)
    (mx, sm) = ranks_max_sum([dt4, float(n_found), float(n_h), float(n_built)])
    if rank == 0:
        out["config4"] = {"note": "8 robots, 28 map pairs of 5000 landmarks through slide_pr_find_transformation_batch (host buffers, device map cache), "
                                  "pairs dealt round-robin over the ranks; second call of the batch (the first one allocates the map slots' buffers)",
                          "n_gpus": world, "pairs": 28, "pairs_per_s": 28 / mx[0], "closures_found": int(sm[1]), "hypotheses_per_s": sm[2] / mx[0],
                          "reference_indexes_built": int(sm[3])}
    # ---- config 5: streaming 300-landmark queries against one 50 000-landmark map (rank 0's GPU: latency)
    if rank == 0:
        nq = 40
        big, queries = synth.config_stream(50000, n_queries=nq, n_sub=300)
        # the accumulated map lives in the device map cache (handed over once, like a robot's map in
        # databaseManager::robotMapDict_); every query is handed over and matched against it
        pr.cache_put(0, 1, big)
        pr.cache_put(1, 0, queries[0])
        pr.findTransformationCached(0, 1)       # builds the map's index (reused by the stream)
        lat, n_found, reuse = [], 0, 0
        for k, q in enumerate(queries):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pr.cache_put(1, k + 1, q)
            f, _, _, inf, _, _ = pr.findTransformationCached(0, 1)
            lat.append((time.perf_counter() - t0) * 1e3)
            n_found += int(f); reuse += int(bool(inf.match.reuse & 2))
        lat = np.array(lat)
        out["config5"] = {"note": f"{nq} submap queries of 300 landmarks against a 50000-landmark map held in the device map cache; per-query latency of "
                                  "handing the query over (slide_pr_map_cache_put) + slide_pr_find_transformation_cached, host buffers",
                          "queries": nq, "p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)),
                          "p99_ms": float(np.percentile(lat, 99)), "closures_found": n_found, "index_reused": reuse,
                          "hypotheses_per_query": int(inf.match.hypotheses_scored)}
    pr.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the generator / config 3 / 4 / 5 extra keys")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
