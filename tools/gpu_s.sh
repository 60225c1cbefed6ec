#!/bin/bash
set -u
mkdir -p gpurun_out
cat > /tmp/intra.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import spr_helpers as H
from slide_slam_b200.place_recognition import PlaceRecognition
maps, cases = H.golden_maps(), H.golden_cases()
ci = cases["prtest_intra_lsq1"]
pr = PlaceRecognition(H.rosparams_from_golden(ci["params"])); pr.inter_loop_closure = False
meas, sub = maps[ci["qry"]], maps[ci["ref"]]
K = 16
rng = np.random.default_rng(0)
subs = []
for k in range(K):
    t = sub.copy(); t[:, 1:3] += rng.normal(0, 0.02, (len(sub), 2)); subs.append(t)
poses = [np.eye(4)] * K
for rep in range(3):
    t0 = time.perf_counter()
    for s in subs: pr.findIntraLoopClosure(meas, s, np.eye(4), np.eye(4))
    t1 = time.perf_counter()
    out = pr.findIntraLoopClosureBatch(meas, subs, np.eye(4), poses)
    t2 = time.perf_counter()
    print(f"{K} candidates: single calls {(t1-t0)*1e3:.2f} ms, batch {(t2-t1)*1e3:.2f} ms, found {sum(o[0] for o in out)}", flush=True)
PY
timeout 120 python /tmp/intra.py > gpurun_out/s_intra.log 2>&1
SLIDE_PR_TRACE=1 timeout 120 python /tmp/intra.py 2>&1 | tail -30 > gpurun_out/s_intra_trace.log
