"""Per-shard cost of the sharded config-3 search on ONE GPU: runs shard r of 8 alone (bound phase, shared incumbent,
verification) and prints its kernel time next to 1/8 of the unsharded search."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from slide_slam_b200.place_recognition import PlaceRecognition  # noqa: E402
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
(ref, qry, _), _ = bench.workload(cfg, 0)
r = bench._ranges(ref, qry)
sref, sqry = ref.copy(), qry.copy()
sref[:, 1:3] -= r["centroid_ref"]; sqry[:, 1:3] -= r["centroid_qry"]
pr = PlaceRecognition(bench.ROS)
pr.prepare(sref, sqry, r["half_x"], r["half_y"])
full, _ = pr.search(); full, _ = pr.search()
print(f"unsharded kernel_ms={full.kernel_ms:.2f} best={full.best_num_inliers}")
seeds = []
for k in range(n):
    s, _ = pr.search(shard_index=k, shard_count=n, bounds_only=True)
    seeds.append((s.best_num_inliers, s.kernel_ms))
inc = max(s[0] for s in seeds)
for k in range(n):
    s, _ = pr.search(shard_index=k, shard_count=n, bounds_only=True)
    v, _ = pr.search(shard_index=k, shard_count=n, incumbent_inliers=inc, reuse_bounds=True)
    print(f"shard {k}/{n}: bound_ms={s.kernel_ms:.2f} verify_ms={v.kernel_ms:.2f} total={s.kernel_ms+v.kernel_ms:.2f} (1/{n} of unsharded = {full.kernel_ms/n:.2f}) seed={s.best_num_inliers} best={v.best_num_inliers} launches={s.gpu_launches}+{v.gpu_launches}")
