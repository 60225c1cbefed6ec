"""Multi-GPU layer of the place-recognition search: one process per GPU (torch.distributed).

The search shards with no data-path collective:
  * one map pair, hypothesis space sharded  -- every rank scores its share of the lattice work
    items (`slide_pr_search` with shard_index / shard_count), then ONE small all-gather of the
    per-rank top-1 record (canonical hypothesis index, inlier count: 16 bytes) and the same
    deterministic merge on every rank: max inliers, ties to the smallest canonical index, i.e. the
    reference's strict '>' first-wins rule (place_recognition.cpp:361).  Every rank then extracts
    the winner's correspondences locally (both maps are replicated).
  * many map pairs (N-robot all-pairs matching) -- pairs are dealt round-robin to the ranks, each
    runs complete searches, results are all-gathered.
NCCL is used when the tensors live on the GPU; the gloo backend (CPU tensors) runs the same code
in the CPU tests.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import capi


@dataclass
class ShardWinner:
    inliers: int
    hyp_index: int
    rank: int


def merge_records(records) -> int:
    """Index of the winning (hyp_index, inliers) record: max inliers, ties to the smallest canonical
    index; records with hyp_index < 0 (empty shards) never win.  Mirrors slide_pr_merge_records."""
    best = -1
    for i, (h, n) in enumerate(records):
        if h < 0:
            continue
        if best < 0 or n > records[best][1] or (n == records[best][1] and h < records[best][0]):
            best = i
    return best


def allgather_winner(local_hyp_index: int, local_inliers: int, device=None, group=None) -> ShardWinner:
    """All-gather of the per-rank top-1 record and deterministic merge (identical on every rank)."""
    world = dist.get_world_size(group)
    rec = torch.tensor([int(local_hyp_index), int(local_inliers)], dtype=torch.int64, device=device)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    recs = [(int(t[0]), int(t[1])) for t in torch.stack(out).cpu()]
    w = merge_records(recs)
    if w < 0:
        return ShardWinner(-10000, -1, -1)  # PR.cpp:125: nothing was scored anywhere
    return ShardWinner(recs[w][1], recs[w][0], w)


def sharded_search(pr, rank: int, world: int, device=None, group=None):
    """One shard of a prepared search.  Pair-join scorer (default engine): one pass.  Lattice kernels: two phases: the bound phase of every shard (with its
    exactly scored seed hypotheses), an all-reduce(max) of the seeds' inlier counts -- the one
    exchange branch-and-bound needs: the incumbent -- and the verification of the shard's
    hypotheses whose bound reaches that incumbent.  Shards that cannot hold the winner verify
    (almost) nothing.  A shard without anything >= the incumbent reports best_hyp_index = -1."""
    if world <= 1:
        return pr.search()[0]
    if getattr(pr, "engine", "lattice") == "join":
        # the pair-join scorer counts every hypothesis of the shard exactly in one pass (lattice blocks dealt
        # round-robin): no bounds, no incumbent, only the final all-gather of the top-1 records
        return pr.search(shard_index=rank, shard_count=world)[0]
    seed, _ = pr.search(shard_index=rank, shard_count=world, bounds_only=True)
    inc = torch.tensor([max(int(seed.best_num_inliers), 0)], dtype=torch.int64, device=device)
    dist.all_reduce(inc, op=dist.ReduceOp.MAX, group=group)
    if seed.search_mode == 0:
        # no bound phase for this problem (no query label occurs in the reference map, >= 65536 query
        # landmarks, a binding compute budget): the first call already searched the shard exhaustively.
        # The same on every rank (the problem is replicated), so the collective above stays matched.
        return seed
    res, _ = pr.search(shard_index=rank, shard_count=world, incumbent_inliers=int(inc.item()), reuse_bounds=True)
    res.gpu_launches += seed.gpu_launches
    res.kernel_ms += seed.kernel_ms
    return res


def sharded_match_maps(pr, reference_objects, query_objects, half_x: float, half_y: float, device=None, group=None):
    """MatchMaps (PR.cpp:98-387) with the hypothesis space sharded over the ranks of `group`.
    Every rank must call it with the same maps.  Returns (winner, ref_idx, qry_idx, R_t, local_result)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pr.prepare(reference_objects, query_objects, half_x, half_y)
    res = sharded_search(pr, rank, world, device=device, group=group)
    win = allgather_winner(res.best_hyp_index, res.best_num_inliers, device=device, group=group)
    if win.hyp_index < 0:
        return win, np.zeros(0, np.int32), np.zeros(0, np.int32), np.eye(3), res
    ext, ri, qi = pr.extract(win.hyp_index)
    if len(ri) != win.inliers:
        raise capi.SlidePrError(capi.ERR_INTERNAL, "sharded winner: brute-force recount disagrees with the indexed count")
    return win, ri, qi, np.array(ext.R_t[:]).reshape(3, 3), res


def pairs_of_rank(n_pairs: int, rank: int, world: int):
    """Round-robin deal of map pairs to ranks (28 pairs on 8 GPUs -> 4,4,4,4,3,3,3,3)."""
    return list(range(rank, n_pairs, world))


def all_pairs_find_transformation(pr, maps, pairs, device=None, group=None):
    """findTransformation for every (ref, qry) index pair, pairs dealt round-robin to the ranks;
    the per-pair [found, inliers, x, y, z, yaw] rows are all-gathered so every rank holds all."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rows = torch.zeros((len(pairs), 6), dtype=torch.float64, device=device)
    for p in pairs_of_rank(len(pairs), rank, world):
        r, q = pairs[p]
        found, xyz_yaw, _tf, info, _ri, _qi = pr.findTransformation(maps[r], maps[q])
        rows[p] = torch.tensor([float(found), float(info.best_num_inliers), *xyz_yaw], dtype=torch.float64)
    # every row is written by exactly one rank, the others hold zeros: a sum is a gather
    dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=group)
    return rows.cpu().numpy()


def records_to_ctypes(records):
    arr = (capi.TopkRecord * len(records))()
    for i, (h, n) in enumerate(records):
        arr[i].hyp_index, arr[i].inliers, arr[i].rank = int(h), int(n), i
    return arr


def merge_records_native(records) -> int:
    """The same merge through the C-ABI (slide_pr_merge_records) -- used by tests to pin both."""
    arr = records_to_ctypes(records)
    return int(capi.lib().slide_pr_merge_records(arr, len(records)))
