"""CPU suite: the multi-rank host logic (record all-gather + deterministic merge, pair dealing)
over the gloo backend with world_size 2 and 3 -- the same code path bench.py runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slide_slam_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cases, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = []
        for per_rank in cases:
            h, n = per_rank[rank]
            w = parallel.allgather_winner(h, n)
            got.append((w.inliers, w.hyp_index, w.rank))
        # all-pairs gather: each rank fills only its own pairs
        rows = torch.zeros((7, 2), dtype=torch.float64)
        for p in parallel.pairs_of_rank(7, rank, world):
            rows[p] = torch.tensor([float(p), float(rank)])
        dist.all_reduce(rows)
        out_q.put((rank, got, rows.numpy().tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_allgather_winner_is_deterministic_and_first_wins(world):
    cases = [
        [(50, 7), (20, 9), (10, 9)][:world],            # max inliers wins
        [(50, 9), (20, 9), (10, 9)][:world],            # tie -> smallest canonical index (PR.cpp:361)
        [(-1, -10000), (33, 0), (-1, -10000)][:world],  # empty shards never win, a 0-inlier hypothesis does
        [(-1, -10000)] * world,                         # nothing scored anywhere
    ]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = []
    for per_rank in cases:
        w = parallel.merge_records(per_rank)
        expect.append((-10000, -1, -1) if w < 0 else (per_rank[w][1], per_rank[w][0], w))
    for rank, got, rows in results:
        assert got == expect                              # identical on every rank
        for p in range(7):
            assert rows[p] == [float(p), float(p % world)]  # pair p was handled by rank p % world
    assert expect[0][1] == (20 if world == 2 else 10)
    assert expect[1][1] == (20 if world == 2 else 10)
    assert expect[2] == (0, 33, 1) and expect[3] == (-10000, -1, -1)


class _FakeResult:
    def __init__(self, inliers, index, launches=1, ms=1.0, search_mode=1):
        self.best_num_inliers, self.best_hyp_index, self.gpu_launches, self.kernel_ms = inliers, index, launches, ms
        self.search_mode = search_mode


class _FakeShardedSearch:
    """Stands in for PlaceRecognition.search: every rank owns some (inliers, index) hypotheses; the
    bound phase reports the rank's seed, the verification everything >= the incumbent."""
    def __init__(self, hyps, seed, can_bound=True):
        self.hyps, self.seed, self.calls, self.can_bound = hyps, seed, [], can_bound

    def search(self, shard_index=0, shard_count=1, bounds_only=False, incumbent_inliers=0, reuse_bounds=False, **kw):
        self.calls.append((shard_index, shard_count, bounds_only, incumbent_inliers, reuse_bounds))
        if bounds_only and not self.can_bound:
            # slide_pr_search: a problem without a bound phase is searched exhaustively by the first call
            best = max(self.hyps, key=lambda h: (h[0], -h[1])) if self.hyps else (-10000, -1)
            return _FakeResult(*best, search_mode=0), None
        if bounds_only:
            return _FakeResult(*self.seed), None
        ok = [h for h in self.hyps if h[0] >= incumbent_inliers]
        if not ok:
            return _FakeResult(-10000, -1), None
        best = max(ok, key=lambda h: (h[0], -h[1]))
        return _FakeResult(*best), None


def _sharded_worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0: weak hypotheses and a weak seed; rank 1: the winner, but its seed missed it
        hyps = [[(12, 40), (9, 3)], [(30, 77), (30, 90), (11, 5)]][rank]
        seed = [(9, 3), (11, 5)][rank]
        pr = _FakeShardedSearch(hyps, seed)
        res = parallel.sharded_search(pr, rank, world)
        win = parallel.allgather_winner(res.best_hyp_index, res.best_num_inliers)
        # the same shards when the problem has no bound phase (e.g. disjoint label sets: every count is 0)
        pr0 = _FakeShardedSearch([[(0, 8)], [(0, 2)]][rank], None, can_bound=False)
        res0 = parallel.sharded_search(pr0, rank, world)
        win0 = parallel.allgather_winner(res0.best_hyp_index, res0.best_num_inliers)
        # default engine (pair-join scorer): one exact pass per shard, no incumbent exchange
        prj = _FakeShardedSearch(hyps, seed)
        prj.engine = "join"
        resj = parallel.sharded_search(prj, rank, world)
        winj = parallel.allgather_winner(resj.best_hyp_index, resj.best_num_inliers)
        out_q.put((rank, pr.calls, (res.best_num_inliers, res.best_hyp_index, res.gpu_launches), (win.inliers, win.hyp_index, win.rank),
                   pr0.calls, (win0.inliers, win0.hyp_index, win0.rank), prj.calls, (winj.inliers, winj.hyp_index, winj.rank)))
    finally:
        dist.destroy_process_group()


def test_two_phase_sharded_search_shares_the_incumbent():
    """bound phase -> all-reduce(max) of the seeds' inlier counts -> verification with that incumbent
    and reuse_bounds; the shard without anything >= the incumbent reports 'none'."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, calls, local, win, calls0, win0, callsj, winj in results:
        assert callsj == [(rank, world, False, 0, False)] and winj == (30, 77, 1)   # pair-join scorer: a single sharded search
        assert calls == [(rank, world, True, 0, False), (rank, world, False, 11, True)]  # incumbent = max(9, 11)
        assert win == (30, 77, 1)                     # ties to the smallest canonical index
        assert calls0 == [(rank, world, True, 0, False)]  # exhaustive fallback: no second call
        assert win0 == (0, 2, 1)                      # all counts 0: the smallest canonical index wins
    assert results[0][2] == (12, 40, 2)               # rank 0 still reports its own best (>= incumbent)
    assert results[1][2] == (30, 77, 2)


def test_merge_matches_the_c_abi():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 9))
        recs = [(int(rng.integers(-1, 6)), int(rng.integers(0, 4))) for _ in range(n)]
        assert parallel.merge_records(recs) == parallel.merge_records_native(recs)


def test_pairs_are_dealt_round_robin():
    counts = [len(parallel.pairs_of_rank(28, r, 8)) for r in range(8)]
    assert counts == [4, 4, 4, 4, 3, 3, 3, 3]
    assert sorted(sum((parallel.pairs_of_rank(28, r, 8) for r in range(8)), [])) == list(range(28))
