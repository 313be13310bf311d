"""Multi-rank host logic on CPU (gloo, world_size 2): pairs shard with no overlap, balanced by cell
count, and the host gather reassembles per-pair results.  Each rank fills its shard with the oracle
(the GPU fill is covered by -m gpu tests); the gathered result must equal a single-rank run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import po


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alignment_algos_b200 import synth, blosum62
    from alignment_algos_b200.shard import shard_pairs, gather_scores
    seqs, pq, pt = synth.pair_workload(5, 60, 10, 60)
    cells = np.array([len(seqs[a]) * len(seqs[b]) for a, b in zip(pq, pt)])
    shards = shard_pairs(cells, world)
    mine = shards[rank]
    _, M = blosum62()
    O = po.Oracle(M, 12, 1, po.SEMI_LOCAL)
    local = np.array([O.fill(seqs[pq[p]], seqs[pt[p]], po.FWD, fast=True)[0][-1, -1] for p in mine], np.float32)
    full = gather_scores(local, mine, len(pq))
    loads = [int(cells[s].sum()) for s in shards]
    if rank == 0:
        ret["full"] = full
        ret["loads"] = loads
        ret["sizes"] = [len(s) for s in shards]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_host_gather():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    from alignment_algos_b200 import synth, blosum62
    seqs, pq, pt = synth.pair_workload(5, 60, 10, 60)
    _, M = blosum62()
    O = po.Oracle(M, 12, 1, po.SEMI_LOCAL)
    want = np.array([O.fill(seqs[a], seqs[b], po.FWD, fast=True)[0][-1, -1] for a, b in zip(pq, pt)], np.float32)
    assert np.array_equal(ret["full"], want)
    assert sum(ret["sizes"]) == len(pq)
    loads = ret["loads"]
    assert abs(loads[0] - loads[1]) <= 0.05 * sum(loads)


def test_shard_pairs_properties():
    from alignment_algos_b200.shard import shard_pairs
    rng = np.random.default_rng(0)
    cells = rng.integers(1, 250000, 1000)
    for world in (1, 2, 4, 8):
        shards = shard_pairs(cells, world)
        allidx = np.concatenate(shards)
        assert len(allidx) == 1000 and len(set(allidx.tolist())) == 1000
        loads = np.array([cells[s].sum() for s in shards], float)
        assert loads.max() / loads.mean() < 1.02


def _rect_worker(rank, world, port, ret):
    # all-vs-all (C4) host logic: every rank scores its own rectangles (here with the oracle), rank 0 gathers the
    # score blocks into the upper triangle -- no collective on the data path, one host gather at the end
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alignment_algos_b200 import synth, blosum62
    from alignment_algos_b200.shard import triangle_rects, shard_rects
    rng = np.random.default_rng(9)
    seqs = synth.random_seqs(rng, 23, 5, 30)
    lens = [len(s) for s in seqs]
    mine = shard_rects(triangle_rects(len(seqs), block=8, sub=4), lens, world)[rank]
    _, M = blosum62()
    O = po.Oracle(M, 12, 1, po.SEMI_LOCAL)
    blocks = []
    for (q0, q1, t0, t1) in mine:
        blk = np.array([[O.fill(seqs[i], seqs[j], po.FWD, fast=True)[0][-1, -1] for j in range(t0, t1)]
                        for i in range(q0, q1)], np.float32)
        blocks.append(((q0, q1, t0, t1), blk))
    gathered = [None] * world
    dist.all_gather_object(gathered, blocks)
    if rank == 0:
        n = len(seqs)
        full = np.full((n, n), np.nan, np.float32)
        seen = np.zeros((n, n), np.int32)
        for part in gathered:
            for (q0, q1, t0, t1), blk in part:
                full[q0:q1, t0:t1] = blk
                seen[q0:q1, t0:t1] += 1
        ret["full"] = full
        ret["seen"] = seen
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_all_vs_all_rectangles_and_gather():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rect_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    from alignment_algos_b200 import synth, blosum62
    rng = np.random.default_rng(9)
    seqs = synth.random_seqs(rng, 23, 5, 30)
    _, M = blosum62()
    O = po.Oracle(M, 12, 1, po.SEMI_LOCAL)
    iu = np.triu_indices(len(seqs), 1)
    assert (ret["seen"][iu] == 1).all(), "every pair i<j must be scored exactly once"
    for i, j in zip(*iu):
        assert ret["full"][i, j] == O.fill(seqs[i], seqs[j], po.FWD, fast=True)[0][-1, -1]


def test_triangle_rects_properties():
    from alignment_algos_b200.shard import triangle_rects, shard_rects
    rng = np.random.default_rng(1)
    for n, block, sub in [(1, 1000, 250), (7, 4, 2), (2300, 1000, 250), (5000, 1000, 250)]:
        rects = triangle_rects(n, block, sub)
        cov = np.zeros((n, n), np.int16)
        for q0, q1, t0, t1 in rects:
            assert 0 <= q0 < q1 <= n and 0 <= t0 < t1 <= n
            cov[q0:q1, t0:t1] += 1
        iu = np.triu_indices(n, 1)
        assert (cov[iu] == 1).all()
        assert cov.max() <= 1 if n > 1 else True
        if n >= 2000:  # wasted work = lower halves of the diagonal sub-squares only
            assert cov.sum() <= 1.12 * len(iu[0])
        lens = rng.integers(100, 501, n)
        for world in (2, 4, 8):
            parts = shard_rects(rects, lens, world)
            assert sum(len(p) for p in parts) == len(rects)
    # balance with the block edge chosen for the world size (what bench.py --workload c4 does)
    from alignment_algos_b200.shard import choose_block
    n = 20000
    lens = rng.integers(100, 501, n)
    cs = np.concatenate([[0], np.cumsum(lens)])
    for world in (1, 2, 4, 8):
        blk = choose_block(n, world)
        parts = shard_rects(triangle_rects(n, blk, max(blk // 4, 1)), lens, world)
        loads = np.array([sum((cs[r[1]] - cs[r[0]]) * (cs[r[3]] - cs[r[2]]) for r in p) for p in parts], float)
        assert loads.max() / loads.mean() < 1.03, (world, blk, loads)
