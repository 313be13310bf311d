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
