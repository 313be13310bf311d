"""Sharding of independent pairs over ranks: no data-path collective, only a final
gather of per-pair scalars.  Pure host logic (numpy / torch.distributed); the fill itself is the C ABI."""
import numpy as np


def shard_pairs(cells, world):
    """Longest-processing-time dealing: pairs sorted by cell count, each given to the least loaded
    rank.  Returns a list of `world` index arrays (ascending pair ids) covering every pair once."""
    cells = np.asarray(cells, dtype=np.int64)
    order = np.argsort(-cells, kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.empty(len(cells), np.int32)
    # dealing in blocks of `world` (snake order) is within one pair of optimal LPT and vectorises
    for start in range(0, len(order), world):
        blk = order[start:start + world]
        ranks = np.argsort(load, kind="stable")[: len(blk)]
        owner[blk] = ranks
        load[ranks] += cells[blk]
    return [np.sort(np.nonzero(owner == r)[0]) for r in range(world)]


def gather_scores(local_values, local_idx, npairs, group=None):
    """Host gather: every rank contributes the scalars of its shard; returns the full array on all ranks."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    objs = [None] * world
    dist.all_gather_object(objs, (np.asarray(local_idx), np.asarray(local_values)), group=group)
    out = np.zeros(npairs, dtype=np.asarray(local_values).dtype)
    seen = np.zeros(npairs, bool)
    for idx, val in objs:
        assert not seen[idx].any(), "a pair was computed by two ranks"
        out[idx] = val
        seen[idx] = True
    assert seen.all(), "a pair was not computed by any rank"
    return out
