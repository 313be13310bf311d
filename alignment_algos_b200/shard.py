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


def triangle_rects(nseq, block=1000, sub=250):
    """Rectangles (q0, q1, t0, t1) of sequence-id ranges whose union covers every unordered pair i < j of
    an all-vs-all job (query = i, template = j).  Off-diagonal blocks are full rectangles; a diagonal
    block is split once more into `sub`-sized pieces whose diagonal squares are computed in full (their
    lower halves are the only wasted work: about sub / (2 nseq) of the job)."""
    rects = []
    edges = list(range(0, nseq, block)) + [nseq]
    for a in range(len(edges) - 1):
        for b in range(a + 1, len(edges) - 1):
            rects.append((edges[a], edges[a + 1], edges[b], edges[b + 1]))
        se = list(range(edges[a], edges[a + 1], sub)) + [edges[a + 1]]
        for x in range(len(se) - 1):
            for y in range(x, len(se) - 1):
                rects.append((se[x], se[x + 1], se[y], se[y + 1]))
    return rects


def choose_block(nseq, world):
    """Block edge for triangle_rects: large enough that a rectangle is a full launch (>= 250 sequences per
    side), small enough that every rank gets a few dozen rectangles to balance (about 6*world blocks per side)."""
    return int(min(1000, max(250, nseq // (6 * max(world, 1)) or 1)))


def rect_is_diagonal(r):
    return r[0] == r[2] and r[1] == r[3]


def shard_rects(rects, lens, world):
    """Deal the rectangles of triangle_rects over ranks, balanced by cell count (LPT).  No collective on
    the data path: every rank scores its own rectangles; the host gathers the score blocks."""
    cs = np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.int64))])
    cells = [(cs[r[1]] - cs[r[0]]) * (cs[r[3]] - cs[r[2]]) for r in rects]
    parts = shard_pairs(cells, world)
    return [[rects[i] for i in p] for p in parts]
