"""Seeded synthetic protein workloads of BASELINE.json's shapes."""
import numpy as np


def random_seqs(rng, n, lo, hi, alphabet_size=20):
    lens = rng.integers(lo, hi + 1, n)
    return [rng.integers(0, alphabet_size, int(L)).astype(np.uint8) for L in lens]


def pair_workload(seed, npairs, lo=100, hi=500):
    """npairs independent pairs: sequences 2p (query) and 2p+1 (template)."""
    rng = np.random.default_rng(seed)
    seqs = random_seqs(rng, 2 * npairs, lo, hi)
    pq = np.arange(0, 2 * npairs, 2, dtype=np.int32)
    pt = pq + 1
    return seqs, pq, pt


def config(name):
    """Named configs of BASELINE.json (C1..C3 shapes)."""
    if name == "c1":
        return pair_workload(1001, 1, 250, 250)
    if name == "c2":
        return pair_workload(1002, 10_000)
    if name == "c3":
        return pair_workload(1003, 100_000)
    raise KeyError(name)
