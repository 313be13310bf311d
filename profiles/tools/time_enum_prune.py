"""UCW enumeration of a batch (aadp_batch_near_optimal) with and without the mask-pruned deletion scans: kernel time and
equality of every alignment.  usage: python profiles/tools/time_enum_prune.py [npairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(7)
seqs, pq, pt = [], [], []
for k in range(n):
    L = int(rng.integers(100, 501))
    s = rng.integers(0, 20, L).astype(np.uint8)
    m = s.copy()
    m[::7] = rng.integers(0, 20, len(m[::7]))
    cut = int(rng.integers(10, L - 10))
    m = np.concatenate([m[:cut], m[cut + int(rng.integers(0, 4)):]])
    seqs += [s, m]
    pq.append(2 * k)
    pt.append(2 * k + 1)
pq, pt = np.array(pq, np.int32), np.array(pt, np.int32)
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
what = a.W_FWD | a.W_REV | a.W_TB | a.W_MASK
delta, K = 0.01, 64
c.fill_batch(res, off, pq, pt, what, delta)
ids = np.arange(n, dtype=np.int64)
res_by = {}
for prune in (1, 0, 1):
    c.set_option("enum_mask_prune", prune)
    c.near_optimal(ids[:64], delta, K)
    c.set_profiling(True)
    got = c.near_optimal(ids, delta, K)
    kms = sum(ms for name, ms, _ in c.profile() if name.startswith("ucw"))
    c.set_profiling(False)
    res_by[prune] = got
    print("prune %d: kernel %.1f ms, %d alignments" % (prune, kms, sum(len(g[2]) for g in got)), flush=True)
same = all(x[0] == y[0] and x[1] == y[1] and len(x[2]) == len(y[2]) and all(s1 == s2 and np.array_equal(p1, p2) for (s1, p1), (s2, p2) in zip(x[2], y[2]))
           for x, y in zip(res_by[1], res_by[0]))
print("identical alignments, scores, order:", same)
