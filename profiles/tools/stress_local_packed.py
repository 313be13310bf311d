"""Randomised comparison of the packed LOCAL kernels (LOC = 1) with the int32 kernels: several integer / dyadic scorings,
random and related pairs, every final score of the batch + full dense matrices (scores and predecessors, both directions)
and local optimal alignments for a sample.  usage: python profiles/tools/stress_local_packed.py [pairs] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
bad = 0
for gi, ge, scale in ((12, 1, 1.0), (10.5, 0.5, 1.0), (8, 2, 1.0), (5.5, 0.25, 0.5), (0, 0, 1.0), (20, 0, 2.0)):
    Ms = (M.astype(np.float32) * np.float32(scale)).astype(np.float32)
    seqs, pq, pt = synth.pair_workload(seed * 100 + int(gi * 4), n, 1, 512)
    seqs = list(seqs)
    for p in range(0, n, 4):   # a quarter of the pairs related
        q = seqs[pq[p]]
        t = np.roll(q.copy(), int(rng.integers(0, 5)))
        idx = rng.integers(0, len(t), max(1, len(t) // 4))
        t[idx] = rng.integers(0, 20, len(idx))
        seqs[pt[p]] = t
    res, off = a.Context.pack(seqs)
    cp, ci = a.Context(0), a.Context(0)
    ci.set_option("packed", 0)
    for c in (cp, ci):
        c.set_scoring(Ms, gi, ge, a.LOCAL)
    what = a.W_FWD | a.W_REV | a.W_TB | a.W_SCORES
    cp.set_profiling(True)
    x = cp.fill_batch(res, off, pq, pt, what)
    used = any("LOC=1" in nm for nm, _, _ in cp.profile())
    cp.set_profiling(False)
    y = ci.fill_batch(res, off, pq, pt, what)
    ok = used and np.array_equal(x["fwd_score"], y["fwd_score"]) and np.array_equal(x["rev_score"], y["rev_score"])
    for p in rng.choice(n, 12, replace=False):
        Lq, Lt = len(seqs[pq[p]]), len(seqs[pt[p]])
        u = cp.fetch_pair(int(p), Lq, Lt, fwd=True, rev=True, mask=False)
        v = ci.fetch_pair(int(p), Lq, Lt, fwd=True, rev=True, mask=False)
        ok = ok and all(u[k] is None or np.array_equal(u[k], v[k]) for k in u)
        for d in (a.FWD, a.REV):
            r1, r2 = cp.optimal(int(p), d, Lq, Lt), ci.optimal(int(p), d, Lq, Lt)
            ok = ok and r1[0] == r2[0] and r1[2] == r2[2] and np.array_equal(r1[1], r2[1])
    cp.close(); ci.close()
    print("gi %g ge %g scale %g: packed local kernels used %s, equal %s" % (gi, ge, scale, used, ok), flush=True)
    bad += 0 if ok else 1
print("mismatching scorings:", bad)
sys.exit(1 if bad else 0)
