"""Times the exact general-gap fp32 path (reference default scoring 4.73 / 0.34) on a C3-shaped sample.
usage: time_general.py [pairs] [general_records 0|1] [related 0|1] [general_budget_mcells]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
related = int(sys.argv[3]) if len(sys.argv) > 3 else 0
seqs, pq, pt = synth.pair_workload(1003, n, 100, 500)
if related:  # every template becomes a mutated copy of its query (25 % substitutions, one shift)
    rng = np.random.default_rng(7)
    seqs = list(seqs)
    for p in range(n):
        q = seqs[pq[p]]
        t = np.roll(q.copy(), 3)
        idx = rng.integers(0, len(t), len(t) // 4)
        t[idx] = rng.integers(0, 20, len(idx))
        seqs[pt[p]] = t
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_option("general_records", rec)
if len(sys.argv) > 4:
    c.set_option("general_budget_mcells", int(sys.argv[4]))
c.set_scoring(M, 4.73, 0.34, a.SEMI_LOCAL)
what = a.W_FWD | a.W_REV | a.W_MASK
c.fill_batch(res, off, pq, pt, what, 0.01)
c.set_profiling(True)
t0 = time.time()
out = c.fill_batch(res, off, pq, pt, what, 0.01)
t1 = time.time()
cells = sum(len(seqs[pq[p]]) * len(seqs[pt[p]]) for p in range(n))
print("records %d related %d pairs %d  wall %.1f ms  %.0f pairs/s  %.2f GCUPS (fwd+rev cells / s)" % (rec, related, n, (t1 - t0) * 1e3, n / (t1 - t0), 2 * cells / (t1 - t0) / 1e9))
prof = c.profile()
by = {}
for name, ms, cu in prof:
    by[name] = by.get(name, 0) + ms
print({k: round(v, 2) for k, v in by.items()})
print("checksum", float(np.sum(out["fwd_score"].astype(np.float64))), float(np.sum(out["rev_score"].astype(np.float64))), int(np.sum(out["nearopt_count"])))
