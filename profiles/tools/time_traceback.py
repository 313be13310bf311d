import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np, torch
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
seqs, pq, pt = synth.pair_workload(1003, 100000, 100, 500)
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
c.fill_batch(res, off, pq, pt, a.W_FWD | a.W_REV | a.W_TB | a.W_MASK, 0.01)
c.set_profiling(True)
for d in (a.FWD, a.REV):
    t0 = time.time()
    aoff, pairs, n, st = c.optimal_all(d, len(pq))
    t1 = time.time()
    print("dir", d, "wall %.1f ms" % ((t1 - t0) * 1e3), "mean len", n.mean(), "bad", int((st != 0).sum()))
print(c.profile())
