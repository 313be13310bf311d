"""LOCAL alignments, C3-shaped batch, fwd+rev + traceback + scores: packed (LOC = 1) kernels against the int32 kernels.
usage: python profiles/tools/time_local.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
seqs, pq, pt = synth.pair_workload(1003, n, 100, 500)
res, off = a.Context.pack(seqs)
cells = sum(len(seqs[pq[p]]) * len(seqs[pt[p]]) for p in range(n))
what = a.W_FWD | a.W_REV | a.W_TB | a.W_SCORES
for packed in (1, 0):
    c = a.Context(0)
    c.set_option("packed", packed)
    c.set_scoring(M, 12, 1, a.LOCAL)
    c.fill_batch(res, off, pq, pt, what)
    c.set_profiling(True)
    c.fill_batch(res, off, pq, pt, what)
    by = {}
    for name, ms, cu in c.profile():
        by[name] = by.get(name, 0) + ms
    tot = sum(by.values())
    print("packed %d: kernels %.2f ms = %.0f GCUPS  %s" % (packed, tot, 2 * cells / tot / 1e6, {k: round(v, 2) for k, v in by.items()}), flush=True)
    c.close()
