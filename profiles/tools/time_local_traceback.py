"""Local optimal alignments of a whole batch (find_max + enumerate_local on the GPU).  usage: time_local_traceback.py [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
seqs, pq, pt = synth.pair_workload(1003, n, 100, 500)
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_scoring(M, 12, 1, a.LOCAL)
c.fill_batch(res, off, pq, pt, a.W_FWD | a.W_REV | a.W_TB | a.W_SCORES)
c.optimal_all_compact(a.FWD, n)
c.set_profiling(True)
t0 = time.time()
coff, cp, cn, cst = c.optimal_all_compact(a.FWD, n)
t1 = time.time()
print("local traceback of %d pairs: call %.1f ms, kernels %s, %d aligned pairs, checksum %d" %
      (n, (t1 - t0) * 1e3, {k: round(v, 2) for k, v, _ in c.profile()}, int(coff[-1]), int(cp[:coff[-1]].astype(np.int64).sum())))
