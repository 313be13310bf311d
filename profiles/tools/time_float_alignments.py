"""Optimal alignments of a whole batch under the reference's default penalties (exact-float mode):
aadp_fill_batch(W_FWD) + aadp_batch_optimal_all_compact.  usage: time_float_alignments.py [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seqs, pq, pt = synth.pair_workload(1003, n, 100, 500)
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_scoring(M, 4.73, 0.34, a.SEMI_LOCAL)
c.fill_batch(res, off, pq, pt, a.W_FWD)
c.optimal_all_compact(a.FWD, n)
t0 = time.time()
out = c.fill_batch(res, off, pq, pt, a.W_FWD)
t1 = time.time()
coff, cpairs, cn, cst = c.optimal_all_compact(a.FWD, n)
t2 = time.time()
print("pairs %d: forward scores %.1f ms (%.0f pairs/s), every optimal alignment %.1f ms (%.0f pairs/s), %d aligned pairs, all ok %s"
      % (n, (t1 - t0) * 1e3, n / (t1 - t0), (t2 - t1) * 1e3, n / (t2 - t1), int(coff[-1]), bool((cst == 0).all())))
