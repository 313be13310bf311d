"""aadp_batch_optimal_all_compact on the resident C3 batch: wall time of the call and its kernels.  usage: [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
seqs, pq, pt = synth.pair_workload(1003, n, 100, 500)
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
c.fill_batch(res, off, pq, pt, a.W_FWD | a.W_REV | a.W_TB | a.W_MASK, 0.01)
coff, cp, cn, cst = c.optimal_all_compact(a.FWD, n)
pin = torch.zeros((int(coff[-1]) + 1024, 2), dtype=torch.int32).pin_memory()
for k in range(3):
    c.set_profiling(True)
    t0 = time.perf_counter()
    coff, cp, cn, cst = c.optimal_all_compact(a.FWD, n, pairs=pin.numpy(), pairs_cap=len(pin))
    t1 = time.perf_counter()
    print("call %.2f ms, kernels %s, aligned pairs %d" % ((t1 - t0) * 1e3, {k2: round(v, 2) for k2, v, _ in c.profile()}, int(coff[-1])), flush=True)
pinr = lambda x: torch.from_numpy(x.copy()).pin_memory()
keep = [pinr(res), pinr(off), pinr(pq), pinr(pt)]
hb = [k.numpy() for k in keep]
what = a.W_FWD | a.W_REV | a.W_TB | a.W_MASK
c.set_profiling(False)
for label, with_ali in (("fill_batch only", False), ("fill_batch + optimal_all_compact", True), ("fill_batch only", False)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        c.fill_batch(*hb, what, 0.01)
        if with_ali:
            c.optimal_all_compact(a.FWD, n, pairs=pin.numpy(), pairs_cap=len(pin))
    torch.cuda.synchronize()
    print("%s: %.2f ms per step" % (label, (time.perf_counter() - t0) * 1e3 / 5), flush=True)
