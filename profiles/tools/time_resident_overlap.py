"""Resident C3 steps (aadp_run_batch, inputs uploaded once) issued from T host threads over T contexts: separates the
cost of running several contexts' kernels concurrently from the cost of the per-step uploads of the end-to-end path.
usage: python profiles/tools/time_resident_overlap.py [pairs] [T,T,...]"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
seqs, pq, pt = synth.pair_workload(1003, n, 100, 500)
res, off = a.Context.pack(seqs)
what = a.W_FWD | a.W_REV | a.W_TB | a.W_MASK
for T in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "3"])]:
    ctxs = []
    for k in range(T):
        c = a.Context(0)
        s = torch.cuda.Stream()
        c.set_stream(s.cuda_stream)
        c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
        c.upload_batch(res, off, pq, pt, what)
        d = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3)] + [torch.empty(n, dtype=torch.int64, device="cuda")]
        c.run_batch(what, 0.01, *[x.data_ptr() for x in d])
        ctxs.append((c, s, d))
    steps = 12
    def worker(c, s, d, k):
        for _ in range(k):
            c.run_batch(what, 0.01, *[x.data_ptr() for x in d])
        s.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(*ctxs[k], steps // T)) for k in range(T)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("resident, threads %d: %.3f ms per step" % (T, (t1 - t0) * 1e3 / (steps // T * T)), flush=True)
    for c, s, d in ctxs:
        c.close()
