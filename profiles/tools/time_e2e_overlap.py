"""End-to-end C3 steps (aadp_fill_batch from pinned host buffers) issued from T host threads over T contexts.
usage: python profiles/tools/time_e2e_overlap.py [pairs]"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
seqs, pq, pt = synth.pair_workload(1003, n, 100, 500)
res, off = a.Context.pack(seqs)
pin = lambda x: torch.from_numpy(x.copy()).pin_memory()
keep = [pin(res), pin(off), pin(pq), pin(pt)]
hb = [k.numpy() for k in keep]
what = a.W_FWD | a.W_REV | a.W_TB | a.W_MASK
for T in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,2,3".split(","))]:
    ctxs, streams = [], []
    shared = torch.cuda.Stream() if os.environ.get('SHARED_STREAM') else None
    for k in range(T):
        c = a.Context(0)
        s = shared or torch.cuda.Stream()
        c.set_stream(s.cuda_stream)
        c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
        c.fill_batch(*hb, what, 0.01)
        c.fill_batch(*hb, what, 0.01)
        ctxs.append(c); streams.append(s)
    steps = 12
    def worker(c, k):
        for _ in range(k):
            c.fill_batch(*hb, what, 0.01)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(ctxs[k], steps // T)) for k in range(T)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("threads %d: %.3f ms per step" % (T, (t1 - t0) * 1e3 / (steps // T * T)), flush=True)
    for c in ctxs:
        c.close()
