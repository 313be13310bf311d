import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import alignment_algos_b200 as a
from alignment_algos_b200 import synth
alpha, M = a.blosum62()
seqs, pq, pt = synth.pair_workload(1003, 100000, 100, 500)
res, off = a.Context.pack(seqs)
pin = lambda x: torch.from_numpy(x.copy()).pin_memory()
keep = [pin(res), pin(off), pin(pq), pin(pt)]
hb = [k.numpy() for k in keep]
what = a.W_FWD | a.W_REV | a.W_TB | a.W_MASK
T = 2
ctxs = []
for k in range(T):
    c = a.Context(0); s = torch.cuda.Stream(); c.set_stream(s.cuda_stream); c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
    c.fill_batch(*hb, what, 0.01); c.fill_batch(*hb, what, 0.01); ctxs.append((c, s))
def worker(c, k):
    for _ in range(k): c.fill_batch(*hb, what, 0.01)
torch.cuda.synchronize()
th = [threading.Thread(target=worker, args=(ctxs[k][0], 5)) for k in range(T)]
[t.start() for t in th]; [t.join() for t in th]
