"""Times the multi-CTA wavefront kernel on one 30k x 30k pair with different products switched on:
score only / + packed traceback / + score matrices and mask.  Shows how much of a row step is stores."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
alpha, M = a.blosum62()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
rng = np.random.default_rng(1005)
seqs = [rng.integers(0, 20, L).astype(np.uint8) for _ in range(2)]
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
for name, what in (("score only", a.W_FWD | a.W_REV), ("+traceback", a.W_FWD | a.W_REV | a.W_TB),
                   ("+traceback+scores+mask", a.W_FWD | a.W_REV | a.W_TB | a.W_MASK)):
    for it in range(3):
        c.set_profiling(it == 2)
        out = c.fill_batch(res, off, [0], [1], what, 0.01)
    print("%-26s" % name, ["%s %.2f ms" % (n, ms) for n, ms, _ in c.profile() if "wave" in n or "mask" in n],
          "optimum", out["fwd_score"][0], out["rev_score"][0])
