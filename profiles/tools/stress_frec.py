"""Randomised comparison of the record-list kernel with the scan kernel (both product kernels; the scan kernel is the one
pinned to the reference by the golden vectors): random scorings, align types, lengths (incl. templates beyond 512 and 1024),
random / related / repetitive sequences; batch scalars for every case, full dense matrices for a sample.
usage: python profiles/tools/stress_frec.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
alpha, M = a.blosum62()
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for case in range(ncase):
    gi = float(rng.choice([4.73, 0.9, 11.3, 2.17, 7.05, 0.0]))
    ge = float(rng.choice([0.34, 0.0, 1.7, 0.61, 0.05]))
    scale = float(rng.choice([1.0, 0.37, 2.9, 0.1]))
    at = int(rng.integers(0, 5))
    Ms = (M.astype(np.float32) * np.float32(scale)).astype(np.float32)
    kind = int(rng.integers(0, 4))
    hi = int(rng.choice([60, 200, 600, 1300, 2040]))
    nseq = 24 if hi > 600 else 60
    seqs = []
    for s in range(nseq):
        L = int(rng.integers(1, hi + 1))
        if kind == 2:   # repetitive: long exact ties
            unit = rng.integers(0, 20, int(rng.integers(1, 4))).astype(np.uint8)
            q = np.resize(unit, L)
        else:
            q = rng.integers(0, 20, L).astype(np.uint8)
        seqs.append(q)
    if kind == 1:       # related: every odd sequence is a mutated copy of its predecessor
        for s in range(1, nseq, 2):
            t = np.roll(seqs[s - 1].copy(), int(rng.integers(0, 4)))
            idx = rng.integers(0, len(t), max(1, len(t) // 5))
            t[idx] = rng.integers(0, 20, len(idx))
            seqs[s] = t
    npairs = 40 if hi > 600 else 200
    pq = rng.integers(0, nseq, npairs).astype(np.int32)
    pt = rng.integers(0, nseq, npairs).astype(np.int32)
    if kind == 1:
        pq[::2] = (pq[::2] // 2) * 2
        pt[::2] = pq[::2] + 1
    res, off = a.Context.pack(seqs)
    cr, cs = a.Context(0), a.Context(0)
    for c, rec in ((cr, 1), (cs, 0)):
        c.set_option("general_records", rec)
        c.set_option("exact_float", 1)
        c.set_scoring(Ms, gi, ge, at)
    what = a.W_FWD | a.W_REV | (0 if at == a.LOCAL else a.W_MASK)
    x, y = cr.fill_batch(res, off, pq, pt, what, 0.02), cs.fill_batch(res, off, pq, pt, what, 0.02)
    ok = all(x[k] is None or np.array_equal(x[k], y[k]) for k in x)
    for p in rng.choice(npairs, 3, replace=False):
        q, t = seqs[pq[p]], seqs[pt[p]]
        if len(q) * len(t) > 400000:
            continue
        u, v = cr.fill_pair(q, t, a.BOTH, delta_ratio=-1.0), cs.fill_pair(q, t, a.BOTH, delta_ratio=-1.0)
        ok = ok and all(u[k] is None or np.array_equal(u[k], v[k]) for k in u)
    cr.close(); cs.close()
    if not ok:
        bad += 1
        print("MISMATCH case", case, dict(gi=gi, ge=ge, scale=scale, at=at, kind=kind, hi=hi), flush=True)
print("cases %d, mismatches %d" % (ncase, bad))
sys.exit(1 if bad else 0)
