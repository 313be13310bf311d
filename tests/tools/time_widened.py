"""Times the rows of SURVEY.md §8(f) built in the third session, next to the reference's CPU code where it travels
(oracle/_ref/libaadp_ref.so):
  f1  aadp_batch_near_optimal      UCW enumeration of a whole batch (one warp per pair over the resident scores)
  f4  aadp_fill_subpair_batch      loop-closure rectangles + Optimal_Subali in one call
  f3  aadp_fill_pair_tabulated     position-dependent gap penalties (HMAP-shaped tables)
usage: python tests/tools/time_widened.py [npairs]   -> one JSON line
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import alignment_algos_b200 as a
from oracle import pyoracle as po   # test infrastructure: the reference side of the comparison only

alpha, M = a.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(7)
out = {}

# ---- f1: related pairs (a sequence against a mutated copy) so that there is something to enumerate
seqs, pq, pt = [], [], []
for k in range(n):
    L = int(rng.integers(100, 501))
    s = rng.integers(0, 20, L).astype(np.uint8)
    m = s.copy()
    m[::7] = rng.integers(0, 20, len(m[::7]))
    cut = int(rng.integers(10, L - 10))
    m = np.concatenate([m[:cut], m[cut + int(rng.integers(0, 4)):]])
    seqs += [s, m]
    pq.append(2 * k)
    pt.append(2 * k + 1)
pq, pt = np.array(pq, np.int32), np.array(pt, np.int32)
res, off = a.Context.pack(seqs)
c = a.Context(0)
c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
what = a.W_FWD | a.W_REV | a.W_TB | a.W_MASK
delta, K = 0.01, 64
c.fill_batch(res, off, pq, pt, what, delta)
ids = np.arange(n, dtype=np.int64)
c.near_optimal(ids[:64], delta, K)
c.set_profiling(True)
t0 = time.time()
got = c.near_optimal(ids, delta, K)
t1 = time.time()
kms = sum(ms for name, ms, _ in c.profile() if name.startswith("ucw"))
c.set_profiling(False)
nali = sum(len(g[2]) for g in got)
over = sum(g[0] == 1 for g in got)
out["f1_ucw"] = {"pairs": n, "delta_ratio": delta, "budget": K, "alignments": nali, "pairs_over_budget": int(over),
                 "kernel_ms": round(kms, 3), "call_ms_incl_d2h_and_python": round((t1 - t0) * 1e3, 1),
                 "pairs_per_s_kernel": round(n / (kms / 1e3))}
if os.path.exists(po.LIB_REF):
    R = po.Reference(alpha, M, 12, 1, a.SEMI_LOCAL)
    ns = 24
    t0 = time.time()
    for p in range(ns):
        R.fill(seqs[pq[p]], seqs[pt[p]], po.FWD)
    tf = time.time() - t0
    t0 = time.time()
    ref_n = 0
    for p in range(ns):
        ref_n += len(R.ucw_alignments(seqs[pq[p]], seqs[pt[p]], delta, 100000))
    te = time.time() - t0 - tf   # ucw_alignments = fill + enumerate
    out["f1_ucw"]["reference_cpu"] = {"pairs": ns, "alignments": ref_n, "enumerate_s_per_pair_1core": round(max(te, 0) / ns, 4),
                                      "fill_s_per_pair_1core": round(tf / ns, 4)}
    gpu_n = sum(len(got[p][2]) for p in range(ns))
    out["f1_ucw"]["same_alignment_count_on_sample"] = bool(gpu_n == ref_n or any(got[p][0] for p in range(ns)))

# ---- f4: loop-closure rectangles: many small rectangles inside the matrices of a few hundred pairs
nl = 5 * n
iq = rng.integers(0, 400, nl).astype(np.int32) * 2
it = iq + 1
rects = np.zeros((nl, 4), np.int32)
Lq_all = np.array([len(seqs[i]) for i in iq]); Lt_all = np.array([len(seqs[i]) for i in it])
q0 = (rng.random(nl) * (Lq_all - 30)).astype(np.int32); t0_ = (rng.random(nl) * (Lt_all - 30)).astype(np.int32)
rects[:, 0], rects[:, 1] = q0, t0_
rects[:, 2] = q0 + rng.integers(2, 26, nl); rects[:, 3] = t0_ + rng.integers(2, 26, nl)
c.fill_subpair_batch(res, off, iq, it, rects, a.FWD)      # first call: buffers grow
c.set_profiling(True)
t0 = time.time()
sc, aoff, pairs, nout, st = c.fill_subpair_batch(res, off, iq, it, rects, a.FWD)
t1 = time.time()
kms4 = sum(ms for name, ms, _ in c.profile())
c.set_profiling(False)
out["f4_subpair_batch"] = {"loops": int(nl), "rect_edge": "2..25", "call_ms": round((t1 - t0) * 1e3, 1), "kernel_ms": round(kms4, 2),
                           "loops_per_s": round(nl / (t1 - t0)), "illegal_start": int((st != 0).sum())}
if os.path.exists(po.LIB_REF):
    ns = 200
    t0 = time.time()
    for k in range(ns):
        R.fill_sub(seqs[iq[k]], seqs[it[k]], tuple(int(x) for x in rects[k]), po.FWD)
    out["f4_subpair_batch"]["reference_cpu_loops_per_s_1core"] = round(ns / (time.time() - t0))
c.close()

# ---- f3: one 300 x 300 pair with HMAP-shaped positional gap tables
sim, dt, itab = po.hmap_like_tables(rng, 300, 300, po.SEMI_LOCAL)
c = a.Context(0)
c.fill_pair_tabulated(sim, dt, itab, False, a.FWD)
t0 = time.time()
for _ in range(5):
    s, _, _ = c.fill_pair_tabulated(sim, dt, itab, False, a.FWD)
tg = (time.time() - t0) / 5
out["f3_tabulated"] = {"pair": "300x300", "gpu_call_ms": round(tg * 1e3, 2)}
if os.path.exists(po.LIB_REF):
    t0 = time.time()
    rs, _, _ = po.reference_fill_tab(sim, dt, itab, False, po.FWD)
    out["f3_tabulated"]["reference_cpu_ms_1core"] = round((time.time() - t0) * 1e3, 1)
    out["f3_tabulated"]["identical"] = bool(np.array_equal(rs, s))
c.close()

# ---- C1 (BASELINE.json configs[0]): ONE pair of 250-residue proteins, fwd+rev fill with dense DPCell-shaped outputs on
# the host, optimal alignment and UCW enumeration (delta 0.01) -- the latency case, next to the reference on one core
rng1 = np.random.default_rng(1001)
q1 = rng1.integers(0, 20, 250).astype(np.uint8)
t1 = q1.copy()
t1[::6] = rng1.integers(0, 20, len(t1[::6]))
c = a.Context(0)
c.set_scoring(M, 12, 1, a.SEMI_LOCAL)
res1, off1 = a.Context.pack([q1, t1])
one = np.array([0], np.int32), np.array([1], np.int32)
def c1_gpu():
    o = c.fill_pair(q1, t1, a.BOTH, delta_ratio=0.01)
    c.fill_batch(res1, off1, one[0], one[1], a.W_FWD | a.W_REV | a.W_TB | a.W_MASK, 0.01)
    rc, ali, sc = c.optimal(0, a.FWD, 250, 250)
    return o, c.near_optimal([0], 0.01, 1024)[0]
c1_gpu()
t0 = time.time()
for _ in range(5):
    o, (st1, thr1, alis1) = c1_gpu()
tg = (time.time() - t0) / 5
out["c1_single_pair"] = {"pair": "250x250 related", "gpu_ms": round(tg * 1e3, 2), "near_optimal_alignments": len(alis1)}
if os.path.exists(po.LIB_REF):
    t0 = time.time()
    rf = R.fill(q1, t1, po.FWD)
    rr = R.fill(q1, t1, po.REV)
    ra = R.ucw_alignments(q1, t1, 0.01, 100000)
    out["c1_single_pair"]["reference_cpu_ms_1core"] = round((time.time() - t0) * 1e3, 1)
    out["c1_single_pair"]["identical"] = bool(np.array_equal(rf[0], o["score_fwd"]) and np.array_equal(rr[0], o["score_rev"])
                                               and len(ra) == len(alis1)
                                               and sorted(s for s, _ in ra) == sorted(s for s, _ in alis1))
c.close()
print(json.dumps(out))
