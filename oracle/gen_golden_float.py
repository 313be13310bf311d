"""Generate tests/golden/reference_vectors_float.npz from the REAL reference (oracle/_ref/libaadp_ref.so)
for scoring that is NOT on a dyadic grid: the reference defaults gap_init 4.73 / gap_extn 0.34
(alib.cpp:17-18) and a non-integer substitution matrix.  These pin the exact general-gap fp32 path
(alignment_algos_b200/csrc/aadp_general.cuh), which must reproduce every rounding of the reference.

Run in the build container (needs /root/reference):   python oracle/gen_golden_float.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from alignment_algos_b200.submatrix import read_matrix, BLOSUM62  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    alpha, M = read_matrix(BLOSUM62)
    Mf = (M.astype(np.float32) * np.float32(0.37)).astype(np.float32)  # non-integer substitution scores
    rng = np.random.default_rng(20261019)
    specs = []
    for at in range(5):
        specs.append(("dflt40_at%d" % at, 38, 45, 4.73, 0.34, at, "blosum"))
        specs.append(("frac30_at%d" % at, 31, 27, 1.9, 0.11, at, "scaled"))
    specs += [("dflt_c1_250", 250, 250, 4.73, 0.34, po.SEMI_LOCAL, "blosum"),
              ("dflt_empty_t", 6, 0, 4.73, 0.34, po.GLOBAL, "blosum"),
              ("dflt_one_one", 1, 1, 4.73, 0.34, po.GLOBAL, "blosum"),
              ("dflt_wide", 12, 140, 4.73, 0.34, po.GLOBAL_LOCAL, "blosum")]
    blob = {"alphabet": np.array(alpha), "sub.blosum": M.astype(np.float32), "sub.scaled": Mf,
            "names": np.array([s[0] for s in specs])}
    for name, Lq, Lt, gi, ge, at, which in specs:
        sub = M.astype(np.float32) if which == "blosum" else Mf
        q = rng.integers(0, 20, Lq).astype(np.uint8)
        t = rng.integers(0, 20, Lt).astype(np.uint8)
        R = po.Reference(alpha, sub, gi, ge, at)
        blob[name + ".q"] = q
        blob[name + ".t"] = t
        blob[name + ".params"] = np.array([gi, ge, at], np.float32)
        blob[name + ".sub"] = np.array(which)
        for d, tag in ((po.FWD, "fwd"), (po.REV, "rev")):
            s, pq, pt, sim = R.fill(q, t, d)
            blob[name + "." + tag + ".score"] = s
            blob[name + "." + tag + ".pq"] = pq.astype(np.int16)
            blob[name + "." + tag + ".pt"] = pt.astype(np.int16)
            rc, pairs, sc = R.optimal(q, t, d)
            blob[name + "." + tag + ".opt_rc"] = np.array([rc], np.int32)
            blob[name + "." + tag + ".opt_pairs"] = pairs.astype(np.int16)
            blob[name + "." + tag + ".opt_score"] = np.array([sc], np.float32)
        if at != po.LOCAL and 2 <= Lq <= 45 and 2 <= Lt <= 45:
            for dr in (0.05, 0.2):
                try:
                    union, n, scores, thr = R.nearopt(q, t, dr, 0, 0)
                except RuntimeError:
                    continue
                if n > 90000:
                    continue
                key = "%s.ucw%02d" % (name, int(dr * 100))
                blob[key + ".union"] = np.packbits(union, axis=None)
                blob[key + ".n"] = np.array([n], np.int64)
                blob[key + ".thr"] = np.array([thr], np.float32)
    path = os.path.join(OUT, "reference_vectors_float.npz")
    np.savez_compressed(path, **blob)
    print("wrote", len(specs), "cases,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
