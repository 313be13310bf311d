"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref/libaadp_ref.so).

Run in the build container (needs /root/reference):   python oracle/gen_golden.py
The reference ships no golden vectors of its own; these fixtures are outputs of
its unmodified code (IEEE build) on seeded inputs, and are what pins parity on the GPU box where
/root/reference does not exist.
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
    os.makedirs(OUT, exist_ok=True)
    alpha, M = read_matrix(BLOSUM62)
    rng = np.random.default_rng(20261018)
    cases = []
    # (name, Lq, Lt, gi, ge, align_type)
    specs = [("toy", 9, 5, 3.0, 1.0, po.GLOBAL), ("toy_semi", 9, 5, 3.0, 1.0, po.SEMI_LOCAL)]
    for at in range(5):
        specs.append(("r40_at%d" % at, 37, 43, 12.0, 1.0, at))
        specs.append(("r70_at%d" % at, 70, 61, 11.0, 1.0, at))
        specs.append(("half_at%d" % at, 33, 29, 10.5, 0.25, at))
    specs += [("empty_q", 0, 7, 12.0, 1.0, po.GLOBAL), ("one_one", 1, 1, 12.0, 1.0, po.SEMI_LOCAL),
              ("one_t", 12, 1, 12.0, 1.0, po.GLOBAL), ("two_two", 2, 2, 2.0, 2.0, po.GLOBAL),
              ("c1_250", 250, 250, 12.0, 1.0, po.SEMI_LOCAL), ("c1_250_global", 250, 250, 12.0, 1.0, po.GLOBAL),
              ("wide_270", 40, 270, 12.0, 1.0, po.SEMI_LOCAL), ("stripe_530", 35, 530, 12.0, 1.0, po.GLOBAL)]
    # the dpmatrix.h:868 case
    b5_q = [alpha.index(c) for c in "AAAACDEC"]
    b5_t = [alpha.index(c) for c in "CDEC"]
    for name, Lq, Lt, gi, ge, at in specs:
        q = rng.integers(0, 20, Lq).astype(np.uint8)
        t = rng.integers(0, 20, Lt).astype(np.uint8)
        cases.append((name, q, t, gi, ge, at))
    cases.append(("b5_revbug", np.array(b5_q, np.uint8), np.array(b5_t, np.uint8), 1.0, 0.0, po.GLOBAL))
    blob = {"alphabet": np.array(alpha), "sub": M, "names": np.array([c[0] for c in cases])}
    for name, q, t, gi, ge, at in cases:
        R = po.Reference(alpha, M, gi, ge, at)
        blob[name + ".q"] = q
        blob[name + ".t"] = t
        blob[name + ".params"] = np.array([gi, ge, at], np.float32)
        for d, tag in ((po.FWD, "fwd"), (po.REV, "rev")):
            s, pq, pt, sim = R.fill(q, t, d)
            blob[name + "." + tag + ".score"] = s
            blob[name + "." + tag + ".pq"] = pq.astype(np.int16)
            blob[name + "." + tag + ".pt"] = pt.astype(np.int16)
            rc, pairs, sc = R.optimal(q, t, d)
            blob[name + "." + tag + ".opt_rc"] = np.array([rc], np.int32)
            blob[name + "." + tag + ".opt_pairs"] = pairs.astype(np.int16)
            blob[name + "." + tag + ".opt_score"] = np.array([sc], np.float32)
        # near-optimal enumeration (UCW) on the small cases only: the branching is exponential
        if at != po.LOCAL and 2 <= len(q) <= 45 and 2 <= len(t) <= 45:
            for dr in (0.05, 0.2):
                try:
                    union, n, scores, thr = R.nearopt(q, t, dr, 0, 0)
                except RuntimeError:
                    continue
                if n > 90000:
                    continue  # user_limit territory (ucw.h:72): union no longer equals the mask
                key = "%s.ucw%02d" % (name, int(dr * 100))
                blob[key + ".union"] = np.packbits(union, axis=None)
                blob[key + ".n"] = np.array([n], np.int64)
                blob[key + ".thr"] = np.array([thr], np.float32)
                blob[key + ".top"] = scores[:16]
    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **blob)
    print("wrote", len(cases), "cases,", os.path.getsize(os.path.join(OUT, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
