"""Generate tests/golden/reference_vectors_sub.npz from the REAL reference: sub-rectangle fills through the
9-argument DPMatrix constructor / build_subdpm (dpmatrix.h:169-189, 319-353).

Run in the build container (needs /root/reference):   python oracle/gen_golden_sub.py
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
    rng = np.random.default_rng(20261020)
    blob = {"sub": M.astype(np.float32)}
    names = []
    k = 0
    for gi, ge in ((12.0, 1.0), (4.73, 0.34)):
        for at in range(5):
            R = po.Reference(alpha, M, gi, ge, at)
            Lq, Lt = int(rng.integers(20, 40)), int(rng.integers(20, 40))
            q = rng.integers(0, 20, Lq).astype(np.uint8)
            t = rng.integers(0, 20, Lt).astype(np.uint8)
            rects = [(0, 0, Lq + 1, Lt + 1), (3, 5, Lq - 2, Lt - 4), (0, 4, Lq - 5, Lt + 1), (6, 0, Lq + 1, Lt - 3),
                     (7, 7, 8, 15), (7, 7, 15, 8), (4, 4, 5, 5)]
            for rect in rects:
                name = "sub%03d" % k
                k += 1
                names.append(name)
                blob[name + ".q"] = q
                blob[name + ".t"] = t
                blob[name + ".params"] = np.array([gi, ge, at], np.float32)
                blob[name + ".rect"] = np.array(rect, np.int32)
                for d, tag in ((po.FWD, "fwd"), (po.REV, "rev")):
                    s, pq, pt = R.fill_sub(q, t, rect, d)
                    blob[name + "." + tag + ".score"] = s
                    blob[name + "." + tag + ".pq"] = pq.astype(np.int16)
                    blob[name + "." + tag + ".pt"] = pt.astype(np.int16)
    blob["names"] = np.array(names)
    path = os.path.join(OUT, "reference_vectors_sub.npz")
    np.savez_compressed(path, **blob)
    print("wrote", len(names), "cases,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
