"""Generate tests/golden/reference_vectors_tab.npz from the REAL reference: the fill of dpmatrix.h:356-1030 driven by a
table-backed Evaluator (oracle/ref_harness.cpp: TableEval) with position-dependent gap penalties shaped like
hmap_eval.h:63-117 and gn2_eval.h:99-158.

Run in the build container (needs /root/reference):   python oracle/gen_golden_tab.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    rng = np.random.default_rng(20261021)
    blob, names = {}, []
    k = 0
    for at in range(5):
        for kind, gen in (("hmap", po.hmap_like_tables), ("gn2", po.gn2_like_tables)):
            for Lq, Lt in ((0, 5), (6, 0), (1, 1), (19, 33), (45, 28)):
                sim, dt, it = gen(rng, Lq, Lt, at)
                name = "tab%03d" % k
                k += 1
                names.append(name)
                blob[name + ".kind"] = np.array(kind)
                blob[name + ".local"] = np.array(int(at == po.LOCAL))
                blob[name + ".sim"] = sim
                blob[name + ".del"] = dt
                blob[name + ".ins"] = it
                for d, tag in ((po.FWD, "fwd"), (po.REV, "rev")):
                    s, pq, pt = po.reference_fill_tab(sim, dt, it, at == po.LOCAL, d)
                    blob[name + "." + tag + ".score"] = s
                    blob[name + "." + tag + ".pq"] = pq.astype(np.int16)
                    blob[name + "." + tag + ".pt"] = pt.astype(np.int16)
    blob["names"] = np.array(names)
    path = os.path.join(OUT, "reference_vectors_tab.npz")
    np.savez_compressed(path, **blob)
    print("wrote", len(names), "cases,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
