"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402  (tests are allowed to use the oracle)

HAVE_REF = os.path.exists(po.LIB_REF) or os.path.isdir("/root/reference")

MODES = [po.GLOBAL_LOCAL, po.GLOBAL, po.LOCAL_GLOBAL, po.LOCAL, po.SEMI_LOCAL]
MODE_NAMES = {0: "global_local", 1: "global", 2: "local_global", 3: "local", 4: "semi_local"}


def golden_cases(g):
    return [str(n) for n in g["names"]]


def golden_case(g, name):
    gi, ge, at = g[name + ".params"]
    return g[name + ".q"], g[name + ".t"], float(gi), float(ge), int(at)


def rand_pair(rng, Lq, Lt, A=20):
    return rng.integers(0, A, Lq).astype(np.uint8), rng.integers(0, A, Lt).astype(np.uint8)


def assert_matrix_equal(name, got, want):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (name, got.shape, want.shape)
    bad = np.argwhere(got != want)
    assert len(bad) == 0, "%s: %d mismatching cells, first at %s: got %s want %s" % (
        name, len(bad), tuple(bad[0]), got[tuple(bad[0])], want[tuple(bad[0])])
