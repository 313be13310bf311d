"""alignment_algos_b200 -- B200-native (sm_100a) DP matrix fill for christang/alignment-algos.

Scope: the forward/reverse affine-gap DP fill of dpmatrix.{h,cpp}, the optimal
tracebacks of optimal.h / optimal_rev.h and the near-optimal cell set consumed by ucw.h / cw.h.
Native code: csrc/ (CUDA kernels + the C ABI of include/aadp.h); host C++ drop-in headers:
include/hmap2/.  This Python package only marshals buffers for tests and bench.py.
"""
from .api import (AadpError, Context, GLOBAL_LOCAL, GLOBAL, LOCAL_GLOBAL, LOCAL, SEMI_LOCAL, FWD, REV, BOTH,
                  REPRO_REV_BUG, W_FWD, W_REV, W_TB, W_SCORES, W_MASK)
from .submatrix import read_matrix, blosum62, AA20, BLOSUM62

__all__ = ["AadpError", "Context", "GLOBAL_LOCAL", "GLOBAL", "LOCAL_GLOBAL", "LOCAL", "SEMI_LOCAL", "FWD", "REV",
           "BOTH", "REPRO_REV_BUG", "W_FWD", "W_REV", "W_TB", "W_SCORES", "W_MASK", "read_matrix", "blosum62",
           "AA20", "BLOSUM62"]
