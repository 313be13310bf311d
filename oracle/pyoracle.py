"""ctypes bindings for the TEST oracles (oracle/liboracle.so and oracle/_ref/libaadp_ref.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product package
(alignment_algos_b200) never does.

`Oracle`  -> the plain-C restatement (oracle/aadp_oracle.c), available everywhere.
`Reference` -> the real reference compiled from /root/reference (oracle/ref_harness.cpp);
              present wherever oracle/_ref/libaadp_ref.so was built (it ships to the GPU box).
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libaadp_ref.so")

GLOBAL_LOCAL, GLOBAL, LOCAL_GLOBAL, LOCAL, SEMI_LOCAL = 0, 1, 2, 3, 4  # alib.h:20-26
FWD, REV = 1, 2  # dpmatrix.h:23-26


def build(force=False):
    """Compile liboracle.so (always) and _ref/libaadp_ref.so (when /root/reference exists)."""
    if force or not os.path.exists(LIB_ORACLE) or (
        os.path.getmtime(LIB_ORACLE) < os.path.getmtime(os.path.join(HERE, "aadp_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (
        force or not os.path.exists(LIB_REF)
        or os.path.getmtime(LIB_REF) < os.path.getmtime(os.path.join(HERE, "ref_harness.cpp"))
    ):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)


class _Scoring(C.Structure):
    _fields_ = [("A", C.c_int), ("sub", C.POINTER(C.c_float)), ("gi", C.c_float),
                ("ge", C.c_float), ("align_type", C.c_int)]


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class Oracle:
    """The C restatement. Sequences are uint8 code arrays without sentinels."""

    def __init__(self, sub, gi, ge, align_type):
        build()
        self.lib = C.CDLL(LIB_ORACLE)
        self.sub = np.ascontiguousarray(sub, dtype=np.float32)
        self.A = self.sub.shape[0]
        self.sc = _Scoring(self.A, _p(self.sub, C.c_float), gi, ge, align_type)
        self.align_type = align_type
        self.lib.orc_threshold.restype = C.c_float
        self.lib.orc_threshold.argtypes = [C.c_float, C.c_float]
        self.lib.orc_nearopt_mask.restype = C.c_long
        self.lib.orc_ucw_cells.restype = C.c_long

    def fill(self, q, t, direction=FWD, repro_rev_bug=True, fast=False, want_sim=False):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        t = np.ascontiguousarray(t, dtype=np.uint8)
        sz1, sz2 = len(q) + 2, len(t) + 2
        score = np.zeros((sz1, sz2), np.float32)
        pq = np.zeros((sz1, sz2), np.int32)
        pt = np.zeros((sz1, sz2), np.int32)
        if fast:
            rc = self.lib.orc_fill_fast(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t),
                                        C.byref(self.sc), direction, int(repro_rev_bug),
                                        _p(score, C.c_float), _p(pq, C.c_int), _p(pt, C.c_int))
            sim = None
        else:
            sim = np.zeros((sz1, sz2), np.float32)
            rc = self.lib.orc_fill(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t),
                                   C.byref(self.sc), direction, int(repro_rev_bug),
                                   _p(score, C.c_float), _p(pq, C.c_int), _p(pt, C.c_int),
                                   _p(sim, C.c_float))
        if rc:
            raise RuntimeError("Illegal bounds building DPM")
        return (score, pq, pt, sim) if want_sim else (score, pq, pt)

    def fill_rec(self, q, t, direction=FWD, repro_rev_bug=True):
        """orc_fill_rec: the record-list fill (CPU model of csrc/aadp_frec.cuh).  Returns (score, pq, pt, stats) with
        stats = (cells, row-walk steps, column-walk steps, cells with an ambiguous column leader)."""
        q = np.ascontiguousarray(q, dtype=np.uint8)
        t = np.ascontiguousarray(t, dtype=np.uint8)
        sz1, sz2 = len(q) + 2, len(t) + 2
        score = np.zeros((sz1, sz2), np.float32)
        pq = np.zeros((sz1, sz2), np.int32)
        pt = np.zeros((sz1, sz2), np.int32)
        st = (C.c_long * 4)()
        rc = self.lib.orc_fill_rec(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), C.byref(self.sc), direction,
                                   int(repro_rev_bug), _p(score, C.c_float), _p(pq, C.c_int), _p(pt, C.c_int), st)
        if rc:
            raise RuntimeError("Illegal bounds building DPM")
        return score, pq, pt, tuple(st)

    def fill_sub(self, q, t, rect, direction=FWD, repro_rev_bug=True):
        """build_subdpm (dpmatrix.h:319-353): rect = (q1_end, t1_end, q2_beg, t2_beg), matrix indices."""
        q = np.ascontiguousarray(q, dtype=np.uint8)
        t = np.ascontiguousarray(t, dtype=np.uint8)
        sz1, sz2 = len(q) + 2, len(t) + 2
        score = np.zeros((sz1, sz2), np.float32)
        pq = np.zeros((sz1, sz2), np.int32)
        pt = np.zeros((sz1, sz2), np.int32)
        rc = self.lib.orc_fill_sub(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), C.byref(self.sc),
                                   direction, int(repro_rev_bug), int(rect[0]), int(rect[1]), int(rect[2]),
                                   int(rect[3]), _p(score, C.c_float), _p(pq, C.c_int), _p(pt, C.c_int))
        if rc:
            raise RuntimeError("Illegal bounds building DPM")
        return score, pq, pt

    def ucw_enumerate(self, q, t, F, sim, thr, max_alignments=20000, pq=None, pt=None, user_limit=100000):
        """orc_ucw_enumerate: (status, [(score, pairs[(len,2)])]) in the reference's depth-first slot order.
        user_limit: ucw.h:72 (beyond it the reference forces optimal paths, ucw.h:115-126; needs pq/pt)."""
        Lq, Lt = len(q), len(t)
        K = int(max_alignments)
        scores = np.zeros(K, np.float32)
        ln = np.zeros(K, np.int32)
        pairs = np.zeros((K, Lq + 2, 2), np.int32)
        st = C.c_int(0)
        self.lib.orc_ucw_enumerate_lim.restype = C.c_long
        n = self.lib.orc_ucw_enumerate_lim(Lq, Lt, C.byref(self.sc), _p(np.ascontiguousarray(F, np.float32), C.c_float),
                                           _p(np.ascontiguousarray(sim, np.float32), C.c_float), C.c_float(thr), C.c_long(K),
                                           _p(scores, C.c_float), _p(ln, C.c_int), _p(pairs, C.c_int), C.byref(st),
                                           _p(np.ascontiguousarray(pq, np.int32), C.c_int) if pq is not None else None,
                                           _p(np.ascontiguousarray(pt, np.int32), C.c_int) if pt is not None else None,
                                           C.c_long(int(user_limit)))
        return st.value, [(float(scores[a]), pairs[a, :ln[a]].copy()) for a in range(n)]

    def cno_enumerate(self, q, t, F, sim, thr, pq, pt, flags=None, max_alignments=20000, user_limit=1000000):
        """orc_cno_enumerate (cw.h): (status, [(score, pairs)]) in the reference's slot order."""
        Lq, Lt = len(q), len(t)
        K = int(max_alignments)
        scores = np.zeros(K, np.float32)
        ln = np.zeros(K, np.int32)
        pairs = np.zeros((K, Lq + 2, 2), np.int32)
        st = C.c_int(0)
        fl = np.ascontiguousarray(flags, np.uint8) if flags is not None else None
        self.lib.orc_cno_enumerate_lim.restype = C.c_long
        n = self.lib.orc_cno_enumerate_lim(Lq, Lt, C.byref(self.sc), _p(np.ascontiguousarray(F, np.float32), C.c_float),
                                           _p(np.ascontiguousarray(sim, np.float32), C.c_float), C.c_float(thr), C.c_long(K),
                                           _p(scores, C.c_float), _p(ln, C.c_int), _p(pairs, C.c_int), C.byref(st),
                                           _p(np.ascontiguousarray(pq, np.int32), C.c_int),
                                           _p(np.ascontiguousarray(pt, np.int32), C.c_int),
                                           _p(fl, C.c_uint8) if fl is not None else None, C.c_long(int(user_limit)))
        return st.value, [(float(scores[a]), pairs[a, :ln[a]].copy()) for a in range(n)]

    @staticmethod
    def fill_tab(sim, del_tab, ins_tab, is_local=False, direction=FWD, repro_rev_bug=True):
        """orc_fill_tab: the literal fill for any evaluator given as tables (see aadp_oracle.h)."""
        build()
        lib = C.CDLL(LIB_ORACLE)
        sim = np.ascontiguousarray(sim, np.float32)
        del_tab = np.ascontiguousarray(del_tab, np.float32)
        ins_tab = np.ascontiguousarray(ins_tab, np.float32)
        sz1, sz2 = sim.shape
        score = np.zeros((sz1, sz2), np.float32)
        pq = np.zeros((sz1, sz2), np.int32)
        pt = np.zeros((sz1, sz2), np.int32)
        rc = lib.orc_fill_tab(_p(sim, C.c_float), sz1 - 2, sz2 - 2, _p(del_tab, C.c_float), _p(ins_tab, C.c_float),
                              int(bool(is_local)), direction, int(repro_rev_bug), _p(score, C.c_float), _p(pq, C.c_int),
                              _p(pt, C.c_int))
        if rc:
            raise RuntimeError("Illegal bounds building DPM")
        return score, pq, pt

    def sim(self, q, t):
        q = np.asarray(q, dtype=np.int64)
        t = np.asarray(t, dtype=np.int64)
        s = np.zeros((len(q) + 2, len(t) + 2), np.float32)
        if len(q) and len(t):
            s[1:-1, 1:-1] = self.sub[q[:, None], t[None, :]]
        return s

    def optimal(self, score, pq, pt, direction=FWD):
        sz1, sz2 = score.shape
        cap = sz1 + sz2 + 8
        pairs = np.zeros((cap, 2), np.int32)
        n = C.c_int(0)
        s = C.c_float(0)
        fn = self.lib.orc_optimal_fwd if direction == FWD else self.lib.orc_optimal_rev
        rc = fn(_p(np.ascontiguousarray(score), C.c_float), _p(np.ascontiguousarray(pq), C.c_int),
                _p(np.ascontiguousarray(pt), C.c_int), sz1, sz2, int(self.align_type == LOCAL),
                _p(pairs, C.c_int), cap, C.byref(n), C.byref(s))
        return rc, pairs[: min(n.value, cap)].copy(), s.value

    def threshold(self, opt, delta_ratio):
        return self.lib.orc_threshold(C.c_float(opt), C.c_float(delta_ratio))

    def nearopt_mask(self, F, R, sim, thr):
        sz1, sz2 = F.shape
        mask = np.zeros((sz1, sz2), np.uint8)
        n = self.lib.orc_nearopt_mask(_p(np.ascontiguousarray(F), C.c_float),
                                      _p(np.ascontiguousarray(R), C.c_float),
                                      _p(np.ascontiguousarray(sim), C.c_float), sz1, sz2,
                                      C.c_float(thr), _p(mask, C.c_uint8))
        return mask, n

    def ucw_cells(self, q, t, F, sim, thr, max_alignments=2_000_000):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        t = np.ascontiguousarray(t, dtype=np.uint8)
        mark = np.zeros(F.shape, np.uint8)
        n = self.lib.orc_ucw_cells(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t),
                                   C.byref(self.sc), _p(np.ascontiguousarray(F), C.c_float),
                                   _p(np.ascontiguousarray(sim), C.c_float), C.c_float(thr),
                                   C.c_long(max_alignments), _p(mark, C.c_uint8))
        return mark, n


def reference_fill_tab(sim, del_tab, ins_tab, is_local=False, direction=FWD):
    """The REAL reference fill (libaadp_ref.so, ref_fill_tab) driven by a table-backed Evaluator."""
    build()
    if not os.path.exists(LIB_REF):
        raise FileNotFoundError(LIB_REF)
    lib = C.CDLL(LIB_REF)
    lib.ref_last_error.restype = C.c_char_p
    sim = np.ascontiguousarray(sim, np.float32)
    del_tab = np.ascontiguousarray(del_tab, np.float32)
    ins_tab = np.ascontiguousarray(ins_tab, np.float32)
    sz1, sz2 = sim.shape
    score = np.zeros((sz1, sz2), np.float32)
    pq = np.zeros((sz1, sz2), np.int32)
    pt = np.zeros((sz1, sz2), np.int32)
    rc = lib.ref_fill_tab(sz1 - 2, sz2 - 2, _p(sim, C.c_float), _p(del_tab, C.c_float), _p(ins_tab, C.c_float),
                          int(bool(is_local)), direction, _p(score, C.c_float), _p(pq, C.c_int), _p(pt, C.c_int))
    if rc:
        raise RuntimeError(lib.ref_last_error().decode())
    return score, pq, pt


def hmap_like_tables(rng, Lq, Lt, align_type, dyadic=False):
    """Tables of an evaluator shaped like hmap_eval.h:63-117: per-template-position gap_init/gap_extn, a gap between
    t1 and t2 costs min(gi[t1],gi[t2]) + min(ge[t1],ge[t2])*(dist-2) -- for deletions AND insertions (the insertion
    takes its parameters from the two template positions it sits between) -- with the free end gaps of the align type;
    similarity is position specific (a profile score)."""
    sz1, sz2 = Lq + 2, Lt + 2
    f32 = np.float32
    if dyadic:
        gi = (rng.integers(16, 64, sz2) / 4.0).astype(f32)
        ge = (rng.integers(1, 12, sz2) / 4.0).astype(f32)
        sim = (rng.integers(-16, 28, (sz1, sz2)) / 2.0).astype(f32)
    else:
        gi = rng.uniform(3, 14, sz2).astype(f32)
        ge = rng.uniform(0.1, 2.5, sz2).astype(f32)
        sim = rng.normal(0.3, 3.0, (sz1, sz2)).astype(f32)
    sim[0, :] = sim[-1, :] = 0
    sim[:, 0] = sim[:, -1] = 0
    del_free = align_type in (LOCAL, SEMI_LOCAL, LOCAL_GLOBAL)
    ins_free = align_type in (LOCAL, SEMI_LOCAL, GLOBAL_LOCAL)
    dt = np.zeros((sz2, sz2), f32)
    for t1 in range(sz2):
        for t2 in range(t1 + 2, sz2):
            if del_free and (t1 == 0 or t2 == sz2 - 1):
                continue
            dt[t1, t2] = f32(min(gi[t1], gi[t2]) + f32(min(ge[t1], ge[t2]) * f32(t2 - t1 - 2)))
    it = np.zeros((sz1 - 1, sz2), f32)
    for ln in range(1, sz1 - 1):
        for t2 in range(1, sz2):
            if ins_free and (t2 == 1 or t2 == sz2 - 1):   # q1 is the Head / q2 is the Tail for these columns
                continue
            it[ln, t2] = f32(min(gi[t2 - 1], gi[t2]) + f32(min(ge[t2 - 1], ge[t2]) * f32(ln - 1)))
    return sim, dt, it


def gn2_like_tables(rng, Lq, Lt, align_type):
    """Tables of an evaluator shaped like gn2_eval.h:99-158: deletions from a pairwise table (8100 beyond a distance
    cutoff, else gi[p2][p1] + ge[p2][p1]*(di-2) + cd[p2][p1]), insertions from per-position vectors
    v_gi[t1] + v_ge[t1]*(di-2) + v_cn[t1]."""
    sz1, sz2 = Lq + 2, Lt + 2
    f32 = np.float32
    sim = rng.normal(0.2, 2.5, (sz1, sz2)).astype(f32)
    sim[0, :] = sim[-1, :] = 0
    sim[:, 0] = sim[:, -1] = 0
    del_free = align_type in (LOCAL, SEMI_LOCAL, LOCAL_GLOBAL)
    ins_free = align_type in (LOCAL, SEMI_LOCAL, GLOBAL_LOCAL)
    dist = rng.uniform(3, 30, (sz2, sz2)).astype(f32)
    vgi = rng.uniform(4, 12, (sz2, sz2)).astype(f32)
    vge = rng.uniform(0.2, 1.5, (sz2, sz2)).astype(f32)
    vcd = rng.uniform(0, 3, (sz2, sz2)).astype(f32)
    dt = np.zeros((sz2, sz2), f32)
    for t1 in range(sz2):
        for t2 in range(t1 + 2, sz2):
            if del_free and (t1 == 0 or t2 == sz2 - 1):
                continue
            p1, p2 = t1, t2 - 2
            gp = f32(8100.0)
            if dist[p2, p1] < 18.0:
                gp = f32(f32(vgi[p2, p1] + f32(vge[p2, p1] * f32(t2 - t1 - 2))) + vcd[p2, p1])
            dt[t1, t2] = gp
    wgi = rng.uniform(4, 12, sz2).astype(f32)
    wge = rng.uniform(0.2, 1.5, sz2).astype(f32)
    wcn = rng.uniform(0, 4, sz2).astype(f32)
    it = np.zeros((sz1 - 1, sz2), f32)
    for ln in range(1, sz1 - 1):
        for t2 in range(1, sz2):
            if ins_free and (t2 == 1 or t2 == sz2 - 1):
                continue
            t1 = t2 - 1
            it[ln, t2] = f32(f32(wgi[t1] + f32(wge[t1] * f32(ln - 1))) + wcn[t1])
    return sim, dt, it


def write_matrix_file(path, alphabet, sub):
    """Write a substitution matrix in the format submatrix.cpp:16-54 parses."""
    with open(path, "w") as f:
        f.write("# written by oracle/pyoracle.py\n")
        f.write("   " + "  ".join(alphabet) + "\n")
        for i, a in enumerate(alphabet):
            f.write(a + " " + " ".join(repr(float(x)) if float(x) != int(x) else str(int(x))
                                       for x in sub[i]) + "\n")


class Reference:
    """The real reference (libaadp_ref.so). Sequences are letter strings (no sentinels)."""

    def __init__(self, alphabet, sub, gi, ge, align_type):
        build()
        if not os.path.exists(LIB_REF):
            raise FileNotFoundError(LIB_REF)
        self.lib = C.CDLL(LIB_REF)
        self.lib.ref_last_error.restype = C.c_char_p
        self.lib.ref_time_fills.restype = C.c_double
        self.alphabet = alphabet
        self.gi, self.ge, self.align_type = float(gi), float(ge), int(align_type)
        fd, self.matrix_file = tempfile.mkstemp(prefix="aadp_sub_", suffix=".txt")
        os.close(fd)
        write_matrix_file(self.matrix_file, alphabet, np.asarray(sub))

    def __del__(self):
        try:
            os.unlink(self.matrix_file)
        except Exception:
            pass

    def letters(self, codes):
        return "".join(self.alphabet[int(c)] for c in codes).encode()

    def _args(self, q, t):
        return (self.letters(q), self.letters(t), self.matrix_file.encode(), C.c_float(self.gi),
                C.c_float(self.ge), self.align_type)

    def fill(self, q, t, direction=FWD):
        sz1, sz2 = len(q) + 2, len(t) + 2
        score = np.zeros((sz1, sz2), np.float32)
        pq = np.zeros((sz1, sz2), np.int32)
        pt = np.zeros((sz1, sz2), np.int32)
        sim = np.zeros((sz1, sz2), np.float32)
        rc = self.lib.ref_fill(*self._args(q, t), direction, _p(score, C.c_float),
                               _p(pq, C.c_int), _p(pt, C.c_int), _p(sim, C.c_float))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return score, pq, pt, sim

    def fill_sub(self, q, t, rect, direction=FWD):
        sz1, sz2 = len(q) + 2, len(t) + 2
        score = np.zeros((sz1, sz2), np.float32)
        pq = np.zeros((sz1, sz2), np.int32)
        pt = np.zeros((sz1, sz2), np.int32)
        rc = self.lib.ref_fill_sub(*self._args(q, t), direction, int(rect[0]), int(rect[1]), int(rect[2]),
                                   int(rect[3]), _p(score, C.c_float), _p(pq, C.c_int), _p(pt, C.c_int))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return score, pq, pt

    def optimal(self, q, t, direction=FWD):
        cap = len(q) + len(t) + 16
        pairs = np.zeros((cap, 2), np.int32)
        n = C.c_int(0)
        s = C.c_float(0)
        if direction == FWD:
            ident = C.c_float(0)
            rc = self.lib.ref_optimal(*self._args(q, t), _p(pairs, C.c_int), cap, C.byref(n),
                                      C.byref(s), C.byref(ident))
        else:
            rc = self.lib.ref_optimal_rev(*self._args(q, t), _p(pairs, C.c_int), cap, C.byref(n),
                                          C.byref(s))
        return rc, pairs[: min(n.value, cap)].copy(), s.value

    def nearopt(self, q, t, delta_ratio, number_suboptimal=0, which=0, flags=None, max_scores=4096):
        sz1, sz2 = len(q) + 2, len(t) + 2
        union = np.zeros((sz1, sz2), np.uint8)
        n = C.c_int(0)
        scores = np.zeros(max_scores, np.float32)
        thr = C.c_float(0)
        rc = self.lib.ref_nearopt(*self._args(q, t), C.c_float(delta_ratio), number_suboptimal,
                                  which, flags.encode() if flags else None, _p(union, C.c_uint8),
                                  C.byref(n), _p(scores, C.c_float), max_scores, C.byref(thr))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return union, n.value, scores[: min(n.value, max_scores)].copy(), thr.value

    def ucw_alignments(self, q, t, delta_ratio, max_alignments=20000, which=0, flags=None):
        """Every alignment of the reference's UnconstrainedNearOptimal (which=0) or ConstrainedNearOptimal (which=1,
        flags = SuboptFlags per template position incl. sentinels, None = all true), sorted by its sortSet."""
        K = int(max_alignments)
        fl = None
        if flags is not None:
            fl = "".join("1" if f else "0" for f in flags).encode()
        cap = K * (len(q) + 2)
        scores = np.zeros(K, np.float32)
        ln = np.zeros(K, np.int32)
        pairs = np.zeros((cap, 2), np.int32)
        n, tot = C.c_int(0), C.c_long(0)
        rc = self.lib.ref_ucw_alignments(*self._args(q, t), C.c_float(delta_ratio), int(which), fl, K, C.c_long(cap), C.byref(n),
                                         C.byref(tot), _p(scores, C.c_float), _p(ln, C.c_int), _p(pairs, C.c_int))
        if rc == 5:
            raise OverflowError("%d alignments" % n.value)
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        out, o = [], 0
        for k in range(n.value):
            out.append((float(scores[k]), pairs[o:o + ln[k]].copy()))
            o += ln[k]
        return out

    def pruned_alignments(self, q, t, delta_ratio, which, flags=None, k_limit=16, sort_limit=100, max_overlap=0.30,
                          user_limit=100000, max_alignments=20000):
        """Every alignment of the reference's KSConstrainedNearOptimal (which=2, kscw.h) or CRConstrainedNearOptimal
        (which=3, crcw.h), sorted by its sortSet.  flags: SuboptFlags per template position incl. sentinels."""
        K = int(max_alignments)
        fl = None
        if flags is not None:
            fl = "".join("1" if f else "0" for f in flags).encode()
        cap = K * (len(q) + len(t) + 4)
        scores = np.zeros(K, np.float32)
        ln = np.zeros(K, np.int32)
        pairs = np.zeros((cap, 2), np.int32)
        n, tot = C.c_int(0), C.c_long(0)
        rc = self.lib.ref_pruned_alignments(*self._args(q, t), C.c_float(delta_ratio), int(which), fl, C.c_uint(k_limit),
                                            C.c_uint(sort_limit), C.c_float(max_overlap), C.c_uint(user_limit), K, C.c_long(cap),
                                            C.byref(n), C.byref(tot), _p(scores, C.c_float), _p(ln, C.c_int), _p(pairs, C.c_int))
        if rc == 5:
            raise OverflowError("%d alignments" % n.value)
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        out, o = [], 0
        for k in range(n.value):
            out.append((float(scores[k]), pairs[o:o + ln[k]].copy()))
            o += ln[k]
        return out

    def time_fills(self, seqs, pair_q, pair_t, what=3, nthreads=1):
        """Time the reference DPMatrix constructor over pairs. Returns (seconds, cells, checksum)."""
        arena = b"".join(self.letters(s) for s in seqs)
        lens = np.array([len(s) for s in seqs], np.int32)
        off = np.zeros(len(seqs), np.int64)
        off[1:] = np.cumsum(lens)[:-1]
        pq = np.ascontiguousarray(pair_q, np.int32)
        pt = np.ascontiguousarray(pair_t, np.int32)
        cells = C.c_double(0)
        chk = C.c_double(0)
        sec = self.lib.ref_time_fills(arena, _p(off, C.c_longlong), _p(lens, C.c_int),
                                      _p(pq, C.c_int), _p(pt, C.c_int), len(pq),
                                      self.matrix_file.encode(), C.c_float(self.gi),
                                      C.c_float(self.ge), self.align_type, what, nthreads,
                                      C.byref(cells), C.byref(chk))
        return sec, cells.value, chk.value
