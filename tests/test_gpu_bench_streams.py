"""GPU parity on the benchmark's OWN input streams (run with -m gpu on a B200).

The parity tests in test_gpu_parity.py use their own seeds and mostly short pairs.  Here the inputs are the seeded
streams bench.py times (synth.pair_workload(1002 / 1003), the seed-1004 sequence set of the all-vs-all workload), cut
to sizes the oracle finishes in seconds but still large enough that the production machinery is live: the pipelined
aadp_fill_batch (>= 32768 pairs: chunked task lists, piece-wise residue upload), couples of unequal query length in
one register, templates up to 500 columns (32 lanes), bin-packed tasks.  Bit-exact bar everywhere."""
import numpy as np
import pytest

from util import po, assert_matrix_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import alignment_algos_b200 as a
    c = a.Context(0)
    yield c
    c.close()


def _c3_prefix(npairs):
    # the first `npairs` pairs of the C3 stream: pair_workload draws the lengths of all sequences first, then the
    # residues sequence by sequence, so the whole stream is generated and cut
    from alignment_algos_b200 import synth
    seqs, pq, pt = synth.pair_workload(1003, 100_000, 100, 500)
    return seqs[: 2 * npairs], pq[:npairs].copy(), pt[:npairs].copy()


def _sample(pq, pt, seqs, n, rng):
    # a sample that over-represents what the short-pair tests do not reach: long templates (31-32 lanes), long
    # queries, and the extremes
    Lq = np.array([len(seqs[i]) for i in pq])
    Lt = np.array([len(seqs[i]) for i in pt])
    pick = set(rng.choice(len(pq), n // 2, replace=False).tolist())
    for key in (Lt, Lq, Lq * Lt, -Lt, -Lq):
        pick.update(np.argsort(key)[-n // 10:].tolist())
    pick.update(np.nonzero((Lt > 480) & (Lq > 330))[0][: n // 8].tolist())
    return sorted(pick)[: n + n // 2]


@pytest.mark.parametrize("with_scores", [False, True], ids=["tb+mask", "tb+scores+mask"])
def test_c3_stream_inside_pipelined_batch(ctx, with_scores):
    import alignment_algos_b200 as a
    _, M20 = a.blosum62()
    npairs = 40960
    seqs, pq, pt = _c3_prefix(npairs)
    res, off = a.Context.pack(seqs)
    ctx.set_scoring(M20, 12, 1, po.SEMI_LOCAL)
    O = po.Oracle(M20, 12, 1, po.SEMI_LOCAL)
    what = a.W_FWD | a.W_REV | a.W_TB | a.W_MASK | (a.W_SCORES if with_scores else 0)  # bench.py's `what` (+ scores)
    out = ctx.fill_batch(res, off, pq, pt, what, 0.01)
    assert_matrix_equal("fwd==rev optimum over the whole batch", out["fwd_score"], out["rev_score"])
    rng = np.random.default_rng(5)
    ids = _sample(pq, pt, seqs, 256, rng)
    assert len(ids) >= 256
    n_unequal = 0
    for p in ids:
        q, t = seqs[pq[p]], seqs[pt[p]]
        F, fq, ft = O.fill(q, t, po.FWD, True, fast=True)
        R, rq, rt = O.fill(q, t, po.REV, True, fast=True)
        tag = "c3 pair %d (%dx%d) " % (p, len(q), len(t))
        assert out["fwd_score"][p] == F[-1, -1] and out["rev_score"][p] == R[0, 0], tag
        thr = O.threshold(float(F[-1, -1]), 0.01)
        mask, cnt = O.nearopt_mask(F, R, O.sim(q, t), thr)
        assert out["threshold"][p] == thr and out["nearopt_count"][p] == cnt, tag
        got = ctx.fetch_pair(int(p), len(q), len(t), fwd=True, rev=True, mask=True, scores=True) if with_scores else None
        if got is None:
            # without W_SCORES the reverse score matrix is never materialised (the mask is fused into the reverse
            # pass): compare everything else
            gf = ctx.fetch_pair(int(p), len(q), len(t), fwd=True, rev=False, mask=True)
            gr = ctx.fetch_pair(int(p), len(q), len(t), fwd=False, rev=True, scores=False)
            assert_matrix_equal(tag + "F", gf["score_fwd"], F)
            assert_matrix_equal(tag + "fq", gf["prevq_fwd"], fq)
            assert_matrix_equal(tag + "ft", gf["prevt_fwd"], ft)
            assert_matrix_equal(tag + "mask", gf["nearopt"], mask)
            assert_matrix_equal(tag + "rq", gr["prevq_rev"], rq)
            assert_matrix_equal(tag + "rt", gr["prevt_rev"], rt)
        else:
            assert_matrix_equal(tag + "F", got["score_fwd"], F)
            assert_matrix_equal(tag + "R", got["score_rev"], R)
            assert_matrix_equal(tag + "fq", got["prevq_fwd"], fq)
            assert_matrix_equal(tag + "ft", got["prevt_fwd"], ft)
            assert_matrix_equal(tag + "rq", got["prevq_rev"], rq)
            assert_matrix_equal(tag + "rt", got["prevt_rev"], rt)
            assert_matrix_equal(tag + "mask", got["nearopt"], mask)
        n_unequal += 1
    assert n_unequal >= 256


def test_c2_stream_all_final_scores(ctx):
    # BASELINE.json configs[1]: all 10 000 forward scores of the seed-1002 stream against the oracle
    import alignment_algos_b200 as a
    from alignment_algos_b200 import synth
    _, M20 = a.blosum62()
    seqs, pq, pt = synth.config("c2")
    res, off = a.Context.pack(seqs)
    ctx.set_scoring(M20, 12, 1, po.SEMI_LOCAL)
    out = ctx.fill_batch(res, off, pq, pt, a.W_FWD)
    O = po.Oracle(M20, 12, 1, po.SEMI_LOCAL)
    want = np.array([O.fill(seqs[pq[p]], seqs[pt[p]], po.FWD, True, fast=True)[0][-1, -1] for p in range(len(pq))],
                    np.float32)
    assert_matrix_equal("c2 forward scores", out["fwd_score"], want)


def test_c4_stream_rectangle(ctx):
    # BASELINE.json configs[3]: a 96 x 80 rectangle of the seed-1004 all-vs-all score matrix (cross mode) against the
    # oracle, rows and columns taken from both ends of the length distribution
    import alignment_algos_b200 as a
    from alignment_algos_b200 import synth
    _, M20 = a.blosum62()
    rng = np.random.default_rng(1004)
    seqs = synth.random_seqs(rng, 20_000, 100, 500)
    res, off = a.Context.pack(seqs)
    L = np.array([len(s) for s in seqs])
    order = np.argsort(L, kind="stable")
    q_ids = np.concatenate([order[:24], order[-24:], np.arange(48)]).astype(np.int32)
    t_ids = np.concatenate([order[:20], order[-20:], np.arange(19_960, 20_000)]).astype(np.int32)
    ctx.set_scoring(M20, 12, 1, po.SEMI_LOCAL)
    got = ctx.cross_scores(res, off, q_ids, t_ids)
    O = po.Oracle(M20, 12, 1, po.SEMI_LOCAL)
    want = np.empty((len(q_ids), len(t_ids)), np.float32)
    for a_, qi in enumerate(q_ids):
        for b_, ti in enumerate(t_ids):
            want[a_, b_] = O.fill(seqs[qi], seqs[ti], po.FWD, True, fast=True)[0][-1, -1]
    assert_matrix_equal("c4 rectangle", np.asarray(got).reshape(want.shape), want)
