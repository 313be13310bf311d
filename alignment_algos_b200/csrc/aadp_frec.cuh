// aadp_frec.cuh -- exact general-gap fp32 fill with RECORD LISTS for sm_100a: the reference's default scoring
// (4.73 / 0.34, alib.cpp:17-18) and any other scoring off the dyadic grid, at a cost per cell that does not grow
// with the length of the reference's scans (dpmatrix.h:459-480).  Same results as aadp_general.cuh's literal scans,
// bit for bit (scores, every DPCell predecessor); a CPU model of the same decisions (orc_fill_rec) is part of the test suite.
//
// Idea.  In real arithmetic the order of the deletion candidates k of a cell (a,b) does not depend on b:
//     D[a-1][k] - gi - ge*(b-k-2) = (D[a-1][k] + ge*k) - const(b),          KEY(k) = D[a-1][k] + ge*k,
// and a running maximum of the key would name the winner.  In fp32 (s = D; s -= pen; s += sim with pen = gi +
// ge*(float)(len-1) rounded twice, dpmatrix.h:460-462, aasubalib.h:37-38) candidates whose keys differ by less than
// the accumulated rounding noise MU can swap places, and the strict '>' of the ascending scan lets the FIRST of
// equal fp32 values win.  Therefore
//   * k is dominated for ever when an earlier k' < k has KEY(k') > KEY(k) + MU (wherever k is a candidate k' is one
//     too and its fp32 value is strictly larger); the others are the RECORDS of the row / column;
//   * for one cell only the records within 2*MU of the running key maximum can win; walking the record list
//     backwards, the walk ends at the first record below max - 2*MU (everything before it is below max - MU);
//   * every visited record is evaluated with the reference's own three fp32 operations and the first maximum
//     (smallest k) is kept -- what the ascending strict-'>' scan does.
// Visited records per cell and scan: 1.0-1.4 on average (statistics of the CPU model), against (m+n)/2 candidates.
// MU = 2^-19 * (max|D| so far + |pen(maxlen)| + |ge|*maxlen + max|sim| + 1): eight times the sum of the error bounds
// of the key (one rounding each for ge*k and the sum), of pen (two roundings) and of the candidate's two roundings.
//
// Mapping.  Row a needs rows < a only, so ONE WARP owns a (pair, direction) and sweeps it row by row with the
// previous row in shared memory.  Lane l owns the key columns [l*K+1, l*K+K] (K = ceil(nt/32)) and computes the cells
// two columns to the right of them, so that the row state of cell (a,b) -- running key maximum and last record among
// 1..b-2 -- is the lane's own running state.  Per row: lane-local key maxima, one warp scan (exclusive prefix
// maximum), the records of every lane written to its segment of the list, then the cells.  The insertion scan of
// cell (a,b) runs over column b-1: its leader (largest key), the runner-up key and the last record live in shared
// memory per column, the record chain of a column in a dense link matrix in HBM that is only followed when the
// leader is not clear (4-10 % of the cells).  No block-wide barrier anywhere.
#pragma once
#include "aadp_general.cuh"
#include <type_traits>

namespace aadp {

// Shared memory per (padded) column: two row buffers, D of the column's two leaders and the third-best key (floats),
// rows of the two leaders, last record, record list (shorts), residue code.  `cap` = 32 * (Kmax | 1) + 8 for the widest
// template of the launch (frec_cap): lane l keeps its K columns at l*Kp .. l*Kp+K-1 with Kp = K | 1, an ODD stride, so
// that the 32 lanes of every access fall into 32 different banks (a stride of K = 16 words was a 16-way conflict).
__host__ __device__ inline int frec_cap(int max_nt) { return 32 * (((max_nt + 31) / 32) | 1) + 8; }
__host__ __device__ inline size_t frec_smem_bytes(int cap) { return (size_t)cap * (4 * 7 + 2 * 2 + 1) + 16 + 64 * 8 + 64 * 4; }

// a cell waiting for the record chain of its column (see the row loop)
struct __align__(4) FrecDeferred { float run; short b, ri; };

__device__ __forceinline__ float frec_key(float d, float ge, int k) { return __fadd_rn(d, __fmul_rn(ge, (float)k)); }

// WIDE = 1: up to 64 key columns per lane (templates of 1025..2048 residues); the record bit mask is then 64 bits wide
template <int TBM, int WIDE>
__global__ void __launch_bounds__(32) frec_fill_kernel(const GeneralParams P, int cap) {
  typedef typename std::conditional<WIDE != 0, unsigned long long, unsigned>::type recmask_t;
  extern __shared__ __align__(16) unsigned char frec_smem[];
  const int item = blockIdx.x;
  const int pair = P.items ? P.items[item] : P.item0 + item;
  const int dsel = blockIdx.y;
  const int rev = P.dirs[dsel];
  const int qs = P.pair_q[pair], ts = P.pair_t[pair];
  const int64_t qo = P.seq_off[qs], to = P.seq_off[ts];
  const int Lq = (int)(P.seq_off[qs + 1] - qo), Lt = (int)(P.seq_off[ts + 1] - to);
  int q0 = 0, t0 = 0, mq1 = Lq + 1, mt1 = Lt + 1;
  if (P.rects) { const int4 r = P.rects[blockIdx.x]; q0 = r.x; t0 = r.y; mq1 = r.z; mt1 = r.w; }
  const int sz2 = Lt + 2, q1 = mq1 - q0, t1 = mt1 - t0, nq = q1 - 1, nt = t1 - 1;
  const int ld = P.compact ? t1 + 1 : sz2, nrows = P.compact ? q1 + 1 : Lq + 2;
  const int r0 = P.compact ? q0 : 0, c0 = P.compact ? t0 : 0;
  const int fin_idx = P.fin_by_item ? P.item0 + item : pair;
  const int64_t base = P.dense_off ? P.dense_off[item] : 0;
  float* D = P.score[dsel] + base;
  int32_t* PQ = TBM ? P.prevq[dsel] + base : nullptr;
  int32_t* PT = TBM ? P.prevt[dsel] + base : nullptr;
  int32_t* LK = reinterpret_cast<int32_t*>(P.pmcol[dsel]) + base;  // record links of the columns
  const uint8_t* qseq = P.residues + qo;
  const uint8_t* tseq = P.residues + to;
  const int lane = threadIdx.x;
  const float gi = P.gi, ge = P.ge;
  const bool local = P.local != 0;

  // shared memory (indexed by the PADDED column position, see frec_cap)
  float* rowA = reinterpret_cast<float*>(frec_smem);
  float* rowB = rowA + cap;
  float* cD1 = rowB + cap;      // D of the column's leader (largest key, smallest row on ties)
  float* cD2 = cD1 + cap;       // D of the runner-up
  float* ckey3 = cD2 + cap;     // third-best key of the column
  float* cK1 = ckey3 + cap;     // rows of leader / runner-up as floats (0 = none): they only
  float* cK2 = cK1 + cap;                              // ever enter float arithmetic (key = D + ge*k, pen(len))
  float* srow_s = cK2 + cap;    // 64 entries: substitution scores of the current query residue
  short* clast = reinterpret_cast<short*>(srow_s + 64);  // last record of the column
  short* rl = clast + cap;      // records (columns) of the previous row, ascending; only the slow path reads it
  FrecDeferred* dq = reinterpret_cast<FrecDeferred*>(rl + cap);  // 64 entries
  uint8_t* tcode = reinterpret_cast<uint8_t*>(dq + 64);

  auto rowof = [&](int a) { return rev ? mq1 - a : q0 + a; };
  auto colof = [&](int b) { return rev ? mt1 - b : t0 + b; };
  auto at = [&](int a, int b) -> int64_t { return (int64_t)(rowof(a) - r0) * ld + (colof(b) - c0); };
  auto clampl = [&](float s) { return (local && s < 0.f) ? 0.f : s; };
  const float* simov = P.simov ? P.simov + base : nullptr;
  auto gdel = [&](int b0, int b1) -> float {
    const int len = b1 - b0 - 1;
    if (len < 1) return 0.f;
    const int x = colof(b0), y = colof(b1);
    if (P.delfree && (min(x, y) == 0 || max(x, y) == Lt + 1)) return 0.f;
    return gg_pen(gi, ge, len);
  };
  auto gins = [&](int a0, int a1) -> float {
    const int len = a1 - a0 - 1;
    if (len < 1) return 0.f;
    const int x = rowof(a0), y = rowof(a1);
    if (P.insfree && (min(x, y) == 0 || max(x, y) == Lq + 1)) return 0.f;
    return gg_pen(gi, ge, len);
  };
  float simf = 0.f;
  {
    const int i = rowof(q1), j = colof(t1);
    if (simov) simf = simov[(int64_t)i * sz2 + j];
    else if (i >= 1 && i <= Lq && j >= 1 && j <= Lt) simf = P.subf[(int)qseq[i - 1] * P.A + (int)tseq[j - 1]];
  }

  // DPCell::DPCell (dpmatrix.cpp:17-25) wherever the fill itself does not write: the whole storage for a
  // sub-rectangle, the four border lines for a whole matrix (every other cell is written below)
  if (P.rects) {
    for (int64_t o = lane; o < (int64_t)nrows * ld; o += 32) {
      D[o] = 0.f;
      if (TBM) { PQ[o] = -1; PT[o] = -1; }
    }
  } else {
    for (int j = lane; j < sz2; j += 32) {
      const int64_t o0 = j, o1 = (int64_t)(Lq + 1) * ld + j;
      D[o0] = 0.f; D[o1] = 0.f;
      if (TBM) { PQ[o0] = -1; PT[o0] = -1; PQ[o1] = -1; PT[o1] = -1; }
    }
    for (int i = 1 + lane; i <= Lq; i += 32) {
      const int64_t o0 = (int64_t)i * ld, o1 = o0 + Lt + 1;
      D[o0] = 0.f; D[o1] = 0.f;
      if (TBM) { PQ[o0] = -1; PT[o0] = -1; PQ[o1] = -1; PT[o1] = -1; }
    }
  }
  __syncwarp();

  auto set_tb = [&](int a, int b, int pa, int pb, float s) {  // dpmatrix.cpp:27-32
    const int64_t o = at(a, b);
    D[o] = s;
    if (TBM) { PQ[o] = rowof(pa); PT[o] = colof(pb); }
  };

  // Special cases #1/#2 (dpmatrix.h:374-390, 712-728): an empty sequence forces one gap; not clamped
  if (nq == 0 || nt == 0) {
    if (lane == 0) {
      float s = 0.f;
      s = __fsub_rn(s, nq == 0 ? gdel(0, t1) : gins(0, q1));
      s = __fadd_rn(s, simf);
      set_tb(q1, t1, 0, 0, s);
      if (P.fin[dsel]) P.fin[dsel][fin_idx] = s;
    }
    return;
  }

  // largest |similarity| the fill can meet (noise bound)
  float smax = 0.f;
  if (simov) {
    for (int i = 1; i <= nq; ++i)
      for (int b = 1 + lane; b <= nt; b += 32) smax = fmaxf(smax, fabsf(simov[(int64_t)rowof(i) * sz2 + colof(b)]));
  } else {
    for (int x = lane; x < P.A * P.A; x += 32) smax = fmaxf(smax, fabsf(P.subf[x]));
  }
  smax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(smax)));
  const int maxlen = max(nq, nt);
  const float wconst = fabsf(gg_pen(gi, ge, maxlen)) + fabsf(__fmul_rn(ge, (float)maxlen)) + smax + 1.0f;

  // lane l owns the key columns l*K+1 .. l*K+K; padded position of column b
  const int K = (nt + 31) >> 5, Kp = K | 1, pad = Kp - K;
  const unsigned kmagic = (unsigned)(0xffffffffu / (unsigned)K) + 1u;  // __umulhi(x, kmagic) == x / K for x < 2^16
  auto ph = [&](int b) -> int { return (b - 1) + (pad ? (int)__umulhi((unsigned)(b - 1), kmagic) : 0); };
  const int k0 = lane * K + 1, k1 = min(nt, k0 + K - 1), p0 = lane * Kp;
  const float NEGK = -3.0e38f;

  // residue codes of the flow columns; column structures
  for (int b = 1 + lane; b <= nt; b += 32) {
    const int p = ph(b);
    tcode[p] = simov ? 0 : tseq[colof(b) - 1];
    cD1[p] = 0.f; cD2[p] = 0.f; ckey3[p] = NEGK; cK1[p] = 0.f; cK2[p] = 0.f; clast[p] = 0;
  }
  __syncwarp();
  auto simrow = [&](int a) -> const float* {
    return simov ? simov + (int64_t)rowof(a) * sz2 : P.subf + (int)qseq[rowof(a) - 1] * P.A;
  };
  auto simat = [&](const float* srow, int b, int p) -> float { return simov ? srow[colof(b)] : srow[(int)tcode[p]]; };

  // boundary row of the flow (dpmatrix.h:408-418, 746-756, 579-589, 920-930)
  float dmaxl = 0.f;
  {
    const float* srow = simrow(1);
    for (int b = 1 + lane; b <= nt; b += 32) {
      const int p = ph(b);
      float s = 0.f;
      if (b >= 2) s = __fsub_rn(s, gdel(0, b));
      s = clampl(__fadd_rn(s, simat(srow, b, p)));
      set_tb(1, b, 0, 0, s);
      rowA[p] = s;
      dmaxl = fmaxf(dmaxl, fabsf(s));
    }
  }
  float dmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(dmaxl)));
  __syncwarp();

  float* cur = rowA;
  float* nxt = rowB;

  for (int a = 2; a <= nq; ++a) {
    const float mu = __fmul_rn(__fadd_rn(dmax, wconst), 1.0f / 524288.0f);
    const float mu2 = __fmul_rn(2.0f, mu);
    // ---- row a-1, seen from the deletion scans of row a: per lane the LEADER of its key columns (largest key,
    // smallest column on ties), the runner-up key; exclusive prefix of that summary over the lanes
    float lk1 = NEGK, ld1 = 0.f, lk2 = NEGK;
    int lc1 = 0;
    for (int k = k0, p = p0; k <= k1; ++k, ++p) {
      const float d = cur[p];
      const float key = frec_key(d, ge, k);
      if (key > lk1) { lk2 = lk1; lk1 = key; ld1 = d; lc1 = k; }
      else lk2 = fmaxf(lk2, key);
    }
    float ik1 = lk1, id1 = ld1, ik2 = lk2;  // inclusive scan: (earlier prefix) merged with (this lane)
    int ic1 = lc1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float uk1 = __shfl_up_sync(0xffffffffu, ik1, o), ud1 = __shfl_up_sync(0xffffffffu, id1, o);
      const float uk2 = __shfl_up_sync(0xffffffffu, ik2, o);
      const int uc1 = __shfl_up_sync(0xffffffffu, ic1, o);
      if (lane >= o) {
        if (ik1 > uk1) { ik2 = fmaxf(uk1, ik2); }                       // the later block keeps the lead
        else { ik2 = fmaxf(uk2, ik1); ik1 = uk1; id1 = ud1; ic1 = uc1; }  // ties: the earlier column leads
      }
    }
    float ek1 = __shfl_up_sync(0xffffffffu, ik1, 1), ed1 = __shfl_up_sync(0xffffffffu, id1, 1);
    float ek2 = __shfl_up_sync(0xffffffffu, ik2, 1);
    int ec1 = __shfl_up_sync(0xffffffffu, ic1, 1);
    if (lane == 0) { ek1 = NEGK; ed1 = 0.f; ek2 = NEGK; ec1 = 0; }
    const float pre = ek1;
    // records of this lane's columns (a bit mask: K <= 64 per lane is guaranteed by the host), their number before
    // this lane (exclusive prefix sum), and the list itself: ONE ascending list for the row, so that a walk is a
    // plain descending index.  Only the slow path (below) reads it.
    recmask_t recmask = 0;
    {
      float run = pre;
      for (int k = k0, p = p0; k <= k1; ++k, ++p) {
        const float key = frec_key(cur[p], ge, k);
        if (key >= __fsub_rn(run, mu)) recmask |= (recmask_t)1 << (k - k0);
        run = fmaxf(run, key);
      }
    }
    int rbase = WIDE ? __popcll((unsigned long long)recmask) : __popc((unsigned)recmask);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, rbase, o);
      if (lane >= o) rbase += u;
    }
    rbase -= WIDE ? __popcll((unsigned long long)recmask) : __popc((unsigned)recmask);
    int jrec = 0;
    for (recmask_t mm = recmask; mm; mm &= mm - 1, ++jrec) {
      const int o = (WIDE ? __ffsll((long long)mm) : __ffs((int)mm)) - 1;
      rl[rbase + jrec] = (short)(k0 + o);
    }
    __syncwarp();

    const float* srow = simrow(a);
    const int ra = rowof(a);
    const int64_t rowbase = (int64_t)(ra - r0) * ld - c0;
    const int64_t lkbase = (int64_t)(rowof(a - 1) - r0) * ld - c0;  // row a-1 of the link matrix
    float rowabs = 0.f;
    auto finish = [&](int bq, int pb, int oa, int ob, float os) {
      if (TBM) {
        const int64_t o = rowbase + colof(bq);
        PQ[o] = rowof(oa);
        PT[o] = colof(ob);
      }
      nxt[pb] = os;
      rowabs = fmaxf(rowabs, fabsf(os));
    };
    // column b-1 receives the candidate of row a-1 (used from row a+1 on)
    auto column_update = [&](int c, int pc, float dc, float key1, float key2, float key3, float d1, int kk1) {
      const int kc = a - 1;
      const float kk = frec_key(dc, ge, kc);
      if (kk >= __fsub_rn(key1, mu)) { LK[lkbase + colof(c)] = (int)clast[pc]; clast[pc] = (short)kc; }
      if (kk > key1) { ckey3[pc] = key2; cD2[pc] = d1; cK2[pc] = (float)kk1; cD1[pc] = dc; cK1[pc] = (float)kc; }
      else if (kk > key2) { ckey3[pc] = key2; cD2[pc] = dc; cK2[pc] = (float)kc; }
      else if (kk > key3) ckey3[pc] = kk;
    };
    // SLOW PATH: a cell whose deletion scan has two or more leaders within the noise, or whose insertion scan has three
    // or more.  Such cells (6-14 %) are queued and worked off by all lanes at once, one queued cell per lane: the record
    // walks -- through the row's list in shared memory, through the column's chain in HBM -- then cost the warp one walk
    // latency per 32 cells instead of one per loop iteration, and the common case stays free of them.
    auto slow_cell = [&](int bq, float run, int ri) {
      const int pb = ph(bq), pc = ph(bq - 1);
      const float simc = simat(srow, bq, pb);
      const float dc = cur[pc];
      int oa = a - 1, ob = bq - 1;
      float os = clampl(__fadd_rn(dc, simc));
      if (bq >= 3) {  // deletions (dpmatrix.h:459-468): records of row a-1 among 1..b-2, last one first
        const float lim = __fsub_rn(run, mu2);
        float bs = 0.f;
        int bk = 0;
        int i = ri;
        if (i >= 0) {
          int kr = (int)rl[i];
          float d = cur[ph(kr)];
          for (;;) {
            // the next record is fetched before this one is evaluated (its key ends the walk)
            --i;
            int kn = 0;
            float dn = 0.f;
            if (i >= 0) { kn = (int)rl[i]; dn = cur[ph(kn)]; }
            if (frec_key(d, ge, kr) < lim) break;
            const float sv = clampl(__fadd_rn(__fsub_rn(d, gg_pen(gi, ge, bq - kr - 1)), simc));
            if (bk == 0 || sv >= bs) { bs = sv; bk = kr; }
            if (i < 0) break;
            kr = kn;
            d = dn;
          }
        }
        if (bk && bs > os) { ob = bk; os = bs; }
      }
      const int kk1 = (int)cK1[pc], kk2 = (int)cK2[pc];
      const float d1 = cD1[pc], d2 = cD2[pc], key3 = ckey3[pc];
      const float key1 = kk1 ? frec_key(d1, ge, kk1) : NEGK, key2 = kk2 ? frec_key(d2, ge, kk2) : NEGK;
      if (a >= 3) {  // insertions (dpmatrix.h:471-480): candidate rows 1..a-2 of column b-1
        const float lim = __fsub_rn(key1, mu2);
        float bs = 0.f;
        int bk = 0;
        if (key3 < lim) {  // at most two candidates can win: both are at hand
          bk = kk1;
          bs = clampl(__fadd_rn(__fsub_rn(d1, gg_pen(gi, ge, a - kk1 - 1)), simc));
          if (kk2 && key2 >= lim) {
            const float s2 = clampl(__fadd_rn(__fsub_rn(d2, gg_pen(gi, ge, a - kk2 - 1)), simc));
            if (s2 > bs || (s2 == bs && kk2 < kk1)) { bs = s2; bk = kk2; }
          }
        } else {  // three or more within the noise: the record chain of the column (dense link matrix in HBM)
          for (int k = (int)clast[pc]; k > 0;) {
            const int64_t o = at(k, bq - 1);
            const float d = D[o];
            const int kn = LK[o];
            if (frec_key(d, ge, k) < lim) break;
            const float sv = clampl(__fadd_rn(__fsub_rn(d, gg_pen(gi, ge, a - k - 1)), simc));
            if (bk == 0 || sv >= bs) { bs = sv; bk = k; }
            k = kn;
          }
        }
        if (bk && bs > os) { oa = bk; ob = bq - 1; os = bs; }
      }
      finish(bq, pb, oa, ob, os);
      column_update(bq - 1, pc, dc, key1, key2, key3, d1, kk1);
    };
    auto flush = [&](int n) {  // the first n (<= 32) queued cells
      if (lane < n) {
        const FrecDeferred e = dq[lane];
        slow_cell((int)e.b, e.run, (int)e.ri);
      }
      __syncwarp();
    };
    int qn = 0;
    if (lane == 0) {
      // boundary column of the flow (dpmatrix.h:420-426, 758-764, 591-599, 932-940)
      float sv = 0.f;
      sv = __fsub_rn(sv, gins(0, a));
      sv = clampl(__fadd_rn(sv, simat(srow, 1, 0)));
      if (TBM) { const int64_t o = rowbase + colof(1); PQ[o] = rowof(0); PT[o] = colof(0); }
      nxt[0] = sv;
      rowabs = fmaxf(rowabs, fabsf(sv));
    }
    // FAST PATH (rows >= 3, cells b >= 3, substitution-table similarity): running leader of the keys 1..k from the
    // exclusive prefix on, one candidate per deletion scan, leader and runner-up of the column for the insertion scan.
    // Row and column indices travel as floats (exact), the column entry is rewritten without branches.
    float rk1 = ek1, rd1 = ed1, rk2 = ek2, rcf = (float)ec1;
    int ri = rbase - 1;  // index of the last record among 1..k
    const bool fast_row = a >= 3 && !simov;
    const float af1 = (float)(a - 1), am2f = (float)(a - 2), gea1 = __fmul_rn(ge, af1);
    const float floorv = local ? 0.f : -INFINITY;
    if (!simov) {
      __syncwarp();
      for (int x = lane; x < P.A; x += 32) srow_s[x] = srow[x];
      __syncwarp();
    }
    float kf = (float)k0;
    for (int j = 0; j < K; ++j, kf += 1.0f) {  // the same trip count for every lane: the queue is filled with warp votes
      const int k = k0 + j, p = p0 + j;
      int bq = 0;
      if (k <= k1) {
        const float d = cur[p];
        const float key = __fadd_rn(d, __fmul_rn(ge, kf));
        const bool lead = key > rk1;
        rk2 = lead ? rk1 : fmaxf(rk2, key);
        rd1 = lead ? d : rd1;
        rcf = lead ? kf : rcf;
        rk1 = lead ? key : rk1;
        ri += (int)((recmask >> j) & (recmask_t)1);
        bq = k + 2;
        // the lane that owns key column nt has no cell of its own there: it takes the first interior cell, b = 2
        if (bq > nt) bq = (k == nt && nt >= 2) ? 2 : 0;
      }
      bool defer = false;
      if (bq) {
        defer = true;
        if (fast_row && bq != 2) {
          // padded positions of columns b and b-1 (pad is non-zero for even K only; then b and b-1 lie at most one lane
          // segment to the right of key column k)
          const int pb = (bq - 1) + pad * (lane + (j + 2 >= K ? 1 : 0));
          const int pc = (bq - 2) + pad * (lane + (j + 1 >= K ? 1 : 0));
          const float d1 = cD1[pc], k1f = cK1[pc], d2 = cD2[pc], k2f = cK2[pc], key3 = ckey3[pc];
          const float key1 = __fadd_rn(d1, __fmul_rn(ge, k1f));  // rows >= 3: the column has a leader
          const float key2 = k2f > 0.f ? __fadd_rn(d2, __fmul_rn(ge, k2f)) : NEGK;
          const float climit = __fsub_rn(key1, mu2);
          if (rk2 < __fsub_rn(rk1, mu2) && (key3 < climit)) {
            defer = false;
            const float simc = srow_s[tcode[pb]];
            const float dc = cur[pc];  // D[a-1][b-1]: the match predecessor, and the new candidate of column b-1
            float os = fmaxf(__fadd_rn(dc, simc), floorv);
            // deletions (dpmatrix.h:459-468): the one leader of row a-1 among 1..b-2; len - 1 = b - 2 - column = k - column
            const float sv = fmaxf(__fadd_rn(__fsub_rn(rd1, __fadd_rn(gi, __fmul_rn(ge, __fsub_rn(kf, rcf)))), simc), floorv);
            const bool wdel = sv > os;
            os = wdel ? sv : os;
            // insertions (dpmatrix.h:471-480): leader and runner-up of column b-1; len - 1 = a - 2 - row
            float bs = fmaxf(__fadd_rn(__fsub_rn(d1, __fadd_rn(gi, __fmul_rn(ge, __fsub_rn(am2f, k1f)))), simc), floorv);
            float bkf = k1f;
            if (key2 >= climit) {  // (a runner-up exists: key2 is NEGK otherwise)
              const float s2 = fmaxf(__fadd_rn(__fsub_rn(d2, __fadd_rn(gi, __fmul_rn(ge, __fsub_rn(am2f, k2f)))), simc), floorv);
              const bool take2 = s2 > bs || (s2 == bs && k2f < k1f);
              bs = take2 ? s2 : bs;
              bkf = take2 ? k2f : bkf;
            }
            const bool wins = bs > os;
            os = wins ? bs : os;
            if (TBM) {
              const int64_t o = rowbase + colof(bq);
              PQ[o] = rowof(wins ? (int)bkf : a - 1);
              PT[o] = colof((wins || !wdel) ? bq - 1 : (int)rcf);
            }
            nxt[pb] = os;
            rowabs = fmaxf(rowabs, fabsf(os));
            // column b-1 receives the candidate of row a-1 (used from row a+1 on)
            const float kk = __fadd_rn(dc, gea1);
            if (kk >= __fsub_rn(key1, mu)) { LK[lkbase + colof(bq - 1)] = (int)clast[pc]; clast[pc] = (short)(a - 1); }
            const bool n1 = kk > key1, n2 = !n1 && kk > key2, n3 = !n1 && !n2 && kk > key3;
            cD1[pc] = n1 ? dc : d1;
            cK1[pc] = n1 ? af1 : k1f;
            cD2[pc] = n1 ? d1 : (n2 ? dc : d2);
            cK2[pc] = n1 ? k1f : (n2 ? af1 : k2f);
            ckey3[pc] = (n1 || n2) ? key2 : (n3 ? kk : key3);
          }
        }
      }
      const unsigned dm = __ballot_sync(0xffffffffu, defer);
      if (dm) {
        if (defer) {
          FrecDeferred de;
          de.run = rk1; de.b = (short)bq; de.ri = (short)ri;
          dq[qn + __popc(dm & ((1u << lane) - 1u))] = de;
        }
        qn += __popc(dm);
        __syncwarp();
        if (qn >= 32) {
          flush(32);
          const int rest = qn - 32;
          FrecDeferred mv;
          if (lane < rest) mv = dq[32 + lane];
          __syncwarp();
          if (lane < rest) dq[lane] = mv;
          __syncwarp();
          qn = rest;
        }
      }
    }
    if (qn) flush(qn);
    rowabs = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(rowabs)));
    dmax = fmaxf(dmax, rowabs);
    __syncwarp();
    // the row goes to HBM in one coalesced sweep
    for (int b = 1 + lane; b <= nt; b += 32) D[rowbase + colof(b)] = nxt[ph(b)];
    float* tsw = cur; cur = nxt; nxt = tsw;
  }

  // final cell (dpmatrix.h:504-534, 844-874, 654-687, 995-1028): match, bottom row (k ascending), right column (k
  // ascending), strict '>'.  Candidates are evaluated by all lanes; the first maximum in that order wins.
  __syncwarp();
  {
    float bs = -3.4e38f;
    int bi = 0x7fffffff;
    for (int k = 1 + lane; k < t1; k += 32) {
      const float s = clampl(__fadd_rn(__fsub_rn(cur[ph(k)], gdel(k, t1)), simf));
      if (s > bs) { bs = s; bi = k; }
    }
    for (int k = 1 + lane; k < q1; k += 32) {
      const float d = (k == nq) ? cur[ph(nt)] : D[at(k, nt)];
      const float s = clampl(__fadd_rn(__fsub_rn(d, gins(k, q1)), simf));
      if (s > bs) { bs = s; bi = t1 - 1 + k; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (s2 > bs || (s2 == bs && i2 < bi)) { bs = s2; bi = i2; }
    }
    if (lane == 0) {
      int oa = nq, ob = nt;
      bool from_col = false;
      float os = clampl(__fadd_rn(cur[ph(nt)], simf));
      if (bi != 0x7fffffff && bs > os) {
        os = bs;
        if (bi <= t1 - 1) { oa = nq; ob = bi; }
        else { oa = bi - (t1 - 1); ob = nt; from_col = true; }
      }
      set_tb(q1, t1, oa, ob, os);
      // dpmatrix.h:868: the global reverse fill records opt_j = t1_m1 (a matrix column) for left-column candidates
      if (TBM && rev && !local && P.repro_rev_bug && from_col) PT[at(q1, t1)] = mt1 - 1;
      if (P.fin[dsel]) P.fin[dsel][fin_idx] = os;
    }
  }
}

}  // namespace aadp
