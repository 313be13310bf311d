/* oracle/aadp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See aadp_oracle.h.
 *
 * Plain-C restatement of the reference DP path. All four reference fills (fwd/rev x
 * global/local, dpmatrix.h:356-1030) are one routine here, written in "flow" coordinates:
 * a flow cell (a,b) is matrix cell (a,b) for the forward fill and (q1-a, t1-b) for the
 * reverse fill, so the anchor is always flow (0,0), the final cell is flow (q1,t1) and every
 * scan runs over ascending flow index -- which is ascending k in the forward fill
 * (dpmatrix.h:459,471) and descending k in the reverse fill (dpmatrix.h:798,810).
 *
 * Float arithmetic keeps the reference's operation order  s = D; s -= gap; s += sim
 * (dpmatrix.h:460-462); compile with -ffp-contract=off (oracle/Makefile).
 */
#include "aadp_oracle.h"

#include <stdlib.h>
#include <string.h>

#define ORC_NULL (-1) /* DPCell::null, dpmatrix.cpp:15 */

/* ---------------------------------------------------------------- evaluator (aasubalib.h) */

static int del_free(int at) { /* aasubalib.h:39-42: case local, semi_local, local_global */
  return at == ORC_LOCAL || at == ORC_SEMI_LOCAL || at == ORC_LOCAL_GLOBAL;
}
static int ins_free(int at) { /* aasubalib.h:65-68: case local, semi_local, global_local */
  return at == ORC_LOCAL || at == ORC_SEMI_LOCAL || at == ORC_GLOBAL_LOCAL;
}

static float affine(const orc_scoring* sc, int len) { /* aasubalib.h:37-38 */
  return sc->gi + sc->ge * (float)(len - 1);
}

/* aasubalib.h:27-51. t_pos1 < t_pos2 are matrix columns; column 0 is Head, sz2-1 is Tail. */
float orc_deletion(const orc_scoring* sc, int sz2, int t_pos1, int t_pos2) {
  int len = t_pos2 - t_pos1 - 1;
  if (len < 1) return 0.f;
  if (del_free(sc->align_type) && (t_pos1 == 0 || t_pos2 == sz2 - 1)) return 0.f;
  return affine(sc, len);
}

/* aasubalib.h:53-77 */
float orc_insertion(const orc_scoring* sc, int sz1, int q_pos1, int q_pos2) {
  int len = q_pos2 - q_pos1 - 1;
  if (len < 1) return 0.f;
  if (ins_free(sc->align_type) && (q_pos1 == 0 || q_pos2 == sz1 - 1)) return 0.f;
  return affine(sc, len);
}

/* simmatrix.h:52-72 + aasubalib.h:17-25: borders 0, interior = substitution score */
static void build_sim(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
                      float* sim) {
  int sz1 = Lq + 2, sz2 = Lt + 2;
  memset(sim, 0, sizeof(float) * (size_t)sz1 * sz2);
  for (int i = 1; i <= Lq; ++i)
    for (int j = 1; j <= Lt; ++j)
      sim[(size_t)i * sz2 + j] = sc->sub[(int)q[i - 1] * sc->A + (int)t[j - 1]];
}

/* ---------------------------------------------------------------- flow-coordinate helpers */

typedef struct {
  int rev, q0, t0; /* anchors of the filled rectangle (dpmatrix.h:319-353); (0,0) for a whole matrix */
  int q1, t1, sz1, sz2, local; /* q1/t1 = FLOW index of the final cell = matrix q1-q0 / t1-t0 */
  int mq1, mt1;                /* matrix row / column of the far anchor */
  const orc_scoring* sc;
  const float* sim;
  float* D;
  int *pq, *pt;
  /* tabulated gap model (orc_fill_tab): any Evaluator whose deletion() ignores the query positions and whose
   * insertion() sees the query only through q_pos2-q_pos1 away from the Head/Tail (hmap_eval.h:63-117,
   * gn2_eval.h:99-158, aasubalib.h:27-77 all do).  NULL = the affine model of sc.                          */
  const float* del_tab; /* sz2*sz2:  [t_pos1*sz2 + t_pos2] = deletion(.,.,t_pos1,t_pos2), t_pos1 < t_pos2 */
  const float* ins_tab; /* sz1*sz2:  [(q_pos2-q_pos1-1)*sz2 + t_pos2] = insertion(q_pos1,q_pos2,t_pos2-1,t_pos2) */
} flow_t;

static int rowof(const flow_t* f, int a) { return f->rev ? f->mq1 - a : f->q0 + a; }
static int colof(const flow_t* f, int b) { return f->rev ? f->mt1 - b : f->t0 + b; }
static size_t at(const flow_t* f, int a, int b) { return (size_t)rowof(f, a) * f->sz2 + colof(f, b); }

/* gap between flow columns b0 < b1 (deletion) / flow rows a0 < a1 (insertion) */
static float gdel(const flow_t* f, int b0, int b1) {
  int x = colof(f, b0), y = colof(f, b1);
  if (f->del_tab) return f->del_tab[(size_t)(x < y ? x : y) * f->sz2 + (x < y ? y : x)];
  return orc_deletion(f->sc, f->sz2, x < y ? x : y, x < y ? y : x);
}
/* b = flow column of the cell that takes the gap: the insertion is evaluated between the template positions
 * (j-1, j) in the forward fill (dpmatrix.h:473) and (j, j+1) in the reverse fill (dpmatrix.h:812)           */
static float gins(const flow_t* f, int a0, int a1, int b) {
  int x = rowof(f, a0), y = rowof(f, a1);
  if (f->ins_tab) return f->ins_tab[(size_t)(a1 - a0 - 1) * f->sz2 + (f->rev ? colof(f, b) + 1 : colof(f, b))];
  return orc_insertion(f->sc, f->sz1, x < y ? x : y, x < y ? y : x);
}
static float clampl(const flow_t* f, float s) { /* dpmatrix.h:580 etc.: s = max(0.f,s) */
  return (f->local && s < 0.f) ? 0.f : s;
}
static void set_tb(const flow_t* f, int a, int b, int pa, int pb, float s) { /* dpmatrix.cpp:27-32 */
  size_t o = at(f, a, b);
  f->D[o] = s;
  f->pq[o] = rowof(f, pa);
  f->pt[o] = colof(f, pb);
}

static int flow_init(flow_t* f, const uint8_t* q, int Lq, const uint8_t* t, int Lt,
                     const orc_scoring* sc, int direction, float* score, int* prev_q, int* prev_t,
                     float* sim) {
  f->rev = (direction == ORC_REV);
  f->sz1 = Lq + 2;
  f->sz2 = Lt + 2;
  f->q0 = f->t0 = 0;
  f->mq1 = f->sz1 - 1;
  f->mt1 = f->sz2 - 1;
  f->q1 = f->sz1 - 1;
  f->t1 = f->sz2 - 1;
  f->local = (sc->align_type == ORC_LOCAL); /* dpmatrix.h:155 */
  f->sc = sc;
  f->del_tab = f->ins_tab = 0;
  f->sim = sim;
  f->D = score;
  f->pq = prev_q;
  f->pt = prev_t;
  build_sim(q, Lq, t, Lt, sc, sim);
  size_t n = (size_t)f->sz1 * f->sz2;
  for (size_t o = 0; o < n; ++o) { /* DPCell::DPCell, dpmatrix.cpp:17-25 */
    score[o] = 0.f;
    prev_q[o] = ORC_NULL;
    prev_t[o] = ORC_NULL;
  }
  return 0;
}

/* Special cases #1/#2 (dpmatrix.h:374-390, 712-728): one sequence is empty. Not clamped in
 * the local variants either (dpmatrix.h:557-573). Returns 1 if handled. */
static int degenerate(flow_t* f) {
  float s;
  if (f->q1 == 1) {
    s = f->D[at(f, 0, 0)];
    s -= gdel(f, 0, f->t1);
    s += f->sim[at(f, f->q1, f->t1)];
    set_tb(f, f->q1, f->t1, 0, 0, s);
    return 1;
  }
  if (f->t1 == 1) {
    s = f->D[at(f, 0, 0)];
    s -= gins(f, 0, f->q1, f->t1);
    s += f->sim[at(f, f->q1, f->t1)];
    set_tb(f, f->q1, f->t1, 0, 0, s);
    return 1;
  }
  return 0;
}

/* boundary row/column of the flow (dpmatrix.h:408-426, 746-764, 579-599, 920-940) */
static void boundary(flow_t* f) {
  float s0 = f->D[at(f, 0, 0)], s;
  s = s0 + f->sim[at(f, 1, 1)];
  set_tb(f, 1, 1, 0, 0, clampl(f, s));
  for (int b = 2; b < f->t1; ++b) {
    s = s0;
    s -= gdel(f, 0, b);
    s += f->sim[at(f, 1, b)];
    set_tb(f, 1, b, 0, 0, clampl(f, s));
  }
  for (int a = 2; a < f->q1; ++a) {
    s = s0;
    s -= gins(f, 0, a, 1);
    s += f->sim[at(f, a, 1)];
    set_tb(f, a, 1, 0, 0, clampl(f, s));
  }
}

/* final cell (dpmatrix.h:504-534, 844-874, 654-687, 995-1028) */
static void final_cell(flow_t* f, int repro_rev_bug) {
  int q1 = f->q1, t1 = f->t1;
  float simf = f->sim[at(f, q1, t1)], s;
  int oa = q1 - 1, ob = t1 - 1, from_col = 0;
  float os = clampl(f, f->D[at(f, oa, ob)] + simf);
  for (int k = 1; k < t1; ++k) {
    s = f->D[at(f, q1 - 1, k)];
    s -= gdel(f, k, t1);
    s += simf;
    s = clampl(f, s);
    if (s > os) { oa = q1 - 1; ob = k; os = s; }
  }
  for (int k = 1; k < q1; ++k) {
    s = f->D[at(f, k, t1 - 1)];
    s -= gins(f, k, q1, t1);
    s += simf;
    s = clampl(f, s);
    if (s > os) { oa = k; ob = t1 - 1; os = s; from_col = 1; }
  }
  set_tb(f, q1, t1, oa, ob, os);
  /* dpmatrix.h:868: the global reverse fill records opt_j = t1_m1 (a matrix column) for the
   * left-column candidates instead of t0_p1. The local variant is correct (dpmatrix.h:1022). */
  if (f->rev && !f->local && repro_rev_bug && from_col) f->pt[at(f, q1, t1)] = f->mt1 - 1;
}

/* ---------------------------------------------------------------- literal O(n^3) fill */

static void literal_fill(flow_t* fp, int repro_rev_bug) {
  flow_t f = *fp;
  const float* sim = f.sim;
  float* score = f.D;
  if (!degenerate(&f)) {
    boundary(&f);
    for (int a = 2; a < f.q1; ++a) {
      for (int b = 2; b < f.t1; ++b) {
        float simc = sim[at(&f, a, b)], s;
        int oa = a - 1, ob = b - 1;
        float os = clampl(&f, score[at(&f, oa, ob)] + simc); /* match, :453-456 */
        for (int k = 1; k < b - 1; ++k) {                     /* deletions, :459-468 */
          s = score[at(&f, a - 1, k)];
          s -= gdel(&f, k, b);
          s += simc;
          s = clampl(&f, s);
          if (s > os) { oa = a - 1; ob = k; os = s; }
        }
        for (int k = 1; k < a - 1; ++k) {                     /* insertions, :471-480 */
          s = score[at(&f, k, b - 1)];
          s -= gins(&f, k, a, b);
          s += simc;
          s = clampl(&f, s);
          if (s > os) { oa = k; ob = b - 1; os = s; }
        }
        set_tb(&f, a, b, oa, ob, os);
      }
    }
    final_cell(&f, repro_rev_bug);
  }
}

int orc_fill(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
             int direction, int repro_rev_bug, float* score, int* prev_q, int* prev_t,
             float* sim_out) {
  if (Lq < 0 || Lt < 0) return 1;
  flow_t f;
  size_t n = (size_t)(Lq + 2) * (Lt + 2);
  float* sim = sim_out ? sim_out : (float*)malloc(sizeof(float) * n);
  flow_init(&f, q, Lq, t, Lt, sc, direction, score, prev_q, prev_t, sim);
  /* dpmatrix.h:306-307: both corner scores are zeroed; bounds check :360 */
  if (f.q1 <= 0 || f.t1 <= 0) { if (!sim_out) free(sim); return 1; }
  literal_fill(&f, repro_rev_bug);
  if (!sim_out) free(sim);
  return 0;
}

/* The same literal fill for ANY evaluator, described by what the fill asks of it: the similarity matrix
 * (simmatrix.h:40-73) and the two gap functions tabulated over every argument combination dpmatrix.h:356-1030
 * can pass (see flow_t).                                                                                   */
int orc_fill_tab(const float* sim, int Lq, int Lt, const float* del_tab, const float* ins_tab, int is_local,
                 int direction, int repro_rev_bug, float* score, int* prev_q, int* prev_t) {
  if (Lq < 0 || Lt < 0) return 1;
  flow_t f;
  f.rev = (direction == ORC_REV);
  f.sz1 = Lq + 2;
  f.sz2 = Lt + 2;
  f.q0 = f.t0 = 0;
  f.mq1 = f.q1 = f.sz1 - 1;
  f.mt1 = f.t1 = f.sz2 - 1;
  f.local = is_local;
  f.sc = 0;
  f.del_tab = del_tab;
  f.ins_tab = ins_tab;
  f.sim = sim;
  f.D = score;
  f.pq = prev_q;
  f.pt = prev_t;
  size_t n = (size_t)f.sz1 * f.sz2;
  for (size_t o = 0; o < n; ++o) {
    score[o] = 0.f;
    prev_q[o] = ORC_NULL;
    prev_t[o] = ORC_NULL;
  }
  literal_fill(&f, repro_rev_bug);
  return 0;
}

/* build_subdpm (dpmatrix.h:319-353): the same literal fill restricted to the rectangle between the anchors
 * (q1_end,t1_end) and (q2_beg,t2_beg); both anchor scores are zeroed (:333-334), cells outside the rectangle
 * keep the DPCell defaults.  The anchors are ordinary residues unless they are the Head / the Tail, so the
 * free end gaps of aasubalib.h apply only there and the final cell adds its real similarity.             */
int orc_fill_sub(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
                 int direction, int repro_rev_bug, int q1_end, int t1_end, int q2_beg, int t2_beg,
                 float* score, int* prev_q, int* prev_t) {
  if (Lq < 0 || Lt < 0) return 1;
  flow_t f;
  size_t n = (size_t)(Lq + 2) * (Lt + 2);
  float* sim = (float*)malloc(sizeof(float) * n);
  flow_init(&f, q, Lq, t, Lt, sc, direction, score, prev_q, prev_t, sim);
  if (q1_end < 0 || t1_end < 0 || q2_beg > Lq + 1 || t2_beg > Lt + 1 || q2_beg <= q1_end || t2_beg <= t1_end) {
    free(sim);
    return 1; /* "Illegal bounds building DPM" (dpmatrix.h:360) or out of the matrix */
  }
  f.q0 = q1_end;
  f.t0 = t1_end;
  f.mq1 = q2_beg;
  f.mt1 = t2_beg;
  f.q1 = q2_beg - q1_end;
  f.t1 = t2_beg - t1_end;
  if (!degenerate(&f)) {
    boundary(&f);
    for (int a = 2; a < f.q1; ++a) {
      for (int b = 2; b < f.t1; ++b) {
        float simc = sim[at(&f, a, b)], s;
        int oa = a - 1, ob = b - 1;
        float os = clampl(&f, score[at(&f, oa, ob)] + simc);
        for (int k = 1; k < b - 1; ++k) {
          s = score[at(&f, a - 1, k)];
          s -= gdel(&f, k, b);
          s += simc;
          s = clampl(&f, s);
          if (s > os) { oa = a - 1; ob = k; os = s; }
        }
        for (int k = 1; k < a - 1; ++k) {
          s = score[at(&f, k, b - 1)];
          s -= gins(&f, k, a, b);
          s += simc;
          s = clampl(&f, s);
          if (s > os) { oa = k; ob = b - 1; os = s; }
        }
        set_tb(&f, a, b, oa, ob, os);
      }
    }
    final_cell(&f, repro_rev_bug);
  }
  free(sim);
  return 0;
}

/* ---------------------------------------------------------------- exact O(n^2) fill */

int orc_fill_fast(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
                  int direction, int repro_rev_bug, float* score, int* prev_q, int* prev_t) {
  if (Lq < 0 || Lt < 0) return 1;
  flow_t f;
  size_t n = (size_t)(Lq + 2) * (Lt + 2);
  float* sim = (float*)malloc(sizeof(float) * n);
  flow_init(&f, q, Lq, t, Lt, sc, direction, score, prev_q, prev_t, sim);
  if (f.q1 <= 0 || f.t1 <= 0) { free(sim); return 1; }
  if (!degenerate(&f)) {
    boundary(&f);
    /* colorg[b] = flow row k (<= a-2) maximising D[k][b] - w(a-k-1), smallest k on ties */
    int* colorg = (int*)malloc(sizeof(int) * (size_t)(f.t1 + 1));
    for (int b = 0; b <= f.t1; ++b) colorg[b] = 0;
    for (int a = 2; a < f.q1; ++a) {
      int roworg = 0; /* flow column k (<= b-2) maximising D[a-1][k] - w(b-k-1) */
      for (int b = 2; b < f.t1; ++b) {
        float simc = sim[at(&f, a, b)], s;
        int oa = a - 1, ob = b - 1;
        float os = clampl(&f, score[at(&f, oa, ob)] + simc);
        if (b >= 3) { /* row running maximum: extend the old origin or open at b-2 */
          int kn = b - 2;
          if (roworg == 0) roworg = kn;
          else {
            float ext = score[at(&f, a - 1, roworg)] - gdel(&f, roworg, b);
            float opn = score[at(&f, a - 1, kn)] - gdel(&f, kn, b);
            if (opn > ext) roworg = kn; /* ties keep the smaller k, as the ascending strict-> scan */
          }
          s = score[at(&f, a - 1, roworg)];
          s -= gdel(&f, roworg, b);
          s += simc;
          s = clampl(&f, s);
          if (s > os) { oa = a - 1; ob = roworg; os = s; }
        }
        if (a >= 3) { /* column running maximum for column b-1 */
          int kn = a - 2, ko = colorg[b - 1];
          if (ko == 0) ko = kn;
          else {
            float ext = score[at(&f, ko, b - 1)] - gins(&f, ko, a, b);
            float opn = score[at(&f, kn, b - 1)] - gins(&f, kn, a, b);
            if (opn > ext) ko = kn;
          }
          colorg[b - 1] = ko;
          s = score[at(&f, ko, b - 1)];
          s -= gins(&f, ko, a, b);
          s += simc;
          s = clampl(&f, s);
          if (s > os) { oa = ko; ob = b - 1; os = s; }
        }
        set_tb(&f, a, b, oa, ob, os);
      }
    }
    free(colorg);
    final_cell(&f, repro_rev_bug);
  }
  free(sim);
  return 0;
}

/* ---------------------------------------------------------------- exact output-sensitive fill for ANY fp32 scoring
 *
 * CPU model of the GPU's record-list kernel (csrc/aadp_frec.cuh); same arithmetic, same decisions.
 *
 * In real arithmetic the order of the deletion candidates k of a cell (a,b) does not depend on b:
 *   D[a-1][k] - gi - ge*(b-k-2) = (D[a-1][k] + ge*k) - const(b),      KEY(k) = D[a-1][k] + ge*k,
 * so a running maximum of the key would do.  In fp32 (dpmatrix.h:460-462: s = D; s -= pen; s += sim, pen = gi +
 * ge*(len-1) rounded twice) candidates whose keys differ by less than the accumulated rounding noise MU can swap
 * places, and the strict '>' of the ascending scan (:463) lets the FIRST of equal fp32 values win.  Hence:
 *   * a candidate k is DOMINATED for ever when an earlier k' < k has KEY(k') > KEY(k) + MU: wherever k is a candidate
 *     k' is one too and its fp32 value is strictly larger.  The others -- KEY(k) >= max(KEY(1..k-1)) - MU -- are the
 *     RECORDS of the row (column); only they can ever win;
 *   * for a given cell only the records within 2*MU of the running key maximum can win: walking the record list
 *     backwards, the walk stops at the first record whose key is below max - 2*MU (everything before it is below
 *     max - MU);
 *   * the visited records are evaluated with the reference's own three fp32 operations and the first maximum
 *     (smallest k) is taken, exactly as the ascending strict-'>' scan would.
 * Cost per cell = size of the group of noise-tied leaders (1-2 on average), not the length of the scan.
 * MU bounds: |key error| + |pen error| + the three roundings of a candidate value, all <= 2^-22 * W with
 * W = max|D| so far + |pen(maxlen)| + ge*maxlen + max|sim| + 1; MU = 2^-19 * W (8x safety).
 * stats[0..3] = cells, row-walk steps, column-walk steps, cells whose column leader was ambiguous.               */
int orc_fill_rec(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc, int direction,
                 int repro_rev_bug, float* score, int* prev_q, int* prev_t, long* stats) {
  if (Lq < 0 || Lt < 0) return 1;
  flow_t f;
  size_t n = (size_t)(Lq + 2) * (Lt + 2);
  float* sim = (float*)malloc(sizeof(float) * n);
  flow_init(&f, q, Lq, t, Lt, sc, direction, score, prev_q, prev_t, sim);
  if (f.q1 <= 0 || f.t1 <= 0) { free(sim); return 1; }
  long st_cells = 0, st_row = 0, st_col = 0, st_amb = 0;
  if (!degenerate(&f)) {
    boundary(&f);
    const int nq = f.q1 - 1, nt = f.t1 - 1;
    float smax = 0.f;
    for (int x = 0; x < sc->A * sc->A; ++x) { float v = sc->sub[x] < 0 ? -sc->sub[x] : sc->sub[x]; if (v > smax) smax = v; }
    const int maxlen = nq > nt ? nq : nt;
    float pmaxabs = affine(sc, maxlen); if (pmaxabs < 0) pmaxabs = -pmaxabs;
    float gemax = sc->ge * (float)maxlen; if (gemax < 0) gemax = -gemax;
    const float wconst = pmaxabs + gemax + smax + 1.0f;
    const float ge = sc->ge;
    const float NEGK = -3.0e38f;
    /* row structures of the previous row */
    float* key = (float*)malloc(sizeof(float) * (size_t)(nt + 2));
    float* pmin = (float*)malloc(sizeof(float) * (size_t)(nt + 2)); /* inclusive prefix maximum of key */
    int* cnt = (int*)malloc(sizeof(int) * (size_t)(nt + 2));       /* records among 1..k */
    int* rl = (int*)malloc(sizeof(int) * (size_t)(nt + 2));
    /* column structures: leader (largest key, smallest row on ties), runner-up key, last record, record links */
    float* ckey = (float*)malloc(sizeof(float) * (size_t)(nt + 2));
    float* c2key = (float*)malloc(sizeof(float) * (size_t)(nt + 2));
    float* cD = (float*)malloc(sizeof(float) * (size_t)(nt + 2));
    int* ck = (int*)malloc(sizeof(int) * (size_t)(nt + 2));
    int* clast = (int*)malloc(sizeof(int) * (size_t)(nt + 2));
    int* link = (int*)calloc((size_t)(nq + 2) * (nt + 2), sizeof(int)); /* link[k][c] = record of column c before row k */
    for (int c = 0; c <= nt + 1; ++c) { ckey[c] = NEGK; c2key[c] = NEGK; cD[c] = 0.f; ck[c] = 0; clast[c] = 0; }
    float dmax = 0.f;
    for (int b = 1; b <= nt; ++b) { float v = score[at(&f, 1, b)]; v = v < 0 ? -v : v; if (v > dmax) dmax = v; }
    for (int a = 2; a <= nq; ++a) {
      /* hand-over of row a-1: keys, prefix maxima, records (uses max|D| over rows <= a-1) */
      { float v = score[at(&f, a - 1, 1)]; v = v < 0 ? -v : v; if (v > dmax) dmax = v; }
      const float mu = (dmax + wconst) * (1.0f / 524288.0f);
      float run = NEGK; int nrec = 0;
      for (int k = 1; k <= nt; ++k) {
        float kk = score[at(&f, a - 1, k)] + ge * (float)k;
        key[k] = kk;
        if (kk >= run - mu) rl[nrec++] = k;
        if (kk > run) run = kk;
        pmin[k] = run;
        cnt[k] = nrec;
      }
      float rowabs = 0.f;
      for (int b = 2; b <= nt; ++b) {
        float simc = sim[at(&f, a, b)], s;
        int oa = a - 1, ob = b - 1;
        float os = clampl(&f, score[at(&f, oa, ob)] + simc);
        st_cells++;
        if (b >= 3) { /* deletions: records of row a-1 among 1..b-2, from the last one backwards */
          const float lim = pmin[b - 2] - 2.0f * mu;
          float bs = 0.f; int bk = 0;
          for (int i = cnt[b - 2] - 1; i >= 0; --i) {
            const int k = rl[i];
            if (key[k] < lim) break;
            st_row++;
            s = score[at(&f, a - 1, k)]; s -= gdel(&f, k, b); s += simc; s = clampl(&f, s);
            if (bk == 0 || s >= bs) { bs = s; bk = k; }
          }
          if (bk && bs > os) { oa = a - 1; ob = bk; os = bs; }
        }
        if (a >= 3) { /* insertions: column b-1, candidate rows 1..a-2 */
          const int c = b - 1;
          float bs; int bk;
          if (c2key[c] < ckey[c] - 2.0f * mu) { /* a clear leader */
            bk = ck[c];
            s = cD[c]; s -= gins(&f, bk, a, b); s += simc; bs = clampl(&f, s);
            st_col++;
          } else {
            st_amb++;
            const float lim = ckey[c] - 2.0f * mu;
            bs = 0.f; bk = 0;
            for (int k = clast[c]; k > 0; k = link[(size_t)k * (nt + 2) + c]) {
              const float dk = score[at(&f, k, c)];
              if (dk + ge * (float)k < lim) break;
              st_col++;
              s = dk; s -= gins(&f, k, a, b); s += simc; s = clampl(&f, s);
              if (bk == 0 || s >= bs) { bs = s; bk = k; }
            }
          }
          if (bk && bs > os) { oa = bk; ob = b - 1; os = bs; }
        }
        set_tb(&f, a, b, oa, ob, os);
        { float v = os < 0 ? -os : os; if (v > rowabs) rowabs = v; }
        /* column b-1 receives the candidate of row a-1 (used from row a+1 on) */
        {
          const int c = b - 1, k = a - 1;
          const float dk = score[at(&f, k, c)];
          const float kk = dk + ge * (float)k;
          if (kk >= ckey[c] - mu) { link[(size_t)k * (nt + 2) + c] = clast[c]; clast[c] = k; }
          if (kk > ckey[c]) { c2key[c] = ckey[c]; ckey[c] = kk; ck[c] = k; cD[c] = dk; }
          else if (kk > c2key[c]) c2key[c] = kk;
        }
      }
      if (rowabs > dmax) dmax = rowabs;
    }
    free(key); free(pmin); free(cnt); free(rl); free(ckey); free(c2key); free(cD); free(ck); free(clast); free(link);
    final_cell(&f, repro_rev_bug);
  }
  if (stats) { stats[0] = st_cells; stats[1] = st_row; stats[2] = st_col; stats[3] = st_amb; }
  free(sim);
  return 0;
}

/* ---------------------------------------------------------------- optimal tracebacks */

/* optimal.h:47-124 */
int orc_optimal_fwd(const float* score, const int* prev_q, const int* prev_t, int sz1, int sz2,
                    int is_local, int* pairs, int max_pairs, int* npairs, float* ali_score) {
  /* the list is built back-to-front with prepend(); collect reversed, then flip */
  int n = 0, ql = sz1 - 1, tl = sz2 - 1, rc = 0;
  int cap = sz1 + sz2 + 4;
  int* tmp = (int*)malloc(sizeof(int) * 2 * (size_t)cap);
#define PUSHF(a, b) do { if (n < cap) { tmp[2 * n] = (a); tmp[2 * n + 1] = (b); } ++n; } while (0)
  if (!is_local) {
    *ali_score = score[(size_t)ql * sz2 + tl];
    PUSHF(ql, tl);
    while (ql > 0) { /* optimal.h:66-71 */
      size_t o = (size_t)ql * sz2 + tl;
      int nq = prev_q[o], nt = prev_t[o];
      ql = nq; tl = nt;
      PUSHF(ql, tl);
      if (ql < 0 || tl < 0 || n > cap) break;
    }
    if (ql != 0 || tl != 0) rc = 3; /* optimal.h:74 */
  } else {
    PUSHF(ql, tl);
    /* find_max, optimal.h:106-124: seed at (sz1-2,sz2-2), row-major scan, strict < */
    int mq = sz1 - 2, mt = sz2 - 2;
    float s = score[(size_t)mq * sz2 + mt];
    for (int i = 0; i < sz1 - 1; ++i)
      for (int j = 0; j < sz2 - 1; ++j)
        if (s < score[(size_t)i * sz2 + j]) { mq = i; mt = j; s = score[(size_t)i * sz2 + j]; }
    ql = mq; tl = mt;
    *ali_score = s;
    PUSHF(ql, tl);
    while (ql > 0) { /* optimal.h:96-102 */
      size_t o = (size_t)ql * sz2 + tl;
      int nq = prev_q[o], nt = prev_t[o];
      ql = nq; tl = nt;
      if (ql < 0 || tl < 0) break;
      if (score[(size_t)ql * sz2 + tl] <= 0.f) break;
      PUSHF(ql, tl);
    }
    if (ql != 0 && tl != 0) PUSHF(0, 0); /* optimal.h:104 */
  }
#undef PUSHF
  int m = n < cap ? n : cap;
  for (int k = 0; k < m && k < max_pairs; ++k) {
    pairs[2 * k] = tmp[2 * (m - 1 - k)];
    pairs[2 * k + 1] = tmp[2 * (m - 1 - k) + 1];
  }
  *npairs = m;
  free(tmp);
  return rc;
}

/* optimal_rev.h:47-131 */
int orc_optimal_rev(const float* score, const int* prev_q, const int* prev_t, int sz1, int sz2,
                    int is_local, int* pairs, int max_pairs, int* npairs, float* ali_score) {
  int n = 0, ql = sz1 - 1, tl = sz2 - 1, qf = 0, tf = 0, rc = 0;
#define PUSHR(a, b) do { if (n < max_pairs) { pairs[2 * n] = (a); pairs[2 * n + 1] = (b); } ++n; } while (0)
  if (!is_local) {
    *ali_score = score[0];
    PUSHR(0, 0);
    int guard = 0;
    while (qf < ql) { /* optimal_rev.h:68-73 */
      size_t o = (size_t)qf * sz2 + tf;
      int nq = prev_q[o], nt = prev_t[o];
      qf = nq; tf = nt;
      PUSHR(qf, tf);
      if (qf < 0 || tf < 0 || ++guard > ql + tl + 4) { rc = 3; break; }
    }
    if (qf != ql || tf != tl) rc = 3; /* optimal_rev.h:76 */
  } else {
    float s = score[0]; /* find_max, optimal_rev.h:114-131 */
    int mq = 0, mt = 0;
    for (int i = sz1 - 1; i > 0; --i)
      for (int j = sz2 - 1; j > 0; --j)
        if (s < score[(size_t)i * sz2 + j]) { mq = i; mt = j; s = score[(size_t)i * sz2 + j]; }
    PUSHR(0, 0);
    qf = mq; tf = mt;
    *ali_score = s;
    PUSHR(qf, tf);
    while (qf < ql) { /* optimal_rev.h:102-108 */
      size_t o = (size_t)qf * sz2 + tf;
      int nq = prev_q[o], nt = prev_t[o];
      qf = nq; tf = nt;
      if (qf < 0 || tf < 0) break;
      if (score[(size_t)qf * sz2 + tf] <= 0.f) break;
      PUSHR(qf, tf);
    }
    if (qf != ql && tf != tl) PUSHR(ql, tl); /* optimal_rev.h:110 */
  }
#undef PUSHR
  *npairs = n;
  return rc;
}

/* ---------------------------------------------------------------- near-optimal cell set */

float orc_threshold(float opt, float delta_ratio) { /* cw.h:86-88 */
  float thr = (1.f - delta_ratio) * opt;
  float alt = opt - 0.1f;
  return thr < alt ? thr : alt;
}

long orc_nearopt_mask(const float* F, const float* R, const float* sim, int sz1, int sz2,
                      float thr, uint8_t* mask) {
  long cnt = 0;
  memset(mask, 0, (size_t)sz1 * sz2);
  for (int i = 1; i < sz1 - 1; ++i)
    for (int j = 1; j < sz2 - 1; ++j) {
      size_t o = (size_t)i * sz2 + j;
      float v = F[o] + R[o];
      v -= sim[o];
      if (v > thr) { mask[o] = 1; ++cnt; }
    }
  return cnt;
}

/* ucw.h:88-191 as a cell-marking recursion. */
typedef struct {
  const orc_scoring* sc;
  const float *F, *sim;
  int sz1, sz2;
  float thr;
  long count, limit;
  uint8_t* mark;
} ucw_t;

static void ucw_branch(ucw_t* u, int q0, int t0, float curr) {
  if (u->count < 0) return;
  int sz2 = u->sz2;
  if (q0 == 1 || t0 == 1) { /* base case, ucw.h:94-101 */
    u->mark[(size_t)q0 * sz2 + t0] = 1;
    u->mark[0] = 1;
    if (++u->count > u->limit) u->count = -1;
    return;
  }
  float r = curr + u->sim[(size_t)q0 * sz2 + t0]; /* ucw.h:141 */
  float f = u->F[(size_t)(q0 - 1) * sz2 + (t0 - 1)];
  int any = 0;
  if (f + r > u->thr) { /* ucw.h:144-150 */
    any = 1;
    ucw_branch(u, q0 - 1, t0 - 1, r);
  }
  for (int i = t0 - 2; i > 0; --i) { /* ucw.h:154-165 */
    f = u->F[(size_t)(q0 - 1) * sz2 + i];
    float g = orc_deletion(u->sc, u->sz2, i, t0);
    if (f + r - g > u->thr) {
      any = 1;
      ucw_branch(u, q0 - 1, i, r - g);
    }
  }
  for (int j = q0 - 2; j > 0; --j) { /* ucw.h:169-180 */
    f = u->F[(size_t)j * sz2 + (t0 - 1)];
    float g = orc_insertion(u->sc, u->sz1, j, q0);
    if (f + r - g > u->thr) {
      any = 1;
      ucw_branch(u, j, t0 - 1, r - g);
    }
  }
  if (any) u->mark[(size_t)q0 * sz2 + t0] = 1;
  /* any == 0 is the opt_path fallback (ucw.h:182-189); it cannot happen in exact arithmetic
   * because the optimal predecessor of a cell that passed always passes; flagged to the caller. */
  else if (u->count >= 0) u->count = -2;
}

long orc_ucw_cells(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
                   const float* F, const float* sim, float thr, long max_alignments,
                   uint8_t* cell_union) {
  (void)q; (void)t;
  ucw_t u;
  u.sc = sc; u.F = F; u.sim = sim; u.sz1 = Lq + 2; u.sz2 = Lt + 2; u.thr = thr;
  u.count = 0; u.limit = max_alignments; u.mark = cell_union;
  memset(cell_union, 0, (size_t)u.sz1 * u.sz2);
  ucw_branch(&u, u.sz1 - 1, u.sz2 - 1, 0.f);
  return u.count;
}

/* ---------------------------------------------------------------- UCW enumeration (ucw.h:63-191) */

/* The same Waterman branching as ucw_branch above, but EMITTING the alignments in the reference's depth-first slot
 * order (slot k of the AlignmentSet before sortSet): path = aligned pairs front to back, score as the reference
 * accumulates it (r = curr + sim; child score r - g; leaf += D[q0][t0].score).                                  */
typedef struct {
  const orc_scoring* sc;
  const float* F;
  const float* sim;
  int sz1, sz2;
  float thr;
  long count, limit, status;
  long user_limit; /* ucw.h:72 / cw.h:76: once this many alignments are complete, every further branch() call forces the
                    * optimal path instead of branching (as.size() = completed + 1 at the entry of a branch() call) */
  int* stack; /* 2 ints per frame */
  int depth;
  float* scores;
  int* ali_len;
  int* pairs; /* limit * (sz1) * 2 ints: fixed slots of sz1 aligned pairs */
  const int *pq, *pt; /* DPCell predecessors of the forward matrix (for opt_path), or NULL */
} ucwe_t;

/* opt_path (ucw.h:194-236): no branching any more, follow the stored predecessors to the base case */
static void ucwe_opt_path(ucwe_t* u, int q0, int t0, float score) {
  int sz2 = u->sz2;
  if (u->count >= u->limit) { u->status = 1; return; }
  int* out = u->pairs + (size_t)u->count * u->sz1 * 2;
  int m = 0, a = q0, b = t0;
  while (b > 1 && a > 1) {
    score += u->sim[(size_t)a * sz2 + b];
    int pa = u->pq[(size_t)a * sz2 + b], pb = u->pt[(size_t)a * sz2 + b];
    float g = (a - pa == 1) ? orc_deletion(u->sc, u->sz2, pb, b) : orc_insertion(u->sc, u->sz1, pa, a);
    score -= g;
    a = pa; b = pb;
    ++m;
  }
  score += u->F[(size_t)a * sz2 + b];
  out[0] = 0; out[1] = 0;
  a = q0; b = t0;
  for (int k = 0; k <= m; ++k) {
    out[2 * (1 + m - k)] = a; out[2 * (1 + m - k) + 1] = b;
    if (k < m) { int pa = u->pq[(size_t)a * sz2 + b], pb = u->pt[(size_t)a * sz2 + b]; a = pa; b = pb; }
  }
  int n = m + 2;
  for (int d = u->depth - 1; d >= 0; --d, ++n) { out[2 * n] = u->stack[2 * d]; out[2 * n + 1] = u->stack[2 * d + 1]; }
  u->ali_len[u->count] = n;
  u->scores[u->count] = score;
  ++u->count;
}

static void ucwe_branch(ucwe_t* u, int q0, int t0, float curr) {
  if (u->status) return;
  int sz2 = u->sz2;
  if (q0 == 1 || t0 == 1) { /* ucw.h:94-101 */
    if (u->count >= u->limit) { u->status = 1; return; }
    int* out = u->pairs + (size_t)u->count * u->sz1 * 2;
    int n = 0;
    out[0] = 0; out[1] = 0; ++n;
    out[2] = q0; out[3] = t0; ++n;
    for (int d = u->depth - 1; d >= 0; --d, ++n) { out[2 * n] = u->stack[2 * d]; out[2 * n + 1] = u->stack[2 * d + 1]; }
    u->ali_len[u->count] = n;
    u->scores[u->count] = curr + u->F[(size_t)q0 * sz2 + t0];
    ++u->count;
    return;
  }
  if (u->count >= u->user_limit) { /* ucw.h:115-126: as.size() > user_limit */
    if (u->pq) ucwe_opt_path(u, q0, t0, curr);
    else u->status = 2;
    return;
  }
  float r = curr + u->sim[(size_t)q0 * sz2 + t0];
  float f = u->F[(size_t)(q0 - 1) * sz2 + (t0 - 1)];
  int any = 0;
  u->stack[2 * u->depth] = q0;
  u->stack[2 * u->depth + 1] = t0;
  ++u->depth;
  if (f + r > u->thr) { any = 1; ucwe_branch(u, q0 - 1, t0 - 1, r); }
  for (int i = t0 - 2; i > 0 && !u->status; --i) {
    f = u->F[(size_t)(q0 - 1) * sz2 + i];
    float g = orc_deletion(u->sc, u->sz2, i, t0);
    if (f + r - g > u->thr) { any = 1; ucwe_branch(u, q0 - 1, i, r - g); }
  }
  for (int j = q0 - 2; j > 0 && !u->status; --j) {
    f = u->F[(size_t)j * sz2 + (t0 - 1)];
    float g = orc_insertion(u->sc, u->sz1, j, q0);
    if (f + r - g > u->thr) { any = 1; ucwe_branch(u, j, t0 - 1, r - g); }
  }
  --u->depth;
  if (!any && !u->status) { /* ucw.h:182-189 */
    if (u->pq) ucwe_opt_path(u, q0, t0, curr);
    else u->status = 2;
  }
}

long orc_ucw_enumerate(int Lq, int Lt, const orc_scoring* sc, const float* F, const float* sim, float thr,
                       long max_alignments, float* scores, int* ali_len, int* pairs, int* status,
                       const int* prev_q, const int* prev_t) {
  return orc_ucw_enumerate_lim(Lq, Lt, sc, F, sim, thr, max_alignments, scores, ali_len, pairs, status, prev_q, prev_t,
                               100000 /* ucw.h:72 */);
}

long orc_ucw_enumerate_lim(int Lq, int Lt, const orc_scoring* sc, const float* F, const float* sim, float thr,
                           long max_alignments, float* scores, int* ali_len, int* pairs, int* status,
                           const int* prev_q, const int* prev_t, long user_limit) {
  ucwe_t u;
  u.user_limit = user_limit;
  u.pq = prev_q; u.pt = prev_t;
  u.sc = sc; u.F = F; u.sim = sim; u.sz1 = Lq + 2; u.sz2 = Lt + 2; u.thr = thr;
  u.count = 0; u.limit = max_alignments; u.status = 0; u.depth = 0;
  u.stack = (int*)malloc(sizeof(int) * 2 * (size_t)(Lq + 3));
  u.scores = scores; u.ali_len = ali_len; u.pairs = pairs;
  ucwe_branch(&u, u.sz1 - 1, u.sz2 - 1, 0.f);
  free(u.stack);
  *status = (int)u.status;
  return u.count;
}

/* ---------------------------------------------------------------- constrained enumeration (cw.h:60-284) */

/* ConstrainedNearOptimal: the Waterman branching of cw.h:94-210 alternating with opt_path (cw.h:213-281), which
 * follows the stored predecessors until the SuboptFlag of the template position changes state (rule #1, cw.h:247-256)
 * -- or to the base case when forced.  The partial alignment is kept as a node list from the final cell downwards
 * (u->stack / u->depth), an alignment is emitted at every base case in the reference's slot order.            */
typedef struct {
  ucwe_t u;
  const uint8_t* subopt; /* sz2 flags, NULL = all true */
} cnoe_t;

static int cno_flag(const cnoe_t* c, int t) { return c->subopt ? (c->subopt[t] != 0) : 1; }
static void cno_branch(cnoe_t* c, int q0, int t0, float score, int force_opt);

static void cno_leaf(ucwe_t* u, int q0, int t0, float score) {
  if (u->count >= u->limit) { u->status = 1; return; }
  int* out = u->pairs + (size_t)u->count * u->sz1 * 2;
  int n = 0;
  out[0] = 0; out[1] = 0; ++n;
  out[2] = q0; out[3] = t0; ++n;
  for (int d = u->depth - 1; d >= 0; --d, ++n) { out[2 * n] = u->stack[2 * d]; out[2 * n + 1] = u->stack[2 * d + 1]; }
  u->ali_len[u->count] = n;
  u->scores[u->count] = score + u->F[(size_t)q0 * u->sz2 + t0];
  ++u->count;
}

static void cno_opt_path(cnoe_t* c, int q0, int t0, float score, int force_opt) { /* cw.h:213-281 */
  ucwe_t* u = &c->u;
  if (u->status) return;
  if (q0 == 1 || t0 == 1) { cno_leaf(u, q0, t0, score); return; }
  int sz2 = u->sz2, depth0 = u->depth;
  int flag = !cno_flag(c, t0);
  int pq = -1, pt = -1;
  while (t0 > 1 && q0 > 1) {
    if (!force_opt && cno_flag(c, t0) == flag) break;
    u->stack[2 * u->depth] = q0; u->stack[2 * u->depth + 1] = t0; ++u->depth;
    score += u->sim[(size_t)q0 * sz2 + t0];
    pq = u->pq[(size_t)q0 * sz2 + t0];
    pt = u->pt[(size_t)q0 * sz2 + t0];
    float g = (q0 - pq == 1) ? orc_deletion(u->sc, u->sz2, pt, t0) : orc_insertion(u->sc, u->sz1, pq, q0);
    score -= g;
    t0 = pt; q0 = pq;
  }
  cno_branch(c, pq, pt, score, force_opt);
  u->depth = depth0;
}

static void cno_branch(cnoe_t* c, int q0, int t0, float curr, int force_opt) { /* cw.h:94-210 */
  ucwe_t* u = &c->u;
  if (u->status) return;
  if (q0 == 1 || t0 == 1) { cno_leaf(u, q0, t0, curr); return; }
  if (force_opt) { cno_opt_path(c, q0, t0, curr, 1); return; }
  if (u->count >= u->user_limit) { cno_opt_path(c, q0, t0, curr, 1); return; } /* cw.h:118-130 */
  int sz2 = u->sz2, any = 0;
  float r = curr + u->sim[(size_t)q0 * sz2 + t0];
  float f = u->F[(size_t)(q0 - 1) * sz2 + (t0 - 1)];
  u->stack[2 * u->depth] = q0; u->stack[2 * u->depth + 1] = t0; ++u->depth;
  if (f + r > u->thr) { any = 1; cno_opt_path(c, q0 - 1, t0 - 1, r, 0); }
  for (int i = t0 - 2; i > 0 && !u->status; --i) {
    f = u->F[(size_t)(q0 - 1) * sz2 + i];
    float g = orc_deletion(u->sc, u->sz2, i, t0);
    if (f + r - g > u->thr) { any = 1; cno_opt_path(c, q0 - 1, i, r - g, 0); }
  }
  for (int j = q0 - 2; j > 0 && !u->status; --j) {
    f = u->F[(size_t)j * sz2 + (t0 - 1)];
    float g = orc_insertion(u->sc, u->sz1, j, q0);
    if (f + r - g > u->thr) { any = 1; cno_opt_path(c, j, t0 - 1, r - g, 0); }
  }
  --u->depth;
  if (!any && !u->status) cno_opt_path(c, q0, t0, curr, 1); /* cw.h:195-201 */
}

long orc_cno_enumerate(int Lq, int Lt, const orc_scoring* sc, const float* F, const float* sim, float thr,
                       long max_alignments, float* scores, int* ali_len, int* pairs, int* status,
                       const int* prev_q, const int* prev_t, const uint8_t* subopt_flags) {
  return orc_cno_enumerate_lim(Lq, Lt, sc, F, sim, thr, max_alignments, scores, ali_len, pairs, status, prev_q, prev_t,
                               subopt_flags, 1000000 /* cw.h:76 */);
}

long orc_cno_enumerate_lim(int Lq, int Lt, const orc_scoring* sc, const float* F, const float* sim, float thr,
                           long max_alignments, float* scores, int* ali_len, int* pairs, int* status,
                           const int* prev_q, const int* prev_t, const uint8_t* subopt_flags, long user_limit) {
  cnoe_t c;
  ucwe_t* u = &c.u;
  u->user_limit = user_limit;
  u->pq = prev_q; u->pt = prev_t;
  u->sc = sc; u->F = F; u->sim = sim; u->sz1 = Lq + 2; u->sz2 = Lt + 2; u->thr = thr;
  u->count = 0; u->limit = max_alignments; u->status = 0; u->depth = 0;
  u->stack = (int*)malloc(sizeof(int) * 2 * (size_t)(Lq + 3));
  u->scores = scores; u->ali_len = ali_len; u->pairs = pairs;
  c.subopt = subopt_flags;
  cno_branch(&c, u->sz1 - 1, u->sz2 - 1, 0.f, 0);
  free(u->stack);
  *status = (int)u->status;
  return u->count;
}
