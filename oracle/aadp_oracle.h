/* oracle/aadp_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's DP hot path (christang/alignment-algos):
 * the general-gap matrix fill of dpmatrix.h driven by AASubstitutionEval (aasubalib.h),
 * the optimal tracebacks of optimal.h / optimal_rev.h and the near-optimal cell set the
 * enumerators of ucw.h / cw.h consume.
 *
 * PARITY PIN: every function here is checked against the real reference compiled from
 * /root/reference (oracle/_ref/libaadp_ref.so, built by oracle/Makefile) in
 * tests/test_oracle_vs_reference.py, and against the committed golden vectors in
 * tests/golden/ (generated from the real reference by oracle/gen_golden.py). The reference
 * ships no tests or golden vectors of its own.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this library.
 */
#ifndef AADP_ORACLE_H
#define AADP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* alib.h:20-26 */
enum { ORC_GLOBAL_LOCAL = 0, ORC_GLOBAL = 1, ORC_LOCAL_GLOBAL = 2, ORC_LOCAL = 3, ORC_SEMI_LOCAL = 4 };
/* dpmatrix.h:23-26 */
enum { ORC_FWD = 1, ORC_REV = 2 };

typedef struct {
  int A;            /* alphabet size; residue codes are 0..A-1 */
  const float* sub; /* A x A row-major substitution scores: sub[q_code*A + t_code] */
  float gi, ge;     /* gap(len) = gi + ge*(len-1), aasubalib.h:37-38 */
  int align_type;   /* alib.h:20-26 */
} orc_scoring;

/* Sequences are residue codes WITHOUT sentinels; matrices are (Lq+2) x (Lt+2) row-major with
 * index 0 = Head '^' and last = Tail '$' exactly as the reference (sequence.cpp:15-16).      */

/* Literal O(Lq*Lt*(Lq+Lt)) restatement of build_{forw,rev}[_local]_dpm_nonlinear_gaps.
 * repro_rev_bug != 0 reproduces dpmatrix.h:868 (opt_j = t1_m1) in the global reverse fill.
 * sim_out may be NULL. Returns 0, or 1 for "Illegal bounds building DPM" (dpmatrix.h:360).  */
int orc_fill(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
             int direction, int repro_rev_bug, float* score, int* prev_q, int* prev_t,
             float* sim_out);

/* The literal fill for ANY Evaluator (evaluator.h:20-147) described by tables: sim = (Lq+2)*(Lt+2) similarity
 * matrix; del_tab[t1*sz2+t2] = deletion(.,.,t1,t2); ins_tab[(q2-q1-1)*sz2 + t2] = insertion(q1,q2,t2-1,t2).
 * Covers hmap_eval.h:63-117 and gn2_eval.h:99-158 style position-dependent gap penalties.                  */
int orc_fill_tab(const float* sim, int Lq, int Lt, const float* del_tab, const float* ins_tab, int is_local,
                 int direction, int repro_rev_bug, float* score, int* prev_q, int* prev_t);

/* build_subdpm + the 9-argument constructor (dpmatrix.h:169-189, 319-353): the literal fill between the
 * anchors (q1_end,t1_end) and (q2_beg,t2_beg) (matrix indices), everything else left at the DPCell defaults. */
int orc_fill_sub(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
                 int direction, int repro_rev_bug, int q1_end, int t1_end, int q2_beg, int t2_beg,
                 float* score, int* prev_q, int* prev_t);

/* Exact O(Lq*Lt) restatement (running maxima with origin tracking).
 * Candidates are re-evaluated as (D[origin] - w(len)) + sim like dpmatrix.h:460-462, so
 * it is bit-identical to orc_fill whenever all scores/penalties lie on one dyadic grid.    */
int orc_fill_fast(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
                  int direction, int repro_rev_bug, float* score, int* prev_q, int* prev_t);

/* Exact fill for ANY fp32 scoring at a cost per cell that does not grow with the scans: record lists (see the
 * comment at the definition).  CPU model of csrc/aadp_frec.cuh; bit-identical to orc_fill for every scoring.
 * stats (may be NULL): cells, row-walk steps, column-walk steps, cells with an ambiguous column leader.       */
int orc_fill_rec(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc, int direction,
                 int repro_rev_bug, float* score, int* prev_q, int* prev_t, long* stats);

/* optimal.h:47-124 on a forward matrix. pairs = 2 ints per aligned pair, front to back.
 * Returns 0, or 3 for "Illegal alignment start pair" (optimal.h:74).                        */
int orc_optimal_fwd(const float* score, const int* prev_q, const int* prev_t, int sz1, int sz2,
                    int is_local, int* pairs, int max_pairs, int* npairs, float* ali_score);

/* optimal_rev.h:47-131 on a reverse matrix. */
int orc_optimal_rev(const float* score, const int* prev_q, const int* prev_t, int sz1, int sz2,
                    int is_local, int* pairs, int max_pairs, int* npairs, float* ali_score);

/* cw.h:86-88 == ucw.h:81-83 == kscw.h:124-126 */
float orc_threshold(float opt_score, float delta_ratio);

/* Near-optimal cell set {interior (i,j): F + R - sim > thr}.
 * mask is sz1*sz2 bytes (0/1). Returns the number of set cells.                             */
long orc_nearopt_mask(const float* F, const float* R, const float* sim, int sz1, int sz2,
                      float thr, uint8_t* mask);

/* Waterman branching of ucw.h:88-191 restated as a cell-marking recursion: marks every cell
 * that lies on any alignment UnconstrainedNearOptimal::enumerate would emit (before sortSet),
 * and counts the alignments. Stops with -1 if more than max_alignments would be produced
 * (the reference switches to opt_path at user_limit=100000, ucw.h:72,115-126).             */
long orc_ucw_cells(const uint8_t* q, int Lq, const uint8_t* t, int Lt, const orc_scoring* sc,
                   const float* F, const float* sim, float thr, long max_alignments,
                   uint8_t* cell_union);

/* aasubalib.h:27-77 restated (positions are matrix indices incl. sentinels). */
float orc_deletion(const orc_scoring* sc, int sz2, int t_pos1, int t_pos2);
float orc_insertion(const orc_scoring* sc, int sz1, int q_pos1, int q_pos2);

/* UnconstrainedNearOptimal::enumerate (ucw.h:63-191) up to, not including, its final sortSet: the alignments in the
 * reference's depth-first slot order.  pairs holds max_alignments fixed slots of (Lq+2) aligned pairs (2 ints each),
 * front to back.  status: 0, 1 = more than max_alignments, 2 = a node without a passing predecessor and no
 * predecessors given.  prev_q/prev_t (the forward DPCell predecessors, may be NULL) enable opt_path (ucw.h:182-236),
 * which only rounding can trigger (a cell that passed has a passing predecessor in exact arithmetic).  The 100000
 * alignment user limit (ucw.h:72,115-126) is not restated.  Returns the number of alignments.                 */
long orc_ucw_enumerate(int Lq, int Lt, const orc_scoring* sc, const float* F, const float* sim, float thr,
                       long max_alignments, float* scores, int* ali_len, int* pairs, int* status,
                       const int* prev_q, const int* prev_t);

/* The same with the reference's alignment limit as a parameter (ucw.h:72 hard-codes 100000, cw.h:76 1000000): once
 * `user_limit` alignments are complete every further branch() forces the optimal path (ucw.h:115-126, cw.h:118-130);
 * prev_q/prev_t are then required.                                                                            */
long orc_ucw_enumerate_lim(int Lq, int Lt, const orc_scoring* sc, const float* F, const float* sim, float thr,
                           long max_alignments, float* scores, int* ali_len, int* pairs, int* status,
                           const int* prev_q, const int* prev_t, long user_limit);
long orc_cno_enumerate_lim(int Lq, int Lt, const orc_scoring* sc, const float* F, const float* sim, float thr,
                           long max_alignments, float* scores, int* ali_len, int* pairs, int* status,
                           const int* prev_q, const int* prev_t, const uint8_t* subopt_flags, long user_limit);

/* ConstrainedNearOptimal::enumerate (cw.h:60-284) up to its final sortSet: branching restricted by SuboptFlags
 * (one byte per template position incl. sentinels, NULL = all true; rule #1 of cw.h:247-256).  Same output layout
 * as orc_ucw_enumerate; prev_q/prev_t are required (the optimal walks between the branch points).           */
long orc_cno_enumerate(int Lq, int Lt, const orc_scoring* sc, const float* F, const float* sim, float thr,
                       long max_alignments, float* scores, int* ali_len, int* pairs, int* status,
                       const int* prev_q, const int* prev_t, const uint8_t* subopt_flags);

#ifdef __cplusplus
}
#endif
#endif
