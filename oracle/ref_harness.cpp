// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C-ABI harness around the UNMODIFIED reference implementation. It #includes the
// reference's own headers from /root/reference (never copied into this repo) and is linked
// against the reference's own translation units (see oracle/Makefile). The result,
// oracle/_ref/libaadp_ref.so, is the ground truth that
//   * pins the C restatement in oracle/aadp_oracle.c,
//   * generates the golden fixtures under tests/golden/ (oracle/gen_golden.py),
//   * serves as the "reference" CPU baseline of bench.py (--impl reference / cpu_baseline).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// What is exercised (reference file:line):
//   DPMatrix ctor + build()            dpmatrix.h:147-165, 291-317
//   forward / reverse / local fills    dpmatrix.h:356-536, 538-689, 691-877, 879-1030
//   SimilarityMatrix                   simmatrix.h:40-73
//   AASubstitutionEval                 aasubalib.h:8-87
//   BlosumMatrix file parser           submatrix.cpp:16-54
//   Optimal::enumerate                 optimal.h:47-124
//   UnconstrainedNearOptimal           ucw.h:63-236
//   ConstrainedNearOptimal             cw.h:67-284
// Optimal_Rev is abstract in the reference (optimal_rev.h:29-30 does not override
// enumerator.h:23-24), so its loop (optimal_rev.h:47-78 / 80-112) is driven here over the
// reference's own rev DPMatrix by following prev_* exactly as that loop does.
//
// Built WITHOUT -ffast-math (IEEE), unlike the reference makefile:6.

#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <chrono>
#include <atomic>

// kscw.h:98-104 and crcw.h:120-126 define debug operator<< overloads for HMAP-only instantiations at namespace scope;
// the four classes only have to be declared for those headers to parse.
class HMAPSequence;
class SMAPSequence;
class Gn2Eval;
class Hmap2Eval;
// kscw.h:188 and crcw.h:268 call min(as.capacity()*2, params->user_limit) with a size_t and an unsigned int: that only
// deduces on ILP32, where the reference was written.  A non-template overload lets the unmodified lines compile on LP64.
#include <cstddef>
inline size_t min(size_t a, const unsigned int& b) { return a < (size_t)b ? a : (size_t)b; }

#include "aa_seq.h"
#include "aasubalib.h"
#include "alib.h"
#include "alignment.h"
#include "cw.h"
#include "dpmatrix.h"
#include "noalib.h"
#include "optimal.h"
#include "sflags.h"
#include "submatrix.h"
#include "ucw.h"
#include "kscw.h"
#include "crcw.h"

typedef AASubstitutionEval<AASequence, AASequence> AAEval;
typedef DPMatrix<AASequence, AASequence, AAEval> AADPM;

// kscw.h:243 and crcw.h:235 stream every operation to cerr; the AAEval instantiation needs its own overloads (found by
// argument-dependent lookup at the point of instantiation).  Same text as the reference's HMAP overloads.
std::ostream& operator<<(std::ostream& os, KSConstrainedNearOptimal<AASequence, AASequence, AAEval>::op_data& op) {
  os << "limit=" << op.limit << ",q0=" << op.q0 << ",t0=" << op.t0 << ",k0=" << op.k0 << ",s=" << op.score
     << ",ns=" << op.new_r << ",t=" << op.thresh;
  return os;
}
std::ostream& operator<<(std::ostream& os, CRConstrainedNearOptimal<AASequence, AASequence, AAEval>::op_data& op) {
  os << "limit=" << op.limit << ",q0=" << op.q0 << ",t0=" << op.t0 << ",k0=" << op.k0 << ",s=" << op.score
     << ",ns=" << op.new_r;
  return os;
}

// A table-driven Evaluator (evaluator.h:20-147): the three scoring functions read what the caller tabulated, so the
// reference's own fill (dpmatrix.h:356-1030) can be driven with position-dependent gap models shaped like
// hmap_eval.h:63-117 / gn2_eval.h:99-158 (which cannot compile here: Troll is not vendored).
class TableEval : public Evaluator<AASequence, AASequence, TableEval> {
 public:
  TableEval(const float* sim_, const float* del_, const float* ins_, int sz2_) : sim(sim_), del(del_), ins(ins_), sz2(sz2_) {}
  float similarity(const AASequence&, const AASequence&, int q_pos, int t_pos) const { return sim[(size_t)q_pos * sz2 + t_pos]; }
  float deletion(const AASequence&, const AASequence&, int, int, int t_pos1, int t_pos2) const {
    return del[(size_t)t_pos1 * sz2 + t_pos2];
  }
  float insertion(const AASequence&, const AASequence&, int q_pos1, int q_pos2, int, int t_pos2) const {
    return ins[(size_t)(q_pos2 - q_pos1 - 1) * sz2 + t_pos2];
  }
  void pre_calculate(const AASequence&, const AASequence&) const {}
  void post_process(SimilarityMatrix&) const {}

 private:
  const float *sim, *del, *ins;
  int sz2;
};
typedef DPMatrix<AASequence, AASequence, TableEval> TabDPM;

namespace {

thread_local std::string g_err;

struct NullBuf : std::streambuf {
  int overflow(int c) { return c; }
};
NullBuf g_nullbuf;

// dpmatrix.h:696 and cw.h:90 write to cerr on every call; silence them for the
// lifetime of the library (restored never: this is a test-only shared object).
struct CerrSilencer {
  CerrSilencer() { std::cerr.rdbuf(&g_nullbuf); }
};
CerrSilencer g_silencer;

void make_seq(AASequence& s, const char* letters) {
  // fastaio.h:126,137 builds sequences as '^' + residues + '$'; FastaRead itself duplicates
  // the last line at EOF, so sequences are appended directly.
  std::string x("^");
  x += letters;
  x += "$";
  s.append(x);
}

AliParams make_params(float gi, float ge, int align_type) {
  AliParams p;
  p.align_type = static_cast<align_t>(align_type);
  p.gap_init_penalty = gi;
  p.gap_extn_penalty = ge;
  return p;
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// Fill one matrix with the reference. direction: 1 = fwd, 2 = rev (dpmatrix.h:23-26).
// Outputs are row-major sz1 x sz2 with sz1 = strlen(q)+2, sz2 = strlen(t)+2. Any may be NULL.
int ref_fill(const char* q, const char* t, const char* matrix_file, float gi, float ge,
             int align_type, int direction, float* score, int* prev_q, int* prev_t,
             float* sim) {
  try {
    AASequence qs, ts;
    make_seq(qs, q);
    make_seq(ts, t);
    AliParams ap = make_params(gi, ge, align_type);
    BlosumMatrix bm(matrix_file);
    AAEval ev(ap, bm);
    AADPM dpm(qs, ts, ev, static_cast<direction_t>(direction), ap.align_type);
    int sz1 = dpm.getQuerySize(), sz2 = dpm.getTemplateSize();
    for (int i = 0; i < sz1; ++i)
      for (int j = 0; j < sz2; ++j) {
        const DPCell* c = dpm.getCell(i, j);
        size_t o = (size_t)i * sz2 + j;
        if (score) score[o] = c->score;
        if (prev_q) prev_q[o] = c->prev_query_idx;
        if (prev_t) prev_t[o] = c->prev_template_idx;
        if (sim) sim[o] = dpm.getSim(i, j);
      }
    return 0;
  } catch (std::string& e) {
    g_err = e;
    return 1;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

// The reference fill driven by TableEval: sim (Lq+2)*(Lt+2), del_tab (Lt+2)^2, ins_tab (Lq+1)*(Lt+2) as in
// include/aadp.h (aadp_fill_pair_tabulated).  align_type only selects local vs. global fills (dpmatrix.h:155).
int ref_fill_tab(int Lq, int Lt, const float* sim, const float* del_tab, const float* ins_tab, int is_local,
                 int direction, float* score, int* prev_q, int* prev_t) {
  try {
    AASequence qs, ts;
    make_seq(qs, std::string((size_t)Lq, 'A').c_str());
    make_seq(ts, std::string((size_t)Lt, 'A').c_str());
    TableEval ev(sim, del_tab, ins_tab, Lt + 2);
    TabDPM dpm(qs, ts, ev, static_cast<direction_t>(direction), is_local ? local : global);
    int sz1 = dpm.getQuerySize(), sz2 = dpm.getTemplateSize();
    for (int i = 0; i < sz1; ++i)
      for (int j = 0; j < sz2; ++j) {
        const DPCell* c = dpm.getCell(i, j);
        size_t o = (size_t)i * sz2 + j;
        if (score) score[o] = c->score;
        if (prev_q) prev_q[o] = c->prev_query_idx;
        if (prev_t) prev_t[o] = c->prev_template_idx;
      }
    return 0;
  } catch (std::string& e) {
    g_err = e;
    return 1;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

// build_subdpm through the reference's 9-argument constructor (dpmatrix.h:169-189; real argument order
// q1_end, t1_end, q2_beg, t2_beg). Same outputs as ref_fill.
int ref_fill_sub(const char* q, const char* t, const char* matrix_file, float gi, float ge,
                 int align_type, int direction, int q1_end, int t1_end, int q2_beg, int t2_beg,
                 float* score, int* prev_q, int* prev_t) {
  try {
    AASequence qs, ts;
    make_seq(qs, q);
    make_seq(ts, t);
    AliParams ap = make_params(gi, ge, align_type);
    BlosumMatrix bm(matrix_file);
    AAEval ev(ap, bm);
    AADPM dpm(qs, ts, ev, q1_end, t1_end, q2_beg, t2_beg, static_cast<direction_t>(direction), ap.align_type);
    int sz1 = dpm.getQuerySize(), sz2 = dpm.getTemplateSize();
    for (int i = 0; i < sz1; ++i)
      for (int j = 0; j < sz2; ++j) {
        const DPCell* c = dpm.getCell(i, j);
        size_t o = (size_t)i * sz2 + j;
        if (score) score[o] = c->score;
        if (prev_q) prev_q[o] = c->prev_query_idx;
        if (prev_t) prev_t[o] = c->prev_template_idx;
      }
    return 0;
  } catch (std::string& e) {
    g_err = e;
    return 1;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

// Optimal alignment through the reference's own Optimal enumerator (optimal.h:47-124) on a
// forward matrix. pairs receives (q,t) index pairs, 2 ints each.
int ref_optimal(const char* q, const char* t, const char* matrix_file, float gi, float ge,
                int align_type, int* pairs, int max_pairs, int* npairs, float* score,
                float* identity) {
  try {
    AASequence qs, ts;
    make_seq(qs, q);
    make_seq(ts, t);
    AliParams ap = make_params(gi, ge, align_type);
    BlosumMatrix bm(matrix_file);
    AAEval ev(ap, bm);
    AADPM dpm(qs, ts, ev, fwd, ap.align_type);
    Optimal<AASequence, AASequence, AAEval> opt(ap.align_type);
    AlignmentSet<AASequence, AASequence, AAEval> as(dpm, opt);
    int n = 0;
    for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = as[0].begin();
         it != as[0].end(); ++it) {
      if (n < max_pairs) {
        pairs[2 * n] = it->query_idx();
        pairs[2 * n + 1] = it->template_idx();
      }
      ++n;
    }
    *npairs = n;
    if (score) *score = as[0].score;
    if (identity) *identity = as[0].identity;
    return 0;
  } catch (std::string& e) {
    g_err = e;
    return 1;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

// Reverse optimal alignment: the loop of optimal_rev.h:47-78 (global) / 80-112 (local) driven
// over the reference's own reverse DPMatrix (the class itself cannot be instantiated).
int ref_optimal_rev(const char* q, const char* t, const char* matrix_file, float gi, float ge,
                    int align_type, int* pairs, int max_pairs, int* npairs, float* score) {
  try {
    AASequence qs, ts;
    make_seq(qs, q);
    make_seq(ts, t);
    AliParams ap = make_params(gi, ge, align_type);
    BlosumMatrix bm(matrix_file);
    AAEval ev(ap, bm);
    AADPM dpm(qs, ts, ev, rev, ap.align_type);
    int q_last = dpm.getQuerySize() - 1, t_last = dpm.getTemplateSize() - 1;
    int n = 0;
    int qf = 0, tf = 0;
#define PUSH(a, b)                 \
  do {                             \
    if (n < max_pairs) {           \
      pairs[2 * n] = (a);          \
      pairs[2 * n + 1] = (b);      \
    }                              \
    ++n;                           \
  } while (0)
    if (ap.align_type != local) {
      if (score) *score = dpm.getCell(0, 0)->score;
      PUSH(0, 0);
      int guard = 0;
      while (qf < q_last) {
        const DPCell* c = dpm.getCell(qf, tf);
        qf = c->prev_query_idx;
        tf = c->prev_template_idx;
        PUSH(qf, tf);
        if (qf < 0 || tf < 0 || ++guard > q_last + t_last + 4) {
          // the :868 bug can send the walk to a cell that was never filled (tb -1,-1)
          *npairs = n;
          g_err = "Illegal alignment start pair";
          return 3;
        }
      }
      *npairs = n;
      if (qf != q_last || tf != t_last) {
        g_err = "Illegal alignment start pair";
        return 3;
      }
    } else {
      float s = dpm.getCell(0, 0)->score;  // optimal_rev.h:114-131 find_max
      int mq = 0, mt = 0;
      for (int i = dpm.getQuerySize() - 1; i > 0; --i)
        for (int j = dpm.getTemplateSize() - 1; j > 0; --j)
          if (s < dpm.getCell(i, j)->score) {
            mq = i;
            mt = j;
            s = dpm.getCell(i, j)->score;
          }
      PUSH(0, 0);
      qf = mq;
      tf = mt;
      if (score) *score = s;
      PUSH(qf, tf);
      while (qf < q_last) {
        const DPCell* c = dpm.getCell(qf, tf);
        qf = c->prev_query_idx;
        tf = c->prev_template_idx;
        if (qf < 0 || tf < 0) break;
        if (dpm.getCell(qf, tf)->score <= 0.f) break;
        PUSH(qf, tf);
      }
      if (qf != q_last && tf != t_last) PUSH(q_last, t_last);
      *npairs = n;
    }
#undef PUSH
    return 0;
  } catch (std::string& e) {
    g_err = e;
    return 1;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

// Near-optimal enumeration with the reference's own enumerators.
//   which = 0: UnconstrainedNearOptimal (ucw.h), 1: ConstrainedNearOptimal (cw.h) with the
//   SuboptFlags given in `flags` (one char '0'/'1' per template position incl. sentinels,
//   NULL = all true; built as nalign.cpp:84 does, not with aa_ali.cpp:86's swapped arguments).
// cell_union (sz1*sz2 bytes, may be NULL) gets 1 for every (q,t) on any enumerated alignment
// *before* sortSet truncates (number_suboptimal is forced huge for the union, then the
// caller's value is applied to the returned score list).
int ref_nearopt(const char* q, const char* t, const char* matrix_file, float gi, float ge,
                int align_type, float delta_ratio, int number_suboptimal, int which,
                const char* flags, unsigned char* cell_union, int* n_alignments,
                float* scores, int max_scores, float* threshold) {
  try {
    AASequence qs, ts;
    make_seq(qs, q);
    make_seq(ts, t);
    AliParams ap = make_params(gi, ge, align_type);
    BlosumMatrix bm(matrix_file);
    AAEval ev(ap, bm);
    AADPM dpm(qs, ts, ev, fwd, ap.align_type);
    NOaliParams np;
    np.delta_ratio = delta_ratio;
    np.number_suboptimal = 0x3fffffff / 32;  // estimateSize()*20 must not overflow (ucw.h:75)
    Optimal<AASequence, AASequence, AAEval> opt(ap.align_type);
    AlignmentSet<AASequence, AASequence, AAEval> as(dpm, opt);
    as.clear();
    int sz1 = dpm.getQuerySize(), sz2 = dpm.getTemplateSize();
    if (which == 0) {
      UnconstrainedNearOptimal<AASequence, AASequence, AAEval> u(np);
      u.enumerate(dpm, as);
    } else {
      SuboptFlags sf(true, (size_t)sz2);
      if (flags)
        for (int j = 0; j < sz2; ++j) sf.Set(j, flags[j] != '0');
      ConstrainedNearOptimal<AASequence, AASequence, AAEval> c(np, sf);
      c.enumerate(dpm, as);
    }
    if (cell_union) {
      memset(cell_union, 0, (size_t)sz1 * sz2);
      for (size_t k = 0; k < as.size(); ++k)
        for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = as[k].begin();
             it != as[k].end(); ++it)
          cell_union[(size_t)it->query_idx() * sz2 + it->template_idx()] = 1;
    }
    *n_alignments = (int)as.size();
    as.sortSet(number_suboptimal > 0 ? number_suboptimal : (int)as.size());
    for (int k = 0; k < (int)as.size() && k < max_scores; ++k) scores[k] = as[k].score;
    if (threshold) {
      float o = dpm.getCell(sz1 - 1, sz2 - 1)->score;
      float thr = (1.f - np.delta_ratio) * o;  // cw.h:86-88, ucw.h:81-83
      thr = std::min(thr, o - 0.1f);
      *threshold = thr;
    }
    return 0;
  } catch (std::string& e) {
    g_err = e;
    return 1;
  } catch (std::bad_alloc&) {
    g_err = "bad_alloc";
    return 4;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

// Every alignment UnconstrainedNearOptimal::enumerate (ucw.h:63-86; which = 0) or ConstrainedNearOptimal::enumerate
// (cw.h:67-92; which = 1 with SuboptFlags `flags`, one char '0'/'1' per template position, NULL = all true) produces, with its score and its aligned pairs, in
// the order the reference leaves them in (sorted by sortSet; number_suboptimal is forced huge so nothing is dropped).
// pairs: concatenated (query_idx, template_idx); alignment k has ali_len[k] of them.  Returns 5 when the buffers are
// too small (n_alignments / total pairs are still reported).
int ref_ucw_alignments(const char* q, const char* t, const char* matrix_file, float gi, float ge, int align_type,
                       float delta_ratio, int which, const char* flags, int max_alignments, long max_pairs, int* n_alignments, long* total_pairs,
                       float* scores, int* ali_len, int* pairs) {
  try {
    AASequence qs, ts;
    make_seq(qs, q);
    make_seq(ts, t);
    AliParams ap = make_params(gi, ge, align_type);
    BlosumMatrix bm(matrix_file);
    AAEval ev(ap, bm);
    AADPM dpm(qs, ts, ev, fwd, ap.align_type);
    NOaliParams np;
    np.delta_ratio = delta_ratio;
    np.number_suboptimal = 0x3fffffff / 32;
    Optimal<AASequence, AASequence, AAEval> opt(ap.align_type);
    AlignmentSet<AASequence, AASequence, AAEval> as(dpm, opt);
    as.clear();
    if (which == 0) {
      UnconstrainedNearOptimal<AASequence, AASequence, AAEval> u(np);
      u.enumerate(dpm, as);
    } else {  // cw.h with SuboptFlags built as nalign.cpp:84 does (one flag per template position)
      int sz2 = dpm.getTemplateSize();
      SuboptFlags sf(true, (size_t)sz2);
      if (flags)
        for (int j = 0; j < sz2; ++j) sf.Set(j, flags[j] != '0');
      ConstrainedNearOptimal<AASequence, AASequence, AAEval> cn(np, sf);
      cn.enumerate(dpm, as);
    }
    *n_alignments = (int)as.size();
    long tot = 0;
    bool fits = (int)as.size() <= max_alignments;
    for (size_t k = 0; k < as.size(); ++k) {
      long len = (long)as[k].size();
      if (fits && tot + len <= max_pairs) {
        scores[k] = as[k].score;
        ali_len[k] = (int)len;
        long o = tot;
        for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = as[k].begin(); it != as[k].end(); ++it, ++o) {
          pairs[2 * o] = it->query_idx();
          pairs[2 * o + 1] = it->template_idx();
        }
      } else {
        fits = false;
      }
      tot += len;
    }
    *total_pairs = tot;
    return fits ? 0 : 5;
  } catch (std::string& e) {
    g_err = e;
    return 1;
  } catch (std::bad_alloc&) {
    g_err = "bad_alloc";
    return 4;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

// KSConstrainedNearOptimal::enumerate (kscw.h:113-134; which = 2) and CRConstrainedNearOptimal::enumerate
// (crcw.h:134-170; which = 3): every alignment in the order the reference leaves them in (sortSet with a huge
// number_suboptimal, so nothing is dropped).  k_limit / sort_limit / max_overlap / user_limit as in NOaliParams
// (noalib.cpp:16-22).  Both enumerators log every operation to cerr; it is silenced for the duration of the call.
int ref_pruned_alignments(const char* q, const char* t, const char* matrix_file, float gi, float ge, int align_type,
                          float delta_ratio, int which, const char* flags, unsigned k_limit, unsigned sort_limit,
                          float max_overlap, unsigned user_limit, int max_alignments, long max_pairs, int* n_alignments,
                          long* total_pairs, float* scores, int* ali_len, int* pairs) {
  std::streambuf* old = std::cerr.rdbuf();
  std::ostringstream sink;
  std::cerr.rdbuf(sink.rdbuf());
  int rc = 0;
  try {
    AASequence qs, ts;
    make_seq(qs, q);
    make_seq(ts, t);
    AliParams ap = make_params(gi, ge, align_type);
    BlosumMatrix bm(matrix_file);
    AAEval ev(ap, bm);
    AADPM dpm(qs, ts, ev, fwd, ap.align_type);
    NOaliParams np;
    np.delta_ratio = delta_ratio;
    np.number_suboptimal = 0x3fffffff / 32;
    np.k_limit = k_limit;
    np.sort_limit = sort_limit;
    np.max_overlap = max_overlap;
    np.user_limit = user_limit;
    Optimal<AASequence, AASequence, AAEval> opt(ap.align_type);
    AlignmentSet<AASequence, AASequence, AAEval> as(dpm, opt);
    as.clear();
    int sz2 = dpm.getTemplateSize();
    SuboptFlags sf(true, (size_t)sz2);
    if (flags)
      for (int j = 0; j < sz2; ++j) sf.Set(j, flags[j] != '0');
    if (which == 2) {
      KSConstrainedNearOptimal<AASequence, AASequence, AAEval> ks(np, sf);
      ks.enumerate(dpm, as);
    } else {
      CRConstrainedNearOptimal<AASequence, AASequence, AAEval> cr(np, sf);
      cr.enumerate(dpm, as);
    }
    *n_alignments = (int)as.size();
    long tot = 0;
    bool fits = (int)as.size() <= max_alignments;
    for (size_t k = 0; k < as.size(); ++k) {
      long len = (long)as[k].size();
      if (fits && tot + len <= max_pairs) {
        scores[k] = as[k].score;
        ali_len[k] = (int)len;
        long o = tot;
        for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = as[k].begin(); it != as[k].end(); ++it, ++o) {
          pairs[2 * o] = it->query_idx();
          pairs[2 * o + 1] = it->template_idx();
        }
      } else {
        fits = false;
      }
      tot += len;
    }
    *total_pairs = tot;
    rc = fits ? 0 : 5;
  } catch (std::string& e) {
    g_err = e;
    rc = 1;
  } catch (std::bad_alloc&) {
    g_err = "bad_alloc";
    rc = 4;
  } catch (...) {
    g_err = "unknown exception";
    rc = 2;
  }
  std::cerr.rdbuf(old);
  return rc;
}

// CPU baseline timing: runs the reference's DPMatrix constructor (fill only, as BASELINE.md §3
// states) over `npairs` pairs given as offsets into one residue arena, on `nthreads` host threads
// (one worker per thread over a shared atomic cursor; the fill is single-threaded and shares no
// mutable state). what: 1 = fwd only, 3 = fwd + rev (two constructions).
// Returns wall seconds; *cells gets sum(Lq*Lt) over the pairs (per direction).
double ref_time_fills(const char* arena, const long long* off, const int* len, const int* pair_q,
                      const int* pair_t, int npairs, const char* matrix_file, float gi, float ge,
                      int align_type, int what, int nthreads, double* cells, double* checksum) {
  AliParams ap = make_params(gi, ge, align_type);
  BlosumMatrix bm(matrix_file);
  AAEval ev(ap, bm);
  std::atomic<int> cursor(0);
  std::vector<double> sums(nthreads, 0.0);
  double ncell = 0;
  for (int p = 0; p < npairs; ++p) ncell += (double)len[pair_q[p]] * len[pair_t[p]];
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int w = 0; w < nthreads; ++w)
    th.emplace_back([&, w]() {
      double acc = 0;
      for (;;) {
        int p = cursor.fetch_add(1);
        if (p >= npairs) break;
        std::string a(arena + off[pair_q[p]], len[pair_q[p]]);
        std::string b(arena + off[pair_t[p]], len[pair_t[p]]);
        AASequence qs, ts;
        make_seq(qs, a.c_str());
        make_seq(ts, b.c_str());
        {
          AADPM f(qs, ts, ev, fwd, ap.align_type);
          acc += f.getCell(f.getQuerySize() - 1, f.getTemplateSize() - 1)->score;
        }
        if (what & 2) {
          AADPM r(qs, ts, ev, rev, ap.align_type);
          acc += r.getCell(0, 0)->score;
        }
      }
      sums[w] = acc;
    });
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  if (cells) *cells = ncell;
  if (checksum) {
    double c = 0;
    for (double s : sums) c += s;
    *checksum = c;
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
