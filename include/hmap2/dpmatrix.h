// hmap2/dpmatrix.h -- drop-in DPMatrix whose fill runs on the GPU.
//
// Same surface as the reference class (dpmatrix.h:23-113): direction_t, DPCell, the 5-argument
// constructor, getCell / getSim / getQuerySize / getTemplateSize / getEvaluator /
// getQuerySequence / getTemplateSequence / getDirection / setEvaluator / reevaluate and the
// score-dump operator<<.  The reference fills the matrix inside build() (dpmatrix.h:291-317) with
// build_{forw,rev}[_local]_dpm_nonlinear_gaps (dpmatrix.h:356-1030); here build() hands the pair to
// aadp_fill_pair (include/aadp.h) and copies the dense result into the DPCell matrix, so every
// enumerator that walks getCell()->prev_* keeps working unchanged.
//
// The 9-argument sub-rectangle constructor (dpmatrix.h:169-189 -> build_subdpm, :319-353) goes through
// aadp_fill_subpair.
//
// Not carried over: the 2-argument constructor (never instantiable in the reference,
// dpmatrix.h:141-142) and the linear-gap stubs (dpmatrix.h:1032-1042).
#ifndef AADP_HMAP2_DPMATRIX_H
#define AADP_HMAP2_DPMATRIX_H

#include <iostream>
#include <string>
#include <vector>

#include "aadp_binding.h"
#include "alib.h"
#include "evaluator.h"
#include "matrix.h"
#include "simmatrix.h"

enum direction_t { fwd = 1, rev = 2 };

struct DPCell {
  int prev_query_idx;
  int prev_template_idx;
  int query_idx;
  int template_idx;
  float score;
  static const int null = -1;
  DPCell() : prev_query_idx(null), prev_template_idx(null), query_idx(null), template_idx(null), score(0.f) {}
  void setTB(int pq, int pt, float s) {
    prev_query_idx = pq;
    prev_template_idx = pt;
    score = s;
  }
};

template <class S1, class S2, class Etype>
class DPMatrix {
 public:
  DPMatrix(const S1& query_seq_, const S2& templ_seq_, const Evaluator<S1, S2, Etype>& eval, direction_t dir = fwd,
           align_t type = global)
      : query_seq(&query_seq_), templ_seq(&templ_seq_), evaluator(&eval), direction(dir), islocal(type == local),
        dpmatrix(0), simmatrix(0), nearopt_delta(-1.f), nearopt_threshold(0.f) {
    allocate();
    build();
  }

  // dpmatrix.h:169-189 (argument order of the definition: q1_end, t1_end, q2_beg, t2_beg)
  DPMatrix(const S1& query_seq_, const S2& templ_seq_, const Evaluator<S1, S2, Etype>& eval, int q1_end, int t1_end,
           int q2_beg, int t2_beg, direction_t dir = fwd, align_t type = global)
      : query_seq(&query_seq_), templ_seq(&templ_seq_), evaluator(&eval), direction(dir), islocal(type == local),
        dpmatrix(0), simmatrix(0), nearopt_delta(-1.f), nearopt_threshold(0.f) {
    allocate();
    build_subdpm(q1_end, t1_end, q2_beg, t2_beg);
  }

  ~DPMatrix() {
    delete dpmatrix;
    delete simmatrix;
  }

  void setEvaluator(const Evaluator<S1, S2, Etype>& eval, direction_t dir) {
    direction = dir;
    evaluator = &eval;
    reevaluate();
  }
  void reevaluate() {  // dpmatrix.h:213-218: reset and refill with the same buffers
    for (int i = 0; i < dpmatrix->rows(); ++i)
      for (int j = 0; j < dpmatrix->cols(); ++j) (*dpmatrix)(i, j).setTB(DPCell::null, DPCell::null, 0.f);
    build();
  }

  const DPCell* getCell(int query_pos, int templ_pos) const { return &(*dpmatrix)(query_pos, templ_pos); }
  direction_t getDirection() const { return direction; }
  int getQuerySize() const { return (int)query_seq->size(); }
  int getTemplateSize() const { return (int)templ_seq->size(); }
  const Evaluator<S1, S2, Etype>* getEvaluator() const { return evaluator; }
  const S1* getQuerySequence() const { return query_seq; }
  const S2* getTemplateSequence() const { return templ_seq; }
  float getSim(int i, int j) const { return (*simmatrix)(i, j); }

  // ---- extension: the near-optimal cell set the Waterman enumerators consume (ucw.h:141-180):
  // mask(i,j) != 0  <=>  F(i,j) + R(i,j) - sim(i,j) > min((1-delta)*opt, opt-0.1f)   (cw.h:86-88).
  // Runs the fused forward+reverse GPU pass for this pair; the matrix itself is left untouched.
  const std::vector<unsigned char>& nearOptimalCells(float delta_ratio, float* threshold = 0) {
    if (nearopt_delta != delta_ratio) {
      std::string alphabet;
      std::vector<float> sub;
      float gi, ge;
      int at;
      describe(&alphabet, &sub, &gi, &ge, &at);
      const std::vector<uint8_t> q = aadp::encode(*query_seq, alphabet), t = aadp::encode(*templ_seq, alphabet);
      aadp_ctx* ctx = aadp::default_context();
      aadp::check(aadp_set_scoring(ctx, sub.data(), (int)alphabet.size(), gi, ge, at, AADP_REPRO_REV_BUG));
      nearopt.assign((size_t)getQuerySize() * getTemplateSize(), 0);
      aadp::check(aadp_fill_pair(ctx, q.data(), (int)q.size(), t.data(), (int)t.size(), 3, delta_ratio, 0, 0, 0, 0, 0, 0,
                                 nearopt.data(), &nearopt_threshold));
      nearopt_delta = delta_ratio;
    }
    if (threshold) *threshold = nearopt_threshold;
    return nearopt;
  }

  // ---- extension: the near-optimal ALIGNMENTS themselves, enumerated on the GPU (aadp_batch_near_optimal =
  // UnconstrainedNearOptimal::branch, ucw.h:88-191) in the reference's depth-first slot order; include/hmap2/ucw.h
  // is the enumerator class on top of it.  budget = maximum number of alignments; *overflow tells when there are more.
  // subopt_flags != 0: the constrained enumeration of cw.h with one flag per template position (sentinels included).
  // The pruned enumerators (kscw.h / crcw.h): forward fill with traceback and scores on the GPU,
  // then aadp_batch_near_optimal_pruned over the resident pair.  variant: AADP_PRUNE_KSORTED / AADP_PRUNE_REDUNDANCY.
  template <class Alignment, class Params>
  void prunedAlignments(int variant, const Params& np, const std::vector<unsigned char>& subopt_flags, int budget,
                        std::vector<Alignment>* out, bool* overflow) {
    std::string alphabet;
    std::vector<float> sub;
    float gi, ge;
    int at;
    describe(&alphabet, &sub, &gi, &ge, &at);
    const std::vector<uint8_t> q = aadp::encode(*query_seq, alphabet), t = aadp::encode(*templ_seq, alphabet);
    std::vector<uint8_t> residues(q);
    residues.insert(residues.end(), t.begin(), t.end());
    residues.push_back(0);
    const int64_t seq_off[3] = {0, (int64_t)q.size(), (int64_t)(q.size() + t.size())};
    const int32_t pq = 0, pt = 1;
    aadp_ctx* ctx = aadp::default_context();
    aadp::check(aadp_set_scoring(ctx, sub.data(), (int)alphabet.size(), gi, ge, at, AADP_REPRO_REV_BUG));
    float fs = 0.f;
    aadp::check(aadp_fill_batch(ctx, residues.data(), seq_off, 2, &pq, &pt, 1, AADP_W_FWD | AADP_W_SCORES | AADP_W_TB,
                                np.delta_ratio, &fs, 0, 0, 0));
    const size_t slot = q.size() + t.size() + 4;
    std::vector<int32_t> paths(2 * (size_t)budget * slot), len((size_t)budget);
    std::vector<float> scores((size_t)budget);
    int32_t n = 0, status = 0;
    aadp::check(aadp_batch_near_optimal_pruned(ctx, 0, variant, subopt_flags.data(), np.delta_ratio, np.k_limit, np.sort_limit,
                                               np.max_overlap, np.user_limit, budget, &n, &status, scores.data(), len.data(),
                                               paths.data(), (int64_t)budget * (int64_t)slot, 0));
    *overflow = status == 1;
    out->clear();
    out->resize((size_t)n);
    size_t at_pair = 0;
    for (int32_t k = 0; k < n; ++k) {
      Alignment& ali = (*out)[(size_t)k];
      ali.score = scores[(size_t)k];
      for (int32_t m = 0; m < len[(size_t)k]; ++m, ++at_pair) ali.append(paths[2 * at_pair], paths[2 * at_pair + 1]);
    }
  }

  template <class Alignment>
  // user_limit: the enumerator's alignment limit (ucw.h:72, cw.h:76); beyond it the GPU walk forces optimal paths
  // exactly as the reference does (ucw.h:115-126, cw.h:118-130).  0 keeps the reference's own value.
  void nearOptimalAlignments(float delta_ratio, int budget, std::vector<Alignment>* out, bool* overflow,
                             const std::vector<unsigned char>* subopt_flags = 0, unsigned int user_limit = 0) {
    std::string alphabet;
    std::vector<float> sub;
    float gi, ge;
    int at;
    describe(&alphabet, &sub, &gi, &ge, &at);
    const std::vector<uint8_t> q = aadp::encode(*query_seq, alphabet), t = aadp::encode(*templ_seq, alphabet);
    std::vector<uint8_t> residues(q);
    residues.insert(residues.end(), t.begin(), t.end());
    residues.push_back(0);
    const int64_t seq_off[3] = {0, (int64_t)q.size(), (int64_t)(q.size() + t.size())};
    const int32_t pq = 0, pt = 1;
    aadp_ctx* ctx = aadp::default_context();
    aadp::check(aadp_set_scoring(ctx, sub.data(), (int)alphabet.size(), gi, ge, at, AADP_REPRO_REV_BUG));
    aadp::check(aadp_set_option(ctx, subopt_flags ? "cw_user_limit" : "ucw_user_limit", (int)user_limit));
    float fs = 0.f;
    aadp::check(aadp_fill_batch(ctx, residues.data(), seq_off, 2, &pq, &pt, 1, AADP_W_FWD | AADP_W_SCORES | AADP_W_TB, delta_ratio,
                                &fs, 0, 0, 0));
    const int64_t id = 0;
    int64_t off[2] = {0, 0};
    const int64_t flag_off[2] = {0, subopt_flags ? (int64_t)subopt_flags->size() : 0};
    aadp::check(aadp_batch_near_optimal(ctx, &id, 1, delta_ratio, budget, 0, 0, 0, 0, off, 0, 0, 0));
    std::vector<int32_t> paths(2 * (size_t)off[1] + 2), len((size_t)budget);
    std::vector<float> scores((size_t)budget);
    int32_t n = 0, status = 0;
    if (subopt_flags)
      aadp::check(aadp_batch_near_optimal_constrained(ctx, &id, 1, subopt_flags->data(), flag_off, delta_ratio, budget, &n, &status,
                                                      scores.data(), len.data(), off, paths.data(), off[1], 0));
    else
      aadp::check(aadp_batch_near_optimal(ctx, &id, 1, delta_ratio, budget, &n, &status, scores.data(), len.data(), off,
                                          paths.data(), off[1], 0));
    if (status == 2) throw std::string("near-optimal enumeration: a cell without a passing predecessor (ucw.h:182-189)");
    *overflow = status == 1;
    out->clear();
    out->resize((size_t)n);
    const int64_t slot = (int64_t)q.size() + 2;
    for (int32_t k = 0; k < n; ++k) {
      Alignment& ali = (*out)[(size_t)k];
      ali.score = scores[(size_t)k];
      for (int32_t m = 0; m < len[(size_t)k]; ++m) ali.append(paths[2 * (k * slot + m)], paths[2 * (k * slot + m) + 1]);
    }
  }

 protected:
  void allocate() {
    const int sz1 = getQuerySize(), sz2 = getTemplateSize();
    dpmatrix = new matrix<DPCell>(sz1, sz2);
    for (int i = 0; i < sz1; ++i)
      for (int j = 0; j < sz2; ++j) {
        (*dpmatrix)(i, j).query_idx = i;
        (*dpmatrix)(i, j).template_idx = j;
      }
  }

  // Evaluators with a DeviceScoring mapping (aadp_binding.h) are described as residues + table + affine gaps;
  // every other Etype runs through the tabulated gap model (build_tabulated below).
  template <bool B>
  struct mapped_tag {};
  void describe(std::string* alphabet, std::vector<float>* sub, float* gi, float* ge, int* at) const {
    describe_impl(alphabet, sub, gi, ge, at, mapped_tag<aadp::DeviceScoring<Etype>::supported>());
  }
  void describe_impl(std::string* alphabet, std::vector<float>* sub, float* gi, float* ge, int* at, mapped_tag<true>) const {
    aadp::DeviceScoring<Etype>::describe(static_cast<const Etype&>(*evaluator), alphabet, sub, gi, ge, at);
  }
  void describe_impl(std::string*, std::vector<float>*, float*, float*, int*, mapped_tag<false>) const {
    throw std::string("DPMatrix: this Evaluator has no residue/affine device mapping (sub-rectangles and the near-optimal "
                      "cell set need one; whole-matrix fills of any Evaluator use the tabulated gap model)");
  }

  // dpmatrix.h:291-317
  void build() {
    delete simmatrix;
    simmatrix = 0;
    evaluator->pre_calculate(*query_seq, *templ_seq);
    simmatrix = new SimilarityMatrix(*query_seq, *templ_seq, *evaluator);

    const int sz1 = getQuerySize(), sz2 = getTemplateSize();
    if (sz1 < 2 || sz2 < 2) throw std::string("Illegal bounds building DPM");  // dpmatrix.h:360-361
    build_fill(mapped_tag<aadp::DeviceScoring<Etype>::supported>());
    nearopt_delta = -1.f;
  }

  // ANY Evaluator: the fill sees an Evaluator only through similarity(i,j),
  // deletion(q,q+1,t1,t2) and insertion(q1,q2,t2-1,t2) (dpmatrix.h:356-1030); tabulate the three on the host over every
  // argument combination the fill can pass and run the reference's own scan over the tables on the GPU
  // (aadp_fill_pair_tabulated).  Position-dependent gap penalties (hmap_eval.h:63-117, gn2_eval.h:99-158) are covered.
  // The tables assume what every reference evaluator satisfies -- deletion ignores the query positions, insertion sees
  // the query only through q2-q1 away from the Head/Tail -- and the assumption is spot-checked here, loudly.
  void build_fill(mapped_tag<false>) {
    const int sz1 = getQuerySize(), sz2 = getTemplateSize(), Lq = sz1 - 2, Lt = sz2 - 2;
    const S1& q = *query_seq;
    const S2& t = *templ_seq;
    const Evaluator<S1, S2, Etype>& ev = *evaluator;
    std::vector<float> sim((size_t)sz1 * sz2), del((size_t)sz2 * sz2, 0.f), ins((size_t)(Lq + 1) * sz2, 0.f);
    for (int i = 0; i < sz1; ++i)
      for (int j = 0; j < sz2; ++j) sim[(size_t)i * sz2 + j] = (*simmatrix)(i, j);
    for (int t1 = 0; t1 < sz2; ++t1)
      for (int t2 = t1 + 1; t2 < sz2; ++t2) {
        // query positions as the fill passes them: boundary row (0,1), final cell (Lq,Lq+1), interior (i-1,i)
        const int qa = t1 == 0 ? 0 : (t2 == sz2 - 1 ? Lq : (Lq >= 1 ? 1 : 0));
        const float v = ev.deletion(q, t, qa, qa + 1, t1, t2);
        del[(size_t)t1 * sz2 + t2] = v;
        if (t1 > 0 && t2 < sz2 - 1 && Lq >= 3 && ((t1 * 31 + t2) % 7 == 0) && ev.deletion(q, t, Lq - 1, Lq, t1, t2) != v)
          throw std::string("DPMatrix: deletion() of this Evaluator depends on the query position; no device mapping");
      }
    for (int len = 0; len <= Lq; ++len)
      for (int t2 = 1; t2 < sz2; ++t2) {
        float v = 0.f;
        if (t2 == 1) v = ev.insertion(q, t, 0, len + 1, 0, 1);                              // Head: dpmatrix.h:421, 861
        else if (t2 == sz2 - 1) v = ev.insertion(q, t, Lq - len, Lq + 1, sz2 - 2, sz2 - 1);  // Tail: dpmatrix.h:520, 759
        else if (len + 2 <= Lq) {
          v = ev.insertion(q, t, 1, len + 2, t2 - 1, t2);                                    // interior: dpmatrix.h:473, 812
          if (len + 3 <= Lq && ((len * 31 + t2) % 7 == 0) && ev.insertion(q, t, 2, len + 3, t2 - 1, t2) != v)
            throw std::string("DPMatrix: insertion() of this Evaluator depends on the query position; no device mapping");
        }
        ins[(size_t)len * sz2 + t2] = v;
      }
    aadp_ctx* ctx = aadp::default_context();
    const size_t n = (size_t)sz1 * sz2;
    std::vector<float> score(n);
    std::vector<int32_t> pq(n), pt(n);
    aadp::check(aadp_fill_pair_tabulated(ctx, sim.data(), Lq, Lt, del.data(), ins.data(), islocal ? 1 : 0, AADP_REPRO_REV_BUG,
                                         direction == fwd ? AADP_FWD : AADP_REV, score.data(), pq.data(), pt.data()));
    for (int i = 0; i < sz1; ++i)
      for (int j = 0; j < sz2; ++j) {
        const size_t o = (size_t)i * sz2 + j;
        (*dpmatrix)(i, j).setTB(pq[o], pt[o], score[o]);
      }
  }

  void build_fill(mapped_tag<true>) {
    const int sz1 = getQuerySize(), sz2 = getTemplateSize();
    std::string alphabet;
    std::vector<float> sub;
    float gi, ge;
    int at;
    describe(&alphabet, &sub, &gi, &ge, &at);
    if (islocal != (at == (int)local)) {
      // The reference takes the 0-clamp from the constructor's align_t (dpmatrix.h:155, 310) and the free end gaps
      // from the evaluator's AliParams (aasubalib.h:27-77); when the two disagree the similarity matrix goes to the
      // exact general-gap fill with the clamp stated separately.
      const size_t n = (size_t)sz1 * sz2;
      std::vector<float> sim(n), score(n);
      std::vector<int32_t> pq(n), pt(n);
      for (int i = 0; i < sz1; ++i)
        for (int j = 0; j < sz2; ++j) sim[(size_t)i * sz2 + j] = (*simmatrix)(i, j);
      aadp::check(aadp_fill_pair_general(aadp::default_context(), sim.data(), sz1 - 2, sz2 - 2, gi, ge, at,
                                         AADP_REPRO_REV_BUG | (islocal ? AADP_CLAMP_ON : AADP_CLAMP_OFF),
                                         direction == fwd ? AADP_FWD : AADP_REV, 0, score.data(), pq.data(), pt.data()));
      for (int i = 0; i < sz1; ++i)
        for (int j = 0; j < sz2; ++j) {
          const size_t o = (size_t)i * sz2 + j;
          (*dpmatrix)(i, j).setTB(pq[o], pt[o], score[o]);
        }
      nearopt_delta = -1.f;
      return;
    }
    const std::vector<uint8_t> q = aadp::encode(*query_seq, alphabet), t = aadp::encode(*templ_seq, alphabet);

    aadp_ctx* ctx = aadp::default_context();
    aadp::check(aadp_set_scoring(ctx, sub.data(), (int)alphabet.size(), gi, ge, at, AADP_REPRO_REV_BUG));
    const size_t n = (size_t)sz1 * sz2;
    std::vector<float> score(n);
    std::vector<int32_t> pq(n), pt(n);
    if (direction == fwd)
      aadp::check(aadp_fill_pair(ctx, q.data(), (int)q.size(), t.data(), (int)t.size(), AADP_FWD, -1.f, score.data(),
                                 pq.data(), pt.data(), 0, 0, 0, 0, 0));
    else
      aadp::check(aadp_fill_pair(ctx, q.data(), (int)q.size(), t.data(), (int)t.size(), AADP_REV, -1.f, 0, 0, 0,
                                 score.data(), pq.data(), pt.data(), 0, 0));
    for (int i = 0; i < sz1; ++i)
      for (int j = 0; j < sz2; ++j) {
        const size_t o = (size_t)i * sz2 + j;
        (*dpmatrix)(i, j).setTB(pq[o], pt[o], score[o]);
      }
    nearopt_delta = -1.f;
  }

  // dpmatrix.h:319-353
  void build_subdpm(int q1_end, int t1_end, int q2_beg, int t2_beg) {
    delete simmatrix;
    simmatrix = 0;
    evaluator->pre_calculate(*query_seq, *templ_seq);
    simmatrix = new SimilarityMatrix(*query_seq, *templ_seq, *evaluator);
    const int sz1 = getQuerySize(), sz2 = getTemplateSize();
    if (sz1 < 2 || sz2 < 2) throw std::string("Illegal bounds building DPM");
    std::string alphabet;
    std::vector<float> sub;
    float gi, ge;
    int at;
    describe(&alphabet, &sub, &gi, &ge, &at);
    const std::vector<uint8_t> q = aadp::encode(*query_seq, alphabet), t = aadp::encode(*templ_seq, alphabet);
    aadp_ctx* ctx = aadp::default_context();
    aadp::check(aadp_set_scoring(ctx, sub.data(), (int)alphabet.size(), gi, ge, at, AADP_REPRO_REV_BUG));
    const size_t n = (size_t)sz1 * sz2;
    std::vector<float> score(n);
    std::vector<int32_t> pq(n), pt(n);
    aadp::check(aadp_fill_subpair(ctx, q.data(), (int)q.size(), t.data(), (int)t.size(), q1_end, t1_end, q2_beg, t2_beg,
                                  direction == fwd ? AADP_FWD : AADP_REV, score.data(), pq.data(), pt.data()));
    for (int i = 0; i < sz1; ++i)
      for (int j = 0; j < sz2; ++j) {
        const size_t o = (size_t)i * sz2 + j;
        (*dpmatrix)(i, j).setTB(pq[o], pt[o], score[o]);
      }
    nearopt_delta = -1.f;
  }

  const S1* query_seq;
  const S2* templ_seq;
  const Evaluator<S1, S2, Etype>* evaluator;
  direction_t direction;
  bool islocal;
  matrix<DPCell>* dpmatrix;
  SimilarityMatrix* simmatrix;
  float nearopt_delta, nearopt_threshold;
  std::vector<unsigned char> nearopt;

 private:
  DPMatrix(const DPMatrix&);
  DPMatrix& operator=(const DPMatrix&);
};

// score dump, one row per query position (dpmatrix.h:116-129)
template <class S1, class S2, class Etype>
std::ostream& operator<<(std::ostream& o, const DPMatrix<S1, S2, Etype>& dpm) {
  for (int i = 0; i < dpm.getQuerySize(); ++i) {
    for (int j = 0; j < dpm.getTemplateSize(); ++j) o << dpm.getCell(i, j)->score << "\t";
    o << std::endl;
  }
  return o;
}

#endif
