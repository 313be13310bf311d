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

  void describe(std::string* alphabet, std::vector<float>* sub, float* gi, float* ge, int* at) const {
    if (!aadp::DeviceScoring<Etype>::supported)
      throw std::string("DPMatrix: this Evaluator has no device scoring model (only AASubstitutionEval is mapped)");
    describe_impl(alphabet, sub, gi, ge, at, aadp::DeviceScoring<Etype>());
  }
  template <class DS>
  void describe_impl(std::string* alphabet, std::vector<float>* sub, float* gi, float* ge, int* at, DS) const {
    DS::describe(static_cast<const Etype&>(*evaluator), alphabet, sub, gi, ge, at);
  }

  // dpmatrix.h:291-317
  void build() {
    delete simmatrix;
    simmatrix = 0;
    evaluator->pre_calculate(*query_seq, *templ_seq);
    simmatrix = new SimilarityMatrix(*query_seq, *templ_seq, *evaluator);

    const int sz1 = getQuerySize(), sz2 = getTemplateSize();
    if (sz1 < 2 || sz2 < 2) throw std::string("Illegal bounds building DPM");  // dpmatrix.h:360-361

    std::string alphabet;
    std::vector<float> sub;
    float gi, ge;
    int at;
    describe(&alphabet, &sub, &gi, &ge, &at);
    if (islocal != (at == (int)local))
      throw std::string("DPMatrix: align_t of the constructor and of the evaluator's AliParams disagree");
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
