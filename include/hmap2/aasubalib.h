// hmap2/aasubalib.h -- substitution-matrix evaluator with affine gaps (reference aasubalib.h:8-87).
#ifndef AADP_HMAP2_AASUBALIB_H
#define AADP_HMAP2_AASUBALIB_H

#include <string>

#include "alib.h"
#include "evaluator.h"
#include "sequence.h"
#include "submatrix.h"

template <class S1, class S2>
class AASubstitutionEval : public Evaluator<S1, S2, AASubstitutionEval<S1, S2> > {
 public:
  AASubstitutionEval(AliParams& p, SubstitutionMatrix& m) : params(&p), sub_matrix(&m) {}

  // aasubalib.h:17-25: sentinels score 0, residues score the substitution matrix entry
  float similarity(const S1& q, const S2& t, int q_pos, int t_pos) const {
    if (q[q_pos]->isHead() || q[q_pos]->isTail() || t[t_pos]->isHead() || t[t_pos]->isTail()) return 0.f;
    return sub_matrix->score(q[q_pos]->olc, t[t_pos]->olc);
  }

  // aasubalib.h:27-51: gap over template positions t_pos1+1 .. t_pos2-1
  float deletion(const S1&, const S2& t, int, int, int t_pos1, int t_pos2) const {
    const int len = t_pos2 - t_pos1 - 1;
    if (len < 1) return 0.f;
    if (end_gaps_free_in_template() && (t[t_pos1]->isHead() || t[t_pos2]->isTail())) return 0.f;
    return affine(len);
  }

  // aasubalib.h:53-77: gap over query positions q_pos1+1 .. q_pos2-1
  float insertion(const S1& q, const S2&, int q_pos1, int q_pos2, int, int) const {
    const int len = q_pos2 - q_pos1 - 1;
    if (len < 1) return 0.f;
    if (end_gaps_free_in_query() && (q[q_pos1]->isHead() || q[q_pos2]->isTail())) return 0.f;
    return affine(len);
  }

  void pre_calculate(const S1&, const S2&) const {}
  void post_process(SimilarityMatrix&) const {}

  // extensions used by the GPU binding (hmap2/dpmatrix.h)
  const AliParams* getParams() const { return params; }
  const SubstitutionMatrix* getSubstitutionMatrix() const { return sub_matrix; }

 private:
  float affine(int len) const { return params->gap_init_penalty + params->gap_extn_penalty * (len - 1); }
  bool end_gaps_free_in_template() const {
    switch (params->align_type) {
      case global: case global_local: return false;
      case local: case semi_local: case local_global: return true;
    }
    throw std::string("Illegal gap style");  // aasubalib.h:49
  }
  bool end_gaps_free_in_query() const {
    switch (params->align_type) {
      case global: case local_global: return false;
      case local: case semi_local: case global_local: return true;
    }
    throw std::string("Illegal gap style");
  }

  AliParams* params;
  SubstitutionMatrix* sub_matrix;
};

#endif
