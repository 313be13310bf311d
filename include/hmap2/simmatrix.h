// hmap2/simmatrix.h -- similarity matrix (reference simmatrix.h:18-73): interior = eval.similarity,
// sentinel rows/columns = 0, then eval.post_process.
#ifndef AADP_HMAP2_SIMMATRIX_H
#define AADP_HMAP2_SIMMATRIX_H

#include "evaluator.h"
#include "matrix.h"

class SimilarityMatrix : public matrix<float> {
 public:
  template <class S1, class S2, class Etype>
  SimilarityMatrix(const S1& qs, const S2& ts, const Evaluator<S1, S2, Etype>& eval)
      : matrix<float>((int)qs.size(), (int)ts.size()) {
    const int ql = rows() - 1, tl = cols() - 1;
    for (int i = 0; i <= ql; ++i)
      for (int j = 0; j <= tl; ++j)
        (*this)(i, j) = (i == 0 || j == 0 || i == ql || j == tl) ? 0.f : eval.similarity(qs, ts, i, j);
    eval.post_process(*this);
  }
};

#endif
