// hmap2/optimal_subali.h -- the optimal alignment through a sub-rectangle (replaces reference
// optimal_subali.h:20-83); the enumerator that goes with the 9-argument DPMatrix constructor / build_subdpm
// (dpmatrix.h:169-189, 319-353).  The loop-closure code of ssss.h:621-633 uses the two together.
#ifndef AADP_HMAP2_OPTIMAL_SUBALI_H
#define AADP_HMAP2_OPTIMAL_SUBALI_H

#include <string>

#include "alib.h"
#include "alignment.h"
#include "enumerator.h"

template <class S1, class S2, class Etype>
class Optimal_Subali : public Enumerator<S1, S2, Etype> {
 public:
  // anchors of the region in the order of the reference constructor: near (q,t), far (q,t)
  Optimal_Subali(int q1, int t1, int q2, int t2) : near_q(q1), near_t(t1), far_q(q2), far_t(t2) {}
  int estimateSize() const { return 1; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    const size_t slot = as.size();
    as.resize(slot + 1);
    AlignedPairList<S1, S2>& ali = as[slot];
    ali.score = dpm.getCell(far_q, far_t)->score;
    ali.append(far_q, far_t);
    aadp::CellPath cells;
    const aadp::WalkEnd end = aadp::follow_predecessors(dpm, far_q, far_t, near_q, false, false, &cells);
    for (size_t k = 0; k < cells.size(); ++k) ali.prepend(cells[k].first, cells[k].second);
    if (end.q != near_q || end.t != near_t) throw std::string("Illegal alignment start pair");  // optimal_subali.h:79
  }

 private:
  int near_q, near_t, far_q, far_t;
};

#endif
