// hmap2/optimal_rev.h -- the optimal alignment of a REVERSE matrix (replaces reference optimal_rev.h:23-131).
// In the reference this class cannot be instantiated: its enumerate(const DPMatrix&, ...) const does not override
// Enumerator::enumerate(DPMatrix&, ...) (optimal_rev.h:29-30 vs enumerator.h:23-24).  Here the override has the
// base signature, so the class works as it was evidently meant to.  Reverse alignments grow at the back.
#ifndef AADP_HMAP2_OPTIMAL_REV_H
#define AADP_HMAP2_OPTIMAL_REV_H

#include <string>

#include "alib.h"
#include "alignment.h"
#include "enumerator.h"

template <class S1, class S2, class Etype>
class Optimal_Rev : public Enumerator<S1, S2, Etype> {
 public:
  Optimal_Rev(align_t type = global) : islocal(type == local) {}
  int estimateSize() const { return 1; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    const size_t slot = as.size();
    as.resize(slot + 1);
    AlignedPairList<S1, S2>& ali = as[slot];
    const int q_end = dpm.getQuerySize() - 1, t_end = dpm.getTemplateSize() - 1;
    aadp::CellPath cells;
    ali.append(0, 0);
    if (!islocal) {
      // optimal_rev.h:57-76: from the Head/Head cell forward to the Tail/Tail anchor.  dpmatrix.h:868 can point the
      // walk at a cell that was never filled (the reference then loops); follow_predecessors stops there.
      ali.score = dpm.getCell(0, 0)->score;
      const aadp::WalkEnd end = aadp::follow_predecessors(dpm, 0, 0, q_end, true, false, &cells);
      for (size_t k = 0; k < cells.size(); ++k) ali.append(cells[k].first, cells[k].second);
      if (end.q != q_end || end.t != t_end) throw std::string("Illegal alignment start pair");
      return;
    }
    // optimal_rev.h:84-110
    int q = 0, t = 0;
    float best = 0.f;
    find_max(dpm, &q, &t, &best);
    ali.score = best;
    ali.append(q, t);
    const aadp::WalkEnd end = aadp::follow_predecessors(dpm, q, t, q_end, true, true, &cells);
    for (size_t k = 0; k < cells.size(); ++k) ali.append(cells[k].first, cells[k].second);
    if (end.q != q_end && end.t != t_end) ali.append(q_end, t_end);
  }

  // optimal_rev.h:114-131: first maximum scanning backwards from the bottom-right corner over rows/columns > 0,
  // seeded with the Head/Head cell
  void find_max(const DPMatrix<S1, S2, Etype>& dpm, int* q, int* t, float* s) const {
    int bq = 0, bt = 0;
    float bs = dpm.getCell(0, 0)->score;
    for (int i = dpm.getQuerySize() - 1; i > 0; --i)
      for (int j = dpm.getTemplateSize() - 1; j > 0; --j) {
        const float v = dpm.getCell(i, j)->score;
        if (bs < v) { bs = v; bq = i; bt = j; }
      }
    *q = bq;
    *t = bt;
    *s = bs;
  }

 private:
  bool islocal;
};

#endif
