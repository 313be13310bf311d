// hmap2/optimal.h -- the optimal alignment of a FORWARD matrix (replaces reference optimal.h:23-124).
// The walk itself is aadp::follow_predecessors (enumerator.h); this class only decides where it starts, where it
// must end and how the visited cells become an AlignedPairList (forward alignments grow at the front).
#ifndef AADP_HMAP2_OPTIMAL_H
#define AADP_HMAP2_OPTIMAL_H

#include <string>

#include "alib.h"
#include "alignment.h"
#include "enumerator.h"

template <class S1, class S2, class Etype>
class Optimal : public Enumerator<S1, S2, Etype> {
 public:
  Optimal(align_t type = global) : islocal(type == local) {}
  int estimateSize() const { return 1; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    const size_t slot = as.size();
    as.resize(slot + 1);
    AlignedPairList<S1, S2>& ali = as[slot];
    const int q_end = dpm.getQuerySize() - 1, t_end = dpm.getTemplateSize() - 1;
    aadp::CellPath cells;
    if (!islocal) {
      // optimal.h:57-74: from the Tail/Tail cell back to the Head/Head anchor, which must be reached exactly
      ali.score = dpm.getCell(q_end, t_end)->score;
      ali.append(q_end, t_end);
      const aadp::WalkEnd end = aadp::follow_predecessors(dpm, q_end, t_end, 0, false, false, &cells);
      for (size_t k = 0; k < cells.size(); ++k) ali.prepend(cells[k].first, cells[k].second);
      if (end.q != 0 || end.t != 0) throw std::string("Illegal alignment start pair");
      return;
    }
    // optimal.h:79-104: the Tail/Tail pair, then the best cell, then predecessors while their score is positive;
    // the anchor is added unless the walk ended in row 0 or column 0
    int q = q_end, t = t_end;
    float best = 0.f;
    ali.append(q_end, t_end);
    find_max(dpm, &q, &t, &best);
    ali.score = best;
    ali.prepend(q, t);
    const aadp::WalkEnd end = aadp::follow_predecessors(dpm, q, t, 0, false, true, &cells);
    for (size_t k = 0; k < cells.size(); ++k) ali.prepend(cells[k].first, cells[k].second);
    if (end.q != 0 && end.t != 0) ali.prepend(0, 0);
  }

  // optimal.h:106-124: first maximum in row-major order over [0,last) x [0,last), seeded with the last interior cell
  void find_max(const DPMatrix<S1, S2, Etype>& dpm, int* q, int* t, float* s) const {
    const int rows = dpm.getQuerySize() - 1, cols = dpm.getTemplateSize() - 1;
    int bq = rows - 1, bt = cols - 1;
    float bs = dpm.getCell(bq, bt)->score;
    for (int cell = 0; cell < rows * cols; ++cell) {
      const float v = dpm.getCell(cell / cols, cell % cols)->score;
      if (bs < v) { bs = v; bq = cell / cols; bt = cell % cols; }
    }
    *q = bq;
    *t = bt;
    *s = bs;
  }

 private:
  bool islocal;
};

#endif
