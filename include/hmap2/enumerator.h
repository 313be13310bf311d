// hmap2/enumerator.h -- the abstract alignment enumerator of the reference (enumerator.h:19-25) plus the one
// piece of logic every traceback enumerator of this port shares: following DPCell::prev_* through a matrix.
#ifndef AADP_HMAP2_ENUMERATOR_H
#define AADP_HMAP2_ENUMERATOR_H

#include <utility>
#include <vector>

template <class S1, class S2, class Etype> class DPMatrix;
template <class S1, class S2, class Etype> class AlignmentSet;

template <class S1, class S2, class Etype>
class Enumerator {
 public:
  virtual ~Enumerator() {}
  virtual int estimateSize() const = 0;
  virtual void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) = 0;
};

namespace aadp {

typedef std::vector<std::pair<int, int> > CellPath;

// Where a predecessor walk ended and why.
struct WalkEnd {
  int q, t;          // last cell reached (may be -1,-1 = DPCell::null)
  bool hit_null;     // a cell without predecessor was read
};

// Walks from (q,t) along the stored predecessors until the query index passes `q_stop` (downwards for forward
// matrices, upwards for reverse ones) and records every cell it moves to.  `positive_only` is the stop rule of
// the local enumerators (optimal.h:96-102, optimal_rev.h:102-108): a predecessor with score <= 0 ends the walk
// and is not recorded.  Forward global walks of the reference are unguarded (optimal.h:66-71); here a null
// predecessor or a path longer than the matrix perimeter ends the walk instead of reading out of bounds.
template <class Matrix>
inline WalkEnd follow_predecessors(const Matrix& dpm, int q, int t, int q_stop, bool upwards, bool positive_only,
                                   CellPath* visited) {
  const int limit = dpm.getQuerySize() + dpm.getTemplateSize() + 4;
  WalkEnd end = {q, t, false};
  for (int steps = 0; upwards ? end.q < q_stop : end.q > q_stop; ++steps) {
    const int pq = dpm.getCell(end.q, end.t)->prev_query_idx, pt = dpm.getCell(end.q, end.t)->prev_template_idx;
    end.q = pq;
    end.t = pt;
    if (pq < 0 || pt < 0) {
      end.hit_null = true;
      if (!positive_only) visited->push_back(std::make_pair(pq, pt));  // the reference appends what it read
      break;
    }
    if (positive_only && dpm.getCell(pq, pt)->score <= 0.f) break;
    visited->push_back(std::make_pair(pq, pt));
    if (steps > limit) break;
  }
  return end;
}

}  // namespace aadp

#endif
