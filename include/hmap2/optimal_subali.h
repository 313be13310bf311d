// hmap2/optimal_subali.h -- optimal traceback inside a sub-rectangle (reference optimal_subali.h:20-83),
// the enumerator that goes with the 9-argument DPMatrix constructor / build_subdpm (dpmatrix.h:169-189,
// 319-353); the loop-closure code of ssss.h:621-633 uses the two together.
#ifndef AADP_HMAP2_OPTIMAL_SUBALI_H
#define AADP_HMAP2_OPTIMAL_SUBALI_H

#include <string>

#include "alib.h"
#include "alignment.h"
#include "enumerator.h"

template <class S1, class S2, class Etype>
class Optimal_Subali : public Enumerator<S1, S2, Etype> {
 public:
  // anchors of the sub-alignment region, in the order of the reference constructor (optimal_subali.h:47-54)
  Optimal_Subali(int q1, int t1, int q2, int t2) : q1_end(q1), t1_end(t1), q2_beg(q2), t2_beg(t2) {}
  int estimateSize() const { return 1; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    const size_t k = as.size();
    as.resize(k + 1);
    int q = q2_beg, t = t2_beg;
    as[k].score = dpm.getCell(q, t)->score;
    as[k].append(q, t);
    while (q > q1_end) {  // optimal_subali.h:71-76: follow the stored predecessors back to the near anchor
      const DPCell* c = dpm.getCell(q, t);
      q = c->prev_query_idx;
      t = c->prev_template_idx;
      as[k].prepend(q, t);
    }
    if (q != q1_end || t != t1_end) throw std::string("Illegal alignment start pair");  // optimal_subali.h:79
  }

 private:
  int q1_end, t1_end, q2_beg, t2_beg;
};

#endif
