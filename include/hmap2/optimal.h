// hmap2/optimal.h -- optimal forward traceback (reference optimal.h:23-124).
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
    if (islocal) {
      enumerate_local(dpm, as);
      return;
    }
    const size_t k = as.size();
    as.resize(k + 1);
    int q = dpm.getQuerySize() - 1, t = dpm.getTemplateSize() - 1;
    as[k].score = dpm.getCell(q, t)->score;
    as[k].append(q, t);
    while (q > 0) {  // optimal.h:66-71: follow the stored predecessors back to the anchor
      const DPCell* c = dpm.getCell(q, t);
      q = c->prev_query_idx;
      t = c->prev_template_idx;
      as[k].prepend(q, t);
    }
    if (q != 0 || t != 0) throw std::string("Illegal alignment start pair");  // optimal.h:74
  }

  void enumerate_local(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    const size_t k = as.size();
    as.resize(k + 1);
    int q = dpm.getQuerySize() - 1, t = dpm.getTemplateSize() - 1;
    float s = 0.f;
    as[k].append(q, t);
    find_max(dpm, &q, &t, &s);
    as[k].score = s;
    as[k].prepend(q, t);
    while (q > 0) {  // optimal.h:96-102: stop at the first non-positive cell
      const DPCell* c = dpm.getCell(q, t);
      q = c->prev_query_idx;
      t = c->prev_template_idx;
      if (q < 0 || t < 0) break;  // the reference reads getCell(-1,-1) here (undefined behaviour)
      if (dpm.getCell(q, t)->score <= 0.f) break;
      as[k].prepend(q, t);
    }
    if (q != 0 && t != 0) as[k].prepend(0, 0);
  }

  // first maximum in row-major order, seeded with the last interior cell (optimal.h:106-124)
  void find_max(const DPMatrix<S1, S2, Etype>& dpm, int* q, int* t, float* s) const {
    *q = dpm.getQuerySize() - 2;
    *t = dpm.getTemplateSize() - 2;
    *s = dpm.getCell(*q, *t)->score;
    for (int i = 0; i < dpm.getQuerySize() - 1; ++i)
      for (int j = 0; j < dpm.getTemplateSize() - 1; ++j)
        if (*s < dpm.getCell(i, j)->score) {
          *q = i;
          *t = j;
          *s = dpm.getCell(i, j)->score;
        }
  }

 private:
  bool islocal;
};

#endif
