// hmap2/optimal_rev.h -- optimal traceback over a REVERSE matrix (reference optimal_rev.h:23-131).
// In the reference this class cannot be instantiated: its enumerate(const DPMatrix&, ...) const does
// not override Enumerator::enumerate(DPMatrix&, ...) (optimal_rev.h:29-30 vs enumerator.h:23-24).
// Here the override has the base signature, so the class works as it was evidently meant to.
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
    if (islocal) {
      enumerate_local(dpm, as);
      return;
    }
    const size_t k = as.size();
    as.resize(k + 1);
    const int q_last = dpm.getQuerySize() - 1, t_last = dpm.getTemplateSize() - 1;
    int q = 0, t = 0;
    as[k].score = dpm.getCell(0, 0)->score;
    as[k].append(0, 0);
    int guard = 0;
    while (q < q_last) {  // optimal_rev.h:68-73
      const DPCell* c = dpm.getCell(q, t);
      q = c->prev_query_idx;
      t = c->prev_template_idx;
      as[k].append(q, t);
      // dpmatrix.h:868 can point the walk at a cell that was never filled; the reference then loops
      if (q < 0 || t < 0 || ++guard > q_last + t_last + 4) break;
    }
    if (q != q_last || t != t_last) throw std::string("Illegal alignment start pair");  // optimal_rev.h:76
  }

  void enumerate_local(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    const size_t k = as.size();
    as.resize(k + 1);
    const int q_last = dpm.getQuerySize() - 1, t_last = dpm.getTemplateSize() - 1;
    int q = 0, t = 0;
    float s = 0.f;
    as[k].append(0, 0);
    find_max(dpm, &q, &t, &s);
    as[k].score = s;
    as[k].append(q, t);
    while (q < q_last) {  // optimal_rev.h:102-108
      const DPCell* c = dpm.getCell(q, t);
      q = c->prev_query_idx;
      t = c->prev_template_idx;
      if (q < 0 || t < 0) break;
      if (dpm.getCell(q, t)->score <= 0.f) break;
      as[k].append(q, t);
    }
    if (q != q_last && t != t_last) as[k].append(q_last, t_last);
  }

  // first maximum scanning from the bottom-right corner (optimal_rev.h:114-131)
  void find_max(const DPMatrix<S1, S2, Etype>& dpm, int* q, int* t, float* s) const {
    *q = 0;
    *t = 0;
    *s = dpm.getCell(0, 0)->score;
    for (int i = dpm.getQuerySize() - 1; i > 0; --i)
      for (int j = dpm.getTemplateSize() - 1; j > 0; --j)
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
