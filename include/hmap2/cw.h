// hmap2/cw.h -- ConstrainedNearOptimal (replaces reference cw.h:25-284): near-optimal alignments with branching
// restricted to the template regions a SuboptFlags object marks; between branch points the optimal path is followed
// (rule #1 of cw.h:247-256: a branch point is where the flag changes state).
//
// enumerate() hands the pair and the flags to the GPU (aadp_batch_near_optimal_constrained) and receives the
// alignments in the reference's slot order with the reference's fp32 scores, then applies the reference's sortSet.
#ifndef AADP_HMAP2_CW_H
#define AADP_HMAP2_CW_H

#include <string>
#include <vector>

#include "alignment.h"
#include "dpmatrix.h"
#include "enumerator.h"
#include "noalib.h"
#include "sflags.h"

template <class S1, class S2, class Etype>
class ConstrainedNearOptimal : public Enumerator<S1, S2, Etype> {
 public:
  typedef AlignedPairList<S1, S2> SingleAlignment;
  typedef AlignedPair<S1, S2> SinglePair;

  ConstrainedNearOptimal(const NOaliParams& p, const SuboptFlags& f) : user_limit(1000000), params(&p), subopt(&f) {}

  unsigned int user_limit;  // cw.h:77

  int estimateSize() const { return params->number_suboptimal; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    if (dpm.getDirection() != fwd) throw std::string("ConstrainedNearOptimal: needs a forward DPMatrix");
    if ((int)subopt->size() < dpm.getTemplateSize()) throw std::string("Sequence flags shorter than template!");
    std::vector<unsigned char> flags((size_t)dpm.getTemplateSize());
    for (size_t j = 0; j < flags.size(); ++j) flags[j] = (*subopt)[(unsigned int)j] ? 1 : 0;
    std::vector<SingleAlignment> found;
    int budget = 4 * params->number_suboptimal;
    if (budget < 1024) budget = 1024;
    for (;;) {  // grow the output budget until the pair fits; beyond user_limit the GPU walk truncates like cw.h:118-130
      bool overflow = false;
      dpm.nearOptimalAlignments(params->delta_ratio, budget, &found, &overflow, &flags, user_limit);
      if (!overflow) break;
      if (budget > 1 << 24) throw std::string("ConstrainedNearOptimal: alignment set does not fit the output budget");
      budget *= 8;
    }
    for (size_t k = 0; k < found.size(); ++k) {
      as.push_back(found[k]);
      if (k == 0) as.back().uid = 0;  // cw.h:83
    }
    as.sortSet(params->number_suboptimal);
  }

 private:
  const NOaliParams* params;
  const SuboptFlags* subopt;
};

#endif
