// hmap2/crcw.h -- CRConstrainedNearOptimal (replaces reference crcw.h:27-594): near-optimal enumeration with "controlled
// redundancy".  The predecessors of a branch point that satisfy Waterman's condition are ranked (at most
// NOaliParams::sort_limit), each is extended along its optimal sub-path to the next SuboptFlags region boundary, and a
// candidate whose sub-path shares more than max_overlap of an already accepted one is dropped; k_limit bounds the
// number of accepted branches and halves down the tree.
//
// enumerate() fills the forward matrix on the GPU and hands the pair to aadp_batch_near_optimal_pruned
// (AADP_PRUNE_REDUNDANCY); see kscw.h.  Same class name, constructor, estimateSize() and enumerate() signature.
#ifndef AADP_HMAP2_CRCW_H
#define AADP_HMAP2_CRCW_H

#include "kscw.h"

template <class S1, class S2, class Etype>
class CRConstrainedNearOptimal : public Enumerator<S1, S2, Etype> {
 public:
  typedef AlignedPairList<S1, S2> SingleAlignment;
  typedef AlignedPair<S1, S2> SinglePair;

  CRConstrainedNearOptimal(const NOaliParams& p, const SuboptFlags& f) : threshold(0.f), params(&p), subopt(&f) {}

  float threshold;  // crcw.h:41 (public in the reference; set by enumerate)

  int estimateSize() const { return params->number_suboptimal; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    const float opt = dpm.getCell(dpm.getQuerySize() - 1, dpm.getTemplateSize() - 1)->score;
    threshold = (1.f - params->delta_ratio) * opt;  // crcw.h:150-151
    if (opt - 0.1f < threshold) threshold = opt - 0.1f;
    aadp::pruned_enumerate(AADP_PRUNE_REDUNDANCY, "CRConstrainedNearOptimal", *params, *subopt, dpm, as);
  }

 private:
  const NOaliParams* params;
  const SuboptFlags* subopt;
};

#endif
