// hmap2/kscw.h -- KSConstrainedNearOptimal (replaces reference kscw.h:26-351): "k-sorted" constrained near-optimal
// enumeration.  At every branch point all predecessors that satisfy Waterman's condition are ranked by f + r - g and only
// the NOaliParams::k_limit best continue (the best keeps the budget, the others half of it); between branch points the
// optimal path is followed until the SuboptFlags state changes.
//
// enumerate() fills the forward matrix on the GPU and hands the pair to aadp_batch_near_optimal_pruned
// (AADP_PRUNE_KSORTED); the alignments come back in the reference's slot order with its fp32 scores and the reference's
// sortSet is applied.  Same class name, constructor, estimateSize() and enumerate() signature.  (The reference header
// only parses in a translation unit that declares the HMAP sequence types and does not compile on LP64 as it stands,
// kscw.h:98-104, 188; this one has neither restriction.)
#ifndef AADP_HMAP2_KSCW_H
#define AADP_HMAP2_KSCW_H

#include <string>
#include <vector>

#include "alignment.h"
#include "dpmatrix.h"
#include "enumerator.h"
#include "noalib.h"
#include "sflags.h"

namespace aadp {
// shared by kscw.h and crcw.h: flags -> bytes, budget loop, slot order -> AlignmentSet
template <class S1, class S2, class Etype>
void pruned_enumerate(int variant, const char* who, const NOaliParams& np, const SuboptFlags& sf, DPMatrix<S1, S2, Etype>& dpm,
                      AlignmentSet<S1, S2, Etype>& as) {
  if (dpm.getDirection() != fwd) throw std::string(who) + ": needs a forward DPMatrix";
  if ((int)sf.size() < dpm.getTemplateSize()) throw std::string("Sequence flags shorter than template!");
  std::vector<unsigned char> flags((size_t)dpm.getTemplateSize());
  for (size_t j = 0; j < flags.size(); ++j) flags[j] = sf[(unsigned int)j] ? 1 : 0;
  std::vector<AlignedPairList<S1, S2> > found;
  long long budget = 4LL * np.number_suboptimal;
  if (budget < 1024) budget = 1024;
  for (;;) {
    bool overflow = false;
    dpm.prunedAlignments(variant, np, flags, (int)budget, &found, &overflow);
    if (!overflow) break;
    if (budget > 8LL * np.user_limit + 65536) throw std::string(who) + ": alignment set does not fit the output budget";
    budget *= 8;
  }
  for (size_t k = 0; k < found.size(); ++k) {
    as.push_back(found[k]);
    as.back().uid = k == 0 ? 1 : (int)k;  // kscw.h:119, 264 / crcw.h:147, 508
  }
  as.sortSet(np.number_suboptimal);
}
}  // namespace aadp

template <class S1, class S2, class Etype>
class KSConstrainedNearOptimal : public Enumerator<S1, S2, Etype> {
 public:
  typedef AlignedPairList<S1, S2> SingleAlignment;
  typedef AlignedPair<S1, S2> SinglePair;

  KSConstrainedNearOptimal(const NOaliParams& p, const SuboptFlags& f) : params(&p), subopt(&f) {}

  int estimateSize() const { return params->number_suboptimal; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    aadp::pruned_enumerate(AADP_PRUNE_KSORTED, "KSConstrainedNearOptimal", *params, *subopt, dpm, as);
  }

 private:
  const NOaliParams* params;
  const SuboptFlags* subopt;
};

#endif
