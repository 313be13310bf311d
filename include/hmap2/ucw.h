// hmap2/ucw.h -- UnconstrainedNearOptimal (replaces reference ucw.h:25-191): every alignment whose score stays within
// delta_ratio of the optimum, found by Waterman-style branching tracebacks over the forward matrix.
//
// The reference recursion (branch, ucw.h:88-191) runs on the host; here enumerate() hands the pair to the GPU
// (aadp_batch_near_optimal: one warp walks the same branching depth first over the resident forward scores) and
// receives the alignments in the reference's depth-first slot order with the reference's fp32 scores, then applies the
// reference's sortSet.  Same class name, constructor, estimateSize() and enumerate() signature.
#ifndef AADP_HMAP2_UCW_H
#define AADP_HMAP2_UCW_H

#include <string>
#include <vector>

#include "alignment.h"
#include "dpmatrix.h"
#include "enumerator.h"
#include "noalib.h"

template <class S1, class S2, class Etype>
class UnconstrainedNearOptimal : public Enumerator<S1, S2, Etype> {
 public:
  typedef AlignedPairList<S1, S2> SingleAlignment;
  typedef AlignedPair<S1, S2> SinglePair;

  UnconstrainedNearOptimal(const NOaliParams& p) : user_limit(100000), params(&p) {}

  unsigned int user_limit;  // ucw.h:72: beyond it the reference stops branching

  int estimateSize() const { return params->number_suboptimal; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    if (dpm.getDirection() != fwd) throw std::string("UnconstrainedNearOptimal: needs a forward DPMatrix");
    // Beyond user_limit the reference forces the optimal path for every further branch (opt_path, ucw.h:115-126);
    // the GPU walk reproduces that truncation, so the set is the reference's in every case.  The output budget is
    // grown until the pair fits: past the limit only the pending branches of the current stack complete, so the final
    // count stays within a small multiple of user_limit.
    std::vector<SingleAlignment> found;
    long long budget = 4LL * params->number_suboptimal;
    if (budget < 1024) budget = 1024;
    for (;;) {
      bool overflow = false;
      dpm.nearOptimalAlignments(params->delta_ratio, (int)budget, &found, &overflow, 0, user_limit);
      if (!overflow) break;
      if (budget > 8LL * user_limit + 65536)
        throw std::string("UnconstrainedNearOptimal: alignment set does not fit the output budget");
      budget *= 8;
    }
    for (size_t k = 0; k < found.size(); ++k) as.push_back(found[k]);  // slot order of ucw.h:77-85
    as.sortSet(params->number_suboptimal);
  }

 private:
  const NOaliParams* params;
};

#endif
