// tests/cxx/pruned_host.cpp -- TEST SHIM: the host walk of the pruned enumerators (alignment_algos_b200/csrc/aadp_pruned.h,
// the code aadp_batch_near_optimal_pruned runs over a GPU-filled pair) behind a plain C entry, so that the CPU test suite
// can drive it with matrices from the oracle and compare it with the reference's own KSConstrainedNearOptimal /
// CRConstrainedNearOptimal (oracle/_ref) without a GPU.
#include "../../alignment_algos_b200/csrc/aadp_pruned.h"

extern "C" int pruned_host_run(int variant, int Lq, int Lt, const float* F, const int32_t* pq, const int32_t* pt, const float* sim,
                               const uint8_t* flags, float gi, float ge, int delfree, int insfree, float delta_ratio,
                               unsigned k_limit, unsigned sort_limit, float max_overlap, unsigned user_limit,
                               long max_alignments, int* n_ali, float* scores, int* ali_len, int* paths, long paths_cap) {
  aadp::PrunedParams Q;
  Q.Lq = Lq; Q.Lt = Lt; Q.F = F; Q.pq = pq; Q.pt = pt; Q.sim = sim; Q.flags = flags;
  Q.gi = gi; Q.ge = ge; Q.delfree = delfree; Q.insfree = insfree; Q.delta_ratio = delta_ratio;
  Q.k_limit = k_limit; Q.sort_limit = sort_limit; Q.max_overlap = max_overlap; Q.user_limit = user_limit;
  Q.max_alignments = max_alignments;
  aadp::PrunedWalk W(Q);
  if (variant == 2) W.run_ksorted(); else W.run_controlled();
  *n_ali = (int)W.as.size();
  long at = 0;
  for (size_t k = 0; k < W.as.size(); ++k) {
    const aadp::PrunedAlignment& a = W.as[k];
    if ((long)(at + a.back.size()) > paths_cap) return 5;
    scores[k] = a.score;
    ali_len[k] = (int)a.back.size();
    for (size_t m = 0; m < a.back.size(); ++m) {
      paths[2 * (at + m)] = a.back[a.back.size() - 1 - m].first;
      paths[2 * (at + m) + 1] = a.back[a.back.size() - 1 - m].second;
    }
    at += (long)a.back.size();
  }
  return W.overflow ? 1 : 0;
}
