// hmap2/optimal_subali.h -- the optimal alignment through a sub-rectangle (replaces reference
// optimal_subali.h:20-83); the enumerator that goes with the 9-argument DPMatrix constructor / build_subdpm
// (dpmatrix.h:169-189, 319-353).  The loop-closure code of ssss.h:621-633 uses the two together.
#ifndef AADP_HMAP2_OPTIMAL_SUBALI_H
#define AADP_HMAP2_OPTIMAL_SUBALI_H

#include <string>
#include <vector>

#include "aadp_binding.h"
#include "alib.h"
#include "alignment.h"
#include "enumerator.h"

template <class S1, class S2, class Etype>
class Optimal_Subali : public Enumerator<S1, S2, Etype> {
 public:
  // anchors of the region in the order of the reference constructor: near (q,t), far (q,t)
  Optimal_Subali(int q1, int t1, int q2, int t2) : near_q(q1), near_t(t1), far_q(q2), far_t(t2) {}
  int estimateSize() const { return 1; }

  void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) {
    const size_t slot = as.size();
    as.resize(slot + 1);
    AlignedPairList<S1, S2>& ali = as[slot];
    ali.score = dpm.getCell(far_q, far_t)->score;
    ali.append(far_q, far_t);
    aadp::CellPath cells;
    const aadp::WalkEnd end = aadp::follow_predecessors(dpm, far_q, far_t, near_q, false, false, &cells);
    for (size_t k = 0; k < cells.size(); ++k) ali.prepend(cells[k].first, cells[k].second);
    if (end.q != near_q || end.t != near_t) throw std::string("Illegal alignment start pair");  // optimal_subali.h:79
  }

 private:
  int near_q, near_t, far_q, far_t;
};

// Batched form of the loop-closure pattern of ssss.h:600-633 / 700-720: instead of one
// 9-argument DPMatrix + one Optimal_Subali per loop, hand over the whole list of loops; every rectangle is filled and
// traced on the GPU in one aadp_fill_subpair_batch call (compact storage: only the rectangles live in HBM).
// out[k] is what `AlignmentSet<..> as(tmp_sub_dpm, opt_subali); as[0]` holds for loop k, identity not assigned.
namespace aadp {
struct LoopRect {
  int q1_end, t1_end, q2_beg, t2_beg;  // argument order of dpmatrix.h:169-175 and optimal_subali.h:36-48
};

template <class S1, class S2, class Etype>
void optimal_subalignments(const S1& query, const S2& templ, const Evaluator<S1, S2, Etype>& eval,
                           const std::vector<LoopRect>& loops, std::vector<AlignedPairList<S1, S2> >* out) {
  if (!DeviceScoring<Etype>::supported)
    throw std::string("optimal_subalignments: this Evaluator has no device scoring model (only AASubstitutionEval is mapped)");
  std::string alphabet;
  std::vector<float> sub;
  float gi, ge;
  int at;
  DeviceScoring<Etype>::describe(static_cast<const Etype&>(eval), &alphabet, &sub, &gi, &ge, &at);
  const std::vector<uint8_t> q = encode(query, alphabet), t = encode(templ, alphabet);
  std::vector<uint8_t> residues(q);
  residues.insert(residues.end(), t.begin(), t.end());
  residues.push_back(0);
  const int64_t seq_off[3] = {0, (int64_t)q.size(), (int64_t)(q.size() + t.size())};
  const size_t n = loops.size();
  std::vector<int32_t> iq(n, 0), it(n, 1), rects(4 * n), n_out(n), status(n);
  for (size_t k = 0; k < n; ++k) {
    rects[4 * k] = loops[k].q1_end;
    rects[4 * k + 1] = loops[k].t1_end;
    rects[4 * k + 2] = loops[k].q2_beg;
    rects[4 * k + 3] = loops[k].t2_beg;
  }
  std::vector<int64_t> off(n + 1, 0);
  std::vector<float> score(n);
  aadp_ctx* ctx = default_context();
  check(aadp_set_scoring(ctx, sub.data(), (int)alphabet.size(), gi, ge, at, AADP_REPRO_REV_BUG));
  check(aadp_fill_subpair_batch(ctx, residues.data(), seq_off, 2, iq.data(), it.data(), rects.data(), (int64_t)n, AADP_FWD, 0,
                                off.data(), 0, 0, 0, 0));
  std::vector<int32_t> pairs(2 * (size_t)off[n] + 2);
  check(aadp_fill_subpair_batch(ctx, residues.data(), seq_off, 2, iq.data(), it.data(), rects.data(), (int64_t)n, AADP_FWD,
                                score.data(), off.data(), pairs.data(), off[n], n_out.data(), status.data()));
  out->clear();
  out->resize(n);
  for (size_t k = 0; k < n; ++k) {
    if (status[k]) throw std::string("Illegal alignment start pair");  // optimal_subali.h:80
    AlignedPairList<S1, S2>& ali = (*out)[k];
    ali.score = score[k];
    for (int32_t m = 0; m < n_out[k]; ++m) ali.append(pairs[2 * (off[k] + m)], pairs[2 * (off[k] + m) + 1]);
  }
}
}  // namespace aadp

#endif
