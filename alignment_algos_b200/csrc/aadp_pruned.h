// aadp_pruned.h -- the two PRUNED near-optimal enumerators of the reference, over a GPU-filled forward matrix
// (SURVEY.md §8 row f2):
//   KSConstrainedNearOptimal  kscw.h:106-351   "k-sorted": at every branch point ALL predecessors that satisfy Waterman's
//                                              condition are collected, ranked by f + r - g, and only the k best continue;
//                                              the best keeps the budget k, the others get k/2
//   CRConstrainedNearOptimal  crcw.h:134-594   "controlled redundancy": the ranked predecessors are extended along their
//                                              optimal sub-paths to the next SuboptFlags region boundary and a candidate is
//                                              dropped when its sub-path shares more than max_overlap of an accepted one
// Both alternate branching with walks along the stored optimal predecessors (DPCell::prev_*), i.e. they consume exactly
// what the fill produces: the forward score matrix, its traceback and the similarity matrix.  The walk of one pair is a
// short, strictly sequential recursion whose width the pruning bounds (k_limit = 16, sort_limit = 100 by default,
// noalib.cpp:19-20); it runs on the host over the dense view of the resident pair (aadp_batch_fetch_pair), like the
// reference's own callers (gn2.cpp, nalign2.cpp) run it over DPMatrix::getCell.  Arithmetic is the reference's fp32 in
// the reference's order (r = curr + sim; sum = f + r - g; child score r - g), so scores are bit-identical.
//
// Ties.  The reference ranks with std::sort / std::partial_sort on the score alone (kscw.h:243-249, crcw.h:313-318):
// which of several equal-score predecessors survive a cut is whatever libstdc++'s introsort leaves in front.  The same
// two library calls are made here on the same sequence of keys, so with the same libstdc++ the result is the same; the
// parity tests compare alignment sets on tie-free cuts and score multisets otherwise.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <utility>
#include <vector>

namespace aadp {

struct PrunedParams {
  int Lq, Lt;
  const float* F;        // (Lq+2)*(Lt+2) forward DPCell::score
  const int32_t* pq;     // DPCell::prev_query_idx
  const int32_t* pt;     // DPCell::prev_template_idx
  const float* sim;      // SimilarityMatrix
  const uint8_t* flags;  // SuboptFlags per template position (Lt+2), null = all true
  float gi, ge;
  int delfree, insfree;  // aasubalib.h:39-42, 65-68
  float delta_ratio;
  unsigned k_limit, sort_limit, user_limit;
  float max_overlap;
  int64_t max_alignments;  // output budget
};

struct PrunedAlignment {
  std::vector<std::pair<int, int> > back;  // aligned pairs from the END of the alignment towards (0,0) ("prepend" = push_back)
  float score = 0.f;
};

class PrunedWalk {
 public:
  explicit PrunedWalk(const PrunedParams& p) : P(p), sz2(p.Lt + 2) {}
  std::vector<PrunedAlignment> as;
  bool overflow = false;
  float threshold = 0.f;

  void run_ksorted() {  // kscw.h:113-134
    begin();
    Op op = {P.k_limit, P.Lq + 1, P.Lt + 1, 0, 0.f, 0.f, 0u};
    ks_branch(op);
  }
  void run_controlled() {  // crcw.h:134-170
    begin();
    // regions[i]: index of the SuboptFlags run position i belongs to (crcw.h:181-186)
    regions.assign((size_t)P.Lt + 2, 0);
    int state = 0;
    for (int i = 0; i + 1 < P.Lt + 2; ++i) {
      if (flag(i + 1) != flag(i)) ++state;
      regions[(size_t)i] = state;
    }
    rows.assign((size_t)P.sort_limit * (size_t)(P.Lt + 1), -1);
    Op op = {P.k_limit, P.Lq + 1, P.Lt + 1, 0, 0.f, 0.f, 0u};
    cr_branch(op);
  }

 private:
  struct Op {  // kscw.h:38-46 / crcw.h:48-56: one candidate predecessor of a branch point
    unsigned limit;
    int q0, t0, k0;
    float score, new_r;
    unsigned index;
    bool operator<(const Op& a) const { return score > a.score; }
  };
  const PrunedParams& P;
  const int sz2;
  std::vector<int> regions;  // CR only
  std::vector<int> rows;     // CR only: sort_limit rows of Lt+1 query indices per template position (crcw.h:176-178)

  float F(int i, int j) const { return P.F[(size_t)i * sz2 + j]; }
  float S(int i, int j) const { return P.sim[(size_t)i * sz2 + j]; }
  bool flag(int t) const { return P.flags ? P.flags[t] != 0 : true; }
  float pen(int len) const { return P.gi + P.ge * (float)(len - 1); }  // aasubalib.h:37-38 (two roundings)
  // deletion(.,.,t1,t2) / insertion(q1,q2,.,.) of AASubstitutionEval (aasubalib.h:27-77)
  float del(int t1, int t2) const {
    const int len = t2 - t1 - 1;
    if (len < 1) return 0.f;
    if (P.delfree && (t1 == 0 || t2 == P.Lt + 1)) return 0.f;
    return pen(len);
  }
  float ins(int q1, int q2) const {
    const int len = q2 - q1 - 1;
    if (len < 1) return 0.f;
    if (P.insfree && (q1 == 0 || q2 == P.Lq + 1)) return 0.f;
    return pen(len);
  }
  float step_gap(int pq, int pt, int q0, int t0) const { return (q0 - pq == 1) ? del(pt, t0) : ins(pq, q0); }

  void begin() {
    as.clear();
    as.push_back(PrunedAlignment());
    overflow = false;
    const float opt = F(P.Lq + 1, P.Lt + 1);
    threshold = std::min((1.f - P.delta_ratio) * opt, opt - 0.1f);  // kscw.h:124-126
  }
  bool room() {
    if ((int64_t)as.size() >= P.max_alignments) { overflow = true; return false; }
    return true;
  }
  void finish(int k, int q0, int t0) {  // base case: kscw.h:144-152
    as[(size_t)k].back.push_back(std::make_pair(q0, t0));
    as[(size_t)k].back.push_back(std::make_pair(0, 0));
    as[(size_t)k].score += F(q0, t0);
  }
  // every predecessor of (q0,t0) that satisfies Waterman's condition, in the reference's scan order
  // (match; deletions t0-2 .. 1; insertions q0-2 .. 1), kscw.h:205-232 / crcw.h:269-297
  void candidates(int q0, int t0, int k0, float curr, unsigned limit, std::vector<Op>* out) const {
    const float r = curr + S(q0, t0);
    float sum = F(q0 - 1, t0 - 1) + r;
    if (sum > threshold) out->push_back(Op{limit, q0 - 1, t0 - 1, k0, sum, r, 0u});
    for (int i = t0 - 2; i > 0; --i) {
      const float g = del(i, t0);
      sum = F(q0 - 1, i) + r - g;
      if (sum > threshold) out->push_back(Op{limit, q0 - 1, i, k0, sum, r - g, 0u});
    }
    for (int j = q0 - 2; j > 0; --j) {
      const float g = ins(j, q0);
      sum = F(j, t0 - 1) + r - g;
      if (sum > threshold) out->push_back(Op{limit, j, t0 - 1, k0, sum, r - g, 0u});
    }
  }

  // ------------------------------------------------------------------ k-sorted (kscw.h)
  void ks_branch(const Op& op) {  // kscw.h:136-269
    if (overflow) return;
    const int q0 = op.q0, t0 = op.t0, k0 = op.k0;
    if (q0 == 1 || t0 == 1) { finish(k0, q0, t0); return; }
    if (as.size() > P.user_limit) { ks_walk(op, true); return; }  // kscw.h:170-181
    const PrunedAlignment curr = as[(size_t)k0];
    std::vector<Op> ops;
    ops.reserve((size_t)(q0 + t0));
    candidates(q0, t0, k0, curr.score, op.limit / 2, &ops);
    if (ops.empty()) {  // kscw.h:236-243: below the threshold after the last extension -> optimal path to the beginning
      Op o = {1u, q0, t0, k0, 0.f, 0.f, 0u};
      ks_walk(o, true);
      return;
    }
    if (ops.size() > op.limit) {
      std::partial_sort(ops.begin(), ops.begin() + op.limit, ops.end());
      ops.erase(ops.begin() + op.limit, ops.end());
    } else {
      std::sort(ops.begin(), ops.end());
    }
    ops[0].limit *= 2;  // only the best branch keeps the full budget (kscw.h:258)
    int k = k0;
    for (size_t n = 0; n < ops.size() && !overflow; ++n) {
      ops[n].k0 = k;
      if ((int)as.size() == k) {
        if (!room()) return;
        as.push_back(curr);
      }
      as[(size_t)k].back.push_back(std::make_pair(q0, t0));
      as[(size_t)k].score = ops[n].new_r;
      ks_walk(ops[n], false);
      k = (int)as.size();
    }
  }
  // opt_path (kscw.h:271-349): follow the optimal predecessors until the SuboptFlags state changes (or, forced / with a
  // budget of 1, to the beginning), then branch again
  void ks_walk(const Op& op, bool force) {
    if (overflow) return;
    int q0 = op.q0, t0 = op.t0;
    const int k0 = op.k0;
    if (op.limit <= 1) force = true;
    if (q0 == 1 || t0 == 1) { finish(k0, q0, t0); return; }
    const bool start = !flag(t0);
    int pq = -1, pt = -1;
    PrunedAlignment& a = as[(size_t)k0];
    while (t0 > 1 && q0 > 1) {
      if (!force && flag(t0) == start) break;
      a.back.push_back(std::make_pair(q0, t0));
      a.score += S(q0, t0);
      pq = P.pq[(size_t)q0 * sz2 + t0];
      pt = P.pt[(size_t)q0 * sz2 + t0];
      a.score -= step_gap(pq, pt, q0, t0);
      t0 = pt;
      q0 = pq;
    }
    Op next = {op.limit, pq, pt, k0, 0.f, 0.f, 0u};
    ks_branch(next);
  }

  // ------------------------------------------------------------------ controlled redundancy (crcw.h)
  void cr_force(const Op& op) {  // force_opt_path, crcw.h:556-592: optimal predecessors down to (0,0)
    int q0 = op.q0, t0 = op.t0;
    PrunedAlignment& a = as[(size_t)op.k0];
    while (t0 > 0 && q0 > 0) {
      a.back.push_back(std::make_pair(q0, t0));
      a.score += S(q0, t0);
      const int pq = P.pq[(size_t)q0 * sz2 + t0], pt = P.pt[(size_t)q0 * sz2 + t0];
      a.score -= step_gap(pq, pt, q0, t0);
      t0 = pt;
      q0 = pq;
    }
    a.back.push_back(std::make_pair(0, 0));
  }
  // regions[t-1] as the reference reads it; for t = 0 the reference reads one int in front of its heap array
  // (crcw.h:411, 436: undefined behaviour) -- with glibc that word is the upper half of the chunk size, i.e. 0
  int region_of(int t) const { return t >= 1 ? regions[(size_t)t - 1] : 0; }
  int& row(size_t i, int col) { return rows[i * (size_t)(P.Lt + 1) + (size_t)col]; }

  void cr_branch(const Op& op) {  // crcw.h:206-336
    if (overflow) return;
    const int q0 = op.q0, t0 = op.t0, k0 = op.k0;
    if (op.limit < 2) { cr_force(op); return; }
    if (as.size() > P.user_limit) { cr_force(op); return; }
    std::vector<Op> ops;
    ops.reserve((size_t)(q0 + t0));
    candidates(q0, t0, k0, as[(size_t)k0].score, op.limit, &ops);
    if (ops.empty()) { cr_force(op); return; }
    if (ops.size() > P.sort_limit) {
      std::partial_sort(ops.begin(), ops.begin() + P.sort_limit, ops.end());
      ops.erase(ops.begin() + P.sort_limit, ops.end());
    } else {
      std::sort(ops.begin(), ops.end());
    }
    cr_filter_and_extend(q0, t0, &ops);
    for (size_t n = 0; n < ops.size() && !overflow; ++n)
      if (ops[n].k0 > -1) cr_branch(ops[n]);
  }

  // crcw.h:338-554
  void cr_filter_and_extend(int q0, int t0, std::vector<Op>* v) {
    std::vector<Op>& ops = *v;
    const size_t n = ops.size();
    std::vector<char> keep(n, 0);
    std::vector<int> end_q(n), end_t(n), len(n), state(n);
    std::vector<float> rs(n);
    for (size_t i = 0; i < n; ++i)
      for (int j = 0; j < t0; ++j) row(i, j) = -1;  // reinit_mem(t0, n)
    // optimal sub-path of every candidate down to the next region boundary (crcw.h:386-420)
    for (size_t i = 0; i < n; ++i) {
      ops[i].index = (unsigned)i;
      int q = ops[i].q0, t = ops[i].t0;
      len[i] = 1;
      state[i] = region_of(t);
      rs[i] = ops[i].new_r;
      while (q > 0 && t > 0 && region_of(t) == state[i]) {
        row(i, t - 1) = q;
        ++len[i];
        const int pq = P.pq[(size_t)q * sz2 + t], pt = P.pt[(size_t)q * sz2 + t];
        rs[i] += S(q, t);
        rs[i] -= step_gap(pq, pt, q, t);
        q = pq;
        t = pt;
      }
      end_q[i] = q;
      end_t[i] = t;
      state[i] = region_of(t);
    }
    // redundancy filter (crcw.h:432-470): candidates in rank order; one is dropped when it shares more than
    // max_overlap * (length of an accepted sub-path ending in the same region) aligned positions with it
    keep[0] = 1;
    unsigned accepted = 1;
    const unsigned lim = ops.back().limit;
    for (size_t i = 1; i < n && accepted < lim; ++i) {
      keep[i] = 1;
      for (size_t j = 0; j < i; ++j) {
        if (keep[i] && keep[j] && state[i] == state[j]) {
          float overlap = 0.f;
          const float overlap_max = P.max_overlap * (float)len[j];
          if (end_q[i] == end_q[j] && end_t[i] == end_t[j]) ++overlap;
          for (int k = t0 - 1; k >= end_t[i]; --k) {
            if (row(i, k) > -1 && row(j, k) > -1 && row(i, k) == row(j, k)) {
              ++overlap;
              if (overlap > overlap_max) { keep[i] = 0; break; }
            }
          }
        }
      }
      if (keep[i]) ++accepted;
    }
    std::vector<Op> kept;
    accepted = 0;
    for (size_t i = 0; i < n && accepted < lim; ++i)
      if (keep[i]) { kept.push_back(ops[i]); ++accepted; }
    ops.swap(kept);
    for (size_t i = 1; i < ops.size(); ++i) ops[i].limit = std::max(2u, lim / 2);  // crcw.h:492-494
    // the accepted sub-paths become alignments (crcw.h:500-541)
    int k = ops[0].k0;
    const PrunedAlignment curr = as[(size_t)k];
    for (size_t i = 0; i < ops.size(); ++i) {
      const size_t src = ops[i].index;
      if (k == (int)as.size()) {
        if (!room()) { for (size_t m = i; m < ops.size(); ++m) ops[m].k0 = -1; return; }
        as.push_back(curr);
      }
      PrunedAlignment& a = as[(size_t)k];
      a.back.push_back(std::make_pair(q0, t0));
      for (int j = t0 - 1; j > end_t[src]; --j) {
        const int aq = row(src, j - 1);
        if (aq > -1) a.back.push_back(std::make_pair(aq, j));
      }
      a.score = rs[src];
      ops[i].q0 = end_q[src];
      ops[i].t0 = end_t[src];
      ops[i].k0 = k;
      if (end_q[src] <= 2 || end_t[src] <= 2) {  // crcw.h:526-529: close to the beginning -> finish along the optimal path
        cr_force(ops[i]);
        ops[i].k0 = -1;
      }
      k = (int)as.size();
    }
  }
};

}  // namespace aadp
