// hmap2/evaluator.h -- CRTP scoring interface (reference evaluator.h:20-147): similarity(), the two
// gap penalties, and the pre/post hooks are forwarded to the derived class without virtual calls.
#ifndef AADP_HMAP2_EVALUATOR_H
#define AADP_HMAP2_EVALUATOR_H

class SimilarityMatrix;

template <class S1, class S2, class Etype>
class Evaluator {
 public:
  // similarity of query position q_pos and template position t_pos (to be maximised)
  float similarity(const S1& q, const S2& t, int q_pos, int t_pos) const { return self().similarity(q, t, q_pos, t_pos); }
  // penalty for skipping the template elements strictly between t_pos1 and t_pos2 while aligning
  // q_pos1->t_pos1 and q_pos2->t_pos2 (evaluator.h:35-52)
  float deletion(const S1& q, const S2& t, int q_pos1, int q_pos2, int t_pos1, int t_pos2) const {
    return self().deletion(q, t, q_pos1, q_pos2, t_pos1, t_pos2);
  }
  // penalty for the extra query elements strictly between q_pos1 and q_pos2 (evaluator.h:53-72)
  float insertion(const S1& q, const S2& t, int q_pos1, int q_pos2, int t_pos1, int t_pos2) const {
    return self().insertion(q, t, q_pos1, q_pos2, t_pos1, t_pos2);
  }
  void post_process(SimilarityMatrix& s) const { self().post_process(s); }
  void pre_calculate(const S1& q, const S2& t) const { self().pre_calculate(q, t); }

 protected:
  Evaluator() {}
  const Etype& self() const { return static_cast<const Etype&>(*this); }
};

#endif
