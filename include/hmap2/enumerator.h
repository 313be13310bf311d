// hmap2/enumerator.h -- abstract alignment enumerator (reference enumerator.h:19-25).
#ifndef AADP_HMAP2_ENUMERATOR_H
#define AADP_HMAP2_ENUMERATOR_H

template <class S1, class S2, class Etype> class DPMatrix;
template <class S1, class S2, class Etype> class AlignmentSet;

template <class S1, class S2, class Etype>
class Enumerator {
 public:
  virtual ~Enumerator() {}
  virtual int estimateSize() const = 0;
  virtual void enumerate(DPMatrix<S1, S2, Etype>& dpm, AlignmentSet<S1, S2, Etype>& as) = 0;
};

#endif
