// hmap2/alignment.h -- result containers of the DP path (reference alignment.h:36-113, 847-959):
// AlignedPair, AlignedPairList (append / prepend / score / identity) and AlignmentSet.  The
// geometry utilities of the reference file (shifts, areas, zig-zag repair, ...) are downstream
// analysis and out of scope.
#ifndef AADP_HMAP2_ALIGNMENT_H
#define AADP_HMAP2_ALIGNMENT_H

#include <algorithm>
#include <list>
#include <string>
#include <utility>
#include <vector>

#include "dpmatrix.h"
#include "enumerator.h"

template <class S1, class S2>
class AlignedPair : public std::pair<int, int> {
 public:
  AlignedPair() : std::pair<int, int>(-1, -1) {}
  AlignedPair(int i, int j) : std::pair<int, int>(i, j) {}
  int query_idx() const { return first; }
  int template_idx() const { return second; }
};

template <class S1, class S2>
class AlignedPairList : public std::list<AlignedPair<S1, S2> > {
 public:
  AlignedPairList() : score(0.f), identity(0.f), significance(9999.f), SSE_CO(0.f), coverage(0.f), uid(-1) {}

  void append(int i, int j) { this->push_back(AlignedPair<S1, S2>(i, j)); }
  void prepend(int i, int j) { this->push_front(AlignedPair<S1, S2>(i, j)); }

  // percent identity over the shorter sequence, sentinels excluded (alignment.h:855-865)
  void calcIdentity(const std::string& query, const std::string& templ) {
    int same = -2;
    const int total = (int)std::min(query.size(), templ.size()) - 2;
    for (typename std::list<AlignedPair<S1, S2> >::const_iterator it = this->begin(); it != this->end(); ++it)
      if (query[it->query_idx()] == templ[it->template_idx()]) ++same;
    identity = float(same) / float(total) * 100.f;
  }

  bool operator<(const AlignedPairList& a) const { return score > a.score; }  // best score first

  float score;
  float identity;
  float significance;
  float SSE_CO;
  float coverage;
  int uid;
};

template <class S1, class S2, class Etype>
class AlignmentSet : public std::vector<AlignedPairList<S1, S2> > {
  typedef AlignedPairList<S1, S2> Alignment;

 public:
  AlignmentSet(DPMatrix<S1, S2, Etype>& dpm, Enumerator<S1, S2, Etype>& en) : dpmatrix(&dpm), enumerator(&en) {
    this->reserve(enumerator->estimateSize());  // alignment.h:934-940
    enumerator->enumerate(*dpmatrix, *this);
    assignIdentity();
  }

  const S1* getQuerySequence() const { return dpmatrix->getQuerySequence(); }
  const S2* getTemplateSequence() const { return dpmatrix->getTemplateSequence(); }
  const DPMatrix<S1, S2, Etype>* getDPMatrix() const { return dpmatrix; }

  // keep the `max` best alignments, best first (alignment.h:922-932)
  void sortSet(int max) {
    if (max >= (int)this->size()) std::sort(this->begin(), this->end());
    else if (max > 0) {
      std::partial_sort(this->begin(), this->begin() + max, this->end());
      this->erase(this->begin() + max, this->end());
    }
  }

  void assignIdentity() {
    for (size_t k = 0; k < this->size(); ++k)
      (*this)[k].calcIdentity(*dpmatrix->getQuerySequence()->getString(), *dpmatrix->getTemplateSequence()->getString());
  }

 private:
  DPMatrix<S1, S2, Etype>* dpmatrix;
  Enumerator<S1, S2, Etype>* enumerator;
};

#endif
