// hmap2/aa_seq.h -- amino-acid sequence (reference aa_seq.h:9-24, aa_seq.cpp:3-23).
#ifndef AADP_HMAP2_AA_SEQ_H
#define AADP_HMAP2_AA_SEQ_H

#include <string>

#include "sequence.h"

class AASequence : public Sequence<SequenceElem*> {
 public:
  AASequence() {}
  ~AASequence() {
    for (size_t i = size(); i-- > 0;) delete (*this)[i];
  }
  // Appends one element per character; the caller supplies the '^' / '$' sentinels as the
  // reference's FASTA reader does (fastaio.h:126,137).
  void append(const std::string& s) {
    int idx = (int)size();
    for (size_t k = 0; k < s.size(); ++k) push_back(new SequenceElem(idx++, s[k]));
    seq_string.clear();
  }
  void append(const char* s) { append(std::string(s)); }

 private:
  AASequence(const AASequence&);             // elements are owned: no copies (reference aa_seq.h:13-14)
  AASequence& operator=(const AASequence&);
};

#endif
