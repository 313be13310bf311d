// hmap2/sequence.h -- sequence containers (reference sequence.h:22-72, sequence.cpp:15-16).
// A sequence is a vector of element POINTERS whose first element is the Head sentinel '^' and whose
// last element is the Tail sentinel '$' (fastaio.h:126,137).
#ifndef AADP_HMAP2_SEQUENCE_H
#define AADP_HMAP2_SEQUENCE_H

#include <string>
#include <vector>

class SequenceElem {
 public:
  int index;
  char olc;  // one-letter code

  SequenceElem() : index(-1), olc(' ') {}
  SequenceElem(int i, char o) : index(i), olc(o) {}

  bool isHead() const { return olc == Head; }
  bool isTail() const { return olc == Tail; }

  static const char Head = '^';
  static const char Tail = '$';
};

template <class elem_t>
class Sequence : public std::vector<elem_t> {
 public:
  Sequence() : seq_length(0) {}

  unsigned int seq_length;  // length without the sentinels
  std::string seq_name;

  char olc(int i) const { return std::vector<elem_t>::at(i)->olc; }

  // Letters of the whole sequence including sentinels; cached like the reference (sequence.h:56-61).
  const std::string* getString() const {
    if (seq_string.empty()) {
      seq_string.reserve(this->size());
      for (typename std::vector<elem_t>::const_iterator it = this->begin(); it != this->end(); ++it)
        seq_string.push_back((*it)->olc);
    }
    return &seq_string;
  }

 protected:
  mutable std::string seq_string;
};

#endif
