// hmap2/sflags.h -- per-template-position flags that say where suboptimal branching is allowed (replaces reference
// sflags.h:24-37, sflags.cpp:14-58).  ConstrainedNearOptimal (cw.h) reads them with operator[].
#ifndef AADP_HMAP2_SFLAGS_H
#define AADP_HMAP2_SFLAGS_H

#include <string>
#include <vector>

#include "sequence.h"

class SuboptFlags : public Sequence<SequenceElem*> {
  typedef Sequence<SequenceElem*> Base;

 public:
  // NOTE the argument order: (flag value, number of template positions) -- nalign.cpp:84; aa_ali.cpp:86 swaps them
  SuboptFlags(bool f, size_t len) : flags(len, f), fill_pos(0) {
    Base::seq_name = "Flags=suboptimal region";
    setString();
  }
  bool operator[](unsigned int i) const { return flags[i]; }
  void append(const std::string& s) {  // '0' clears, anything else sets (sflags.cpp:24-33)
    for (std::string::const_iterator it = s.begin(); it != s.end(); ++it) {
      if (fill_pos >= flags.size()) throw std::string("Sequence flags longer than template!");
      flags[fill_pos++] = (*it != '0');
    }
  }
  void append(const char* cs) { append(std::string(cs)); }
  void Set(unsigned int i, bool b) {
    if (i > flags.size()) throw std::string("Subopt index out of range");
    flags[i] = b;
    seq_string.replace(i, 1, b ? "1" : "0");
  }
  size_t size() const { return flags.size(); }
  void setString() {
    seq_string.clear();
    for (size_t i = 0; i < flags.size(); ++i) seq_string.push_back(flags[i] ? '1' : '0');
  }

 private:
  std::vector<bool> flags;
  size_t fill_pos;
};

#endif
