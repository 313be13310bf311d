// hmap2/aadp_binding.h -- glue between the C++ template API and the C ABI (include/aadp.h).
// Nothing here computes: it owns a per-thread aadp context, turns aadp error codes into the
// std::string exceptions the reference throws (dpmatrix.h:361, optimal.h:74), and describes how an
// Evaluator maps onto the device scoring model.
#ifndef AADP_HMAP2_BINDING_H
#define AADP_HMAP2_BINDING_H

#include <stdint.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "../aadp.h"
#include "aasubalib.h"

namespace aadp {

inline void check(int rc) {
  if (rc) throw std::string(aadp_last_error());
}

// One context per thread (a context is bound to one device + stream and is not thread-safe).
// AADP_DEVICE selects the device (default 0).
inline aadp_ctx* default_context() {
  struct Holder {
    aadp_ctx* c;
    Holder() : c(0) {}
    ~Holder() { if (c) aadp_destroy(c); }
  };
  static thread_local Holder h;
  if (!h.c) {
    int dev = 0;
    if (const char* e = getenv("AADP_DEVICE")) dev = atoi(e);
    h.c = aadp_create(dev);
    if (!h.c) throw std::string(aadp_last_error());
  }
  return h.c;
}

// What the device needs to know about an evaluator: a residue alphabet with a dense substitution
// table and the affine gap model of aasubalib.h:27-77.  Evaluators with position-dependent gap
// penalties (hmap_eval.h, gn2_eval.h) have no mapping yet: the primary
// template refuses them loudly instead of silently computing something else.
template <class Etype>
struct DeviceScoring {
  static const bool supported = false;
};

template <class S1, class S2>
struct DeviceScoring<AASubstitutionEval<S1, S2> > {
  static const bool supported = true;
  static void describe(const AASubstitutionEval<S1, S2>& e, std::string* alphabet, std::vector<float>* sub, float* gi,
                       float* ge, int* align_type) {
    *alphabet = e.getSubstitutionMatrix()->getAlphabet();
    *sub = e.getSubstitutionMatrix()->dense();
    *gi = e.getParams()->gap_init_penalty;
    *ge = e.getParams()->gap_extn_penalty;
    *align_type = (int)e.getParams()->align_type;
  }
};

// residue letters (without the sentinels) -> codes in `alphabet`
template <class S>
inline std::vector<uint8_t> encode(const S& seq, const std::string& alphabet) {
  std::vector<uint8_t> out;
  const int n = (int)seq.size();
  if (n < 2 || !seq[0]->isHead() || !seq[n - 1]->isTail())
    throw std::string("sequence must start with '^' and end with '$'");
  out.reserve(n - 2);
  for (int i = 1; i < n - 1; ++i) {
    const size_t k = alphabet.find(seq[i]->olc);
    if (k == std::string::npos) throw std::string("Letter not in substitution matrix: ") + seq[i]->olc;
    out.push_back((uint8_t)k);
  }
  return out;
}

}  // namespace aadp

#endif
