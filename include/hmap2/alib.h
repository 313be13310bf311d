// hmap2/alib.h -- alignment parameters (drop-in for the reference's alib.h:20-44, alib.cpp:16-27).
// Written for the B200 build; only the fields the DP hot path reads are kept. The rc-file /
// ParamStore loaders of the reference are host configuration plumbing and out of scope.
#ifndef AADP_HMAP2_ALIB_H
#define AADP_HMAP2_ALIB_H

#include <string>

enum align_t {        // alib.h:20-26 -- treatment of end gaps
  global_local = 0,   // end gaps free in the query, penalised in the template
  global = 1,         // all end gaps penalised
  local_global = 2,   // end gaps free in the template, penalised in the query
  local = 3,          // local alignment (scores clamped at 0)
  semi_local = 4      // all end gaps free
};

class AliParams {
 public:
  AliParams()
      : align_type(default_align_type()),
        gap_init_penalty(default_gap_init_penalty()),
        gap_extn_penalty(default_gap_extn_penalty()) {}

  // alib.cpp:16-18
  static align_t default_align_type() { return semi_local; }
  static float default_gap_init_penalty() { return 4.73f; }
  static float default_gap_extn_penalty() { return 0.34f; }

  align_t align_type;
  float gap_init_penalty;
  float gap_extn_penalty;
  std::string submatrix_fn;
};

#endif
