// hmap2/noalib.h -- near-optimal alignment parameters (reference noalib.h:18-43, noalib.cpp:16-36).
#ifndef AADP_HMAP2_NOALIB_H
#define AADP_HMAP2_NOALIB_H

class NOaliParams {
 public:
  NOaliParams()
      : number_suboptimal(200), subopt_per_round(200), delta_ratio(0.01f), k_limit(16), sort_limit(100),
        user_limit(100000), max_overlap(0.30f), final_overlap(0.30f), rounds(4) {}

  int number_suboptimal;
  int subopt_per_round;
  float delta_ratio;
  unsigned int k_limit;
  unsigned int sort_limit;
  unsigned int user_limit;
  float max_overlap;
  float final_overlap;
  unsigned int rounds;
};

#endif
