// tests/cxx/pruned_demo.cpp -- one source, two builds:
//   pruned_demo : compiled against include/hmap2/ (this repo; forward fill on the GPU, pruned walk via libaadp.so)
//   pruned_ref  : compiled against /root/reference (the unmodified kscw.h / crcw.h, CPU)
// Runs KSConstrainedNearOptimal and CRConstrainedNearOptimal as the reference's HMAP drivers do (gn2.cpp:214-222,
// nalign2.cpp) -- here for AASubstitutionEval -- and prints every alignment of the sorted AlignmentSet.
//
//   usage: demo <matrix file> <align_type 0..4> <gi> <ge> <delta> <k_limit> <sort_limit> <max_overlap> <flags> <query> <template>
//   flags: one '0'/'1' per template position including the two sentinels, or "-" for all true
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#ifdef REF_BUILD
// the reference headers need these to parse / compile for AASubstitutionEval on LP64 (see oracle/ref_harness.cpp)
class HMAPSequence;
class SMAPSequence;
class Gn2Eval;
class Hmap2Eval;
#include <cstddef>
inline size_t min(size_t a, const unsigned int& b) { return a < (size_t)b ? a : (size_t)b; }
#endif

#include "aa_seq.h"
#include "aasubalib.h"
#include "alib.h"
#include "alignment.h"
#include "dpmatrix.h"
#include "noalib.h"
#include "optimal.h"
#include "sflags.h"
#include "submatrix.h"
#include "kscw.h"
#include "crcw.h"

typedef AASubstitutionEval<AASequence, AASequence> AAEval;
typedef DPMatrix<AASequence, AASequence, AAEval> Matrix;

#ifdef REF_BUILD
std::ostream& operator<<(std::ostream& os, KSConstrainedNearOptimal<AASequence, AASequence, AAEval>::op_data&) { return os; }
std::ostream& operator<<(std::ostream& os, CRConstrainedNearOptimal<AASequence, AASequence, AAEval>::op_data&) { return os; }
#endif

template <class Set>
static void print_set(const char* tag, Set& as) {
  std::printf("%s n %d\n", tag, (int)as.size());
  for (size_t k = 0; k < as.size(); ++k) {
    std::printf("%s %.6g pairs", tag, as[k].score);
    for (typename Set::value_type::const_iterator it = as[k].begin(); it != as[k].end(); ++it)
      std::printf(" %d:%d", it->query_idx(), it->template_idx());
    std::printf("\n");
  }
}

int main(int argc, char** argv) {
  if (argc != 12) {
    std::fprintf(stderr, "usage: %s matrix align_type gi ge delta k_limit sort_limit max_overlap flags query template\n", argv[0]);
    return 2;
  }
  std::ostringstream sink;  // the reference enumerators log every operation to cerr
  std::streambuf* old = std::cerr.rdbuf(sink.rdbuf());
  int rc = 0;
  try {
    AliParams params;
    params.submatrix_fn = argv[1];
    params.align_type = static_cast<align_t>(std::atoi(argv[2]));
    params.gap_init_penalty = (float)std::atof(argv[3]);
    params.gap_extn_penalty = (float)std::atof(argv[4]);
    NOaliParams np;
    np.delta_ratio = (float)std::atof(argv[5]);
    np.k_limit = (unsigned)std::atoi(argv[6]);
    np.sort_limit = (unsigned)std::atoi(argv[7]);
    np.max_overlap = (float)std::atof(argv[8]);
    np.number_suboptimal = 100000;
    AASequence query, templ;
    query.append(std::string("^") + argv[10] + "$");
    templ.append(std::string("^") + argv[11] + "$");
    BlosumMatrix blosum(params.submatrix_fn.c_str());
    AAEval eval(params, blosum);
    Matrix forward(query, templ, eval, fwd, params.align_type);
    SuboptFlags sf(true, (size_t)forward.getTemplateSize());
    const std::string fl = argv[9];
    if (fl != "-")
      for (size_t j = 0; j < fl.size() && (int)j < forward.getTemplateSize(); ++j) sf.Set((unsigned)j, fl[j] != '0');
    Optimal<AASequence, AASequence, AAEval> opt(params.align_type);
    {
      AlignmentSet<AASequence, AASequence, AAEval> as(forward, opt);
      as.clear();
      KSConstrainedNearOptimal<AASequence, AASequence, AAEval> ks(np, sf);
      ks.enumerate(forward, as);
      print_set("KS", as);
    }
    {
      AlignmentSet<AASequence, AASequence, AAEval> as(forward, opt);
      as.clear();
      CRConstrainedNearOptimal<AASequence, AASequence, AAEval> cr(np, sf);
      cr.enumerate(forward, as);
      print_set("CR", as);
    }
  } catch (std::string& e) {
    std::printf("EXCEPTION %s\n", e.c_str());
    rc = 1;
  }
  std::cerr.rdbuf(old);
  return rc;
}
