// tests/cxx/refpatch_demo.cpp -- the UNMODIFIED reference (its own headers and sources, compiled where they
// lie under /root/reference) with ONE change made from the outside: the two member functions that run the DP
// fill, DPMatrix::build() (dpmatrix.h:291-317) and DPMatrix::build_subdpm() (dpmatrix.h:319-353), are
// explicitly specialised for the protein instantiation and hand the fill to libaadp.so
// (aadp_fill_pair_general: the similarity matrix the reference itself built + the affine gap model).
// Everything downstream is the reference's own code: Optimal, UnconstrainedNearOptimal (ucw.h),
// ConstrainedNearOptimal (cw.h), AlignmentSet.  This is "Option B" of INTEGRATION.md, done without editing a
// reference file.
//
// Built twice from this one source (tests/cxx/Makefile):
//   refpatch_gpu : with -DREFPATCH_GPU -> the specialisations below are compiled in, the fill runs on the GPU
//   refpatch_cpu : without            -> the pure reference
// tests/test_cxx_dropin.py requires identical output (matrices, optimal alignment, every near-optimal
// alignment with its score).
//
//   usage: refpatch_{gpu,cpu} <matrix file> <align_type 0..4> <gi> <ge> <delta_ratio> <query> <template>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "aa_seq.h"
#include "aasubalib.h"
#include "alib.h"
#include "alignment.h"
#include "cw.h"
#include "dpmatrix.h"
#include "noalib.h"
#include "optimal.h"
#include "sflags.h"
#include "submatrix.h"
#include "ucw.h"

typedef AASubstitutionEval<AASequence, AASequence> AAEval;
typedef DPMatrix<AASequence, AASequence, AAEval> Matrix;

#ifdef REFPATCH_GPU
#include "aadp.h"

// AASubstitutionEval keeps its AliParams private (aasubalib.h:82-85); a maintainer would add a getter, the
// demo passes them on the side.
static const AliParams* g_params = 0;

static aadp_ctx* gpu_ctx() {
  static aadp_ctx* c = aadp_create(0);
  if (!c) throw std::string(aadp_last_error());
  return c;
}

static void gpu_fill(const matrix<float>& sim, int sz1, int sz2, direction_t dir, const int* rect, matrix<DPCell>& out) {
  std::vector<float> s((size_t)sz1 * sz2), score((size_t)sz1 * sz2);
  std::vector<int32_t> pq((size_t)sz1 * sz2), pt((size_t)sz1 * sz2);
  for (int i = 0; i < sz1; ++i)
    for (int j = 0; j < sz2; ++j) s[(size_t)i * sz2 + j] = const_cast<matrix<float>&>(sim)(i, j);
  if (aadp_fill_pair_general(gpu_ctx(), s.data(), sz1 - 2, sz2 - 2, g_params->gap_init_penalty, g_params->gap_extn_penalty,
                             (int)g_params->align_type, AADP_REPRO_REV_BUG, dir == fwd ? AADP_FWD : AADP_REV, rect,
                             score.data(), pq.data(), pt.data()))
    throw std::string(aadp_last_error());
  for (int i = 0; i < sz1; ++i)
    for (int j = 0; j < sz2; ++j) {
      const size_t o = (size_t)i * sz2 + j;
      if (pq[o] != DPCell::null || score[o] != 0.f) out[i][j].setTB(pq[o], pt[o], score[o]);
    }
}

template <>
void Matrix::build() {
  if (simmatrix) delete simmatrix;
  evaluator->pre_calculate(*query_seq, *templ_seq);
  simmatrix = new SimilarityMatrix(*query_seq, *templ_seq, *evaluator);
  const int sz1 = (int)query_seq->size(), sz2 = (int)templ_seq->size();
  if (direction == rev && !islocal) std::cerr << "starting to build rev non-local" << std::endl;  // dpmatrix.h:696
  gpu_fill(*simmatrix, sz1, sz2, direction, 0, *dpmatrix);
}

template <>
void Matrix::build_subdpm(int q1_end, int t1_end, int q2_beg, int t2_beg) {
  if (simmatrix) delete simmatrix;
  evaluator->pre_calculate(*query_seq, *templ_seq);
  simmatrix = new SimilarityMatrix(*query_seq, *templ_seq, *evaluator);
  const int sz1 = (int)query_seq->size(), sz2 = (int)templ_seq->size();
  const int rect[4] = {q1_end, t1_end, q2_beg, t2_beg};
  gpu_fill(*simmatrix, sz1, sz2, direction, rect, *dpmatrix);
}
#endif

static void dump(const char* tag, const Matrix& m) {
  for (int i = 0; i < m.getQuerySize(); ++i)
    for (int j = 0; j < m.getTemplateSize(); ++j) {
      const DPCell* c = m.getCell(i, j);
      std::printf("%s %d %d %.9g %d %d %.9g\n", tag, i, j, c->score, c->prev_query_idx, c->prev_template_idx, m.getSim(i, j));
    }
}

static void print_set(const char* tag, AlignmentSet<AASequence, AASequence, AAEval>& as) {
  std::printf("%s n %d\n", tag, (int)as.size());
  for (size_t k = 0; k < as.size(); ++k) {
    std::printf("%s %d score %.9g pairs", tag, (int)k, as[k].score);
    for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = as[k].begin(); it != as[k].end(); ++it)
      std::printf(" %d:%d", it->query_idx(), it->template_idx());
    std::printf("\n");
  }
}

int main(int argc, char** argv) {
  if (argc != 8) {
    std::fprintf(stderr, "usage: %s matrix align_type gi ge delta query template\n", argv[0]);
    return 2;
  }
  try {
    AliParams params;
    params.submatrix_fn = argv[1];
    params.align_type = static_cast<align_t>(std::atoi(argv[2]));
    params.gap_init_penalty = (float)std::atof(argv[3]);
    params.gap_extn_penalty = (float)std::atof(argv[4]);
#ifdef REFPATCH_GPU
    g_params = &params;
#endif
    NOaliParams np;
    np.delta_ratio = (float)std::atof(argv[5]);
    np.number_suboptimal = 0x3fffffff / 32;  // estimateSize()*20 must not overflow (ucw.h:75)
    AASequence query, templ;
    query.append(std::string("^") + argv[6] + "$");
    templ.append(std::string("^") + argv[7] + "$");
    BlosumMatrix blosum(params.submatrix_fn.c_str());
    AAEval eval(params, blosum);

    Matrix forward(query, templ, eval, fwd, params.align_type);
    dump("F", forward);
    Matrix reverse(query, templ, eval, rev, params.align_type);
    dump("R", reverse);
    if (query.size() > 9 && templ.size() > 9) {
      Matrix sub(query, templ, eval, 2, 3, (int)query.size() - 3, (int)templ.size() - 2, fwd, params.align_type);
      dump("S", sub);
    }
    Optimal<AASequence, AASequence, AAEval> opt(params.align_type);
    AlignmentSet<AASequence, AASequence, AAEval> best(forward, opt);
    print_set("OPT", best);
    if (params.align_type != local) {
      AlignmentSet<AASequence, AASequence, AAEval> as(forward, opt);
      as.clear();
      UnconstrainedNearOptimal<AASequence, AASequence, AAEval> ucw(np);
      ucw.enumerate(forward, as);   // ucw.h:63-191, the reference's own recursion over the GPU-filled matrix
      print_set("UCW", as);
      AlignmentSet<AASequence, AASequence, AAEval> cs(forward, opt);
      cs.clear();
      SuboptFlags sf(true, (size_t)templ.size());
      ConstrainedNearOptimal<AASequence, AASequence, AAEval> cno(np, sf);
      cno.enumerate(forward, cs);   // cw.h:94-284
      print_set("CNO", cs);
    }
  } catch (std::string& e) {
    std::printf("ERROR %s\n", e.c_str());
    return 1;
  }
  return 0;
}
