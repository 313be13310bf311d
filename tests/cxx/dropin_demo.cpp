// tests/cxx/dropin_demo.cpp -- one source, two builds:
//   dropin_demo : compiled against include/hmap2/ (this repo; the fill runs on the GPU via libaadp.so)
//   ref_demo    : compiled against /root/reference (the unmodified reference headers + sources, CPU)
// It follows the reference driver aa_ali.cpp:63-90 (sequences -> BlosumMatrix -> AASubstitutionEval ->
// DPMatrix -> Optimal -> AlignmentSet) and dumps every DPCell plus the optimal alignment, so the
// two builds can be compared byte for byte (tests/test_cxx_dropin.py).
//
//   usage: demo <matrix file> <align_type 0..4> <gi> <ge> <query letters> <template letters>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "aa_seq.h"
#include "aasubalib.h"
#include "alib.h"
#include "alignment.h"
#include "cw.h"
#include "dpmatrix.h"
#include "optimal.h"
#include "noalib.h"
#include "optimal_subali.h"
#include "ucw.h"
#include "submatrix.h"
#ifdef AADP_HMAP2_DPMATRIX_H
#include "optimal_rev.h"  // abstract (un-instantiable) in the reference: optimal_rev.h:29-30
#endif

typedef AASubstitutionEval<AASequence, AASequence> AAEval;
typedef DPMatrix<AASequence, AASequence, AAEval> Matrix;

static void dump(const char* tag, const Matrix& m) {
  for (int i = 0; i < m.getQuerySize(); ++i)
    for (int j = 0; j < m.getTemplateSize(); ++j) {
      const DPCell* c = m.getCell(i, j);
      std::printf("%s %d %d %.6g %d %d %.6g\n", tag, i, j, c->score, c->prev_query_idx, c->prev_template_idx, m.getSim(i, j));
    }
}

int main(int argc, char** argv) {
  if (argc != 7) {
    std::fprintf(stderr, "usage: %s matrix align_type gi ge query template\n", argv[0]);
    return 2;
  }
  try {
    AliParams params;
    params.submatrix_fn = argv[1];
    params.align_type = static_cast<align_t>(std::atoi(argv[2]));
    params.gap_init_penalty = (float)std::atof(argv[3]);
    params.gap_extn_penalty = (float)std::atof(argv[4]);
    AASequence query, templ;
    query.append(std::string("^") + argv[5] + "$");
    templ.append(std::string("^") + argv[6] + "$");
    BlosumMatrix blosum(params.submatrix_fn.c_str());
    AAEval eval(params, blosum);

    Matrix forward(query, templ, eval, fwd, params.align_type);
    dump("F", forward);
    Matrix reverse(query, templ, eval, rev, params.align_type);
    dump("R", reverse);

    Optimal<AASequence, AASequence, AAEval> opt(params.align_type);
    AlignmentSet<AASequence, AASequence, AAEval> alignments(forward, opt);
    std::printf("OPT score %.6g identity %.6g pairs", alignments[0].score, alignments[0].identity);
    for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = alignments[0].begin();
         it != alignments[0].end(); ++it)
      std::printf(" %d:%d", it->query_idx(), it->template_idx());
    std::printf("\n");
    // every near-optimal alignment within 20 % of the optimum (ucw.h): this build enumerates on the GPU
    // (aadp_batch_near_optimal), the reference recurses on the host.  '@' lines are compared as a sorted set (the
    // reference's final sortSet is not stable on equal scores).
    if (params.align_type != local && query.size() * templ.size() < 4000) {
      NOaliParams np;
      np.delta_ratio = 0.2f;
      np.number_suboptimal = 50000;
      UnconstrainedNearOptimal<AASequence, AASequence, AAEval> ucw(np);
      AlignmentSet<AASequence, AASequence, AAEval> near(forward, ucw);
      std::printf("@UCW count %d\n", (int)near.size());
      for (size_t k = 0; k < near.size(); ++k) {
        std::printf("@UCW score %.6g identity %.6g pairs", near[k].score, near[k].identity);
        for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = near[k].begin(); it != near[k].end(); ++it)
          std::printf(" %d:%d", it->query_idx(), it->template_idx());
        std::printf("\n");
      }
    }
    // ... and the constrained variant (cw.h) with suboptimal regions marked on every second block of four template
    // positions (SuboptFlags as nalign.cpp:84-91 builds them)
    if (params.align_type != local && query.size() * templ.size() < 4000) {
      NOaliParams np;
      np.delta_ratio = 0.25f;
      np.number_suboptimal = 50000;
      SuboptFlags flags(true, templ.size());
      for (unsigned int j = 0; j < templ.size(); ++j) flags.Set(j, (j / 4) % 2 == 0);
      ConstrainedNearOptimal<AASequence, AASequence, AAEval> cno(np, flags);
      AlignmentSet<AASequence, AASequence, AAEval> near(forward, cno);
      std::printf("@CNO count %d\n", (int)near.size());
      for (size_t k = 0; k < near.size(); ++k) {
        std::printf("@CNO score %.6g identity %.6g pairs", near[k].score, near[k].identity);
        for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = near[k].begin(); it != near[k].end(); ++it)
          std::printf(" %d:%d", it->query_idx(), it->template_idx());
        std::printf("\n");
      }
    }
    // sub-rectangle fill through the 9-argument constructor (dpmatrix.h:169-189 -> build_subdpm)
    if (query.size() > 9 && templ.size() > 9) {
      Matrix sub(query, templ, eval, 2, 3, (int)query.size() - 3, (int)templ.size() - 2, fwd, params.align_type);
      dump("S", sub);
      // ... and the optimal alignment through that region (optimal_subali.h; used together in ssss.h:621-633)
      Optimal_Subali<AASequence, AASequence, AAEval> osub(2, 3, (int)query.size() - 3, (int)templ.size() - 2);
      AlignmentSet<AASequence, AASequence, AAEval> subali(sub, osub);
      std::printf("SUBOPT score %.6g pairs", subali[0].score);
      for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = subali[0].begin(); it != subali[0].end(); ++it)
        std::printf(" %d:%d", it->query_idx(), it->template_idx());
      std::printf("\n");
      Matrix subr(query, templ, eval, 2, 3, (int)query.size() - 3, (int)templ.size() - 2, rev, params.align_type);
      dump("T", subr);
      // a list of loops closed one after the other, as ssss.h:600-633 does: per loop a sub-matrix and its optimal
      // sub-alignment.  This build hands the whole list to the GPU in one call; the output must not differ.
      const int nq = (int)query.size(), nt = (int)templ.size();
      const int loops[5][4] = {{0, 0, 4, 5}, {3, 2, nq - 4, nt - 3}, {nq - 6, nt - 5, nq - 1, nt - 1}, {1, 1, 2, 6}, {2, 4, 7, 5}};
#ifdef AADP_HMAP2_DPMATRIX_H
      std::vector<aadp::LoopRect> lr(5);
      for (int k = 0; k < 5; ++k) { lr[k].q1_end = loops[k][0]; lr[k].t1_end = loops[k][1]; lr[k].q2_beg = loops[k][2]; lr[k].t2_beg = loops[k][3]; }
      std::vector<AlignedPairList<AASequence, AASequence> > closed;
      aadp::optimal_subalignments(query, templ, eval, lr, &closed);
#endif
      for (int k = 0; k < 5; ++k) {
#ifdef AADP_HMAP2_DPMATRIX_H
        const AlignedPairList<AASequence, AASequence>& la = closed[k];
#else
        Matrix lm(query, templ, eval, loops[k][0], loops[k][1], loops[k][2], loops[k][3], fwd, params.align_type);
        Optimal_Subali<AASequence, AASequence, AAEval> lo(loops[k][0], loops[k][1], loops[k][2], loops[k][3]);
        AlignmentSet<AASequence, AASequence, AAEval> ls(lm, lo);
        const AlignedPairList<AASequence, AASequence>& la = ls[0];
#endif
        std::printf("LOOP %d score %.6g pairs", k, la.score);
        for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = la.begin(); it != la.end(); ++it)
          std::printf(" %d:%d", it->query_idx(), it->template_idx());
        std::printf("\n");
      }
    }
#ifdef AADP_HMAP2_DPMATRIX_H
    // extras of this build: reverse traceback and the near-optimal cell set
    try {
      Optimal_Rev<AASequence, AASequence, AAEval> ropt(params.align_type);
      AlignmentSet<AASequence, AASequence, AAEval> ra(reverse, ropt);
      std::printf("#REV score %.6g pairs", ra[0].score);
      for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = ra[0].begin(); it != ra[0].end(); ++it)
        std::printf(" %d:%d", it->query_idx(), it->template_idx());
      std::printf("\n");
    } catch (std::string& e) {
      std::printf("#REV error %s\n", e.c_str());
    }
    if (params.align_type != local) {
      float thr = 0.f;
      const std::vector<unsigned char>& cells = forward.nearOptimalCells(0.05f, &thr);
      long n = 0;
      for (size_t k = 0; k < cells.size(); ++k) n += cells[k];
      std::printf("#NEAROPT delta 0.05 threshold %.6g cells %ld\n", thr, n);
    }
#endif
  } catch (std::string& e) {
    std::printf("ERROR %s\n", e.c_str());
    return 1;
  }
  return 0;
}
