// tests/cxx/tabeval_demo.cpp -- a user-defined Evaluator with POSITION-DEPENDENT gap penalties through the DPMatrix
// template API.  The reference's own evaluators of that kind (hmap_eval.h:63-117,
// gn2_eval.h:99-158) need the un-vendored Troll library and cannot compile here, so this source defines two
// evaluators with the same gap-function structure over plain AASequences:
//   style 1 (HMAP-like): per-template-position gap_init/gap_extn, a gap between two template positions costs
//                        min(gi[t1],gi[t2]) + min(ge[t1],ge[t2])*(dist-2), for deletions and insertions alike;
//   style 2 (GN2-like):  deletions from a pairwise (t1,t2) table with a 8100 cutoff, insertions
//                        v_gi[t1] + v_ge[t1]*(di-2) + v_cn[t1].
// Built twice from this one source (tests/cxx/Makefile):
//   tabeval_demo : against include/hmap2/ -- DPMatrix tabulates the evaluator and fills on the GPU
//                  (aadp_fill_pair_tabulated)
//   tabeval_ref  : against the unmodified reference headers and sources -- the CPU fill of dpmatrix.h
// tests/test_cxx_dropin.py requires identical output.
//
//   usage: tabeval_{demo,ref} <matrix file> <align_type 0..4> <gi> <ge> <style 1|2> <query> <template>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "aa_seq.h"
#include "alib.h"
#include "alignment.h"
#include "dpmatrix.h"
#include "evaluator.h"
#include "optimal.h"
#include "submatrix.h"

template <class S1, class S2>
class PositionalGapEval : public Evaluator<S1, S2, PositionalGapEval<S1, S2> > {
 public:
  PositionalGapEval(AliParams& p, SubstitutionMatrix& m, int style_) : params(&p), sub_matrix(&m), style(style_) {}

  float similarity(const S1& q, const S2& t, int q_pos, int t_pos) const {
    if (q[q_pos]->isHead() || q[q_pos]->isTail() || t[t_pos]->isHead() || t[t_pos]->isTail()) return 0.f;
    const float w = 0.5f + 0.125f * (float)(t_pos % 5);  // a profile column weight
    return sub_matrix->score(q[q_pos]->olc, t[t_pos]->olc) * w;
  }

  float gap_init(int t_pos) const { return params->gap_init_penalty + 0.37f * (float)((t_pos * 7) % 5); }
  float gap_extn(int t_pos) const { return params->gap_extn_penalty + 0.11f * (float)((t_pos * 3) % 4); }

  float deletion(const S1& q, const S2& t, int q_pos1, int q_pos2, int t_pos1, int t_pos2) const {
    const int dist = t_pos2 - t_pos1;
    if (dist < 2) return 0;
    float gp;
    if (style == 1) {
      const float gi = std::min(gap_init(t_pos1), gap_init(t_pos2));
      const float ge = std::min(gap_extn(t_pos1), gap_extn(t_pos2));
      gp = gi + ge * (dist - 2);
    } else {
      const int p1 = t_pos1, p2 = t_pos2 - 2;
      gp = 8100.f;
      if ((p1 * 13 + p2 * 7) % 11 != 0)
        gp = (params->gap_init_penalty + 0.25f * (float)((p1 + 2 * p2) % 7)) +
             (params->gap_extn_penalty + 0.05f * (float)((3 * p1 + p2) % 5)) * (dist - 2) + 0.3f * (float)((p1 * p2) % 3);
    }
    switch (params->align_type) {
      case global:
      case global_local:
        return gp;
      case local:
      case semi_local:
      case local_global:
        if (t[t_pos1]->isHead() || t[t_pos2]->isTail()) return 0;
        return gp;
      default:
        throw std::string("Illegal gap style");
    }
  }

  float insertion(const S1& q, const S2& t, int q_pos1, int q_pos2, int t_pos1, int t_pos2) const {
    const int dist = q_pos2 - q_pos1;
    if (dist < 2) return 0;
    float gp;
    if (style == 1) {
      const float gi = std::min(gap_init(t_pos1), gap_init(t_pos2));
      const float ge = std::min(gap_extn(t_pos1), gap_extn(t_pos2));
      gp = gi + ge * (dist - 2);
    } else {
      gp = gap_init(t_pos1) + gap_extn(t_pos1) * (dist - 2) + 0.2f * (float)(t_pos1 % 6);
    }
    switch (params->align_type) {
      case global:
      case local_global:
        return gp;
      case local:
      case semi_local:
      case global_local:
        if (q[q_pos1]->isHead() || q[q_pos2]->isTail()) return 0;
        return gp;
      default:
        throw std::string("Illegal gap style");
    }
  }

  void pre_calculate(const S1&, const S2&) const {}
  void post_process(SimilarityMatrix&) const {}

 private:
  AliParams* params;
  SubstitutionMatrix* sub_matrix;
  int style;
};

typedef PositionalGapEval<AASequence, AASequence> PGEval;
typedef DPMatrix<AASequence, AASequence, PGEval> Matrix;

static void dump(const char* tag, const Matrix& m) {
  for (int i = 0; i < m.getQuerySize(); ++i)
    for (int j = 0; j < m.getTemplateSize(); ++j) {
      const DPCell* c = m.getCell(i, j);
      std::printf("%s %d %d %.9g %d %d %.9g\n", tag, i, j, c->score, c->prev_query_idx, c->prev_template_idx, m.getSim(i, j));
    }
}

int main(int argc, char** argv) {
  if (argc != 8) {
    std::fprintf(stderr, "usage: %s matrix align_type gi ge style query template\n", argv[0]);
    return 2;
  }
  try {
    AliParams params;
    params.submatrix_fn = argv[1];
    params.align_type = static_cast<align_t>(std::atoi(argv[2]));
    params.gap_init_penalty = (float)std::atof(argv[3]);
    params.gap_extn_penalty = (float)std::atof(argv[4]);
    const int style = std::atoi(argv[5]);
    AASequence query, templ;
    query.append(std::string("^") + argv[6] + "$");
    templ.append(std::string("^") + argv[7] + "$");
    BlosumMatrix blosum(params.submatrix_fn.c_str());
    PGEval eval(params, blosum, style);

    Matrix forward(query, templ, eval, fwd, params.align_type);
    dump("F", forward);
    Matrix reverse(query, templ, eval, rev, params.align_type);
    dump("R", reverse);
    Optimal<AASequence, AASequence, PGEval> opt(params.align_type);
    AlignmentSet<AASequence, AASequence, PGEval> alignments(forward, opt);
    std::printf("OPT score %.9g pairs", alignments[0].score);
    for (std::list<AlignedPair<AASequence, AASequence> >::const_iterator it = alignments[0].begin();
         it != alignments[0].end(); ++it)
      std::printf(" %d:%d", it->query_idx(), it->template_idx());
    std::printf("\n");
  } catch (std::string& e) {
    std::printf("ERROR %s\n", e.c_str());
    return 1;
  }
  return 0;
}
