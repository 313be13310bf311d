// hmap2/matrix.h -- dense row-major 2-D array with the access surface the DP path uses from the
// reference's valarray matrix (matrix.h:150-235): (r,c), [r][c], rows(), cols(), size().
#ifndef AADP_HMAP2_MATRIX_H
#define AADP_HMAP2_MATRIX_H

#include <vector>

template <class val_t>
class matrix {
 public:
  matrix(int nrows, int ncols) : nr(nrows), nc(ncols), v((size_t)nrows * ncols) {}

  int size() const { return nr * nc; }
  int rows() const { return nr; }
  int cols() const { return nc; }

  val_t operator()(int r, int c) const { return v[(size_t)r * nc + c]; }
  val_t& operator()(int r, int c) { return v[(size_t)r * nc + c]; }

  val_t* operator[](int r) { return &v[(size_t)r * nc]; }
  const val_t* operator[](int r) const { return &v[(size_t)r * nc]; }

  val_t* data() { return v.data(); }
  const val_t* data() const { return v.data(); }

 private:
  int nr, nc;
  std::vector<val_t> v;
};

#endif
