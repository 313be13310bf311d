// hmap2/submatrix.h -- substitution matrices (reference submatrix.h:19-48, submatrix.cpp:16-54).
#ifndef AADP_HMAP2_SUBMATRIX_H
#define AADP_HMAP2_SUBMATRIX_H

#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

class SubstitutionMatrix {
 public:
  bool hasLetter(char x) const { return alphabet.find(x) != std::string::npos; }

  // score of aligning letter a to letter b. The reference dereferences map::find without a check
  // (submatrix.h:36-38, undefined behaviour for unknown letters); this build throws instead.
  float score(char a, char b) const {
    std::map<char, std::map<char, float> >::const_iterator r = sub_matrix.find(a);
    if (r == sub_matrix.end()) throw std::string("Letter not in substitution matrix: ") + a;
    std::map<char, float>::const_iterator c = r->second.find(b);
    if (c == r->second.end()) throw std::string("Letter not in substitution matrix: ") + b;
    return c->second;
  }

  // extension used by the GPU binding: the alphabet and the dense A x A table
  const std::string& getAlphabet() const { return alphabet; }
  std::vector<float> dense() const {
    const size_t n = alphabet.size();
    std::vector<float> d(n * n);
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) d[i * n + j] = score(alphabet[i], alphabet[j]);
    return d;
  }

 protected:
  std::string alphabet;
  std::map<char, std::map<char, float> > sub_matrix;
};

// BLOSUM-format file: '#' comment lines, one header line of letters, then one row per letter
// ("<label> v v v ...").
class BlosumMatrix : public SubstitutionMatrix {
 public:
  explicit BlosumMatrix(const char* filename) {
    std::ifstream in(filename);
    if (!in.good()) throw std::string("File not found (substitution matrix) ") + filename;  // submatrix.cpp:24-26
    std::string line;
    while (std::getline(in, line))
      if (line.empty() || line[0] != '#') break;
    for (size_t k = 0; k < line.size(); ++k)
      if (line[k] != ' ' && line[k] != '\n' && line[k] != '\r' && line[k] != '\t') alphabet.push_back(line[k]);
    const size_t n = alphabet.size();
    for (size_t i = 0; i < n; ++i) {
      std::string label;
      in >> label;
      for (size_t j = 0; j < n; ++j) {
        float v = 0.f;
        in >> v;
        sub_matrix[alphabet[i]][alphabet[j]] = v;
      }
    }
  }
};

#endif
