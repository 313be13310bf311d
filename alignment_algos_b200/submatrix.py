"""Substitution-matrix file reader with the behaviour of the reference's BlosumMatrix
(submatrix.cpp:16-54): '#' comment lines, one header line of letters (blanks ignored), then one
row per letter: `<label> v v v ...`.  Returns (alphabet, float32 matrix)."""
import os

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
BLOSUM62 = os.path.join(DATA, "BLOSUM62")
AA20 = "ARNDCQEGHILKMFPSTWYV"  # synthetic alphabet, codes 0..19


def read_matrix(path):
    if not os.path.exists(path):
        raise FileNotFoundError("File not found (substitution matrix) " + path)  # submatrix.cpp:24-26
    with open(path) as f:
        lines = f.read().split("\n")
    k = 0
    while k < len(lines) and lines[k].startswith("#"):
        k += 1
    alphabet = "".join(ch for ch in lines[k] if ch not in " \n\r\t")
    toks = " ".join(lines[k + 1:]).split()
    n = len(alphabet)
    m = np.zeros((n, n), np.float32)
    pos = 0
    for i in range(n):
        pos += 1  # row label (submatrix.cpp:48)
        for j in range(n):
            m[i, j] = float(toks[pos])
            pos += 1
    return alphabet, m


def blosum62(alphabet=AA20):
    """BLOSUM62 restricted to `alphabet` (default: the 20 standard residues)."""
    full, m = read_matrix(BLOSUM62)
    idx = [full.index(ch) for ch in alphabet]
    return alphabet, np.ascontiguousarray(m[np.ix_(idx, idx)])
