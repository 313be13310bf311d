"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/aadp.h declares; compute calls fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "aadp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(aadp_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from alignment_algos_b200._lib import lib, EXPORTS
    L = lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), "libaadp.so does not export " + name
    assert sorted(EXPORTS) == declared, "python binding list and header disagree"
    assert b"sm_100a" in L.aadp_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import alignment_algos_b200 as a
    with pytest.raises(a.AadpError) as e:
        a.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "alignment_algos_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "aadp_oracle" not in src and "libaadp_ref" not in src, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read()


def test_tb_row_bytes_helper():
    from alignment_algos_b200._lib import lib
    L = lib()
    for Lt, want in [(1, 16), (8, 16), (32, 16), (33, 32), (100, 64), (500, 256), (512, 256), (513, 272)]:
        assert L.aadp_tb_row_bytes(Lt) == want


def test_submatrix_reader_matches_reference_format(tmp_path):
    from alignment_algos_b200.submatrix import read_matrix, blosum62
    p = tmp_path / "m.txt"
    p.write_text("# c1\n# c2\n  A  C  G  T\nA 1 -1 -1 -1\nC -1 1 -1 -1\nG -1 -1 1 -1\nT -1 -1 -1 1.5\n")
    alpha, m = read_matrix(str(p))
    assert alpha == "ACGT" and m.shape == (4, 4) and m[3, 3] == 1.5 and m[0, 1] == -1
    a20, b = blosum62()
    assert len(a20) == 20 and b[a20.index("W"), a20.index("W")] == 11 and (b == b.T).all()


def test_driver_scripts_have_no_undefined_names():
    # bench.py and __graft_entry__.py only run on the GPU box; a name that exists nowhere (a refactoring slip) must be
    # caught here, on the CPU
    import builtins
    import symtable
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for fn in ("bench.py", "__graft_entry__.py"):
        st = symtable.symtable(open(os.path.join(root, fn)).read(), fn, "exec")
        mod_names = {s.get_name() for s in st.get_symbols()}
        bad = []

        def walk(t):
            for ch in t.get_children():
                for s in ch.get_symbols():
                    n = s.get_name()
                    if s.is_global() and s.is_referenced() and n not in mod_names and not hasattr(builtins, n):
                        bad.append((ch.get_name(), n))
                walk(ch)
        walk(st)
        assert not bad, "%s: undefined names %s" % (fn, bad)
