import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def blosum():
    from alignment_algos_b200.submatrix import read_matrix, BLOSUM62
    return read_matrix(BLOSUM62)


@pytest.fixture(scope="session")
def golden_float():
    # reference outputs for scoring that is NOT on a dyadic grid (oracle/gen_golden_float.py)
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_float.npz"))


@pytest.fixture(scope="session")
def golden_sub():
    # reference outputs of the 9-argument constructor / build_subdpm (oracle/gen_golden_sub.py)
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_sub.npz"))


@pytest.fixture(scope="session")
def golden_tab():
    # reference fills driven by a table-backed Evaluator with position-dependent gaps (oracle/gen_golden_tab.py)
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_tab.npz"))
