import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_csr(z):
    import scipy.sparse as sp
    return sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))


@pytest.fixture(scope="session")
def golden_hex():
    return load_golden("svg_hex_k6")


@pytest.fixture(scope="session")
def golden_random():
    return load_golden("svg_random_k8")


@pytest.fixture(scope="session")
def golden_square():
    return load_golden("svg_square_k8")


@pytest.fixture(scope="session")
def golden_pca_aa():
    return load_golden("pca_aa_hex")
