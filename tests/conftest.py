import os
import sys

import numpy as np
import pytest
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))

    def __getitem__(self, k):
        return self.z[k]

    def csr(self, prefix):
        return sparse.csr_matrix((self.z[prefix + "_data"], self.z[prefix + "_indices"], self.z[prefix + "_indptr"]),
                                 shape=tuple(self.z[prefix + "_shape"]))


@pytest.fixture(scope="session")
def moons():
    return Golden("twomoons500")


@pytest.fixture(scope="session")
def blobs():
    return Golden("blobs2000")


@pytest.fixture(scope="session")
def plap():
    return Golden("plaplace2000")


@pytest.fixture(scope="session")
def small():
    return Golden("small300")


def rel_err(a, b):
    """The parity metric of SURVEY.md 8d: max|a-b| / max|b|."""
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) / np.max(np.abs(b)))
