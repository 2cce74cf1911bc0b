import os
import sys

import numpy as np
import pytest
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def csr_unpack(d, prefix):
    return sps.csr_matrix((d[prefix + "_data"], d[prefix + "_indices"], d[prefix + "_indptr"]),
                          shape=tuple(d[prefix + "_shape"]))


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def golden_bpp():
    return load_golden("bpp.npz")


@pytest.fixture(scope="session")
def golden_als():
    return load_golden("als_small.npz")


@pytest.fixture(scope="session")
def golden_als3():
    return load_golden("als_three.npz")


@pytest.fixture(scope="session")
def golden_online():
    return load_golden("online.npz")


@pytest.fixture(scope="session")
def golden_pre():
    return load_golden("preprocess.npz")


@pytest.fixture(scope="session")
def golden_hals():
    return load_golden("hals.npz")


@pytest.fixture(scope="session")
def golden_quantile():
    return load_golden("quantile.npz")


QUANTILE_CASES = {
    "a": dict(quantiles=50, min_cells=20, max_sample=1000, knn_k=20, refine_knn=True, do_center=False, dims_use=None,
              rand_seed=1),
    "b": dict(quantiles=20, min_cells=5, max_sample=60, knn_k=15, refine_knn=True, do_center=False, dims_use=None,
              rand_seed=7),
    "c": dict(quantiles=30, min_cells=10, max_sample=100, knn_k=20, refine_knn=False, do_center=True,
              dims_use=[0, 2, 3, 5], rand_seed=3),
}
