import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_steps():
    return load_golden("steps_small.npz")


@pytest.fixture(scope="session")
def golden_identity():
    return load_golden("steps_identity.npz")


@pytest.fixture(scope="session")
def golden_bisect():
    return load_golden("bisect_small.npz")


@pytest.fixture(scope="session")
def golden_variants():
    return load_golden("variants_small.npz")


@pytest.fixture(scope="session")
def golden_fits():
    return load_golden("fits_small.npz")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
