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


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device and the built library: skip them (with the reason) instead of failing
    when a plain `pytest` runs on a CPU box.  `-m gpu` on the B200 box runs them all."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device (espm_b200 has no CPU fallback)"
    except Exception as exc:      # pragma: no cover
        reason = "torch unavailable: %s" % exc
    if reason is None:
        from espm_b200 import _lib
        if not os.path.exists(_lib.LIBPATH):
            reason = "%s is not built (python -m espm_b200.build)" % _lib.LIBPATH
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
