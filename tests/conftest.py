import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def pb():
    """The product binding; -m gpu tests fail loudly (no fallback) if the CUDA library is missing."""
    import pies_b200
    pies_b200.lib()
    return pies_b200


@pytest.fixture(scope="session")
def ref():
    """The compiled unmodified reference, when it travelled with the snapshot (oracle/_ref)."""
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/libpies_ref.so not present")
    return refapi


def bbox_diag(p):
    p = np.asarray(p)
    return float(np.linalg.norm(p.max(0) - p.min(0)))
