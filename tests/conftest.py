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


def noise_floor(scene, tick):
    """Largest position difference the UNMODIFIED reference shows against itself at `tick` of `scene` when its positions
    are perturbed by at most 1e-6 (about one fp32 ulp) before every tick (tests/golden/sensitivity.py, committed numbers
    in tests/golden/sensitivity.json): the floor an equally valid fp32 evaluation cannot be expected to beat."""
    import json
    with open(os.path.join(GOLDEN, "sensitivity.json")) as f:
        return float(json.load(f)[scene]["every_tick"][str(tick)])


def translation_floor(scene, tick):
    """Largest position difference between the UNMODIFIED reference and ITSELF on the same bodies translated by about one
    body size in x / z (tests/golden/sensitivity.py::translation_floor).  On TetGen meshes with sliver tets the reference's
    fp32 Cholesky is far from translation invariant, which bounds how closely any implementation can agree with it."""
    import json
    with open(os.path.join(GOLDEN, "sensitivity.json")) as f:
        return float(json.load(f)[scene + "_translation"]["max_abs_diff"][str(tick)])
