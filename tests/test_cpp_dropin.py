"""Drop-in check of the C++ surface: ONE host source (tests/cpp/host_demo.cpp, written only against the
reference's public Pies::Solver API) is compiled against the reference (CPU) and against this repo's
Include/Pies/Solver.h + libpies_b200.so (B200); both binaries must produce the same scene.
The binaries are built by __graft_entry__.build() where the reference's glm is available and travel
to the GPU box (tests/cpp/_build/, git-ignored)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bbox_diag

BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
OURS = os.path.join(BUILD, "host_demo_b200")
REF = os.path.join(BUILD, "host_demo_ref")


def _run(binary, scene, ticks, tmp_path):
    out = os.path.join(str(tmp_path), os.path.basename(binary) + "_" + scene + ".bin")
    subprocess.run([binary, scene, str(ticks), out], check=True, stdout=subprocess.PIPE, timeout=300)
    raw = open(out, "rb").read()
    n, nt, nl = np.frombuffer(raw[:12], np.uint32)
    v = np.frombuffer(raw[12:12 + 16 * n], np.float32).reshape(n, 4)
    off = 12 + 16 * int(n)
    tris = np.frombuffer(raw[off:off + 12 * nt], np.uint32)
    lines = np.frombuffer(raw[off + 12 * int(nt):off + 12 * int(nt) + 4 * int(nl)], np.uint32)
    return v[:, :3], v[:, 3], tris, lines


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip("%s not built (needs the reference's glm at build time)" % os.path.relpath(p, ROOT))


def test_header_compiles_against_the_reference_api():
    """The demo host compiled against BOTH headers from the same source (done in build())."""
    _need(OURS, REF)
    assert os.access(OURS, os.X_OK) and os.access(REF, os.X_OK)


def test_reference_build_of_the_demo_runs(tmp_path):
    _need(REF)
    pos, radius, tris, lines = _run(REF, "twobox", 4, tmp_path)
    assert pos.shape == (54, 3) and len(tris) == 3 * 96 and np.isfinite(pos).all()


@pytest.mark.gpu
@pytest.mark.parametrize("scene,ticks", [("twobox", 12), ("sheet", 10), ("shapes", 6), ("boxes", 10), ("tetcube", 12)])
def test_same_host_source_same_result(tmp_path, scene, ticks):
    """Positions within 1e-4 x bbox diagonal (north_star tolerance); radii, triangles and lines exact."""
    _need(OURS, REF)
    g = _run(OURS, scene, ticks, tmp_path)
    r = _run(REF, scene, ticks, tmp_path)
    assert g[0].shape == r[0].shape and len(g[0]) > 0
    assert (g[1] == r[1]).all() and (g[2] == r[2]).all() and (g[3] == r[3]).all()
    tol = 1e-4 * bbox_diag(r[0])
    assert np.abs(g[0] - r[0]).max() <= tol, (scene, float(np.abs(g[0] - r[0]).max()), tol)
