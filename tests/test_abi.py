"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/pies_b200.h declares, mirrors the reference's PODs, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pies_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pies_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported(pb):
    L = C.CDLL(pb.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/pies_b200.h but not exported: %s" % missing


def test_binding_covers_every_symbol(pb):
    from pies_b200 import solver
    assert sorted(solver._SIGNATURES) == declared_symbols()


def test_pod_layouts_match_reference(pb):
    # Pies::SolverOptions is 14 x 4 bytes (Solver.h:23-38); Solver::Vertex is 36 bytes (Solver.h:42-49)
    assert C.sizeof(pb.SolverOptions) == 56
    assert pb.VERTEX_DTYPE.itemsize == 36


def test_default_options_match_reference(pb):
    o = pb.SolverOptions()
    pb.lib().pies_b200_default_options(C.byref(o))
    expect = dict(fixedTimestepSize=0.012, timeSubsteps=1, iterations=4, collisionStabilizationIterations=4,
                  collisionThresholdDistance=0.1, collisionThickness=0.05, gravity=10.0, damping=0.006, friction=0.01,
                  staticFrictionThreshold=0.0, floorHeight=0.0, gridSpacing=2.0, threadCount=8, solver=1)
    for k, v in expect.items():
        assert getattr(o, k) == pytest.approx(v, rel=1e-6), k


def test_default_options_match_compiled_reference(pb, ref):
    a, b = pb.SolverOptions(), ref.RefOptions()
    pb.lib().pies_b200_default_options(C.byref(a))
    ref.lib().pref_default_options(C.byref(b))
    assert bytes(a) == bytes(b)


def test_null_arguments_are_rejected(pb):
    L = pb.lib()
    assert L.pies_b200_create(None, -1, None) == -1
    assert L.pies_b200_tick(None, 0.0) == -1
    assert L.pies_b200_vertex_count(None) == 0
    assert L.pies_b200_get_vertices(None) is None


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pb):
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    with pytest.raises(pb.PiesError) as e:
        pb.Solver()
    assert "no CPU fallback" in str(e.value)
    import numpy as np
    with pytest.raises(pb.PiesError):
        pb.probe_ccd(np.zeros((1, 18), np.float32), 0.1)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under pies_b200/ or include/ may import, link or name it."""
    bad = []
    for base in ("pies_b200", "include", "Include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle[/.]|libpies_ref|libpies_oracle|refapi", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_sell_copy_is_the_same_matrix(pb):
    """Host logic of the CG mat-vec's matrix layout (pies_b200/csrc/system.cpp, buildSell): the sliced-ELLPACK copy must
    hold exactly the CSR entries, row by row in CSR order, padded with zeros, rows permuted only inside 256-row windows
    and sorted by length there, slices padded to their longest row."""
    from pies_b200 import solver
    rng = np.random.default_rng(7)
    for n in (1, 31, 32, 33, 256, 257, 1000):
        lens = rng.integers(0, 20, n)
        lens[rng.integers(0, n)] = 45                      # one long row
        row_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        col = rng.integers(0, n, row_ptr[-1]).astype(np.int32)
        val = rng.standard_normal(row_ptr[-1]).astype(np.float32)
        sp, sr, sc, sv = solver.probe_sell(row_ptr, col, val)
        ns = (n + 31) // 32
        assert sp[0] == 0 and len(sp) == ns + 1 and (np.diff(sp.astype(np.int64)) % 32 == 0).all()
        rows = sr[sr != 0xFFFFFFFF]
        assert sorted(rows.tolist()) == list(range(n))                      # a permutation of the rows
        for w0 in range(0, n, 256):                                         # ... inside its own window, longest first
            w = sr[w0:min(w0 + 256, 32 * ns)]
            w = w[w != 0xFFFFFFFF]
            assert w.min() >= w0 and w.max() < min(n, w0 + 256)
            assert (np.diff(lens[w]) <= 0).all()
        x = rng.standard_normal(n)
        y_csr = np.array([np.dot(val[row_ptr[r]:row_ptr[r + 1]].astype(np.float64), x[col[row_ptr[r]:row_ptr[r + 1]]]) for r in range(n)])
        y = np.zeros(n)
        for s in range(ns):
            width = (int(sp[s + 1]) - int(sp[s])) // 32
            for lane in range(32):
                r = sr[32 * s + lane]
                idx = int(sp[s]) + 32 * np.arange(width) + lane
                if r == 0xFFFFFFFF:
                    assert (sv[idx] == 0).all()
                    continue
                m = lens[r]
                assert width >= m
                assert (sc[idx[:m]] == col[row_ptr[r]:row_ptr[r + 1]]).all() and (sv[idx[:m]] == val[row_ptr[r]:row_ptr[r + 1]]).all()
                assert (sv[idx[m:]] == 0).all()
                y[r] = np.dot(sv[idx].astype(np.float64), x[sc[idx]])
        assert np.allclose(y, y_csr, rtol=0, atol=1e-12)
