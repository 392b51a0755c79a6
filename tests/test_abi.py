"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/pies_b200.h declares, mirrors the reference's PODs, and refuses to run without a GPU."""
import ctypes as C
import os
import re

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
