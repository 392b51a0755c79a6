"""Reduced-size versions of BASELINE.json configs[3] (S4: shape-matching + goal-matching bodies with hull triangles,
CCD and friction) and configs[4] (S5: TetGen soft bodies dropped onto the floor and each other), run live against the
compiled unmodified reference (oracle/_ref travels with the snapshot).  Bars: identical collision
lists and positions within 1e-4 x bbox diagonal while the scene is still deterministic-comparable (first contacts);
afterwards aggregate properties only (the reference itself is chaotic there, see test_solver_gpu.py)."""
import numpy as np
import pytest

from conftest import bbox_diag

pytestmark = pytest.mark.gpu


def test_config4_reduced_shape_goal_ccd_friction(pb, ref):
    from pies_b200 import scenes
    kw = dict(bodies=8, per_side=2, cx=3, cy=4, cz=4, pitch=2.2, y0=0.3, goal_bodies=1)
    r = ref.RefSolver(iterations=6)
    g = pb.Solver(iterations=6)
    _, regions = scenes.build_s4(r, **kw)
    scenes.build_s4(g, **kw)
    assert (g.getTriangles() == r.getTriangles()).all() and len(g.getTriangles()) == 8 * 84
    diag = bbox_diag(r.getVertices())
    seen_tri = 0
    for t in range(1, 40):
        m = scenes.s4_region_script(regions, t)
        r.updateFixedRegions(m); g.updateFixedRegions(m)
        r.tick(); g.tick()
        nt, nf = r.count("tri_collision"), r.count("static_collision")
        st = g.stats()
        if seen_tri <= 1:   # up to and including the tick after the first body-body contacts
            assert (st.triCollisions, st.staticCollisions) == (nt, nf), t
            assert (g.triCollisions() == r.triCollisions()).all(), t
            err = np.abs(g.positions - r.getVertices()).max()
            assert err <= 1e-4 * diag, (t, err, 1e-4 * diag)
        if nt:
            seen_tri += 1
    assert seen_tri >= 2, "the scene never reached body-body contact"
    p = g.positions
    assert np.isfinite(p).all() and not g.simFailed
    assert abs(p[:, 1].mean() - r.getVertices()[:, 1].mean()) <= 0.02 * diag
    assert 0.5 * nt <= g.stats().triCollisions <= 2.0 * nt + 8


def test_config5_reduced_tetgen_bodies(pb, ref):
    from pies_b200 import scenes
    kw = dict(bodies=8, per_side=2, n=3, side=2.0, pitch=2.3, y0=0.25)
    r = ref.RefSolver(iterations=10)
    g = pb.Solver(iterations=10)
    scenes.build_s5(r, g, **kw)
    assert len(g.getVertices()) == r.count("node") and (g.getTriangles() == r.getTriangles()).all()
    diag = bbox_diag(r.getVertices())
    seen_tri = 0
    for t in range(1, 37):
        r.tick(); g.tick()
        nt, nf = r.count("tri_collision"), r.count("static_collision")
        st = g.stats()
        if seen_tri <= 1:
            assert (st.triCollisions, st.staticCollisions) == (nt, nf), t
            assert (g.triCollisions() == r.triCollisions()).all(), t
            err = np.abs(g.positions - r.getVertices()).max()
            assert err <= 1e-4 * diag, (t, err, 1e-4 * diag)
        if nt:
            seen_tri += 1
    assert seen_tri >= 2
    p = g.positions
    assert np.isfinite(p).all() and not g.simFailed and p[:, 1].min() >= -1e-3
    assert abs(p[:, 1].mean() - r.getVertices()[:, 1].mean()) <= 0.02 * diag
