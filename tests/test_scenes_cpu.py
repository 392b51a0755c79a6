"""Host-side scene builders (pies_b200/scenes.py) checked without a GPU: geometry of the synthetic scenes of
SURVEY section 8(d) and, where the compiled reference is present, that it accepts them."""
import numpy as np
import pytest

from pies_b200 import scenes


@pytest.mark.parametrize("dims", [(2, 2, 2), (3, 4, 4), (4, 8, 8)])
def test_lattice_hull_is_closed_and_outward(dims):
    cx, cy, cz = dims
    tris = scenes.lattice_hull(cx, cy, cz, offset=5).astype(np.int64) - 5
    quads = (cx - 1) * (cy - 1) + (cx - 1) * (cz - 1) + (cy - 1) * (cz - 1)
    assert len(tris) == 4 * quads                       # two faces per direction, two triangles per quad
    pos = np.array([[i, j, k] for i in range(cx) for j in range(cy) for k in range(cz)], float) * 0.5
    surface = (pos == 0).any(1) | (pos[:, 0] == 0.5 * (cx - 1)) | (pos[:, 1] == 0.5 * (cy - 1)) | (pos[:, 2] == 0.5 * (cz - 1))
    assert set(np.unique(tris)) == set(np.flatnonzero(surface))   # exactly the surface nodes (interior nodes are in no triangle)
    n = np.cross(pos[tris[:, 1]] - pos[tris[:, 0]], pos[tris[:, 2]] - pos[tris[:, 0]])
    assert (np.einsum("ij,ij->i", n, pos[tris[:, 0]] - pos.mean(0)) > 0).all()   # outward normals
    edges = {}
    for a, b, c in tris:
        for e in ((a, b), (b, c), (c, a)):
            edges[e] = edges.get(e, 0) + 1
    assert all(v == 1 for v in edges.values()) and all((b, a) in edges for a, b in edges)   # closed, consistently oriented


def test_s3_and_s4_layouts():
    t = scenes.s3_translations(20834)
    assert t.shape == (20834, 3) and t[:, 1].min() >= 0.5 and t[:, 1].max() < 0.5 + 2.5 * 21
    assert np.unique(np.round(t[:, [0, 2]] / 3.0).astype(int), axis=0).shape[0] == 1024     # 32 x 32 columns
    t4 = scenes.s4_translations(15625, 25)
    assert np.unique(np.round(t4 / 5.0 - [0, 0.2, 0]).astype(int), axis=0).shape[0] == 15625  # 25^3 lattice, pitch 5
    j = scenes.lcg_jitter(1000)
    assert j.min() >= 0 and j.max() < 0.01 and (scenes.lcg_jitter(1000) == j).all()            # seeded, reproducible
    m = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (3, 1))
    moved = scenes.s4_region_script(m, 10).reshape(3, 4, 4)
    assert np.allclose(moved[:, 3, 0], 0.2) and np.allclose(moved[:, :3, :3], np.eye(3))       # rigid translation only


def test_reference_accepts_reduced_s4_and_s5(ref):
    r = ref.RefSolver(iterations=2)
    _, regions = scenes.build_s4(r, bodies=2, per_side=2, cx=3, cy=4, cz=4, pitch=2.2, y0=0.3, goal_bodies=1)
    assert r.count("node") == 96 and r.count("triangle") == 2 * 84 and regions.shape == (1, 16)
    r.updateFixedRegions(scenes.s4_region_script(regions, 1))
    r.tick()
    assert np.isfinite(r.getVertices()).all() and not r.simFailed
    r5 = ref.RefSolver(iterations=2)
    scenes.build_s5(r5, None, bodies=2, per_side=2, n=3, side=2.0, pitch=2.3, y0=0.25)
    assert r5.count("tet") > 0 and r5.count("tet") == r5.count("volume") and r5.count("triangle") == 2 * 108
    r5.tick()
    assert np.isfinite(r5.getVertices()).all()
