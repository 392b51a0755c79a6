"""PBD path parity (GPU, through the C ABI) against golden data from the unmodified reference
(reference Solver::tickPBD, Src/Solver.cpp:40-160; fixtures: tests/golden/make_golden.py pbd).

The reference's PBD iteration is numerically unstable on chains (its distance projection moves only node 0
of a link; a hanging or self-overlapping rope diverges within ~15-25 ticks IN THE REFERENCE), so chain
scenes are compared inside the window before its own divergence; the distance-box scene is stable and is
compared for 60 ticks through floor contact and box-on-box node-node collisions."""
import numpy as np
import pytest

from conftest import bbox_diag, golden
from pies_b200 import scenes

pytestmark = pytest.mark.gpu
TOL = 1e-4  # x scene bbox diagonal (north_star tolerance)


def _check(s, g, prefix, ticks):
    tol = TOL * bbox_diag(g["%s_pos%d" % (prefix, ticks[0])])
    t = 0
    for k in ticks:
        while t < k:
            s.tick(); t += 1
        err = float(np.abs(s.positions - g["%s_pos%d" % (prefix, k)]).max())
        assert err <= tol, (prefix, k, err, tol)
        verr = float(np.abs(s.velocities - g["%s_vel%d" % (prefix, k)]).max())
        assert verr <= tol / 0.012, (prefix, k, verr)
    return tol


def test_pbd_boxes_floor_and_node_node_collisions(pb):
    g = golden("pbd")
    s = pb.Solver(**scenes.S2_OPTIONS)
    scenes.build_pbd_boxes(s)
    _check(s, g, "boxes", (1, 10, 25, 40, 60))
    assert s.stats().collisionProjections > 0 and not s.simFailed


def test_pbd_node_hash_occupancy_bit_exact(pb):
    """Cells keyed by exact (x,y,z), members ascending: identical to the reference's SpatialHash<Node>
    (parallelBulkInsert + NodeCompRange) on the reference's own tick-40 positions."""
    g = golden("pbd")
    s = pb.Solver(**scenes.S2_OPTIONS)
    scenes.build_pbd_boxes(s)
    pos = g["boxes_pos40"]
    s.setState(pos, pos, np.zeros_like(pos))
    s.detectNodes()
    cells, counts, members = s.nodeOccupancy()
    # the fixture is already sorted by (x, y, z) with members in bucket order (oracle/ref_driver.cpp)
    assert cells.shape == g["boxes_occ_cells"].shape
    assert (cells == g["boxes_occ_cells"]).all()
    assert (counts == g["boxes_occ_counts"]).all()
    assert (members == g["boxes_occ_members"]).all()


def test_pbd_rope_helix_before_the_reference_diverges(pb):
    g = golden("pbd")
    s = pb.Solver(**scenes.S2_OPTIONS)
    scenes.build_rope(s, n=2000, helix_radius=2.0)
    _check(s, g, "rope", (1, 10, 20))


def test_pbd_rope_spiral_with_active_self_collisions(pb):
    g = golden("pbd")
    s = pb.Solver(**scenes.S2_OPTIONS)
    scenes.build_rope(s, n=1500, shape="spiral", pinned=False)
    _check(s, g, "spiral", (1, 2, 3, 5, 8))
    assert s.stats().collisionProjections > 0


def test_pbd_sheets_distance_and_bend(pb):
    g = golden("pbd")
    s = pb.Solver(**scenes.S2_OPTIONS)
    s.createSheet((0.0, 2.0, 0.0), 0.5, 1.0, 0.8)
    s.createBendSheet((12.0, 2.0, 0.0), 0.5, 0.6)
    _check(s, g, "sheets", (1, 5, 10))


def test_pbd_live_against_reference(pb, ref):
    """Same scene on both sides, live (oracle/_ref travels with the snapshot): 30 ticks of the box scene."""
    g = pb.Solver(**scenes.S2_OPTIONS); r = ref.RefSolver(**scenes.S2_OPTIONS)
    scenes.build_pbd_boxes(g); scenes.build_pbd_boxes(r)
    for _ in range(30):
        g.tick(); r.tick()
    tol = TOL * bbox_diag(r.positions)
    assert float(np.abs(g.positions - r.positions).max()) <= tol


def test_pbd_run_to_run_bit_stable(pb):
    out = []
    for _ in range(2):
        s = pb.Solver(**scenes.S2_OPTIONS)
        scenes.build_pbd_boxes(s)
        for _ in range(30):
            s.tick()
        out.append(s.positions.copy())
    assert (out[0] == out[1]).all()


def test_pbd_rope_100k_properties(pb):
    """Config 2 at full size (100 000-node chain): finite, links stay near rest length while in free fall, and
    the even-then-odd link order runs as two colour batches (projections counted)."""
    s = pb.Solver(**scenes.S2_OPTIONS)
    n = scenes.build_rope(s, n=100000)
    for _ in range(5):
        s.tick()
    p = s.positions
    assert np.isfinite(p).all()
    d = np.linalg.norm(p[1:] - p[:-1], axis=1)
    assert abs(d - 0.5).max() < 1e-2
    st = s.stats()
    assert st.projectionsLastTick >= 4 * (n - 1)


def test_pbd_with_tets_is_refused(pb):
    """The reference produces NaN on the first PBD tick of a tet body; we refuse instead."""
    s = pb.Solver(solver="PBD")
    s.createTetBox((0, 1, 0), 1.0, (0, 0, 0), 1000.0, 1.0, False)
    with pytest.raises(pb.PiesError):
        s.tick()
