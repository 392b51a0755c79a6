"""CPU suite: the numpy restatement (oracle/port.py) pinned against the golden fixtures that the UNMODIFIED
compiled reference produced (tests/golden/make_golden.py), and live against oracle/_ref where it is present."""
import numpy as np
import pytest

from conftest import golden
from oracle import port


def test_distance_projection_against_reference(ref):
    rng = np.random.default_rng(5)
    n = 4000
    pos = rng.normal(size=(n, 2, 3)).astype(np.float32)
    pos[:16, 1] = pos[:16, 0]                       # coincident nodes: dir falls back to (1,0,0)
    rest = rng.uniform(0.1, 2.0, n).astype(np.float32)
    out = np.empty((n, 6), np.float32)
    ref.lib().pref_probe_distance(n, np.ascontiguousarray(pos.reshape(n, 6)), rest, out)
    got = port.distance_projection(pos, rest).reshape(n, 6)
    # fp32 both sides; the only freedom is how length() rounds
    assert np.abs(got - out).max() <= 4e-6
    assert (got[:, 3:] == pos[:, 1]).all()          # node 1 never moves (Constraints.cpp:35-36)


def test_tet_strain_projection_against_golden():
    g = golden("projections")
    keep = np.ones(len(g["pos"]), bool)
    keep[g["flat"]] = False                          # det F ~ 0: the sign test is decided by rounding
    got = port.tet_strain_projection(g["pos"], g["qinv"], 0.8, 1.0)
    err = np.abs(got - g["strain"])[keep].max()
    assert err <= 2e-5, err                          # fp64 LAPACK SVD vs Eigen fp32 JacobiSVD
    assert (g["strain"][:, :3] == 0).all() and (got[:, :3] == 0).all()   # projected[0] = 0 (Constraints.cpp:124)


@pytest.mark.parametrize("key,lo,hi", [("volume", 1.0, 1.0), ("volume_09_11", 0.9, 1.1)])
def test_tet_volume_projection_against_golden(key, lo, hi):
    g = golden("projections")
    keep = np.ones(len(g["pos"]), bool)
    keep[g["flat"]] = False                          # sigma3 ~ 0: computeD divides by a vanishing gradient
    got = port.tet_volume_projection(g["pos"], g["qinv"], lo, hi)
    scale = np.abs(g[key]).max(axis=1)[keep]
    err = (np.abs(got - g[key]).max(axis=1)[keep] / np.maximum(scale, 1.0)).max()
    assert err <= 5e-5, err


def test_volume_projection_restores_the_volume():
    """Size-independent property: with omega in [1,1] the corrected gradient has |determinant| 1."""
    g = golden("projections")
    out = g["volume"].reshape(-1, 4, 3)
    Fhat = np.stack([out[:, 1], out[:, 2], out[:, 3]], axis=2).astype(np.float64)
    keep = np.ones(len(out), bool); keep[g["flat"]] = False
    det = np.linalg.det(Fhat)[keep]                  # inverted inputs stay inverted (no sign fix, :206-255)
    assert np.abs(np.abs(det) - 1.0).max() <= 2e-3


def test_triangle_ranges_bit_exact_against_golden():
    g = golden("ranges")
    mins, lens = port.tri_cell_range(g["tri_pos"], g["tri_prev"])
    assert (mins == g["tri_min"]).all()
    assert (lens == g["tri_len"]).all()
    assert (g["tri_len"][16:24] == 0).any(axis=1).all()      # over the 50-cell cap: empty range
    assert (g["tri_len"][:16, 0] == 0).all()                 # integer-plane quirk


def test_node_ranges_bit_exact_against_golden():
    g = golden("ranges")
    mins, lens = port.node_cell_range(g["node_pos"], g["node_radius"], 2.0)
    assert (mins == g["node_min"]).all()
    assert (lens == g["node_len"]).all()
    assert (g["node_len"][:8] == 0).all()


def test_triangle_hash_occupancy_against_golden():
    """Two tet boxes: cell -> members multiset of the triangle hash, from the fixture's
    node state and the tetbox factory's triangle list."""
    c = golden("collisions")
    f = golden("factories")
    tris = f["tetbox_tris"].reshape(-1, 3).astype(np.int64)
    nper = len(f["tetbox_pos"])
    tris = np.concatenate([tris, tris + nper])
    pos, prev = c["occ_pos"], c["occ_prev"]
    mins, lens = port.tri_cell_range(pos[tris], prev[tris])
    cells, counts, members = port.cell_occupancy(mins, lens)
    assert (cells == c["occ_cells"]).all()
    assert (counts == c["occ_counts"]).all()
    assert (members == c["occ_members"]).all()


def test_node_hash_occupancy_against_golden():
    """PBD boxes at tick 40: node hash occupancy (NodeCompRange + parallelBulkInsert)."""
    p = golden("pbd")
    radius = np.full(len(p["boxes_pos40"]), 0.5, np.float32)   # createBox(scale 1): radius 0.5 * scale
    mins, lens = port.node_cell_range(p["boxes_pos40"], radius, 2.0)
    cells, counts, members = port.cell_occupancy(mins, lens)
    assert (cells == p["boxes_occ_cells"]).all()
    assert (counts == p["boxes_occ_counts"]).all()
    assert (members == p["boxes_occ_members"]).all()


def test_free_fall_restatement_against_golden():
    """A tet box in free fall (factories.npz, `tetbox`): rest-state constraints, no contact for the first ticks,
    so the tick is the inertial prediction + velocity update."""
    f = golden("factories")
    pos0, vel0 = f["tetbox_pos"], f["tetbox_vel"]
    traj = f["tetbox_traj"]                           # (ticks, n, 3) positions after tick 1..
    pos, vel = port.free_fall(pos0, vel0, 1)
    diag = float(np.linalg.norm(pos0.max(0) - pos0.min(0)))
    assert np.abs(pos - traj[0]).max() <= 1e-4 * diag


def test_collision_matrix_against_the_reference_system(ref):
    """The reference's own S + C_t of a contact tick minus its S of a contact-free tick is the restated collision matrix
    (duplicates of the list counted with their multiplicity, 1e4 per point-triangle copy and per floor copy), and the
    streamable CSR form the CUDA mat-vec reads is the same matrix."""
    from oracle import port

    def scene(s):
        s.createTetBox((0.1, 0.3, 0.1), 1.0, (0, 0, 0), 1000.0, 1.0, False)
        s.createTetBox((0.4, 2.6, 0.3), 1.0, (0, -5, 0), 1000.0, 1.0, False)

    r = ref.RefSolver(iterations=10)
    scene(r)
    n = 54

    def dense():
        rows, cols, vals, _, _ = r.system()
        m = np.zeros((n, n))
        np.add.at(m, (rows, cols), vals.astype(np.float64))
        return m

    r.tick(2)
    assert r.count("tri_collision") == 0 and r.count("static_collision") == 0
    S = dense()
    assert np.allclose(S, S.T)
    seen = 0
    for t in range(2, 16):
        r.tick()
        tri, floor = r.triCollisions(), r.staticCollisions()
        if not len(tri) and not len(floor):
            continue
        seen += 1
        C = dense() - S
        want = port.collision_matrix(n, tri, floor)
        assert np.abs(C - want).max() <= 1e-3 * max(1.0, np.abs(want).max()), t      # fp32 sums of S + multiples of 1e4
        ptr, col, val, diag = port.collision_csr(n, tri, floor)
        csr = np.diag(diag)
        for i in range(n):
            np.add.at(csr[i], col[ptr[i]:ptr[i + 1]], val[ptr[i]:ptr[i + 1]])
        assert np.array_equal(csr, want), t
        assert ptr[-1] == 6 * len(np.unique(tri, axis=0)) if len(tri) else ptr[-1] == 0
    assert seen >= 8
