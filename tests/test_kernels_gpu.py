"""Per-kernel parity (GPU, through the C ABI probes) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py) and, when oracle/_ref travelled, against it live."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _proj_err(got, want):
    # tolerance relative to the size of the projected tet (entries are O(|F| * edge))
    scale = np.maximum(np.abs(want).max(axis=1), 1.0)
    return np.abs(got - want).max(axis=1) / scale


def test_tet_strain_projection_matches_reference(pb):
    g = golden("projections")
    got = pb.probe_tet_projection(g["pos"], g["qinv"], 0.8, 1.0)
    err = _proj_err(got, g["strain"])
    ok = np.ones(len(err), bool)
    ok[g["flat"]] = False  # det F == 0: the sign of the collapsed direction is arbitrary in any SVD (documented)
    # fp32 tolerance: 2e-5 relative (JacobiSVD vs our two-sided Jacobi differ by rounding only)
    assert err[ok].max() < 2e-5, (err[ok].argmax(), err[ok].max())
    assert (got[:, :3] == 0).all()  # projected[0] is the zero differential coordinate (Constraints.cpp:124)


def test_tet_volume_projection_matches_reference(pb):
    g = golden("projections")
    for key, lo, hi in (("volume", 1.0, 1.0), ("volume_09_11", 0.9, 1.1)):
        got = pb.probe_volume_projection(g["pos"], g["qinv"], lo, hi)
        err = _proj_err(got, g[key])
        ok = np.ones(len(err), bool)
        ok[g["flat"]] = False
        assert err[ok].max() < 5e-5, (key, err[ok].argmax(), err[ok].max())


def test_flat_tets_stay_finite(pb):
    g = golden("projections")
    got = pb.probe_tet_projection(g["pos"][g["flat"]], g["qinv"][g["flat"]], 0.8, 1.0)
    assert np.isfinite(got).all()


def test_projection_live_against_reference(pb, ref):
    rng = np.random.default_rng(3)
    n = 4096
    rest = rng.normal(size=(n, 4, 3)).astype(np.float32)
    vol = np.abs(np.linalg.det(rest[:, 1:] - rest[:, :1]))
    rest[vol < 0.3] = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    qinv = np.empty((n, 9), np.float32)
    ref.lib().pref_probe_qinv(n, np.ascontiguousarray(rest.reshape(n, 12)), qinv)
    defo = np.ascontiguousarray((rest + 0.3 * rng.normal(size=rest.shape)).astype(np.float32).reshape(n, 12))
    want = np.empty((n, 12), np.float32)
    ref.lib().pref_probe_tet(n, defo, qinv, 0.8, 1.0, want)
    assert _proj_err(pb.probe_tet_projection(defo, qinv, 0.8, 1.0), want).max() < 2e-5
    ref.lib().pref_probe_volume(n, defo, qinv, 1.0, 1.0, want)
    assert _proj_err(pb.probe_volume_projection(defo, qinv, 1.0, 1.0), want).max() < 5e-5


def test_ccd_decisions_bit_exact(pb):
    """Hit/miss of pointTriangleCCD is integer work: bit-exact.  t agrees to 1e-5 where the cubic is solved."""
    g = golden("ccd")
    hit, t = pb.probe_ccd(g["queries"], 0.1)
    assert (hit == g["hit"]).all(), np.flatnonzero(hit != g["hit"])[:10]
    both = hit == 1
    assert both.sum() > 50
    assert np.abs(t[both] - g["t"][both]).max() < 1e-5


def test_edge_edge_ccd_decisions_bit_exact(pb):
    """edgeEdgeCCD (reference CollisionDetection.cpp:304-418, never emitted by the reference's tick: SURVEY F13) restated
    with the reference's shadowing bug; hit/miss bit-exact on 6 000 fixture queries (static hits, crossings, parallel
    edges, degenerate input), t to 5e-5 where the cubic is solved (1.3e-5 measured)."""
    g = golden("ccd")
    hit, t = pb.probe_edge_ccd(g["edge_queries"])
    keep = g["edge_compare"] == 1      # all but the exactly-parallel queries whose outcome is a root at the interval's end
    assert keep.sum() > 5500 and keep[1500:2000].sum() > 20
    assert (hit[keep] == g["edge_hit"][keep]).all(), np.flatnonzero(keep & (hit != g["edge_hit"]))[:10]
    both = keep & (hit == 1)
    assert both.sum() > 200 and (hit == 0).sum() > 200
    moving = both & (g["edge_t"] != 1.0)
    assert moving.sum() > 20
    assert np.abs(t[both] - g["edge_t"][both]).max() < 5e-5   # closed-form cubic here, companion-matrix eigenvalues there


def test_cell_ranges_bit_exact(pb):
    g = golden("ranges")
    mins, lens = pb.probe_tri_range(g["tri_pos"], g["tri_prev"])
    assert (mins == g["tri_min"]).all() and (lens == g["tri_len"]).all()
    assert (lens[:16].min(axis=1) == 0).all()  # integer-plane triangles are not inserted
    assert (lens[16:24] == 0).all()            # over the 50-cell cap: empty range
    mins, lens = pb.probe_node_range(g["node_pos"], g["node_radius"], 2.0)
    assert (mins == g["node_min"]).all() and (lens == g["node_len"]).all()


@pytest.mark.parametrize("n,bits", [(1, 8), (31, 16), (2048, 24), (2049, 40), (100003, 64), (327680, 17), (327681, 17),
                                    (1 << 20, 33)])
def test_radix_sort_is_a_stable_sort(pb, n, bits):
    rng = np.random.default_rng(n)
    mask = np.uint64((1 << bits) - 1) if bits < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) & mask
    keys[: n // 3] = keys[n // 3: 2 * (n // 3)][: n // 3] if n >= 3 else keys[: n // 3]  # force duplicates
    vals = np.arange(n, dtype=np.uint32)
    k, v = pb.probe_sort_pairs(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    assert (k == keys[order]).all()
    assert (v == vals[order]).all()  # equal keys keep their input order => buckets come out in ascending element index


def test_empty_inputs(pb):
    assert pb.probe_tet_projection(np.zeros((0, 12), np.float32), np.zeros((0, 9), np.float32), 0.8, 1.0).shape == (0, 12)
    hit, t = pb.probe_ccd(np.zeros((0, 18), np.float32), 0.1)
    assert len(hit) == 0
    k, v = pb.probe_sort_pairs(np.zeros(0, np.uint64), np.zeros(0, np.uint32), 16)
    assert len(k) == 0
