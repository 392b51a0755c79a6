"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Run where /root/reference exists:   make -C oracle ref && python tests/golden/make_golden.py
The fixtures travel with the repo, so the parity tests do not need /root/reference (nor oracle/_ref)
at run time.  Every array comes out of the reference's own functions (through oracle/ref_driver.cpp).
"""
import contextlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.refapi import RefSolver, lib  # noqa: E402
from pies_b200 import scenes  # noqa: E402


@contextlib.contextmanager
def quiet():  # TetGen prints its statistics to stdout
    sys.stdout.flush()
    fd = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        os.dup2(fd, 1)
        os.close(devnull)
        os.close(fd)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024.0))


def projections():
    rng = np.random.default_rng(7)
    n = 768
    rest = rng.normal(size=(n, 4, 3)).astype(np.float32)
    # keep the rest tets well conditioned (like mesh elements): reject slivers
    vol = np.abs(np.linalg.det(rest[:, 1:] - rest[:, :1]))
    rest[vol < 0.3] = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    qinv = np.empty((n, 9), np.float32)
    lib().pref_probe_qinv(n, np.ascontiguousarray(rest.reshape(n, 12)), qinv)
    defo = (rest + 0.25 * rng.normal(size=rest.shape)).astype(np.float32)
    defo[:64] = rest[:64]                                   # rest state
    defo[64:128] = rest[64:128] * 1.3                      # uniform stretch (repeated sigma)
    # inverted, with distinct singular values (a pure reflection has sigma1=sigma2=sigma3 and det<0, where the
    # axis the reference un-inverts is an artefact of Eigen's sweep order — SURVEY App. A)
    defo[128:192] = rest[128:192] * np.array([1.1, -0.9, 0.7], np.float32)
    defo[192:256, 3] = defo[192:256, 0] + 0.3 * (defo[192:256, 1] - defo[192:256, 0]) + 0.4 * (defo[192:256, 2] - defo[192:256, 0])  # flat
    defo[256:320] = rest[256:320] * 0.5                    # compressed
    strain = np.empty((n, 12), np.float32)
    volume = np.empty((n, 12), np.float32)
    lib().pref_probe_tet(n, np.ascontiguousarray(defo.reshape(n, 12)), qinv, 0.8, 1.0, strain)
    lib().pref_probe_volume(n, np.ascontiguousarray(defo.reshape(n, 12)), qinv, 1.0, 1.0, volume)
    volume2 = np.empty((n, 12), np.float32)
    lib().pref_probe_volume(n, np.ascontiguousarray(defo.reshape(n, 12)), qinv, 0.9, 1.1, volume2)
    save("projections", pos=defo.reshape(n, 12), qinv=qinv, strain=strain, volume=volume, volume_09_11=volume2,
         flat=np.arange(192, 256))


def ccd_and_ranges():
    rng = np.random.default_rng(11)
    m = 6000
    q = rng.normal(scale=0.5, size=(m, 18)).astype(np.float32)
    q[:, 9:] = q[:, :9] + 0.12 * rng.normal(size=(m, 9)).astype(np.float32)
    q[:500, 9:] = q[:500, :9]                              # static configurations
    q[500:520, :] = 0.0                                    # fully degenerate
    hit = np.empty(m, np.int32); t = np.empty(m, np.float32)
    lib().pref_probe_ccd(m, q, 0.1, hit, t)
    n = 3000
    base = rng.uniform(-40, 40, size=(n, 1, 3))
    tri = base + rng.normal(scale=1.0, size=(n, 3, 3))
    tri[16:24] = base[16:24] + rng.normal(scale=60.0, size=(8, 3, 3))   # longer than the 50-cell cap
    swept = tri + rng.normal(scale=0.3, size=tri.shape)
    tri[:16, :, 0] = np.round(base[:16, :, 0]); swept[:16, :, 0] = tri[:16, :, 0]  # in an integer plane (F6): length 0
    p = np.ascontiguousarray(tri.reshape(n, 9).astype(np.float32))
    o = np.ascontiguousarray(swept.reshape(n, 9).astype(np.float32))
    tm = np.empty((n, 3), np.int64); tl = np.empty((n, 3), np.uint32)
    lib().pref_probe_tri_range(n, p, o, tm, tl)
    pn = rng.uniform(-40, 40, size=(n, 3)).astype(np.float32)
    rad = rng.uniform(0.05, 1.0, n).astype(np.float32)
    rad[:8] = 200.0                                        # over the 50-cell cap
    nm = np.empty((n, 3), np.int64); nl = np.empty((n, 3), np.uint32)
    lib().pref_probe_node_range(n, pn, rad, 2.0, nm, nl)
    # edgeEdgeCCD (dead code in the reference's tick): edges (a,b) x (c,d) relative to a; far apart at the end
    # of the step (its static test fires below 0.5), crossing, parallel (the det == 0 branch) and degenerate ones
    me = 6000
    qe = rng.normal(scale=1.5, size=(me, 18)).astype(np.float32)
    qe[:, 9:] = qe[:, :9] + 0.6 * rng.normal(size=(me, 9)).astype(np.float32)
    qe[:1500, 9:] *= 0.2                                   # end state close together: static proximity hits
    # c-d exactly parallel to a-b at the end state (det == 0: the only way into the reference's interval branch), c a
    # quarter unit off the line so that overlapping ranges are static hits.  Exactly parallel at t = 1 also means coplanar
    # at t = 1, i.e. a root of the cubic exactly at the END of the interval: when the static test does not fire, hit or
    # miss hinges on whether the root finder lands on 1 - ulp or 1 + ulp, so those queries are excluded from the
    # comparison (edge_compare = 0); the ones the static test decides are kept.
    par = slice(1500, 2000)
    qe[par, 9:12] = np.array([2.0, 0.0, 0.0], np.float32)
    cx = rng.choice([-3.0, -1.0, 0.5, 1.0, 3.0], size=500).astype(np.float32)
    qe[par, 12:15] = np.stack([cx, np.full(500, 0.25, np.float32), np.full(500, 0.125, np.float32)], axis=1)
    qe[par, 15:18] = qe[par, 12:15] + np.array([1.0, 0.0, 0.0], np.float32) * rng.choice([-2.0, 0.5, 1.0, 4.0], size=(500, 1)).astype(np.float32)
    qe[2000:2020, :] = 0.0
    ehit = np.empty(me, np.int32); et = np.empty(me, np.float32)
    lib().pref_probe_edge_ccd(me, qe, ehit, et)
    compare = np.ones(me, np.int32)
    compare[1500:2000] = ((ehit[1500:2000] == 1) & (et[1500:2000] == 1.0)).astype(np.int32)
    # a hit at exactly t = 1 in that slice is either the static test or the boundary root; keep only those the static test
    # explains: the closest points of the two parallel segments are at most sqrt(0.25^2 + 0.125^2) = 0.28 < 0.5 apart
    # when their ranges along the line overlap, further than 0.5 otherwise
    lo = np.minimum(qe[par, 12], qe[par, 15]); hi = np.maximum(qe[par, 12], qe[par, 15])
    overlap = (hi >= 0.0) & (lo <= 2.0)
    compare[1500:2000] &= overlap.astype(np.int32)
    save("ccd", queries=q, hit=hit, t=t, edge_queries=qe, edge_hit=ehit, edge_t=et, edge_compare=compare)
    save("ranges", tri_pos=p, tri_prev=o, tri_min=tm, tri_len=tl, node_pos=pn, node_radius=rad, node_min=nm, node_len=nl)


FACTORIES = {
    "tetbox": lambda s: s.createTetBox((0.25, 3.0, -1.5), 1.0, (0.5, 0.0, -0.25), 1000.0, 2.0, False),
    "tetbox_hinged": lambda s: s.createTetBox((0.0, 1.0, 0.0), 0.5, (0, 0, 0), 500.0, 1.0, True),
    "box": lambda s: s.createBox((1.0, 2.0, 3.0), 0.5, 100.0),
    "sheet": lambda s: s.createSheet((0.0, 5.0, 0.0), 0.25, 0.5, 200.0),
    "bendsheet": lambda s: s.createBendSheet((0.0, 4.0, 0.0), 0.5, 50.0),
    "shapebox": lambda s: s.createShapeMatchingBox((0.0, 2.0, 0.0), 3, 4, 5, 1.0, (0, 0, 0), 800.0),
    "shapesheet": lambda s: s.createShapeMatchingSheet((0.0, 6.0, 0.0), 0.2, (0, 0, 0), 300.0),
}


def factories():
    """Topology + rest data every factory produces, then a short trajectory (PD, default options)."""
    out = {}
    for name, fn in FACTORIES.items():
        r = RefSolver()
        fn(r)
        rad, im = r.nodeScalars()
        out[name + "_pos"] = r.positions; out[name + "_vel"] = r.velocities
        out[name + "_radius"] = rad; out[name + "_invmass"] = im
        out[name + "_tris"] = r.getTriangles(); out[name + "_lines"] = r.getLines()
        out[name + "_counts"] = np.array([r.count(k) for k in ("position", "distance", "tet", "volume", "bend", "shape", "goal")])
        traj = []
        for _ in range(10):
            r.tick()
            traj.append(r.positions)
        out[name + "_traj"] = np.stack(traj)
        out[name + "_vel10"] = r.velocities
    save("factories", **out)


def two_box(s):
    s.createTetBox((0.1, 0.3, 0.1), 1.0, (0, 0, 0), 1000.0, 1.0, False)
    s.createTetBox((0.4, 2.6, 0.3), 1.0, (0, -5, 0), 1000.0, 1.0, False)


def collisions():
    """SURVEY Appendix B scene: states fed to detection + the lists/occupancy the reference finds."""
    r = RefSolver(iterations=10)
    two_box(r)
    out = {}
    counts = []
    for t in range(40):
        # state entering tick t: detection inside the tick sees pos + h*v against prev
        if t in (3, 10, 14, 30):
            pos, prev, vel = r.positions, r.prevPositions, r.velocities
        r.tick()
        counts.append((r.count("tri_collision"), r.count("static_collision")))
        if t in (3, 10, 14, 30):
            # reproduce the detection input: positions after the inertia step (Solver.cpp:229-238)
            h = np.float32(0.012)
            out["t%d_pos" % t] = (pos + h * vel).astype(np.float32)
            out["t%d_prev" % t] = prev
            out["t%d_tri" % t] = r.triCollisions()
            out["t%d_floor" % t] = r.staticCollisions()
        if t in (0, 9, 39):
            out["traj%d_pos" % (t + 1)] = r.positions
            out["traj%d_vel" % (t + 1)] = r.velocities
    out["counts"] = np.asarray(counts)
    # occupancy of the triangle hash for the final state
    r2 = RefSolver(iterations=10)
    two_box(r2)
    r2.tick(12)
    cells, cnts, members = r2.triOccupancy()
    out["occ_pos"] = r2.positions; out["occ_prev"] = r2.prevPositions
    out["occ_cells"] = cells; out["occ_counts"] = cnts; out["occ_members"] = members
    save("collisions", **out)


def tetgen_cube():
    """S1 at reduced size (config 1): TetGen cube falling on the floor, PD, strain + volume."""
    with quiet():
        r = RefSolver()
        pts, tets, faces = scenes.add_tetgen_cube(r, n=6, origin=(0.0, 0.4, 0.0))
    out = dict(points=pts, tets=tets, faces=faces)
    traj = {}
    for t in range(1, 61):
        r.tick()
        if t in (1, 10, 30, 60):
            traj["pos%d" % t] = r.positions; traj["vel%d" % t] = r.velocities
            traj["ncoll%d" % t] = np.array([r.count("tri_collision"), r.count("static_collision")])
    out.update(traj)
    save("tetgen_cube", **out)


def stack():
    """Reduced S3 (config 3 at 2x2 columns x 4 layers = 16 bodies), iterations=10: positions at K = 1, 10, 40, 44
    (free fall, floor contact, first body-body contacts), 50 and 100 (after the impacts, chaotic: a 1e-6
    perturbation of the reference at tick 40 changes its own tick-48 positions by 1.9e-2 and tick-50 by 8e-2)."""
    r = RefSolver(**scenes.S3_OPTIONS)
    scenes.build_s3(r, bodies=16, nx=2, nz=2)
    out = {}
    counts = []
    for t in range(1, 101):
        r.tick()
        counts.append((r.count("tri_collision"), r.count("static_collision")))
        if t in (1, 10, 40, 44, 50, 100):
            out["pos%d" % t] = r.positions; out["vel%d" % t] = r.velocities
            out["ncoll%d" % t] = np.array([r.count("tri_collision"), r.count("static_collision")])
    out["counts"] = np.asarray(counts)   # per tick: a threshold contact caught by only one side shows up here first
    save("stack16", **out)


def clusters():
    """Shape matching + goal matching (config 4 ingredients): regions driven by a scripted transform."""
    r = RefSolver(iterations=6)
    for k in range(3):
        r.createShapeMatchingBox((3.0 * k, 1.0 + 0.1 * k, 0.0), 3, 3, 3, 1.0, (0, 0, 0), 1000.0)
    region = np.eye(4, dtype=np.float32)
    region[3, :3] = (0.5, 1.5, 0.5)          # column-major: translation in the 4th column
    r.addFixedRegions(region.reshape(1, 16), 1000.0)
    lr = np.eye(4, dtype=np.float32)
    lr[3, :3] = (3.5, 1.6, 0.5)
    r.addLinkedRegions(lr.reshape(1, 16), 500.0)
    out = {"goal_ids": r.goal(0)[0], "shape_count": np.array([r.count("shape")])}
    xforms = []
    for t in range(1, 21):
        m = region.copy()
        m[3, :3] += (0.02 * t, 0.01 * t, 0.0)
        xforms.append(m.reshape(16))
        r.updateFixedRegions(m.reshape(1, 16))
        r.tick()
        if t in (1, 10, 20):
            out["pos%d" % t] = r.positions; out["vel%d" % t] = r.velocities
    out["xforms"] = np.stack(xforms)
    out["region"] = region.reshape(16); out["linked"] = lr.reshape(16)
    save("clusters", **out)


def pbd():
    """PBD path (config 2 ingredients): distance boxes colliding (node-node hash), a pinned 2 000-node rope coiling
    on the floor, cloth sheets with bend constraints; plus the node-hash occupancy on a reference state."""
    out = {}
    r = RefSolver(**scenes.S2_OPTIONS)
    scenes.build_pbd_boxes(r)
    for t in range(1, 61):
        r.tick()
        if t in (1, 10, 25, 40, 60):
            out["boxes_pos%d" % t] = r.positions; out["boxes_vel%d" % t] = r.velocities
        if t == 40:
            cells, cnts, members = r.nodeOccupancy()
            out["boxes_occ_cells"] = cells; out["boxes_occ_counts"] = cnts; out["boxes_occ_members"] = members
    r = RefSolver(**scenes.S2_OPTIONS)
    scenes.build_rope(r, n=2000, helix_radius=2.0)
    for t in range(1, 21):      # the reference blows up once the hanging chain starts to land (tick ~25): stop before
        r.tick()
        if t in (1, 10, 20):
            out["rope_pos%d" % t] = r.positions; out["rope_vel%d" % t] = r.velocities
    r = RefSolver(**scenes.S2_OPTIONS)
    scenes.build_rope(r, n=1500, shape="spiral", pinned=False)
    for t in range(1, 9):       # overlapping arms: collisions from tick 1; the reference's own iteration diverges by tick ~15
        r.tick()
        if t in (1, 2, 3, 5, 8):
            out["spiral_pos%d" % t] = r.positions; out["spiral_vel%d" % t] = r.velocities
    r = RefSolver(**scenes.S2_OPTIONS)
    r.createSheet((0.0, 2.0, 0.0), 0.5, 1.0, 0.8)
    r.createBendSheet((12.0, 2.0, 0.0), 0.5, 0.6)
    for t in range(1, 11):      # NaN in the reference before tick 20
        r.tick()
        if t in (1, 5, 10):
            out["sheets_pos%d" % t] = r.positions; out["sheets_vel%d" % t] = r.velocities
    r = RefSolver(**scenes.S2_OPTIONS)
    r.createTetBox((0.0, 1.0, 0.0), 0.5, (0, 0, 0), 0.5, 1.0, True)   # hinged: position constraints; tets ignored? no: F5
    out["hinged_counts"] = np.array([r.count(k) for k in ("position", "distance", "tet", "bend")])
    save("pbd", **out)


def s1_full():
    """S1 at full size (config 1): the side-8 cube with a 20 x 20-quad surface per face (~10 k tets after the reference's
    TetGen run), strain + volume w 1000, default options, dropped from y0 = 3.07 onto the floor (first floor contacts
    around tick 64).  The mesh itself is part of the fixture, so the GPU side needs no TetGen and no oracle."""
    with quiet():
        r = RefSolver()
        pts, tets, faces = scenes.add_tetgen_cube(r, n=20, origin=(0.0, 3.07, 0.0))
    out = dict(points=pts, tets=tets, faces=faces)
    for t in range(1, 101):
        r.tick()
        if t in (1, 10, 60, 70, 80, 100):
            out["pos%d" % t] = r.positions; out["vel%d" % t] = r.velocities
            out["ncoll%d" % t] = np.array([r.count("tri_collision"), r.count("static_collision")])
    save("s1_full", **out)


def s5_pair():
    """Two S5 bodies at full resolution (config 5: the side-8 cube with a 24 x 24-quad surface, 16 437 tets / 4 508
    nodes each): one resting just above the floor, the second dropped onto it from half a unit above, offset in x and z,
    so the run covers floor contact, body-body point-triangle contacts and a 9 k-node island."""
    with quiet():
        r = RefSolver()
        pts, tets, faces = scenes.add_tetgen_cube(r, n=24, origin=(0.0, 0.07, 0.0))
        pts2, tets2, faces2 = scenes.add_tetgen_cube(r, n=24, origin=(1.3, 8.57, 0.9))
    # TetGen runs once per body: the two meshes need not be translates of each other, so both are kept
    out = dict(points=pts, tets=tets, faces=faces, points2=pts2, tets2=tets2, faces2=faces2)
    for t in range(1, 61):
        r.tick()
        if t in (1, 10, 30, 40, 50, 60):
            out["pos%d" % t] = r.positions; out["vel%d" % t] = r.velocities
            out["ncoll%d" % t] = np.array([r.count("tri_collision"), r.count("static_collision")])
    save("s5_pair", **out)


if __name__ == "__main__":
    import sys
    lib().pref_srand(1)
    if len(sys.argv) > 1:          # regenerate selected fixtures only, e.g. `make_golden.py stack`
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    projections()
    ccd_and_ranges()
    factories()
    collisions()
    tetgen_cube()
    stack()
    clusters()
    pbd()
    s1_full()
    s5_pair()
