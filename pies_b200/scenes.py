"""Seeded synthetic scenes of SURVEY.md §8(d), built through the host API.

Every builder takes a solver-like object exposing the reference's factory names
(pies_b200.Solver or the oracle's RefSolver), so both sides are constructed by the same code.
"""
import numpy as np


def lcg_jitter(n, scale=0.01, seed=12345):
    """n x 3 jitter in [0, scale) from a fixed 32-bit LCG (so no triangle sits in an integer plane)."""
    out = np.empty(3 * n, dtype=np.float64)
    x = seed & 0xFFFFFFFF
    for i in range(3 * n):
        x = (1664525 * x + 1013904223) & 0xFFFFFFFF
        out[i] = (x >> 8) / float(1 << 24)
    return (out.reshape(n, 3) * scale).astype(np.float32)


def s3_translations(bodies, nx=32, nz=32, pitch=3.0, y0=0.5, dy=2.5):
    """Box i sits at (pitch*(i%nx), y0 + dy*(i//(nx*nz)), pitch*((i//nx)%nz)) + jitter."""
    i = np.arange(bodies)
    t = np.stack([pitch * (i % nx), y0 + dy * (i // (nx * nz)), pitch * ((i // nx) % nz)], axis=1).astype(np.float32)
    return (t + lcg_jitter(bodies)).astype(np.float32)


def build_s3(s, bodies=20834, nx=32, nz=32, w=1000.0, mass=1.0):
    """S3: `bodies` x createTetBox(scale 1, w, mass, hinged=False) in stacked columns.
    Full size: 20 834 bodies = 562 518 nodes, 1 000 032 tets (2 000 064 tet-type constraints), 1 000 032 triangles."""
    if hasattr(s, "reserve"):  # oracle only: makes the factory's exact-size reserve() calls no-ops
        s.reserve(nodes=27 * bodies, tets=48 * bodies, vols=48 * bodies, tris=48 * bodies)
    for t in s3_translations(bodies, nx, nz):
        s.createTetBox(t, 1.0, (0.0, 0.0, 0.0), w, mass, False)
    return bodies


S3_OPTIONS = dict(iterations=10, timeSubsteps=1, solver="PD")


def s3_algorithmic_bytes(n_nodes, n_tet_constraints, n_pt, n_floor):
    """Local step + RHS assembly bytes per PD iteration, SURVEY §8(d): 40 B/node + 176 B per tet-type
    projection + 136 B per point-triangle + 48 B per floor contact."""
    return 40 * n_nodes + 176 * n_tet_constraints + 136 * n_pt + 48 * n_floor


def cube_surface(side=8.0, n=8, origin=(0.0, 3.07, 0.0)):
    """Closed triangulated surface of an axis-aligned cube, n x n quads per face, shared vertices
    (input of Solver::addTriMeshVolume for S1/S5)."""
    index = {}
    verts = []

    def vid(i, j, k):
        key = (i, j, k)
        if key not in index:
            index[key] = len(verts)
            verts.append((origin[0] + side * i / n, origin[1] + side * j / n, origin[2] + side * k / n))
        return index[key]

    tris = []
    for a in range(n):
        for b in range(n):
            for fixed in (0, n):
                quads = [((fixed, a, b), (fixed, a + 1, b), (fixed, a + 1, b + 1), (fixed, a, b + 1)),
                         ((a, fixed, b), (a + 1, fixed, b), (a + 1, fixed, b + 1), (a, fixed, b + 1)),
                         ((a, b, fixed), (a + 1, b, fixed), (a + 1, b + 1, fixed), (a, b + 1, fixed))]
                for q in quads:
                    p = [vid(*c) for c in q]
                    tris.append((p[0], p[1], p[2]))
                    tris.append((p[0], p[2], p[3]))
    return np.asarray(verts, np.float32), np.asarray(tris, np.uint32)


def add_tetgen_cube(ref, ours=None, side=8.0, n=8, origin=(0.0, 3.07, 0.0), density=1.0, strain_w=1000.0,
                    min_strain=0.8, max_strain=1.0, volume_w=1000.0, compression=1.0, stretching=1.0):
    """S1-style body: the reference tetrahedralises the cube with its vendored TetGen; the resulting mesh
    (points, tets, boundary faces) is handed to our side through add_tet_mesh_volume so both sides
    simulate the same mesh.  Returns (points, tets, boundary_tris) local to the body."""
    verts, tris = cube_surface(side, n, origin)
    n0, t0, v0, f0 = ref.count("node"), ref.count("tet"), ref.count("volume"), ref.count("triangle")
    ref.addTriMeshVolume(verts, tris.reshape(-1), (0, 0, 0), density, strain_w, min_strain, max_strain, volume_w,
                         compression, stretching)
    points = ref.positions[n0:]
    tets = (ref.tets()[0][t0:] if strain_w != 0 else ref.volumes()[0][v0:]) - n0
    # our API expects the boundary faces as the reference stores them (already re-wound)
    faces = ref.getTriangles()[f0:] - n0
    if ours is not None:
        ours.addTetMeshVolume(points, tets, faces, (0, 0, 0), density, strain_w, min_strain, max_strain, volume_w,
                              compression, stretching)
    return points, tets.astype(np.uint32), faces.astype(np.uint32)


# ---- PBD scenes (config 2 and its small relatives) ------------------------------------------------------
S2_OPTIONS = dict(iterations=4, timeSubsteps=1, solver="PBD", gridSpacing=2.0)


def rope_points(n, radius=0.25, helix_radius=4.0, rise_per_turn=1.0, y0=1.0):
    """n points spaced 2*radius apart (arc length) on a helix: consecutive nodes touch exactly, turns are
    `rise_per_turn` apart so nothing overlaps at rest."""
    spacing = 2.0 * radius
    turn_len = np.sqrt((2 * np.pi * helix_radius) ** 2 + rise_per_turn ** 2)
    s = spacing * np.arange(n, dtype=np.float64)
    ang = 2 * np.pi * s / turn_len
    return np.stack([helix_radius * np.cos(ang), y0 + rise_per_turn * s / turn_len, helix_radius * np.sin(ang)],
                    axis=1).astype(np.float32)


def spiral_points(n, radius=0.25, arm_gap=0.45, y0=1.0):
    """n points spaced 2*radius apart on a flat Archimedean spiral whose arms are `arm_gap` apart: with
    arm_gap < 2*radius neighbouring arms overlap, so node-node collisions are active from the first tick.
    (The reference's PBD blows up when a hanging chain lands on the floor - its distance projection moves only
    node 0 of a link, Constraints.cpp:11-37 - while a flat coil that lands all at once stays bounded.)"""
    b = arm_gap / (2 * np.pi)
    th = np.empty(n)
    t = 2 * np.pi * 2.0          # start two turns out so the first turn is not degenerate
    for i in range(n):
        th[i] = t
        t += 2.0 * radius / (b * np.sqrt(1.0 + t * t))
    r = b * th
    return np.stack([r * np.cos(th), np.full(n, y0), r * np.sin(th)], axis=1).astype(np.float32)


def build_rope(s, n=100000, radius=0.25, w=1.0, pinned=True, shape="helix", **kw):
    """S2: n-node distance-constraint chain, invMass 1, links created even-then-odd (so the reference's
    sequential sweep is a 2-colour sweep), first node position-constrained."""
    pts = rope_points(n, radius, **kw) if shape == "helix" else spiral_points(n, radius, **kw)
    links = np.concatenate([np.arange(0, n - 1, 2), np.arange(1, n - 1, 2)]).astype(np.uint32)
    if hasattr(s, "appendNodes"):       # pies_b200.Solver (bulk)
        s.appendNodes(pts, None, radius, 1.0)
        s.appendDistanceConstraints(np.stack([links, links + 1], axis=1), w)
        if pinned:
            s.appendPositionConstraints(np.array([0], np.uint32), 1.0)
    else:                               # oracle RefSolver (white-box, one call per element)
        if hasattr(s, "reserve"):
            s.reserve(nodes=n, dist=n)
        for p in pts:
            s.appendNode(p, (0, 0, 0), radius, 1.0)
        for a in links:
            s.appendDistance(int(a), int(a) + 1, w)
        if pinned:
            s.appendPosition(0, 1.0)
    return n


def build_pbd_boxes(s):
    """Distance-constraint boxes (createBox): one resting on the floor, one dropped onto it, one falling beside them."""
    s.createBox((0.0, 0.6, 0.0), 1.0, 0.5)
    s.createBox((0.3, 5.9, 0.2), 1.0, 0.5)
    s.createBox((9.0, 3.0, 1.0), 1.0, 0.5)


# ---- config 4 / config 5, any size ------------------------------------------
def lattice_hull(cx, cy, cz, offset=0):
    """Outward-wound surface triangles of a createShapeMatchingBox body (node id = offset + (i*cy + j)*cz + k,
    PrimitiveUtilities.cpp:1003-1013).  The factory itself emits no triangles, so without
    these config 4 has no point-triangle CCD and no friction."""
    def nid(i, j, k):
        return offset + (i * cy + j) * cz + k
    tris = []
    def quad(a, b, c, d, flip):
        if flip:
            b, d = d, b
        tris.append((a, b, c)); tris.append((a, c, d))
    for j in range(cy - 1):
        for k in range(cz - 1):
            for i, flip in ((0, False), (cx - 1, True)):
                quad(nid(i, j, k), nid(i, j, k + 1), nid(i, j + 1, k + 1), nid(i, j + 1, k), flip)
    for i in range(cx - 1):
        for k in range(cz - 1):
            for j, flip in ((0, False), (cy - 1, True)):
                quad(nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j, k + 1), nid(i, j, k + 1), flip)
    for i in range(cx - 1):
        for j in range(cy - 1):
            for k, flip in ((0, False), (cz - 1, True)):
                quad(nid(i, j, k), nid(i, j + 1, k), nid(i + 1, j + 1, k), nid(i + 1, j, k), flip)
    return np.asarray(tris, np.uint32)


def s4_translations(bodies, per_side, pitch=5.0, y0=1.0):
    i = np.arange(bodies)
    t = np.stack([pitch * (i % per_side), y0 + pitch * (i // (per_side * per_side)), pitch * ((i // per_side) % per_side)],
                 axis=1).astype(np.float32)
    return (t + lcg_jitter(bodies, seed=4242)).astype(np.float32)


def build_s4(s, bodies=15625, per_side=25, cx=4, cy=8, cz=8, w=1000.0, pitch=5.0, y0=1.0, goal_bodies=244, goal_w=1000.0):
    """S4: `bodies` x createShapeMatchingBox(cx, cy, cz) (one cluster per body, factory scale 0.5, invMass 0.1) on a
    lattice of the given pitch, hull triangles appended on both sides, and one unit-cube-derived goal region around
    each of the first `goal_bodies` bodies (addFixedRegions, w = goal_w).  Returns (translations, region matrices);
    the caller scripts updateFixedRegions."""
    trans = s4_translations(bodies, per_side, pitch, y0)
    n_per = cx * cy * cz
    hull = lattice_hull(cx, cy, cz)
    for b, t in enumerate(trans):
        s.createShapeMatchingBox(t, cx, cy, cz, 0.5, (0.0, 0.0, 0.0), w)
        tri = hull + np.uint32(b * n_per)
        if hasattr(s, "appendTriangles"):
            s.appendTriangles(tri)
        else:
            for a, bb, c in tri:
                s.appendTriangle(int(a), int(bb), int(c))
    ext = 0.5 * np.array([cx - 1, cy - 1, cz - 1], np.float32)
    regions = []
    for t in trans[:goal_bodies]:
        m = np.zeros((4, 4), np.float32)          # column-major storage: m[c, r]
        m[0, 0], m[1, 1], m[2, 2], m[3, 3] = ext[0] + 0.5, ext[1] + 0.5, ext[2] + 0.5, 1.0
        m[3, :3] = t + 0.5 * ext                  # unit cube centred on the body, 0.25 larger on every side
        regions.append(m.reshape(16))
    regions = np.stack(regions) if regions else np.zeros((0, 16), np.float32)
    if len(regions):
        s.addFixedRegions(regions, goal_w)
    return trans, regions


def s4_region_script(regions, tick, speed=0.02):
    """The scripted rigid motion of the goal regions: a slow translation along +x with a vertical sway."""
    out = regions.reshape(-1, 4, 4).copy()
    out[:, 3, 0] += speed * tick
    out[:, 3, 1] += 0.5 * speed * np.sin(0.2 * tick)
    return out.reshape(-1, 16)


def build_s5(ref, ours=None, bodies=512, per_side=8, n=24, side=8.0, pitch=10.0, y0=3.07):
    """S5: `bodies` x addTriMeshVolume of the side-8 cube with an n x n-quad surface per face (n = 24: 16 437 tets each),
    strain w 1000 [0.8, 1] + volume w 1000 [1, 1], on a lattice of the given pitch above the floor.  The reference
    tetrahedralises (its vendored TetGen); our side receives the same meshes."""
    i = np.arange(bodies)
    origins = np.stack([pitch * (i % per_side), y0 + pitch * (i // (per_side * per_side)), pitch * ((i // per_side) % per_side)],
                       axis=1).astype(np.float32) + lcg_jitter(bodies, seed=5151)
    for o in origins:
        add_tetgen_cube(ref, ours, side=side, n=n, origin=tuple(float(x) for x in o))
    return origins


def cube24_mesh():
    """The config-5 body as the reference's TetGen run leaves it (side-8 cube, 24 x 24 quads per face: 4 518 nodes,
    16 546 tets, 6 912 boundary triangles), committed under pies_b200/data/ so that large S5 scenes can be built without
    TetGen (generated by tests/golden/make_golden.py::s5_pair: first body, origin moved to 0)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "cube24.npz"))
    return g["points"], g["tets"], g["faces"]


def s5_origins(bodies=512, per_side=8, pitch=10.0, y0=3.07):
    i = np.arange(bodies)
    o = np.stack([pitch * (i % per_side), y0 + pitch * (i // (per_side * per_side)), pitch * ((i // per_side) % per_side)],
                 axis=1).astype(np.float32)
    return (o + lcg_jitter(bodies, seed=5151)).astype(np.float32)


def build_s5_replicated(s, bodies=512, per_side=8, pitch=10.0, y0=3.07):
    """S5 from ONE committed TetGen mesh replicated on the lattice of SURVEY section 8(d) (the reference re-runs TetGen per
    body; its meshes differ by a handful of tets from body to body, which no throughput figure depends on)."""
    pts, tets, faces = cube24_mesh()
    for o in s5_origins(bodies, per_side, pitch, y0):
        s.addTetMeshVolume(pts + o, tets, faces, (0, 0, 0), 1.0, 1000.0, 0.8, 1.0, 1000.0, 1.0, 1.0)
    return bodies
