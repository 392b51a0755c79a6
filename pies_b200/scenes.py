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
