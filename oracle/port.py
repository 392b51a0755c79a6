"""oracle/port.py — numpy restatement of the closed-form pieces of the Pies solver loop.

TEST INFRASTRUCTURE ONLY.  Nothing under pies_b200/ may import this module (tests/test_abi.py greps for it);
it exists so that the CPU suite can check the committed golden fixtures (generated from the UNMODIFIED,
compiled reference by tests/golden/make_golden.py) against an independent statement of the same algorithm,
and so that GPU parity tests have a checker on boxes where oracle/_ref did not travel.

Parity pin: every function below is pinned against tests/golden/*.npz (outputs of the reference's own
functions) in tests/test_oracle_port.py.  Pieces of the path that are NOT restated here (JacobiSVD sweep
order on degenerate inputs, the cubic root finder of the CCD, SimplicialLLT) are validated only through the
compiled reference oracle/_ref.

Citations are reference paths under /root/reference.
"""
import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------------
# Local step projections (Src/Constraints.cpp)
# --------------------------------------------------------------------------------------------------
def distance_projection(pos, rest):
    """DistanceConstraintProjection::operator(), Src/Constraints.cpp:11-37.
    pos: (n, 2, 3) fp32, rest: (n,) -> projected (n, 2, 3).  Only node 0 moves (the b side is commented out in
    the reference), by -(rest - dist) * dir with dir = (1,0,0) when dist <= 1e-5."""
    pos = np.asarray(pos, F32).reshape(-1, 2, 3)
    rest = np.asarray(rest, F32)
    a, b = pos[:, 0], pos[:, 1]
    diff = (b - a).astype(F32)
    dist = np.sqrt((diff * diff).sum(axis=1, dtype=F32), dtype=F32)
    safe = np.where(dist > F32(1e-5), dist, F32(1.0))
    d = np.where((dist > F32(1e-5))[:, None], diff / safe[:, None], np.array([1, 0, 0], F32)[None, :]).astype(F32)
    disp = (rest - dist).astype(F32)
    out = pos.copy()
    out[:, 0] = a + (-disp)[:, None] * d
    return out


def _deformation_gradient(pos, qinv):
    """F = P * Qinv with P's columns the edge vectors x2-x1, x3-x1, x4-x1 (Constraints.cpp:85-91) and Qinv in
    glm column-major order (9 floats = 3 columns)."""
    pos = np.asarray(pos, np.float64).reshape(-1, 4, 3)
    P = np.stack([pos[:, 1] - pos[:, 0], pos[:, 2] - pos[:, 0], pos[:, 3] - pos[:, 0]], axis=2)  # columns
    Q = np.asarray(qinv, np.float64).reshape(-1, 3, 3).transpose(0, 2, 1)                       # column-major -> matrix
    return P @ Q


def _pack_projection(Fhat):
    """projected = (0, col0, col1, col2) of the corrected gradient (Constraints.cpp:113-127: the Eigen matrix is
    filled from the TRANSPOSE of F and the result transposed back, so the net effect is proj(F))."""
    n = Fhat.shape[0]
    out = np.zeros((n, 4, 3), F32)
    out[:, 1] = Fhat[:, :, 0]
    out[:, 2] = Fhat[:, :, 1]
    out[:, 3] = Fhat[:, :, 2]
    return out.reshape(n, 12)


def tet_strain_projection(pos, qinv, min_strain, max_strain):
    """TetrahedralConstraintProjection::operator(), Constraints.cpp:76-128: clamp the singular values of F to
    [min, max], negate the smallest one when det F < 0, rebuild U S V^T.  (U S V^T = sum s_i u_i v_i^T does not
    depend on the SVD's sign/ordering conventions while the clamped values follow their own singular vectors.)"""
    F = _deformation_gradient(pos, qinv)
    U, s, Vt = np.linalg.svd(F)            # descending, like Eigen's JacobiSVD
    sc = np.clip(s, min_strain, max_strain)
    neg = np.linalg.det(F) < 0.0
    sc[neg, 2] *= -1.0
    return _pack_projection((U * sc[:, None, :]) @ Vt)


def compute_d(sigma, omega_min, omega_max, iters=10):
    """computeD, Constraints.cpp:186-204: 10 fixed iterations of the volume correction, in fp32 like glm."""
    sigma = np.asarray(sigma, F32)
    D = np.zeros_like(sigma)
    for _ in range(iters):
        sp = (sigma + D).astype(F32)
        prod = (sp[:, 0] * sp[:, 1] * sp[:, 2]).astype(F32)
        omega = np.clip(prod, F32(omega_min), F32(omega_max))
        C = (prod - omega).astype(F32)
        g = np.stack([sp[:, 1] * sp[:, 2], sp[:, 0] * sp[:, 2], sp[:, 0] * sp[:, 1]], axis=1).astype(F32)
        gd = (g * D).sum(axis=1, dtype=F32)
        gg = (g * g).sum(axis=1, dtype=F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            D = ((gd - C)[:, None] * g / gg[:, None]).astype(F32)
    return D


def tet_volume_projection(pos, qinv, min_omega, max_omega):
    """VolumeConstraintProjection::operator(), Constraints.cpp:206-255: sigma += computeD(sigma), no sign fix."""
    F = _deformation_gradient(pos, qinv)
    U, s, Vt = np.linalg.svd(F)
    s2 = s + compute_d(s.astype(F32), min_omega, max_omega).astype(np.float64)
    return _pack_projection((U * s2[:, None, :]) @ Vt)


# --------------------------------------------------------------------------------------------------
# Spatial hash cell ranges (Src/Solver.cpp:877-979) — integer work, bit-exact
# --------------------------------------------------------------------------------------------------
def _capped(mins, lens, cap):
    bad = (lens > cap).any(axis=1)
    mins = mins.copy(); lens = lens.copy()
    mins[bad] = 0; lens[bad] = 0            # `return {};`
    return mins, lens


def tri_cell_range(pos, prev, cap=50):
    """Solver::TriCompRange::operator(), Solver.cpp:942-979.  pos/prev: (n, 3 corners, 3) fp32.  The range spans
    the swept AABB (positions and previous positions); min = floor(min), length = ceil(max) - min.  Note the
    reference does NOT divide by the grid scale here (x1..x3 are computed and unused), and a triangle lying in
    an integer plane gets length 0 on that axis.  Ranges longer than 50 cells are dropped."""
    pos = np.asarray(pos, F32).reshape(-1, 3, 3)
    prev = np.asarray(prev, F32).reshape(-1, 3, 3)
    both = np.concatenate([pos, prev], axis=1)
    lo = both.min(axis=1); hi = both.max(axis=1)
    mins = np.floor(lo).astype(np.int64)
    # static_cast<uint32_t>(ceil(max) - float(minX)): the subtraction is done in float
    lens = (np.ceil(hi) - mins.astype(F32)).astype(F32).astype(np.int64).astype(np.uint32)
    return _capped(mins, lens, cap)


def node_cell_range(pos, radius, scale, cap=50):
    """Solver::NodeCompRange::operator(), Solver.cpp:877-901: padded sphere AABB in grid units;
    length = ceil(fract(min) + 2 r)."""
    pos = np.asarray(pos, F32).reshape(-1, 3)
    r = ((np.asarray(radius, F32) + F32(0.5)) / F32(scale)).astype(F32)
    p = (pos / F32(scale)).astype(F32)
    lo = (p - r[:, None]).astype(F32)
    fl = np.floor(lo)
    mins = fl.astype(np.int64)
    fract = (lo - fl).astype(F32)
    lens = np.ceil((fract + (F32(2.0) * r)[:, None]).astype(F32)).astype(np.int64).astype(np.uint32)
    return _capped(mins, lens, cap)


def cell_occupancy(mins, lens):
    """SpatialHash::parallelBulkInsert, Include/Pies/SpatialHash.h:67-79,129-189: every element is appended to the
    bucket of each cell of its range.  Returns (cells sorted by (x,y,z), counts, members concatenated in
    ascending element index within a cell) — the canonical form the fixtures use (oracle/ref_driver.cpp)."""
    cells = {}
    for e in range(len(mins)):
        lx, ly, lz = (int(v) for v in lens[e])
        mx, my, mz = (int(v) for v in mins[e])
        for i in range(lx):
            for j in range(ly):
                for k in range(lz):
                    cells.setdefault((mx + i, my + j, mz + k), []).append(e)
    keys = sorted(cells)
    counts = np.array([len(cells[k]) for k in keys], np.uint32)
    members = np.array([m for k in keys for m in sorted(cells[k])], np.uint32)
    return np.array(keys, np.int64).reshape(-1, 3), counts, members


# --------------------------------------------------------------------------------------------------
# Node update streams (Src/Solver.cpp) — used by the closed-form free-fall check
# --------------------------------------------------------------------------------------------------
def free_fall(pos, vel, ticks, h=0.012, gravity=10.0, damping=0.006, floor=0.0):
    """A PD tick on a body whose constraints are at rest and that touches nothing reduces to
    Solver.cpp:229-238 (q = x + h v) and :386-395 (v = (1-d)(q - x)/h + h f invMass with f = -g m): the
    constrained minimiser is the inertial prediction itself.  fp32 like the reference."""
    pos = np.asarray(pos, F32).copy(); vel = np.asarray(vel, F32).copy()
    h = F32(h)
    for _ in range(ticks):
        q = (pos + h * vel).astype(F32)
        vel = (F32(1.0 - damping) * (q - pos) / h).astype(F32)
        vel[:, 1] = vel[:, 1] - h * F32(gravity)
        pos = q
    return pos, vel


# --------------------------------------------------------------------------------------------------
# Collision terms of the global system (Src/CollisionConstraint.cpp)
# --------------------------------------------------------------------------------------------------
PT_WEIGHT = F32(10000.0)      # PointTriangleCollisionConstraint::w, CollisionConstraint.h:32
FLOOR_WEIGHT = F32(10000.0)   # StaticCollisionConstraint::w, CollisionConstraint.h:78


def collision_matrix(n, tri_list, floor_list):
    """C_t, what the collision constraints of one substep add to the system matrix (Solver.cpp:242-262).
    tri_list: (m, 4) entries (a, b, c, d) of the point-triangle list WITH its duplicates (one copy per shared cell,
    SURVEY F7): each adds w * A^T A with A = [0; -1 1 0 0; -1 0 1 0; -1 0 0 1] (CollisionConstraint.cpp:74-83,176-184),
    i.e. w * [[3,-1,-1,-1],[-1,1,0,0],[-1,0,1,0],[-1,0,0,1]] on rows/columns (a, b, c, d).
    floor_list: node ids of the floor list with duplicates: each adds w on the diagonal (CollisionConstraint.cpp:442-445).
    Returns a dense (n, n) float64 matrix (sums of exactly representable multiples of 1e4)."""
    C = np.zeros((n, n), np.float64)
    ata = np.array([[3, -1, -1, -1], [-1, 1, 0, 0], [-1, 0, 1, 0], [-1, 0, 0, 1]], np.float64) * float(PT_WEIGHT)
    for e in np.asarray(tri_list, np.int64).reshape(-1, 4):
        C[np.ix_(e, e)] += ata
    for v in np.asarray(floor_list, np.int64).reshape(-1):
        C[v, v] += float(FLOOR_WEIGHT)
    return C


def collision_csr(n, tri_list, floor_list):
    """The same matrix in the streamable form the CUDA mat-vec reads (pies_b200/csrc/detect.cu, k_ccsr_fill): distinct
    contacts with weight = copies * w, off-diagonals per node in (contact, slot) order, diagonal summed separately.
    Returns (cPtr, cCol, cVal, cDiag)."""
    tri = np.asarray(tri_list, np.int64).reshape(-1, 4)
    uniq, copies = (np.unique(tri, axis=0, return_counts=True) if len(tri) else (tri, np.zeros(0, np.int64)))
    rows = [[] for _ in range(n)]
    diag = np.zeros(n, np.float64)
    for (a, b, c, d), k in zip(uniq, copies):
        w = float(PT_WEIGHT) * k
        rows[a] += [(b, -w), (c, -w), (d, -w)]; diag[a] += 3 * w
        for v in (b, c, d):
            rows[v].append((a, -w)); diag[v] += w
    for v in np.asarray(floor_list, np.int64).reshape(-1):
        diag[v] += float(FLOOR_WEIGHT)
    ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    col = np.array([c for r in rows for c, _ in r], np.int64)
    val = np.array([v for r in rows for _, v in r], np.float64)
    return ptr, col, val, diag
