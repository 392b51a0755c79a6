"""ctypes binding of oracle/_ref/libpies_ref.so — TEST INFRASTRUCTURE ONLY.

The library is the unmodified reference (nithinp7/Pies @ 2e552ea) compiled by
oracle/Makefile plus the white-box driver oracle/ref_driver.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (pies_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpies_ref.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


class RefOptions(C.Structure):
    """Field-for-field mirror of Pies::SolverOptions (Include/Pies/Solver.h:23-38)."""
    _fields_ = [
        ("fixedTimestepSize", C.c_float), ("timeSubsteps", C.c_uint32),
        ("iterations", C.c_uint32), ("collisionStabilizationIterations", C.c_uint32),
        ("collisionThresholdDistance", C.c_float), ("collisionThickness", C.c_float),
        ("gravity", C.c_float), ("damping", C.c_float), ("friction", C.c_float),
        ("staticFrictionThreshold", C.c_float), ("floorHeight", C.c_float),
        ("gridSpacing", C.c_float), ("threadCount", C.c_uint32), ("solver", C.c_uint32),
    ]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libpies_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        sig = {
            "pref_default_options": (None, [C.POINTER(RefOptions)]),
            "pref_create": (vp, [C.POINTER(RefOptions)]),
            "pref_destroy": (None, [vp]),
            "pref_srand": (None, [C.c_uint32]),
            "pref_create_tet_box": (None, [vp, _f32p, C.c_float, _f32p, C.c_float, C.c_float, C.c_int]),
            "pref_create_box": (None, [vp, _f32p, C.c_float, C.c_float]),
            "pref_create_sheet": (None, [vp, _f32p, C.c_float, C.c_float, C.c_float]),
            "pref_create_shape_matching_box": (None, [vp, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _f32p, C.c_float]),
            "pref_create_shape_matching_sheet": (None, [vp, _f32p, C.c_float, _f32p, C.c_float]),
            "pref_create_bend_sheet": (None, [vp, _f32p, C.c_float, C.c_float]),
            "pref_add_nodes": (None, [vp, C.c_uint32, _f32p]),
            "pref_add_tri_mesh_volume": (None, [vp, C.c_uint32, _f32p, C.c_uint32, _u32p, _f32p] + [C.c_float] * 7),
            "pref_add_fixed_regions": (None, [vp, C.c_uint32, _f32p, C.c_float]),
            "pref_update_fixed_regions": (None, [vp, C.c_uint32, _f32p]),
            "pref_add_linked_regions": (None, [vp, C.c_uint32, _f32p, C.c_float]),
            "pref_clear": (None, [vp]),
            "pref_set_release_hinge": (None, [vp, C.c_int]),
            "pref_reserve": (None, [vp] + [C.c_size_t] * 5),
            "pref_append_node": (C.c_uint32, [vp, _f32p, _f32p, C.c_float, C.c_float]),
            "pref_append_distance": (None, [vp, C.c_uint32, C.c_uint32, C.c_float]),
            "pref_append_position": (None, [vp, C.c_uint32, C.c_float]),
            "pref_append_tet": (None, [vp, _u32p, C.c_float, C.c_float, C.c_float]),
            "pref_append_volume": (None, [vp, _u32p, C.c_float, C.c_float, C.c_float]),
            "pref_append_bend": (None, [vp, _u32p, C.c_float]),
            "pref_append_triangle": (None, [vp, C.c_uint32, C.c_uint32, C.c_uint32]),
            "pref_append_shape": (None, [vp, C.c_uint32, _u32p, C.c_float]),
            "pref_tick": (None, [vp, C.c_uint32]),
            "pref_sim_failed": (C.c_int, [vp]),
            "pref_detect": (None, [vp]),
            "pref_get_node_vec": (None, [vp, C.c_int, _f32p]),
            "pref_set_node_vec": (None, [vp, C.c_int, _f32p]),
            "pref_get_node_scalars": (None, [vp, _f32p, _f32p]),
            "pref_get_vertices": (None, [vp, _f32p]),
            "pref_get_triangles": (None, [vp, _u32p]),
            "pref_get_lines": (None, [vp, _u32p]),
            "pref_get_tets": (None, [vp, _u32p, _f32p, _f32p, _f32p, _f32p]),
            "pref_get_volumes": (None, [vp, _u32p, _f32p, _f32p, _f32p, _f32p]),
            "pref_get_distances": (None, [vp, _u32p, _f32p, _f32p]),
            "pref_get_positions_c": (None, [vp, _u32p, _f32p, _f32p]),
            "pref_get_bends": (None, [vp, _u32p, _f32p, _f32p]),
            "pref_shape_size": (C.c_uint32, [vp, C.c_uint32]),
            "pref_get_shape": (None, [vp, C.c_uint32, _u32p, _f64p, _f64p, C.POINTER(C.c_float)]),
            "pref_goal_size": (C.c_uint32, [vp, C.c_uint32]),
            "pref_get_goal": (None, [vp, C.c_uint32, _u32p, _f32p, C.POINTER(C.c_float)]),
            "pref_get_tri_collisions": (None, [vp, _u32p]),
            "pref_get_static_collisions": (None, [vp, _u32p]),
            "pref_stiffness_nnz": (C.c_int64, [vp]),
            "pref_system_nnz": (C.c_int64, [vp]),
            "pref_get_system": (None, [vp, _i32p, _i32p, _f32p, _f32p, _f32p]),
            "pref_probe_tet": (None, [C.c_uint32, _f32p, _f32p, C.c_float, C.c_float, _f32p]),
            "pref_probe_volume": (None, [C.c_uint32, _f32p, _f32p, C.c_float, C.c_float, _f32p]),
            "pref_probe_qinv": (None, [C.c_uint32, _f32p, _f32p]),
            "pref_probe_bend": (None, [C.c_uint32, _f32p, _f32p, _f32p, _f32p]),
            "pref_probe_bend_angle": (None, [C.c_uint32, _f32p, _f32p]),
            "pref_probe_distance": (None, [C.c_uint32, _f32p, _f32p, _f32p]),
            "pref_probe_ccd": (None, [C.c_uint32, _f32p, C.c_float, _i32p, _f32p]),
            "pref_probe_edge_ccd": (None, [C.c_uint32, _f32p, _i32p, _f32p]),
            "pref_probe_node_range": (None, [C.c_uint32, _f32p, _f32p, C.c_float, _i64p, _u32p]),
            "pref_probe_tri_range": (None, [C.c_uint32, _f32p, _f32p, _i64p, _u32p]),
            "pref_tri_occupancy_build": (C.c_uint64, [vp, C.POINTER(C.c_uint64)]),
            "pref_node_occupancy_build": (C.c_uint64, [vp, C.POINTER(C.c_uint64)]),
            "pref_tri_occupancy_get": (None, [_i64p, _u32p, _u32p]),
        }
        for name in ("node", "triangle", "line_index", "tet", "volume", "distance", "position",
                     "bend", "shape", "goal", "tri_collision", "static_collision"):
            sig["pref_%s_count" % name] = (C.c_uint32, [vp])
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _v3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(3))


class RefSolver:
    """The reference Pies::Solver, driven white-box.  Method names follow Solver.h."""

    def __init__(self, **opts):
        L = lib()
        o = RefOptions()
        L.pref_default_options(C.byref(o))
        for k, v in opts.items():
            if k == "solver":
                v = {"PBD": 0, "PD": 1}.get(v, v)
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, v)
        self.options = o
        self.h = C.c_void_p(L.pref_create(C.byref(o)))

    def close(self):
        if self.h:
            lib().pref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- factories (public reference API) --
    def createTetBox(self, t, scale, v0, w, mass, hinged=False):
        lib().pref_create_tet_box(self.h, _v3(t), scale, _v3(v0), w, mass, int(hinged))

    def createBox(self, t, scale, w):
        lib().pref_create_box(self.h, _v3(t), scale, w)

    def createSheet(self, t, scale, mass, k):
        lib().pref_create_sheet(self.h, _v3(t), scale, mass, k)

    def createShapeMatchingBox(self, t, cx, cy, cz, scale, v0, w):
        lib().pref_create_shape_matching_box(self.h, _v3(t), cx, cy, cz, scale, _v3(v0), w)

    def createShapeMatchingSheet(self, t, scale, v0, w):
        lib().pref_create_shape_matching_sheet(self.h, _v3(t), scale, _v3(v0), w)

    def createBendSheet(self, t, scale, w):
        lib().pref_create_bend_sheet(self.h, _v3(t), scale, w)

    def addNodes(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        lib().pref_add_nodes(self.h, len(xyz), xyz)

    def addTriMeshVolume(self, verts, idx, v0, density, strainStiffness, minStrain, maxStrain,
                         volumeStiffness, compression, stretching):
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
        lib().pref_add_tri_mesh_volume(self.h, len(verts), verts, len(idx), idx, _v3(v0), density,
                                       strainStiffness, minStrain, maxStrain, volumeStiffness,
                                       compression, stretching)

    def addFixedRegions(self, mats, w):
        m = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 16)
        lib().pref_add_fixed_regions(self.h, len(m), m, w)

    def updateFixedRegions(self, mats):
        m = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 16)
        lib().pref_update_fixed_regions(self.h, len(m), m)

    def addLinkedRegions(self, mats, w):
        m = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 16)
        lib().pref_add_linked_regions(self.h, len(m), m, w)

    def clear(self):
        lib().pref_clear(self.h)

    # -- white-box builders --
    def reserve(self, nodes=0, dist=0, tets=0, vols=0, tris=0):
        lib().pref_reserve(self.h, nodes, dist, tets, vols, tris)

    def appendNode(self, pos, vel=(0, 0, 0), radius=0.1, invMass=1.0):
        return lib().pref_append_node(self.h, _v3(pos), _v3(vel), radius, invMass)

    def appendDistance(self, a, b, w):
        lib().pref_append_distance(self.h, a, b, w)

    def appendPosition(self, a, w):
        lib().pref_append_position(self.h, a, w)

    def appendTet(self, ids, w, minStrain=0.8, maxStrain=1.0):
        lib().pref_append_tet(self.h, np.asarray(ids, dtype=np.uint32), w, minStrain, maxStrain)

    def appendVolume(self, ids, w, compression=1.0, stretching=1.0):
        lib().pref_append_volume(self.h, np.asarray(ids, dtype=np.uint32), w, compression, stretching)

    def appendBend(self, ids, w):
        lib().pref_append_bend(self.h, np.asarray(ids, dtype=np.uint32), w)

    def appendTriangle(self, a, b, c):
        lib().pref_append_triangle(self.h, a, b, c)

    def appendShape(self, ids, w):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        lib().pref_append_shape(self.h, len(ids), ids, w)

    # -- stepping / state --
    def tick(self, n=1):
        lib().pref_tick(self.h, n)

    def detect(self):
        lib().pref_detect(self.h)

    @property
    def simFailed(self):
        return bool(lib().pref_sim_failed(self.h))

    def count(self, what):
        return getattr(lib(), "pref_%s_count" % what)(self.h)

    def _vec(self, which):
        out = np.empty((self.count("node"), 3), dtype=np.float32)
        lib().pref_get_node_vec(self.h, which, out)
        return out

    positions = property(lambda self: self._vec(0))
    prevPositions = property(lambda self: self._vec(1))
    velocities = property(lambda self: self._vec(2))

    def setState(self, pos=None, prev=None, vel=None):
        for which, a in ((0, pos), (1, prev), (2, vel)):
            if a is not None:
                lib().pref_set_node_vec(self.h, which, np.ascontiguousarray(a, dtype=np.float32))

    def nodeScalars(self):
        n = self.count("node")
        r = np.empty(n, np.float32); m = np.empty(n, np.float32)
        lib().pref_get_node_scalars(self.h, r, m)
        return r, m

    def getVertices(self):
        out = np.empty((self.count("node"), 3), dtype=np.float32)
        lib().pref_get_vertices(self.h, out)
        return out

    def getTriangles(self):
        out = np.empty((self.count("triangle"), 3), dtype=np.uint32)
        lib().pref_get_triangles(self.h, out)
        return out

    def getLines(self):
        out = np.empty(self.count("line_index"), dtype=np.uint32)
        lib().pref_get_lines(self.h, out)
        return out

    def tets(self):
        n = self.count("tet")
        ids = np.empty((n, 4), np.uint32); q = np.empty((n, 9), np.float32)
        w = np.empty(n, np.float32); a = np.empty(n, np.float32); b = np.empty(n, np.float32)
        lib().pref_get_tets(self.h, ids, q, w, a, b)
        return ids, q, w, a, b

    def volumes(self):
        n = self.count("volume")
        ids = np.empty((n, 4), np.uint32); q = np.empty((n, 9), np.float32)
        w = np.empty(n, np.float32); a = np.empty(n, np.float32); b = np.empty(n, np.float32)
        lib().pref_get_volumes(self.h, ids, q, w, a, b)
        return ids, q, w, a, b

    def distances(self):
        n = self.count("distance")
        ids = np.empty((n, 2), np.uint32); r = np.empty(n, np.float32); w = np.empty(n, np.float32)
        lib().pref_get_distances(self.h, ids, r, w)
        return ids, r, w

    def positionConstraints(self):
        n = self.count("position")
        ids = np.empty(n, np.uint32); t = np.empty((n, 3), np.float32); w = np.empty(n, np.float32)
        lib().pref_get_positions_c(self.h, ids, t, w)
        return ids, t, w

    def bends(self):
        n = self.count("bend")
        ids = np.empty((n, 4), np.uint32); a = np.empty(n, np.float32); w = np.empty(n, np.float32)
        lib().pref_get_bends(self.h, ids, a, w)
        return ids, a, w

    def shape(self, i):
        n = lib().pref_shape_size(self.h, i)
        ids = np.empty(n, np.uint32); mat = np.empty((n, 3), np.float64); q = np.empty(9, np.float64)
        w = C.c_float()
        lib().pref_get_shape(self.h, i, ids, mat, q, C.byref(w))
        return ids, mat, q, w.value

    def goal(self, i):
        n = lib().pref_goal_size(self.h, i)
        ids = np.empty(n, np.uint32); mat = np.empty((n, 3), np.float32)
        w = C.c_float()
        lib().pref_get_goal(self.h, i, ids, mat, C.byref(w))
        return ids, mat, w.value

    def system(self):
        """White-box: the system of the last global step, S + C_t as (rows, cols, vals) triplets (Solver.cpp:242-262),
        its last right-hand side and the solver's answer (both n x 3)."""
        nnz, n = lib().pref_system_nnz(self.h), self.count("node")
        rows = np.empty(nnz, np.int32); cols = np.empty(nnz, np.int32); vals = np.empty(nnz, np.float32)
        rhs = np.empty((n, 3), np.float32); state = np.empty((n, 3), np.float32)
        lib().pref_get_system(self.h, rows, cols, vals, rhs, state)
        return rows, cols, vals, rhs, state

    def triCollisions(self):
        out = np.empty((self.count("tri_collision"), 4), dtype=np.uint32)
        lib().pref_get_tri_collisions(self.h, out)
        return out

    def staticCollisions(self):
        out = np.empty(self.count("static_collision"), dtype=np.uint32)
        lib().pref_get_static_collisions(self.h, out)
        return out

    def triOccupancy(self):
        """Sorted (cells[n,3], counts[n], members[total]) of the triangle hash for the current state."""
        total = C.c_uint64()
        n = lib().pref_tri_occupancy_build(self.h, C.byref(total))
        cells = np.empty((n, 3), np.int64); counts = np.empty(n, np.uint32)
        members = np.empty(total.value, np.uint32)
        lib().pref_tri_occupancy_get(cells, counts, members)
        return cells, counts, members

    def nodeOccupancy(self):
        total = C.c_uint64()
        n = lib().pref_node_occupancy_build(self.h, C.byref(total))
        cells = np.empty((n, 3), np.int64); counts = np.empty(n, np.uint32)
        members = np.empty(total.value, np.uint32)
        lib().pref_tri_occupancy_get(cells, counts, members)
        return cells, counts, members
