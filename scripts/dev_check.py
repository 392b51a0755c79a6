"""Developer parity smoke on a GPU box: CUDA path vs oracle/_ref on small cases (not a test; prints)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import pies_b200 as pb
from oracle.refapi import RefSolver, lib as reflib

rng = np.random.default_rng(0)

def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))

# ---- probes ----
n = 2000
rest = rng.normal(size=(n, 4, 3)).astype(np.float32)
qinv = np.empty((n, 9), np.float32)
reflib().pref_probe_qinv(n, rest.reshape(n, 12).copy(), qinv)
defo = (rest + 0.3 * rng.normal(size=rest.shape)).astype(np.float32)
defo[:50] = rest[:50]
defo[50:100, 1] = defo[50:100, 0] + 1e-3 * (defo[50:100, 1] - defo[50:100, 0])  # nearly flat
defo[100:150] *= np.array([1, -1, 1], np.float32)  # inverted
ref = np.empty((n, 12), np.float32)
reflib().pref_probe_tet(n, defo.reshape(n, 12).copy(), qinv, 0.8, 1.0, ref)
got = pb.probe_tet_projection(defo, qinv, 0.8, 1.0)
err = np.abs(got - ref).max(axis=1)
print("tet strain proj: max abs err", err.max(), "median", np.median(err), "worst idx", err.argmax())
reflib().pref_probe_volume(n, defo.reshape(n, 12).copy(), qinv, 1.0, 1.0, ref)
got = pb.probe_volume_projection(defo, qinv, 1.0, 1.0)
err = np.abs(got - ref).max(axis=1)
print("tet volume proj: max abs err", err.max(), "median", np.median(err), "worst idx", err.argmax())

# ccd
m = 20000
q = rng.normal(scale=0.5, size=(m, 18)).astype(np.float32)
q[:, 9:] = q[:, :9] + 0.1 * rng.normal(size=(m, 9)).astype(np.float32)
hit_r = np.empty(m, np.int32); t_r = np.empty(m, np.float32)
reflib().pref_probe_ccd(m, q, 0.1, hit_r, t_r)
hit_g, t_g = pb.probe_ccd(q, 0.1)
print("ccd: ref hits", hit_r.sum(), "gpu hits", hit_g.sum(), "mismatch", int((hit_r != hit_g).sum()),
      "max |dt| on common", float(np.abs(t_r - t_g)[(hit_r == 1) & (hit_g == 1)].max(initial=0)))

# ranges
p = rng.uniform(-30, 30, size=(5000, 9)).astype(np.float32)
o = (p + rng.normal(scale=0.3, size=p.shape)).astype(np.float32)
p[:10] = np.round(p[:10])  # integer-plane quirk F6
o[:10] = p[:10]
mr = np.empty((5000, 3), np.int64); lr = np.empty((5000, 3), np.uint32)
reflib().pref_probe_tri_range(5000, p, o, mr, lr)
mg, lg = pb.probe_tri_range(p, o)
print("tri range mismatches:", int((mr != mg).any(axis=1).sum()), int((lr != lg).any(axis=1).sum()))
pn = rng.uniform(-30, 30, size=(5000, 3)).astype(np.float32); rad = rng.uniform(0.05, 1.0, 5000).astype(np.float32)
reflib().pref_probe_node_range(5000, pn, rad, 2.0, mr, lr)
mg, lg = pb.probe_node_range(pn, rad, 2.0)
print("node range mismatches:", int((mr != mg).any(axis=1).sum()), int((lr != lg).any(axis=1).sum()))

# sort
k = rng.integers(0, 1 << 40, size=300001, dtype=np.uint64); v = np.arange(len(k), dtype=np.uint32)
ks, vs = pb.probe_sort_pairs(k, v, 40)
order = np.argsort(k, kind="stable")
print("sort ok:", bool((ks == k[order]).all() and (vs == v[order]).all()))

# ---- scenes ----
def build_both(fn, **opts):
    reflib().pref_srand(1)
    r = RefSolver(**opts); fn(r)
    g = pb.Solver(**opts); fn(g)
    import os
    g.setTuning(pcgTolerance=float(os.environ.get("PIES_TOL", "1e-7")))
    return r, g

def compare_ticks(name, r, g, ticks, every=1):
    diag = float(np.linalg.norm(r.positions.max(0) - r.positions.min(0)))
    for t in range(ticks):
        r.tick(); g.tick()
        if (t + 1) % every == 0 or t == ticks - 1:
            pr, pg = r.positions, g.positions
            vr, vg = r.velocities, g.velocities
            st = g.stats()
            print("%s tick %3d: pos err %.3e (tol %.3e) vel err %.3e | ref coll %d/%d gpu %d/%d | pcg %d res %.2e" % (
                name, t + 1, np.abs(pr - pg).max(), 1e-4 * diag, np.abs(vr - vg).max(),
                r.count("tri_collision"), r.count("static_collision"), st.triCollisions, st.staticCollisions,
                st.pcgIterationsLastTick, st.pcgLastRelResidual))

def one_box(s):
    s.createTetBox((0, 3, 0), 1.0, (0, 0, 0), 1000.0, 1.0, False)
r, g = build_both(one_box)
print("scene tris equal:", bool((r.getTriangles() == g.getTriangles()).all()), "pos equal:", bool((r.positions == g.positions).all()))
compare_ticks("onebox", r, g, 100, every=10)

def two_box(s):
    s.createTetBox((0.1, 0.3, 0.1), 1.0, (0, 0, 0), 1000.0, 1.0, False)
    s.createTetBox((0.4, 2.6, 0.3), 1.0, (0, -5, 0), 1000.0, 1.0, False)
r, g = build_both(two_box, iterations=10)
compare_ticks("twobox", r, g, 100, every=5)
