"""Diagnosis: reduced config 4 (shape-matching bodies with hull triangles).  Feeds the reference's state entering the first
contact tick to the GPU detection pass and prints the contacts only one side finds, with the geometry of each."""
import sys
import numpy as np
sys.path.insert(0, ".")
import pies_b200 as pb
from oracle import refapi
from pies_b200 import scenes

kw = dict(bodies=8, per_side=2, cx=3, cy=4, cz=4, pitch=2.2, y0=0.3, goal_bodies=1)
r = refapi.RefSolver(iterations=6)
d = pb.Solver(iterations=6)
_, regions = scenes.build_s4(r, **kw)
scenes.build_s4(d, **kw)
d.tick()
H = np.float32(0.012)
for t in range(1, 41):
    r.updateFixedRegions(scenes.s4_region_script(regions, t))
    pos, prev, vel = r.positions, r.prevPositions, r.velocities
    r.tick()
    if not r.count("tri_collision"):
        continue
    q = (pos + H * vel).astype(np.float32)
    d.setState(q, prev, None)
    d.detect()
    ours, ref = d.triCollisions(), r.triCollisions()
    so = {}
    for e in map(tuple, ours): so[e] = so.get(e, 0) + 1
    sr = {}
    for e in map(tuple, ref): sr[e] = sr.get(e, 0) + 1
    print("tick", t, "ours", len(ours), "ref", len(ref), "unique ours", len(so), "unique ref", len(sr))
    for e in sorted(set(so) | set(sr)):
        if so.get(e, 0) != sr.get(e, 0):
            a, b, c, dd = e
            n0 = np.cross(prev[c] - prev[b], prev[dd] - prev[b]); n1 = np.cross(q[c] - q[b], q[dd] - q[b])
            n0 /= np.linalg.norm(n0); n1 /= np.linalg.norm(n1)
            d0 = float(np.dot(n0, prev[a] - prev[b])); d1 = float(np.dot(n1, q[a] - q[b]))
            ab, ac = q[c] - q[b], q[dd] - q[b]
            M = np.stack([ab, ac, n1], 1).astype(np.float64)
            bary = np.linalg.solve(M, (q[a] - q[b]).astype(np.float64))
            print("   ", e, "copies ours", so.get(e, 0), "ref", sr.get(e, 0), "| n0.ap0 %.7f n1.ap1 %.7f bary (%.7f, %.7f) sum %.7f" % (d0, d1, bary[0], bary[1], bary[0] + bary[1]))
            for nm, P in (("q", q), ("prev", prev)):
                T = np.stack([P[b], P[c], P[dd]])
                lo, hi = np.floor(np.minimum(q[[b, c, dd]].min(0), prev[[b, c, dd]].min(0))), np.ceil(np.maximum(q[[b, c, dd]].max(0), prev[[b, c, dd]].max(0)))
            print("        tri cells", lo, hi, "point q", q[a], "prev", prev[a])
    if t >= 39:
        break
