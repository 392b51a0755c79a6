import sys, os
import numpy as np
sys.path.insert(0, ".")
import pies_b200 as pb
from pies_b200 import scenes
from oracle.refapi import RefSolver
bodies = int(os.environ.get("BODIES", "16")); nx = int(os.environ.get("NX", "2"))
r = RefSolver(**scenes.S3_OPTIONS); scenes.build_s3(r, bodies=bodies, nx=nx, nz=nx)
g = pb.Solver(**scenes.S3_OPTIONS); scenes.build_s3(g, bodies=bodies, nx=nx, nz=nx)
g.setTuning(pcgTolerance=float(os.environ.get("PIES_TOL", "1e-7")))
diag = float(np.linalg.norm(r.positions.max(0) - r.positions.min(0)))
for t in range(1, 121):
    r.tick(); g.tick()
    if t % 4 == 0 or t > 60:
        pr, pg = r.positions, g.positions
        e = np.abs(pr - pg).max(axis=1)
        st = g.stats()
        rt, gt = r.triCollisions(), g.triCollisions()
        same = len(rt) == len(gt) and bool((rt == gt).all())
        print("tick %3d err %.3e (1e-4 diag %.3e) worst node %d body %d | ref %d/%d gpu %d/%d lists_equal %s | pcg %d" % (
            t, e.max(), 1e-4 * diag, e.argmax(), e.argmax() // 27, len(rt), r.count("static_collision"), st.triCollisions, st.staticCollisions, same, st.pcgIterationsLastTick))
