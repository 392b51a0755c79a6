"""Island-local vs grid-wide global solve on the 16-body stack: per-tick difference and error against the golden reference."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import pies_b200 as pb
from pies_b200 import scenes
g = np.load("tests/golden/stack16.npz")
def make(**tune):
    s = pb.Solver(**scenes.S3_OPTIONS); scenes.build_s3(s, bodies=16, nx=2, nz=2)
    if tune: s.setTuning(**tune)
    return s
runs = {"grid": make(islandSolves=False), "isl": make(), "noWarp": make(islandTiersOff=1), "t2+": make(islandTiersOff=3), "t3": make(islandTiersOff=7, islandBigTier=True),
        "isl_tol1e-8": make(pcgTolerance=1e-8)}
for t in range(1, 47):
    P = {}
    for k, s in runs.items():
        s.tick(); P[k] = s.positions
    if t >= 38 or t in (1, 10):
        ref = g["pos%d" % t] if ("pos%d" % t) in g.files else None
        line = "t %2d" % t
        for k, s in runs.items():
            st = s.stats()
            line += " | %s d=%.2e" % (k, np.abs(P[k] - P["grid"]).max())
            if ref is not None: line += " e=%.2e" % np.abs(P[k] - ref).max()
            line += " pt %d it %d tiers %s gw %d" % (st.triCollisions, st.pcgIterationsLastTick, list(st.islandsTier), st.islandsGlobal)
        print(line, flush=True)
