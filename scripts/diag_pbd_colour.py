"""Colour-batched PBD node-node response: per-tick wall time and the executor's own debug lines (PIES_DEBUG_PBD=1)."""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np
import pies_b200 as pb
from pies_b200 import scenes
n = int(os.environ.get("SIZE", "20000")); ticks = int(os.environ.get("TICKS", "3"))
s = pb.Solver(**scenes.S2_OPTIONS)
scenes.build_rope(s, n=n, shape=os.environ.get("SHAPE", "spiral"), pinned=False)
s.setTuning(pbdColourBatches=True)
for t in range(ticks):
    t0 = time.time(); s.tick(); dt = time.time() - t0
    st = s.stats()
    print("n=%d tick %d wall %.1f ms dev %.1f ms launches %d visits %d failed %s" % (n, t + 1, 1e3 * dt, st.msTick, st.kernelLaunchesLastTick, st.collisionProjections, s.simFailed), flush=True)
p = s.positions
print("finite", bool(np.isfinite(p).all()), "link length median %.4f max %.4f" % (np.median(np.linalg.norm(p[1:] - p[:-1], axis=1)), np.linalg.norm(p[1:] - p[:-1], axis=1).max()))
