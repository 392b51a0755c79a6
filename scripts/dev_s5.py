"""Phase timeline of a (reduced) S5 scene: TetGen cube bodies dropped onto the floor and each other."""
import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import pies_b200 as pb
from pies_b200 import scenes
bodies = int(os.environ.get("BODIES", "64")); ticks = int(os.environ.get("TICKS", "100"))
t0 = time.time()
s = pb.Solver(**scenes.S3_OPTIONS)
scenes.build_s5_replicated(s, bodies, per_side=int(os.environ.get("PER_SIDE", "4")))
print("scene built in %.2fs, nodes %d" % (time.time() - t0, len(s.getVertices())), flush=True)
s.setTuning(profilePhases=True, islandBigTier=bool(int(os.environ.get("BIG_TIER", "0"))), islandSolves=not int(os.environ.get("NO_ISLANDS", "0")))
t0 = time.time(); s.tick(); print("first tick (incl. system build + upload) %.2fs" % (time.time() - t0), flush=True)
for t in range(1, ticks):
    s.tick()
    st = s.stats()
    if t % 10 == 0 or t < 3:
        print("tick %3d dev %.2fms | local %.2f global %.2f (islands %.2f) detect %.2f contact %.2f | pcg %d cap %d | pt %d floor %d | launches %d | islands %s grid-wide %d (%d nodes)" % (
            t, st.msTick, st.msLocal, st.msGlobal, st.msIslandKernels, st.msDetect, st.msContact, st.pcgIterationsLastTick, st.pcgCapHits,
            st.triCollisions, st.staticCollisions, st.kernelLaunchesLastTick, list(st.islandsTier), st.islandsGlobal, st.islandNodesGlobal), flush=True)
p = s.positions
print("finite:", bool(np.isfinite(p).all()), "y range", p[:, 1].min(), p[:, 1].max(), "failed", s.simFailed)
