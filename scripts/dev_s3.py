import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import pies_b200 as pb
from pies_b200 import scenes
bodies = int(os.environ.get("BODIES", "20834"))
ticks = int(os.environ.get("TICKS", "60"))
t0 = time.time()
s = pb.Solver(**scenes.S3_OPTIONS)
scenes.build_s3(s, bodies)
print("scene built in %.2fs, nodes %d" % (time.time() - t0, len(s.getVertices())), flush=True)
s.setTuning(profilePhases=True, dataflowSweepsOnly=bool(int(os.environ.get("DATAFLOW_ONLY", "0"))),
            islandSolves=not int(os.environ.get("NO_ISLANDS", "0")), islandTiersOff=int(os.environ.get("TIERS_OFF", "0")), islandBigTier=bool(int(os.environ.get("BIG_TIER", "0"))))
t0 = time.time(); s.tick(); print("first tick (incl. system build + upload) %.2fs" % (time.time() - t0), flush=True)
for t in range(1, ticks):
    t1 = time.time(); s.tick(); wall = time.time() - t1
    st = s.stats()
    if t % 5 == 0 or t < 3:
        print("tick %3d wall %.1fms dev %.2fms | local %.2f (tet %.2f/%d) global %.2f (islands %.2f) detect %.2f contact %.2f other %.2f | pcg %d res %.1e cap %d | pt %d floor %d | launches %d | clusters mid %d large %d | islands %s grid-wide %d (%d nodes)" % (
            t, wall * 1e3, st.msTick, st.msLocal, st.msTetKernel, st.tetKernelLaunches, st.msGlobal, st.msIslandKernels, st.msDetect, st.msContact, st.msOther,
            st.pcgIterationsLastTick, st.pcgLastRelResidual, st.pcgCapHits, st.triCollisions, st.staticCollisions, st.kernelLaunchesLastTick, st.reserved & 0xffff, st.reserved >> 16,
            list(st.islandsTier), st.islandsGlobal, st.islandNodesGlobal), flush=True)
p = s.positions
print("finite:", bool(np.isfinite(p).all()), "y range", p[:, 1].min(), p[:, 1].max(), "failed", s.simFailed)
