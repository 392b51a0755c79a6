import os, sys
sys.path.insert(0, ".")
import numpy as np
import pies_b200 as pb
from pies_b200 import scenes
skip = int(os.environ.get("SKIP", "60"))
for tol in (1e-7, 1e-6, 1e-5):
    s = pb.Solver(**scenes.S3_OPTIONS)
    scenes.build_s3(s)
    s.setTuning(pcgTolerance=tol, profilePhases=True)
    for _ in range(skip):
        s.tick()
    if tol == 1e-7:
        os.environ["PIES_DEBUG_PCG"] = "1"
    s.tick()
    os.environ.pop("PIES_DEBUG_PCG", None)
    st = s.stats()
    print("tol %g: tick %d pcg %d global %.2f ms tick %.2f ms pt %d" % (tol, skip + 1, st.pcgIterationsLastTick, st.msGlobal, st.msTick, st.triCollisions), flush=True)
    if tol == 1e-7:
        c = s.triCollisions()
        u, cnt = np.unique(c, axis=0, return_counts=True)
        print("contacts %d unique %d max mult %d; unique (a) nodes %d, unique tris %d" % (len(c), len(u), cnt.max(), len(np.unique(c[:, 0])), len(np.unique(c[:, 1:], axis=0))))
        pairs = np.unique(np.stack([c[:, 0] // 27, c[:, 1] // 27], 1), axis=0)
        print("body pairs in contact:", len(pairs))
    del s
