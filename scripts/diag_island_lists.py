"""Per tick: islands per list (warp / dense / cta128 / cta320 / cta512 / grid-wide) and the phase times, S3 ticks FROM..TO."""
import os, sys
sys.path.insert(0, ".")
os.environ.setdefault("PIES_B200_ISLAND_TRACE", "1")
import numpy as np
import pies_b200 as pb
from pies_b200 import scenes
s = pb.Solver(**scenes.S3_OPTIONS)
scenes.build_s3(s, int(os.environ.get("BODIES", "20834")))
s.setTuning(profilePhases=True)
lo, hi = int(os.environ.get("FROM", "60")), int(os.environ.get("TO", "120"))
for t in range(1, hi + 1):
    s.tick()
    if t >= lo and (t - lo) % int(os.environ.get("EVERY", "2")) == 0:
        st = s.stats()
        n = {name: len(s.debugIslandTrace(slot)) for slot, name in ((0, "warp"), (5, "dense"), (6, "dense192"), (4, "cta128"), (1, "cta320"), (2, "cta512"))}
        rows = {}
        for slot, name in ((5, "dense"), (6, "dense192"), (4, "cta128"), (1, "cta320"), (2, "cta512")):
            tr = s.debugIslandTrace(slot)
            if len(tr):
                rows[name] = "rows %d..%d it<=%d" % (tr[:, 0].min(), tr[:, 0].max(), tr[:, 1].max())
        print("tick %3d  ms %.2f  detect %.2f local %.2f global %.2f contact %.2f | pcg %3d | %s grid-wide %d | %s"
              % (t, st.msTick, st.msDetect, st.msLocal, st.msGlobal, st.msContact, st.pcgIterationsLastTick, n, st.islandsGlobal, rows), flush=True)
