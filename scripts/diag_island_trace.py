"""Island-size / iteration / clock distribution of the CTA island lists at a given S3 tick (PIES_B200_ISLAND_TRACE=1)."""
import os, sys
sys.path.insert(0, ".")
os.environ.setdefault("PIES_B200_ISLAND_TRACE", "1")
import numpy as np
import torch
import pies_b200 as pb
from pies_b200 import scenes
bodies = int(os.environ.get("BODIES", "20834"))
s = pb.Solver(**scenes.S3_OPTIONS)
scenes.build_s3(s, bodies)
for t in range(1, 1 + int(os.environ.get("TICKS", "75"))):
    s.tick()
    if t in (61, 75):
        st = s.stats()
        print("tick", t, "pcg", st.pcgIterationsLastTick, "islands", list(st.islandsTier), "msGlobal", st.msGlobal)
        for slot, name in ((4, "cta128"), (1, "cta320"), (0, "warp")):
            tr = s.debugIslandTrace(slot)
            if not len(tr):
                continue
            m, it, clk, nnz = tr[:, 0].astype(np.int64), tr[:, 1].astype(np.int64), tr[:, 2].astype(np.int64), tr[:, 3].astype(np.int64)
            print(" %s: %d islands rows mean %.0f max %d | iters mean %.1f max %d | clocks mean %.0f max %d sum %.3g | nnz/row %.1f"
                  % (name, len(tr), m.mean(), m.max(), it.mean(), it.max(), clk.mean(), clk.max(), clk.sum(), nnz.sum() / max(m.sum(), 1)))
            edges = [0, 32, 64, 96, 128, 160, 192, 224, 256, 320, 448, 640]
            for lo, hi in zip(edges[:-1], edges[1:]):
                sel = (m > lo) & (m <= hi)
                if sel.any():
                    print("   rows %3d..%3d: %4d islands, iters %.1f (max %d), clocks/island %.0f (max %d), clocks/iter %.0f, nnz/row %.1f"
                          % (lo + 1, hi, sel.sum(), it[sel].mean(), it[sel].max(), clk[sel].mean(), clk[sel].max(),
                             clk[sel].sum() / max(it[sel].sum(), 1), nnz[sel].sum() / m[sel].sum()))
