"""How the PBD tick scales with the rope length (config 2): ms per tick, node-node visits, launches."""
import sys, time, os
sys.path.insert(0, ".")
import pies_b200 as pb
from pies_b200 import scenes
for shape in ("spiral", "helix"):
    for n in (1000, 2000, 4000, 8000):
        s = pb.Solver(**scenes.S2_OPTIONS)
        scenes.build_rope(s, n=n, shape=shape)
        t0 = time.time(); s.tick(); first = time.time() - t0
        ms = []
        for _ in range(4):
            t0 = time.time(); s.tick(); ms.append(1e3 * (time.time() - t0))
        st = s.stats()
        print("%s n=%d first %.2fs ticks %s ms | dev %.1f ms | proj %d collision visits %d launches %d" % (
            shape, n, first, ["%.1f" % m for m in ms], st.msTick, st.projectionsLastTick, st.collisionProjections, st.kernelLaunchesLastTick), flush=True)
        if max(ms) > 3000:
            break
