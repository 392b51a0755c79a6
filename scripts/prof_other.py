"""Profiling target for configs 2 and 4: builds the (reduced) scene, runs SKIP ticks, brackets TICKS ticks with
cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).  WORKLOAD=s2|s4, SIZE = nodes (s2) / bodies (s4)."""
import os, sys
sys.path.insert(0, ".")
import torch
import pies_b200 as pb
from pies_b200 import scenes
wl = os.environ.get("WORKLOAD", "s4"); skip = int(os.environ.get("SKIP", "20")); ticks = int(os.environ.get("TICKS", "1"))
if wl == "s2":
    n = int(os.environ.get("SIZE", "100000"))
    s = pb.Solver(**scenes.S2_OPTIONS); scenes.build_rope(s, n=n, shape="spiral"); script = None
else:
    b = int(os.environ.get("SIZE", "1000")); side = max(1, round(b ** (1 / 3)))
    s = pb.Solver(iterations=4); _, regions = scenes.build_s4(s, bodies=b, per_side=side, goal_bodies=min(244, b))
    script = lambda t: s.updateFixedRegions(scenes.s4_region_script(regions, t))
for t in range(1, skip + 1):
    if script: script(t)
    s.tick()
torch.cuda.synchronize(); torch.cuda.profiler.start()
for t in range(skip + 1, skip + ticks + 1):
    if script: script(t)
    s.tick()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
st = s.stats()
print(wl, "tick", skip + ticks, "ms", st.msTick, "pt", st.triCollisions, "floor", st.staticCollisions, "proj", st.projectionsLastTick, "launches", st.kernelLaunchesLastTick)
