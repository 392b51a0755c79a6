"""Runs S3 to tick SKIP unprofiled, then brackets TICKS ticks with cudaProfilerStart/Stop
(use with `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, ".")
import torch
import pies_b200 as pb
from pies_b200 import scenes
bodies = int(os.environ.get("BODIES", "20834")); skip = int(os.environ.get("SKIP", "60")); ticks = int(os.environ.get("TICKS", "1"))
s = pb.Solver(**scenes.S3_OPTIONS)
scenes.build_s3(s, bodies)
if os.environ.get("PIES_TOL"):
    s.setTuning(pcgTolerance=float(os.environ["PIES_TOL"]))
for _ in range(skip):
    s.tick()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(ticks):
    s.tick()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
st = s.stats()
print("tick", skip + ticks, "pt", st.triCollisions, "floor", st.staticCollisions, "pcg", st.pcgIterationsLastTick, "ms", st.msTick)
