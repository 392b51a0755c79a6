"""Small contact scene for compute-sanitizer: the 4 x 12 stack (islands of every list from tick ~30 on), TICKS ticks."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import pies_b200 as pb
from pies_b200 import scenes
s = pb.Solver(**scenes.S3_OPTIONS)
scenes.build_s3(s, bodies=48, nx=2, nz=2)
for t in range(int(os.environ.get("TICKS", "46"))):
    s.tick()
st = s.stats()
print("ticks done; contacts", st.triCollisions, st.staticCollisions, "islands", list(st.islandsTier), "finite", bool(np.isfinite(s.positions).all()))
