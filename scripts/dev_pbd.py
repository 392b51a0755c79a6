"""PBD debugging on the GPU box: error messages, per-tick times, visit counts."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import pies_b200 as pb
from pies_b200 import scenes
import os
def golden(name):
    return np.load(os.path.join("tests", "golden", name + ".npz"))

def trial(name, build, ticks, **opts):
    s = pb.Solver(**opts)
    build(s)
    t0 = time.time()
    try:
        for t in range(ticks):
            t1 = time.time(); s.tick()
            if t < 3: print("   tick %d: %.1f ms wall, %.2f ms dev" % (t, 1e3 * (time.time() - t1), s.stats().msTick), flush=True)
    except Exception as e:
        print(name, "FAILED at tick", t, ":", e)
        return None
    st = s.stats()
    print("%s: %d ticks in %.3fs (%.2f ms/tick, last dev %.2f ms), launches/tick %d, visits %d, finite %s" % (
        name, ticks, time.time() - t0, 1e3 * (time.time() - t0) / ticks, st.msTick, st.kernelLaunchesLastTick,
        st.collisionProjections, bool(np.isfinite(s.positions).all())), flush=True)
    return s

def sheets(s):
    s.createSheet((0.0, 2.0, 0.0), 0.5, 1.0, 0.8)
    s.createBendSheet((12.0, 2.0, 0.0), 0.5, 0.6)
def sheets_pd(s):
    s.createSheet((0.0, 4.0, 0.0), 1.0, 1.0, 100.0)
    s.createBendSheet((15.0, 4.0, 0.0), 1.0, 100.0)
g = golden("pbd")
s = trial("pbd sheets", sheets, 10, **scenes.S2_OPTIONS)
if s is not None:
    print("  err vs ref tick10:", np.abs(s.positions - g["sheets_pos10"]).max())
trial("pbd sheet only", lambda s: s.createSheet((0.0, 2.0, 0.0), 0.5, 1.0, 0.8), 10, **scenes.S2_OPTIONS)
trial("pbd bendsheet only", lambda s: s.createBendSheet((12.0, 2.0, 0.0), 0.5, 0.6), 10, **scenes.S2_OPTIONS)
trial("pd sheet+bendsheet", sheets_pd, 10, iterations=10)
trial("pd sheet", lambda s: s.createSheet((0.0, 4.0, 0.0), 1.0, 1.0, 100.0), 10, iterations=10)
trial("pd bendsheet", lambda s: s.createBendSheet((15.0, 4.0, 0.0), 1.0, 100.0), 10, iterations=10)
trial("pbd boxes", scenes.build_pbd_boxes, 60, **scenes.S2_OPTIONS)
trial("pbd rope 2000", lambda s: scenes.build_rope(s, n=2000, helix_radius=2.0), 20, **scenes.S2_OPTIONS)
t0 = time.time()
trial("pbd rope 100k", lambda s: scenes.build_rope(s, n=100000), 5, **scenes.S2_OPTIONS)
print("rope 100k total incl. build %.2fs" % (time.time() - t0))
trial("pbd spiral 100k", lambda s: scenes.build_rope(s, n=100000, shape="spiral", pinned=False), 3, **scenes.S2_OPTIONS)
