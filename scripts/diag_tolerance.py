"""Error against the golden reference trajectories as a function of the PCG tolerance (two-box collision, TetGen cube on
the floor): separates solver-tolerance error from ordering / rounding differences."""
import sys
import numpy as np
sys.path.insert(0, ".")
import pies_b200 as pb
g1 = np.load("tests/golden/collisions.npz"); g2 = np.load("tests/golden/tetgen_cube.npz")
def two_box(s):
    s.createTetBox((0.1, 0.3, 0.1), 1.0, (0, 0, 0), 1000.0, 1.0, False)
    s.createTetBox((0.4, 2.6, 0.3), 1.0, (0, -5, 0), 1000.0, 1.0, False)
for isl in (True, False):
    for tol in (1e-7, 1e-8, 1e-9, 1e-10):
        s = pb.Solver(iterations=10); two_box(s); s.setTuning(pcgTolerance=tol, islandSolves=isl, pcgMaxIterations=400)
        diag = float(np.linalg.norm(g1["traj1_pos"].max(0) - g1["traj1_pos"].min(0)))
        line = "two_box islands %d tol %.0e:" % (isl, tol)
        its = 0
        for t in range(1, 41):
            try:
                s.tick()
            except pb.PiesError as e:
                line += " [t%d %s]" % (t, str(e)[-60:])
            its += s.stats().pcgIterationsLastTick
            if t in (1, 10, 40):
                line += " K=%d err/tol %.3f" % (t, np.abs(s.positions - g1["traj%d_pos" % t]).max() / (1e-4 * diag))
        print(line, "| iterations", its, "cap", s.stats().pcgCapHits, flush=True)
        s = pb.Solver(); s.addTetMeshVolume(g2["points"], g2["tets"], g2["faces"], (0, 0, 0), 1.0, 1000.0, 0.8, 1.0, 1000.0, 1.0, 1.0)
        s.setTuning(pcgTolerance=tol, islandSolves=isl, pcgMaxIterations=400)
        diag = float(np.linalg.norm(g2["points"].max(0) - g2["points"].min(0)))
        line = "tetgen_cube islands %d tol %.0e:" % (isl, tol)
        its = 0
        for t in range(1, 61):
            try:
                s.tick()
            except pb.PiesError as e:
                line += " [t%d %s]" % (t, str(e)[-60:])
            its += s.stats().pcgIterationsLastTick
            if t in (1, 10, 30, 60):
                st = s.stats()
                line += " t=%d err/tol %.3f (%d,%d vs %s)" % (t, np.abs(s.positions - g2["pos%d" % t]).max() / (1e-4 * diag), st.triCollisions, st.staticCollisions, tuple(g2["ncoll%d" % t]))
        print(line, "| iterations", its, "tiers", list(s.stats().islandsTier), "gw", s.stats().islandsGlobal, flush=True)
