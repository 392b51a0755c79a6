"""Throw-away GPU diagnostics for the failures of the r01c run (sheet scenes, partitioned-vs-single drift)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pies_b200 as pb
from oracle import refapi
from pies_b200 import multigpu, scenes


def sheets(which):
    def build(s):
        if which in ("both", "sheet"):
            s.createSheet((0.0, 4.0, 0.0), 1.0, 1.0, 100.0)
        if which in ("both", "bend"):
            s.createBendSheet((15.0, 4.0, 0.0), 1.0, 100.0)
    return build


def pd_sheets():
    for which in ("sheet", "bend", "both"):
        print("== PD sheets:", which, flush=True)
        g = pb.Solver(iterations=10); r = refapi.RefSolver(iterations=10)
        sheets(which)(g); sheets(which)(r)
        for t in range(1, 11):
            try:
                g.tick()
            except Exception as e:
                print("  tick", t, "ours raised", e); break
            r.tick()
            st = g.stats()
            gp = g.getVertices()["position"]; rp = r.getVertices()
            err = np.abs(gp - rp)
            i = int(err.max(axis=1).argmax())
            print("  t %2d err %.3e at node %d ours %s ref %s | tri %d/%d floor %d/%d failed %s/%s pcg %d res %.1e" % (
                t, err.max(), i, gp[i], rp[i], st.triCollisions, r.count("tri_collision"), st.staticCollisions,
                r.count("static_collision"), g.simFailed, r.simFailed, st.pcgIterationsLastTick, st.pcgLastRelResidual), flush=True)


def pbd_sheets():
    for which in ("sheet", "bend", "both"):
        print("== PBD sheets:", which, flush=True)
        def build(s):
            if which in ("both", "sheet"):
                s.createSheet((0.0, 2.0, 0.0), 0.5, 1.0, 0.8)
            if which in ("both", "bend"):
                s.createBendSheet((12.0, 2.0, 0.0), 0.5, 0.6)
        g = pb.Solver(**scenes.S2_OPTIONS); r = refapi.RefSolver(**scenes.S2_OPTIONS)
        build(g); build(r)
        for t in range(1, 11):
            try:
                g.tick()
            except Exception as e:
                print("  tick", t, "ours raised", e); break
            r.tick()
            gp = g.positions; rp = r.positions
            err = np.abs(gp - rp)
            i = int(err.max(axis=1).argmax())
            print("  t %2d err %.3e at node %d ours %s ref %s finite %s/%s" % (t, err.max(), i, gp[i], rp[i], np.isfinite(gp).all(), np.isfinite(rp).all()), flush=True)


def mgpu():
    from test_multi_gpu import row_specs, OPTS
    for world, halo, pitch in ((2, 0.5, 3.0), (2, 1.0, 2.05), (2, 3.2, 2.05)):
        print("== lockstep world %d halo %.1f pitch %.2f" % (world, halo, pitch), flush=True)
        specs = row_specs(columns=6, pitch=pitch)
        a = pb.Solver(**OPTS); b = pb.Solver(**OPTS)
        for sp in specs:
            multigpu.apply_spec(a, sp); multigpu.apply_spec(b, sp)
        b.setTuning(pcgTolerance=0.7e-7)   # chaos probe: the same scene with a slightly different CG stop
        ranks = [multigpu.SlabSolver(specs, rank=r, world=world, halo=halo, device=0, snap=0.5, **OPTS) for r in range(world)]
        print("   ghosts", sum(int((~r.owned).sum()) for r in ranks))
        for t in range(1, 41):
            a.tick(); b.tick()
            multigpu.tick_lockstep(ranks)
            pos, prev, vel = multigpu.gather_lockstep(ranks)
            ea = float(np.abs(pos - a.positions).max()); eb = float(np.abs(b.positions - a.positions).max())
            # collision lists in global node ids
            la = a.triCollisions()
            parts = []
            for r in ranks:
                lt = r.solver.triCollisions()
                if len(lt):
                    gl = r.l2g[lt.astype(np.int64)]
                    own = r.owned[lt[:, 0].astype(np.int64)]
                    parts.append(gl[own])
            lp = np.concatenate(parts) if parts else np.zeros((0, 4), np.int64)
            same_set = sorted(map(tuple, la.tolist())) == sorted(map(tuple, lp.tolist()))
            if t <= 6 or t % 4 == 0 or not same_set:
                print("  t %2d part-vs-single %.3e | tol-perturbed-vs-single %.3e | contacts single %d part(owned) %d same multiset %s" % (
                    t, ea, eb, len(la), len(lp), same_set), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["pd", "pbd", "mgpu"]
    if "pd" in what: pd_sheets()
    if "pbd" in what: pbd_sheets()
    if "mgpu" in what: mgpu()
