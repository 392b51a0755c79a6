"""Throw-away: PBD bend-sheet divergence probe."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pies_b200 as pb
from oracle import refapi
from pies_b200 import scenes

f32 = np.float32
def bend_terms(pos, ids, angle):
    x1, x2, x3, x4 = (pos[ids[:, k]].astype(f32) for k in range(4))
    p2, p3, p4 = x2 - x1, x3 - x1, x4 - x1
    c23, c24 = np.cross(p2, p3).astype(f32), np.cross(p2, p4).astype(f32)
    l23 = np.sqrt((c23 * c23).sum(1, dtype=f32)); l24 = np.sqrt((c24 * c24).sum(1, dtype=f32))
    n1, n2 = c23 / l23[:, None], c24 / l24[:, None]
    d = (n1 * n2).sum(1, dtype=f32)
    q3 = (np.cross(p2, n2) + np.cross(n1, p2) * d[:, None]) / l23[:, None]
    q4 = (np.cross(p2, n1) + np.cross(n2, p2) * d[:, None]) / l24[:, None]
    q2 = -((np.cross(p3, n2) + np.cross(n1, p3) * d[:, None]) / l23[:, None]) - ((np.cross(p4, n1) + np.cross(n2, p4) * d[:, None]) / l24[:, None])
    q1 = -q2 - q3 - q4
    qq = sum((q * q).sum(1) for q in (q1, q2, q3, q4)).astype(f32)
    return d, qq, l23, l24

for iters in (1, 4):
    opts = dict(scenes.S2_OPTIONS); opts["iterations"] = iters
    print("== PBD bend sheet, iterations", iters, flush=True)
    g = pb.Solver(**opts); r = refapi.RefSolver(**opts)
    for s in (g, r):
        s.createBendSheet((12.0, 2.0, 0.0), 0.5, 0.6)
    ids, angle = r.bends()[:2]
    ids = np.asarray(ids).reshape(-1, 4).astype(np.int64)
    print("   bends", len(ids), "rest angles", np.unique(np.asarray(angle))[:6])
    for t in range(1, 9):
        try:
            g.tick()
        except Exception as e:
            print("  tick", t, "ours raised", e)
            break
        r.tick()
        gp, rp = g.positions, r.positions
        err = np.abs(gp - rp); i = int(err.max(1).argmax())
        d, qq, l23, l24 = bend_terms(rp, ids, angle)
        print("  t %d err %.3e node %d ours %s ref %s | ref-state: d in [%.9f, %.9f] d<-1: %d qq in [%.3e, %.3e] qq in (0.5e-5,2e-5): %d qq>=1e-5: %d  max|pos| ours %.3e ref %.3e" % (
            t, err.max(), i, gp[i], rp[i], d.min(), d.max(), int((d < -1).sum()), qq.min(), qq.max(), int(((qq > 0.5e-5) & (qq < 2e-5)).sum()),
            int((qq >= 1e-5).sum()), np.abs(gp).max(), np.abs(rp).max()), flush=True)
