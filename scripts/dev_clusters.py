import os, sys
sys.path.insert(0, ".")
import numpy as np, scipy.sparse as sp
from scipy.sparse.csgraph import connected_components
import pies_b200 as pb
from pies_b200 import scenes
s = pb.Solver(**scenes.S3_OPTIONS)
scenes.build_s3(s)
n = len(s.getVertices())
t = 0
for stop in (61, 111, 145):
    while t < stop:
        s.tick(); t += 1
    c = s.triCollisions().astype(np.int64)
    rows = np.concatenate([c[:, 0]] * 3); cols = np.concatenate([c[:, 1], c[:, 2], c[:, 3]])
    g = sp.coo_matrix((np.ones(len(rows), np.int8), (rows, cols)), shape=(n, n))
    nc, lab = connected_components(g, directed=False)
    touched = np.zeros(n, bool); touched[np.unique(c)] = True
    sizes = np.bincount(lab[touched]); sizes = sizes[sizes > 0]
    ent = np.bincount(lab[c[:, 0]], minlength=len(np.bincount(lab)))
    ent = ent[ent > 0]
    order = np.argsort(-sizes)
    print("tick %d: %d entries, %d clusters; nodes/cluster mean %.1f max %d; >32: %d clusters, >64: %d, >128: %d, >256: %d" % (
        t, len(c), len(sizes), sizes.mean(), sizes.max(), (sizes > 32).sum(), (sizes > 64).sum(), (sizes > 128).sum(), (sizes > 256).sum()))
    lab_sizes = np.bincount(lab[touched], minlength=lab.max() + 1)
    esz = lab_sizes[lab[c[:, 0]]]
    print("   entries in clusters <=32: %d, 33-64: %d, 65-128: %d, 129-256: %d, >256: %d; max entries in one cluster %d" % (
        (esz <= 32).sum(), ((esz > 32) & (esz <= 64)).sum(), ((esz > 64) & (esz <= 128)).sum(), ((esz > 128) & (esz <= 256)).sum(), (esz > 256).sum(), ent.max()))
