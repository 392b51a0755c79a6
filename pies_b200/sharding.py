"""Host-side partitioning of multi-body scenes across ranks.

Bodies are independent except through contacts, so a multi-body scene is cut into `world`
slabs along x with equal constraint counts; each rank owns the bodies of its slab.  Bodies whose
swept extent reaches within `halo` of a cut are reported as boundary bodies: their surface nodes
are what a rank has to publish to its slab neighbour before collision detection.
pies_b200/multigpu.py builds the per-rank solvers and exchange lists from this partition; the halo exchange itself runs
inside libpies_b200.so (csrc/halo.cu).
"""
import numpy as np


def slab_partition(centroid_x, weight, world):
    """Returns (owner[b], cuts[world-1]): bodies sorted by centroid x are split into `world`
    contiguous groups of (nearly) equal total weight.  Deterministic; ties broken by body index."""
    centroid_x = np.asarray(centroid_x, dtype=np.float64)
    weight = np.asarray(weight, dtype=np.float64)
    order = np.lexsort((np.arange(len(centroid_x)), centroid_x))
    csum = np.cumsum(weight[order])
    total = csum[-1] if len(csum) else 0.0
    owner = np.empty(len(centroid_x), dtype=np.int32)
    bounds = np.searchsorted(csum, total * np.arange(1, world) / world, side="left")
    # never cut inside a group of bodies sharing the same x (columns of a stack stay together)
    cuts = []
    for b in bounds:
        while 0 < b < len(order) and centroid_x[order[b]] == centroid_x[order[b - 1]]:
            b += 1
        cuts.append(min(b, len(order)))
    start = 0
    for rank, end in enumerate(cuts + [len(order)]):
        owner[order[start:end]] = rank
        start = end
    cut_x = [0.5 * (centroid_x[order[c - 1]] + centroid_x[order[c]]) if 0 < c < len(order) else np.inf for c in cuts]
    return owner, np.asarray(cut_x)


def boundary_bodies(min_x, max_x, owner, cut_x, halo):
    """Bodies that must publish surface nodes to a neighbouring slab: their [min_x - halo, max_x + halo]
    interval crosses a cut.  Returns a list (per cut) of (left_rank_bodies, right_rank_bodies)."""
    out = []
    for k, c in enumerate(cut_x):
        near = (np.asarray(max_x) + halo >= c) & (np.asarray(min_x) - halo <= c)
        out.append((np.flatnonzero(near & (owner == k)), np.flatnonzero(near & (owner == k + 1))))
    return out
