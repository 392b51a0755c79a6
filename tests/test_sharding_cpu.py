"""CPU tests of the host-side multi-GPU logic: world_size-2 gloo processes agree on the slab
partition of a multi-body scene, cover every body exactly once, and see matching halos."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from pies_b200 import scenes, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = scenes.s3_translations(96, nx=8, nz=4)
    cx = t[:, 0] + 1.0
    owner, cuts = sharding.slab_partition(cx, np.full(len(cx), 96.0), world)
    mine = np.flatnonzero(owner == rank)
    # every rank publishes how many bodies it owns and the checksum of their ids
    info = torch.tensor([len(mine), int(mine.sum())], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, info)
    halos = sharding.boundary_bodies(t[:, 0], t[:, 0] + 2.0, owner, cuts, halo=1.1)
    left, right = halos[0]
    # the two sides of a cut exchange halo sizes: what rank 0 sends is what rank 1 expects
    send = torch.tensor([len(left) if rank == 0 else len(right)], dtype=torch.int64)
    recv = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(recv, send)
    # max-over-ranks timing reduction used by bench.py
    tmax = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    out[rank] = dict(counts=[int(g[0]) for g in gathered], sums=[int(g[1]) for g in gathered],
                     halo=[int(r[0]) for r in recv], expect_halo=[len(left), len(right)], tmax=float(tmax[0]),
                     owner=owner.tolist())
    dist.destroy_process_group()


def test_slab_partition_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert a["owner"] == b["owner"]                      # deterministic, identical on every rank
    assert sum(a["counts"]) == 96 and a["counts"] == b["counts"]
    assert sum(a["sums"]) == sum(range(96))              # every body owned exactly once
    assert abs(a["counts"][0] - a["counts"][1]) <= 12    # balanced to within one column of bodies
    assert a["halo"] == a["expect_halo"] == b["halo"]
    assert a["halo"][0] > 0 and a["halo"][1] > 0
    assert a["tmax"] == 2.0


def test_slab_partition_properties():
    sys.path.insert(0, ROOT)
    from pies_b200 import sharding
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 100, 1000)
    w = rng.integers(1, 10, 1000).astype(float)
    for world in (1, 2, 4, 8):
        owner, cuts = sharding.slab_partition(x, w, world)
        assert set(owner.tolist()) == set(range(world))
        assert len(cuts) == world - 1
        loads = np.array([w[owner == r].sum() for r in range(world)])
        assert loads.max() - loads.min() <= 2 * w.max() * world
        for r in range(world - 1):   # slabs are ordered along x
            assert x[owner == r].max() <= x[owner == r + 1].min()
