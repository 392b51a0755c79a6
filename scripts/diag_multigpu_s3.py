"""Diagnosis: the slab-partitioned S3 bench scene on N ranks, per-tick state of every rank
(torchrun --nproc-per-node N scripts/diag_multigpu_s3.py)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
import pies_b200 as pb
from pies_b200 import scenes, multigpu

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.Stream(device=lr); torch.cuda.set_stream(stream)
bodies = int(os.environ.get("BODIES", "20834")); ticks = int(os.environ.get("TICKS", "120"))
trans = scenes.s3_translations(bodies * world, nx=32 * world)
specs = [multigpu.tetbox(t) for t in trans]
drv = multigpu.SlabSolver(specs, rank=rank, world=world, halo=float(os.environ.get("HALO", "4.0")), device=lr, dist=dist, snap=0.5,
                          **scenes.S3_OPTIONS)
s = drv.solver
if os.environ.get("DATAFLOW_ONLY"):
    s.setTuning(dataflowSweepsOnly=True)
for t in range(1, ticks + 1):
    drv.tick()
    if t % 5 == 0 or t > int(os.environ.get("VERBOSE_FROM", "1000")):
        st = s.stats()
        p = s.positions
        ok = drv.check_halo(repartition=False)
        print("rank %d tick %3d pt %7d floor %6d pcg %4d res %.1e failed %d finite %s y [%.3f, %.3f] x [%.2f, %.2f] halo_ok %s mid %d large %d" % (
            rank, t, st.triCollisions, st.staticCollisions, st.pcgIterationsLastTick, st.pcgLastRelResidual, s.simFailed,
            bool(np.isfinite(p).all()), p[:, 1].min(), p[:, 1].max(), p[:, 0].min(), p[:, 0].max(), ok, st.reserved & 0xffff, st.reserved >> 16), flush=True)
dist.destroy_process_group()
