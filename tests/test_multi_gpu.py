"""Slab-partitioned solver vs the single-GPU solver on the same scene (there is no reference for
partitioning, so the test is N-rank result == 1-rank result).

 * `lockstep` tests emulate 2 and 3 ranks inside one process on ONE GPU (same plan, same send/recv lists, ghost
   rows copied tensor-to-tensor instead of through NCCL), so the halo logic is exercised on single-GPU boxes;
 * `nccl` tests spawn one process per GPU and need >= 2 GPUs (skipped otherwise).

Tolerance: 1e-4 x scene bbox diagonal, the north_star's position tolerance, inside the window in which the scene
is not yet chaotic.  The partitioned solve is not bit-identical by construction: each rank's CG stops on its own
residual norm, and a ghost body misses the contacts with its far-side neighbours within one iteration (corrected
by the per-iteration halo overwrite).  Measured on a B200 (scripts/diag_r01c.py, profiles/r01d_diag_multigpu.log):

 * the row-of-columns scene below amplifies ANY perturbation once the stacks land (ticks > ~16): the single-GPU
   solver run with pcgTolerance 0.7e-7 instead of 1e-7 drifts from itself by 1.5e-3 at tick 20 and 0.15 at tick 40,
   exactly like the partitioned run, so later ticks are checked for sanity only;
 * a ghost layer ONE body deep (halo 1.0 at pitch 2.05) leaves a coupling error of ~2e-4 per tick on glued
   columns; TWO bodies deep (halo 3.2) stays at the perturbation noise level (3e-5 at tick 5).  The tests and
   bench.py therefore use a two-body halo."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, bbox_diag

sys.path.insert(0, ROOT)
from pies_b200 import multigpu  # noqa: E402

pytestmark = pytest.mark.gpu
OPTS = dict(iterations=10, timeSubsteps=1, solver="PD")


def row_specs(columns=4, layers=2, pitch=2.05, seed=3):
    """Columns of boxes `pitch` apart along x (boxes are 2 wide: pitch 2.05 leaves a 0.05 gap, inside the 0.1 collision
    threshold, so neighbouring columns are in contact from the first tick), `layers` boxes per column falling onto
    each other and the floor."""
    rng = np.random.default_rng(seed)
    specs = []
    for layer in range(layers):
        for c in range(columns):
            t = np.array([pitch * c, 0.3 + 2.4 * layer, 0.0]) + rng.uniform(0, 0.01, 3)
            specs.append(multigpu.tetbox(t, v0=(0.0, -1.0 * layer, 0.0)))
    return specs


def _single(pb, specs, ticks):
    s = pb.Solver(**OPTS)
    for sp in specs:
        multigpu.apply_spec(s, sp)
    out = {}
    for t in range(1, max(ticks) + 1):
        s.tick()
        if t in ticks:
            out[t] = (s.positions.copy(), s.velocities.copy(), s.stats().triCollisions)
    return out


@pytest.mark.parametrize("world,halo,pitch", [(2, 3.2, 2.05), (3, 3.2, 2.05), (2, 0.5, 3.0)])
def test_lockstep_ranks_match_single_solver(pb, world, halo, pitch):
    strict = (1, 5, 10, 15)          # before the scene turns chaotic (see module docstring)
    ticks = strict + (30,)
    specs = row_specs(columns=6, pitch=pitch)
    ref = _single(pb, specs, ticks)
    ranks = [multigpu.SlabSolver(specs, rank=r, world=world, halo=halo, device=0, snap=0.5, **OPTS) for r in range(world)]
    ghosts = sum(int((~r.owned).sum()) for r in ranks)
    assert (ghosts > 0) == (pitch < 2.5)
    diag = bbox_diag(ref[1][0])
    for t in range(1, max(ticks) + 1):
        multigpu.tick_lockstep(ranks)
        if t in ticks:
            pos, prev, vel = multigpu.gather_lockstep(ranks)
            err = float(np.abs(pos - ref[t][0]).max())
            if t in strict:
                assert err <= 1e-4 * diag, (world, t, err, 1e-4 * diag, ref[t][2])
            else:   # chaotic regime: same bodies in the same places, nothing blown up or tunnelled
                assert np.isfinite(pos).all() and pos[:, 1].min() >= -1e-3 and err <= 0.05 * diag, (world, t, err)
    if pitch < 2.5:
        assert ref[max(strict)][2] > 0        # the strict comparison did cover cross-cut contacts
        owned_contacts = sum(sum(r.solver.countOwnedContacts()) for r in ranks)
        assert owned_contacts > 0


def test_one_body_halo_is_not_enough(pb):
    """Documents the coupling error of a one-body-deep ghost layer (why the default is two): it stays bounded
    (< 1e-3 x diagonal after 5 ticks) but above the strict tolerance of the two-body halo."""
    specs = row_specs(columns=6, pitch=2.05)
    ref = _single(pb, specs, (5,))
    ranks = [multigpu.SlabSolver(specs, rank=r, world=2, halo=1.0, device=0, snap=0.5, **OPTS) for r in range(2)]
    for _ in range(5):
        multigpu.tick_lockstep(ranks)
    pos, _, _ = multigpu.gather_lockstep(ranks)
    err = float(np.abs(pos - ref[5][0]).max())
    assert err <= 1e-3 * bbox_diag(ref[5][0]), err


def test_world1_slab_solver_is_the_plain_solver(pb):
    specs = row_specs(columns=3)
    ref = _single(pb, specs, (10,))
    s = multigpu.SlabSolver(specs, rank=0, world=1, device=0, **OPTS)
    for _ in range(10):
        s.tick()
    assert (s.solver.positions == ref[10][0]).all()
    assert s.projections_last_tick() == 10 * (96 * len(specs) + s.solver.stats().triCollisions + s.solver.stats().staticCollisions)


def test_repartition_keeps_the_trajectory(pb):
    """Forcing a repartition (rebuild from the gathered state) mid-run does not disturb the trajectory."""
    specs = row_specs(columns=4)
    ref = _single(pb, specs, (12,))
    s = multigpu.SlabSolver(specs, rank=0, world=1, device=0, **OPTS)
    for _ in range(6):
        s.tick()
    lo, hi = s.body_extents_x()
    s._build(lo, hi, s.gather_state())
    for _ in range(6):
        s.tick()
    diag = bbox_diag(ref[12][0])
    assert float(np.abs(s.solver.positions - ref[12][0]).max()) <= 1e-4 * diag


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_lockstep_with_tetgen_bodies(pb):
    """The slab plan handles any body kind: two TetGen cubes (the committed config-1 test mesh) side by side within
    contact reach plus tet boxes, two emulated ranks against the single solver."""
    from conftest import golden
    g = golden("tetgen_cube")
    specs = []
    for k, dx in enumerate((0.0, 8.06, 16.12)):    # side-8 cubes 0.06 apart: neighbours are in contact from the start
        pts = g["points"] + np.array([dx, 0.0, 0.0], np.float32)
        specs.append(multigpu.tetmesh(pts, g["tets"], g["faces"]))
    specs += row_specs(columns=2, layers=1, pitch=2.05, seed=5)
    for sp in specs[3:]:
        sp["t"] = sp["t"] + np.array([26.0, 0.0, 0.0], np.float32); sp["lo"] = sp["lo"] + [26.0, 0, 0]; sp["hi"] = sp["hi"] + [26.0, 0, 0]
    ticks = (1, 5, 10)
    ref = _single(pb, specs, ticks)
    ranks = [multigpu.SlabSolver(specs, rank=r, world=2, halo=9.0, device=0, **OPTS) for r in range(2)]
    assert sum(int((~r.owned).sum()) for r in ranks) > 0
    diag = bbox_diag(ref[1][0])
    for t in range(1, max(ticks) + 1):
        multigpu.tick_lockstep(ranks)
        if t in ticks:
            pos, _, _ = multigpu.gather_lockstep(ranks)
            assert float(np.abs(pos - ref[t][0]).max()) <= 1e-4 * diag, t
    assert ref[10][2] > 0


def _nccl_worker(rank, world, port, out, native=True, pitch=2.05):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    if native == "every":   # A/B switch of the library (read once per process): never skip the per-iteration exchanges
        os.environ["PIES_B200_HALO_EVERY_ITERATION"] = "1"
        native = True
    from pies_b200 import multigpu as mg
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    specs = row_specs(columns=6, pitch=pitch)
    s = mg.SlabSolver(specs, rank=rank, world=world, halo=3.2, device=rank, dist=dist, snap=0.5, check_every=4,
                      native_halo=native, **OPTS)
    assert s.native_halo == native
    for _ in range(10):
        s.tick()
    ok = s.check_halo(repartition=False)
    pos, prev, vel = s.gather_state()
    if rank == 0:
        out["pos"] = pos; out["halo_bytes"] = s.halo_bytes; out["halo_ok"] = ok
        out["exchanges"] = s.solver.stats().haloExchangesLastTick
    if native:
        s.solver.haloDestroy()
    dist.destroy_process_group()


@pytest.mark.parametrize("native,pitch", [(True, 2.05), (True, 3.0), ("every", 3.0), (False, 2.05)])
def test_nccl_two_ranks_match_single_solver(pb, native, pitch):
    """native: the halo exchange inside libpies_b200.so (ncclSend / ncclRecv from pies_b200_tick, csrc/halo.cu);
    otherwise the same lists through torch.distributed point-to-point ops between the phases of the tick.
    pitch 2.05: neighbouring columns touch, so islands mix owned and ghost rows and every PD iteration exchanges;
    pitch 3.0: columns apart — the ranks agree collectively that no island mixes and skip the per-iteration exchanges
    ("every" forces them: same positions either way)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    specs = row_specs(columns=6, pitch=pitch)
    ref = _single(pb, specs, (10,))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, _free_port(), out, native, pitch), nprocs=2, join=True)
    diag = bbox_diag(ref[10][0])
    assert float(np.abs(out["pos"] - ref[10][0]).max()) <= 1e-4 * diag
    assert out["halo_bytes"] > 0
    if native is True and pitch > 2.5:
        assert out["exchanges"] == 1       # only the three-plane exchange of the substep
    elif native is not False:
        assert out["exchanges"] == 11      # one 3-plane exchange per substep + one per PD iteration


def test_counting_owned_contacts_does_not_disturb_the_solve(pb):
    """bench.py asks every rank for its owned contacts after every tick (countOwnedContacts).  That query once wrote
    its result over the ticket counter of the CG kernels' grid reductions, so from the first contact on the dot products
    were never reduced and the solves stopped at once (seen as ~2 CG iterations per tick in a 2-GPU bench).  A solver
    that is queried every tick must follow the trajectory of one that is not, bit for bit."""
    specs = row_specs(columns=3, layers=2)
    plain = _single(pb, specs, ticks=(30,))
    s = pb.Solver(**OPTS)
    for sp in specs:
        multigpu.apply_spec(s, sp)
    s.tick()
    s.setOwnedNodes(np.ones(len(s.getVertices()), np.uint8))
    seen = 0
    for t in range(2, 31):
        s.tick()
        nt, nf = s.countOwnedContacts()
        st = s.stats()
        assert (nt, nf) == (st.triCollisions, st.staticCollisions)
        seen = max(seen, nt)
    assert seen > 0
    assert (s.positions == plain[30][0]).all() and (s.velocities == plain[30][1]).all()
