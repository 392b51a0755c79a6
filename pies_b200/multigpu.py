"""Slab-partitioned multi-GPU driver for multi-body scenes.

One process per GPU.  Bodies (whole, never split: a connected body's global solve couples all its nodes) are
sorted by centroid x and cut into `world` slabs of equal constraint count.  Rank r runs an ordinary
pies_b200.Solver on

    owned bodies  +  ghost bodies = other ranks' bodies within `halo` of r's owned extent,

created in GLOBAL body order, so local node / triangle order is the global order restricted to the local
set, and the collision lists keep the reference's canonical order (pies_b200_set_triangle_order).  Ghost
bodies are simulated in full (all their constraints), which makes every owned<->ghost contact part of the
same coupled block as in the single-GPU solve; what a ghost lacks is the contacts with ITS far-side
neighbours, and that error is removed by overwriting the ghost nodes with their owners' values

  * at every substep start: position, previous position, velocity (24+12 B per node),
  * after every PD iteration: position (the north_star's "halo exchanged per iteration"),

inside the library: `pies_b200_halo_*` (csrc/halo.cu) packs the rows, exchanges them with ncclSend / ncclRecv in one
group and unpacks them, all on the solver's stream, from within `pies_b200_tick` — this module only computes the
partition and the exchange lists and hands the NCCL unique id around (over the caller's `torch.distributed` group,
which is rendezvous plumbing only).  Without NCCL (the gloo CPU tests, the single-GPU lockstep emulation) the same
lists drive a `torch.distributed` point-to-point exchange between the phases of the tick.
No data-path collective exists besides this exchange and a one-int failure reduction per tick.  Every `check_every`
ticks the ranks compare the bodies' current extents with the ghost sets (`check_halo`); if a contact partner could be
missing from a rank's ghost set the scene is repartitioned from the gathered state.

Nothing here touches the oracle; the solver behind it is the CUDA library only.
"""
import numpy as np

from .sharding import slab_partition

# ------------------------------------------------------------------------------------------------
# body specs: what a rank needs to know about a body without building it — node / triangle / static projection counts
# (per PD iteration) and the x extent — plus what it takes to build it through the reference's factories
def tetbox(t, scale=1.0, v0=(0.0, 0.0, 0.0), w=1000.0, mass=1.0):
    """createTetBox, non-hinged: 3x3x3 nodes, 48 tets -> 48 strain + 48 volume constraints, 48 triangles."""
    t = np.asarray(t, np.float32)
    return dict(kind="tetbox", t=t, scale=float(scale), v0=tuple(v0), w=float(w), mass=float(mass),
                nodes=27, tris=48, proj=96, lo=t.astype(np.float64), hi=t.astype(np.float64) + 2.0 * scale)


def tetmesh(points, tets, faces, v0=(0.0, 0.0, 0.0), density=1.0, strain_w=1000.0, min_strain=0.8, max_strain=1.0,
            volume_w=1000.0, compression=1.0, stretching=1.0):
    """A tetrahedralised body as Solver::addTriMeshVolume leaves it after TetGen (config 1 / config 5 bodies):
    points, tets (4 ids) and re-wound boundary faces (3 ids), local to the body."""
    points = np.ascontiguousarray(points, np.float32)
    per_tet = (1 if strain_w != 0 else 0) + (1 if volume_w != 0 else 0)
    return dict(kind="tetmesh", points=points, tets=np.ascontiguousarray(tets, np.uint32), faces=np.ascontiguousarray(faces, np.uint32),
                v0=tuple(v0), args=(density, strain_w, min_strain, max_strain, volume_w, compression, stretching),
                nodes=len(points), tris=len(faces), proj=per_tet * len(tets),
                lo=points.min(0).astype(np.float64), hi=points.max(0).astype(np.float64))


def apply_spec(solver, spec):
    if spec["kind"] == "tetbox":
        solver.createTetBox(spec["t"], spec["scale"], spec["v0"], spec["w"], spec["mass"], False)
    elif spec["kind"] == "tetmesh":
        solver.addTetMeshVolume(spec["points"], spec["tets"], spec["faces"], spec["v0"], *spec["args"])
    else:
        raise ValueError("unknown body kind %r" % spec["kind"])


class SlabPlan:
    """Deterministic partition + halo plan, identical on every rank (pure numpy)."""

    def __init__(self, counts, lo_x, hi_x, world, halo, thread_count=8, snap=0.0):
        """counts: per body (nodes, triangles, static projections per PD iteration)."""
        counts = np.asarray(counts, np.int64).reshape(-1, 3)
        nb = len(counts)
        self.world, self.halo, self.thread_count = world, float(halo), int(thread_count)
        self.nodes, self.tris, self.proj = counts[:, 0].copy(), counts[:, 1].copy(), counts[:, 2].copy()
        self.node_off = np.concatenate([[0], np.cumsum(self.nodes)])
        self.tri_off = np.concatenate([[0], np.cumsum(self.tris)])
        lo_x = np.asarray(lo_x, np.float64); hi_x = np.asarray(hi_x, np.float64)
        cx = 0.5 * (lo_x + hi_x)
        if snap > 0.0:   # bodies whose centroids fall in the same `snap`-wide bin are never separated (stacked columns)
            cx = np.round(cx / snap) * snap
        self.owner, self.cuts = slab_partition(cx, self.proj, world)
        self.local = []        # per rank: sorted global body ids (owned + ghosts)
        self.ext = []          # per rank: x extent of the owned bodies
        for r in range(world):
            own = self.owner == r
            if not own.any():
                self.local.append(np.zeros(0, np.int64)); self.ext.append((np.inf, -np.inf)); continue
            e = (lo_x[own].min(), hi_x[own].max())
            ghost = (~own) & (hi_x + halo >= e[0]) & (lo_x - halo <= e[1])
            self.local.append(np.flatnonzero(own | ghost))
            self.ext.append(e)
        self.n_bodies = nb

    # ---- per-rank views ----
    def local_node_offsets(self, r):
        return np.concatenate([[0], np.cumsum(self.nodes[self.local[r]])])

    def owned_node_mask(self, r):
        loc = self.local[r]
        return np.repeat((self.owner[loc] == r).astype(np.uint8), self.nodes[loc])

    def local_to_global_nodes(self, r):
        loc = self.local[r]
        return np.concatenate([self.node_off[b] + np.arange(self.nodes[b]) for b in loc]) if len(loc) else np.zeros(0, np.int64)

    def triangle_order(self, r):
        """Position of every local triangle in the global canonical order (thread t handles triangles t, t+T, ...;
        per-thread lists concatenated in thread order, reference Solver.cpp:714,852-873) restricted to the rank."""
        loc = self.local[r]
        if not len(loc):
            return np.zeros(0, np.uint32)
        gt = np.concatenate([self.tri_off[b] + np.arange(self.tris[b]) for b in loc])
        n, T = int(self.tri_off[-1]), self.thread_count
        th, k = gt % T, gt // T
        full, rem = n // T, n % T
        grank = th * full + np.minimum(th, rem) + k
        order = np.empty(len(gt), np.uint32)
        order[np.argsort(grank, kind="stable")] = np.arange(len(gt), dtype=np.uint32)
        return order

    def exchange_lists(self, r):
        """(send, recv): dicts peer -> local node indices.  recv[p]: r's ghost nodes owned by p; send[p]: r's owned
        nodes that are ghosts on p.  Both sides enumerate the shared bodies in global order, so buffers line up."""
        send, recv = {}, {}
        off_r = self.local_node_offsets(r)
        pos_r = {int(b): i for i, b in enumerate(self.local[r])}
        for p in range(self.world):
            if p == r:
                continue
            ghosts_here = [int(b) for b in self.local[r] if self.owner[b] == p]
            if ghosts_here:
                recv[p] = np.concatenate([off_r[pos_r[b]] + np.arange(self.nodes[b]) for b in ghosts_here])
            ghosts_there = [int(b) for b in self.local[p] if self.owner[b] == r]
            if ghosts_there:
                send[p] = np.concatenate([off_r[pos_r[b]] + np.arange(self.nodes[b]) for b in ghosts_there])
        return send, recv

    def missing_ghosts(self, lo_x, hi_x):
        """Bodies that, at the given CURRENT extents, should be ghosts of some rank but are not in its local set."""
        lo_x = np.asarray(lo_x, np.float64); hi_x = np.asarray(hi_x, np.float64)
        missing = []
        for r in range(self.world):
            own = self.owner == r
            if not own.any():
                continue
            e = (lo_x[own].min(), hi_x[own].max())
            # contacts need actual proximity: half the build-time halo is the alarm threshold
            need = (~own) & (hi_x + 0.5 * self.halo >= e[0]) & (lo_x - 0.5 * self.halo <= e[1])
            have = np.zeros(self.n_bodies, bool); have[self.local[r]] = True
            miss = np.flatnonzero(need & ~have)
            if len(miss):
                missing.append((r, miss))
        return missing


class HaloExchange:
    """Ghost <- owner copies of rows of [n, 4] float32 state tensors over torch.distributed point-to-point ops."""

    def __init__(self, plan, rank, device, dist=None, group=None):
        import torch
        self.torch, self.dist, self.group, self.rank = torch, dist, group, rank
        send, recv = plan.exchange_lists(rank)
        self.send = {p: torch.as_tensor(ix, dtype=torch.long, device=device) for p, ix in sorted(send.items())}
        self.recv = {p: torch.as_tensor(ix, dtype=torch.long, device=device) for p, ix in sorted(recv.items())}
        self.bytes_per_plane = 16 * (sum(len(v) for v in self.send.values()) + sum(len(v) for v in self.recv.values()))
        self.active = bool(self.send or self.recv)
        self._cache = {}

    def _plan_for(self, k, like):
        """Send / receive buffers and the P2P op list for k planes, built once (the exchange runs 11 times per tick:
        allocation and op construction would otherwise be most of its host time)."""
        torch, dist = self.torch, self.dist
        if k not in self._cache:
            sbuf = {p: torch.empty((k, len(ix), 4), dtype=like.dtype, device=like.device) for p, ix in self.send.items()}
            rbuf = {p: torch.empty((k, len(ix), 4), dtype=like.dtype, device=like.device) for p, ix in self.recv.items()}
            ops = [dist.P2POp(dist.isend, sbuf[p], p, group=self.group) for p in self.send]
            ops += [dist.P2POp(dist.irecv, rbuf[p], p, group=self.group) for p in self.recv]
            self._cache[k] = (sbuf, rbuf, ops)
        return self._cache[k]

    def __call__(self, planes):
        """planes: list of [n,4] tensors; ghost rows are overwritten with the owners' rows."""
        if not self.active or self.dist is None:
            return 0
        torch, dist = self.torch, self.dist
        k = len(planes)
        sbuf, rbuf, ops = self._plan_for(k, planes[0])
        for p, ix in self.send.items():
            for j, pl in enumerate(planes):
                torch.index_select(pl, 0, ix, out=sbuf[p][j])
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for p, ix in self.recv.items():
            for j, pl in enumerate(planes):
                pl.index_copy_(0, ix, rbuf[p][j])
        return k * self.bytes_per_plane


class _DevPlane:
    """Exposes a raw device pointer as a CUDA array so torch can alias it without copying."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n, 4), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


class SlabSolver:
    """PD solver of a multi-body scene sharded across the ranks of a torch.distributed group."""

    def __init__(self, specs, rank=0, world=1, halo=1.0, device=0, dist=None, group=None, snap=0.0, check_every=0,
                 native_halo=None, **options):
        """check_every = k > 0: every k ticks all ranks check the ghost layer and repartition if needed (collective).
        native_halo: exchange ghost rows inside the library over NCCL (default: whenever the group's backend is nccl)."""
        import torch
        from .solver import Solver
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world, self.halo, self.device, self.options = rank, world, float(halo), device, dict(options)
        self.snap = float(snap)
        self.check_every, self.ticks = int(check_every), 0
        if native_halo is None:
            native_halo = bool(dist is not None and world > 1 and dist.get_backend(group) == "nccl")
        self.native_halo = bool(native_halo)
        self.specs = list(specs)
        self._Solver = Solver
        self.repartitions = 0
        lo = np.array([s["lo"][0] for s in self.specs]); hi = np.array([s["hi"][0] for s in self.specs])
        self._build(lo, hi, state=None)

    # ---- construction / repartition ----
    def _build(self, lo_x, hi_x, state):
        torch = self.torch
        opts = dict(self.options)
        self.plan = SlabPlan([(s["nodes"], s["tris"], s["proj"]) for s in self.specs], lo_x, hi_x, self.world, self.halo,
                             thread_count=opts.get("threadCount", 8), snap=self.snap)
        r = self.rank
        old = getattr(self, "solver", None)
        if old is not None:      # repartition: the old communicator goes first (collective, like its creation)
            if self.native_halo:
                old.haloDestroy()
            old.close()
        self.solver = s = self._Solver(device=self.device, **opts)
        for b in self.plan.local[r]:
            apply_spec(s, self.specs[int(b)])
        self.l2g = self.plan.local_to_global_nodes(r)
        self.owned = self.plan.owned_node_mask(r).astype(bool)
        if state is not None:   # repartition: rest data comes from the factories, the state from the running scene
            s.setState(state[0][self.l2g], state[1][self.l2g], state[2][self.l2g])
        if self.world > 1:
            s.setTriangleOrder(self.plan.triangle_order(r))
            s.setOwnedNodes(self.owned.astype(np.uint8))
        s.setStream(torch.cuda.current_stream(self.device).cuda_stream)
        qp, pp, vp, n = s.deviceState()
        dev = torch.device("cuda", self.device)
        self.n = n
        if n:
            self.q = torch.as_tensor(_DevPlane(qp, n), device=dev)
            self.prev = torch.as_tensor(_DevPlane(pp, n), device=dev)
            self.vel = torch.as_tensor(_DevPlane(vp, n), device=dev)
        else:
            self.q = self.prev = self.vel = torch.zeros((0, 4), device=dev)
        self.halo_x = HaloExchange(self.plan, r, dev, self.dist if self.world > 1 else None, self.group)
        if self.native_halo and self.world > 1:
            # rank 0 draws the NCCL unique id; the caller's process group only carries these 128 bytes
            uid = [self._Solver.haloUniqueId() if r == 0 else None]
            self.dist.broadcast_object_list(uid, src=0, group=self.group)
            s.haloInit(r, self.world, uid[0])
            send, recv = self.plan.exchange_lists(r)
            s.haloSetLists(send, recv)
        loc = self.plan.local[r]
        own_b = self.plan.owner[loc] == r
        self.owned_static = int(self.plan.proj[loc][own_b].sum())
        self.body_of_node = torch.as_tensor(np.repeat(np.arange(len(loc)), self.plan.nodes[loc]), dtype=torch.long, device=dev)
        self.owned_bodies = loc[own_b]
        self.owned_body_local = torch.as_tensor(np.flatnonzero(own_b), dtype=torch.long, device=dev)
        self.halo_bytes = 0

    # ---- stepping ----
    def planes(self, which):
        return [getattr(self, w) for w in which]

    def tick(self):
        s, o = self.solver, self.solver.getOptions()
        if self.check_every and self.ticks and self.ticks % self.check_every == 0:
            self.check_halo(repartition=True)   # collective: the same tick on every rank
            s = self.solver
        self.ticks += 1
        if self.native_halo and self.world > 1:
            s.tick()                             # the library exchanges the halos itself (csrc/halo.cu)
            self.halo_bytes += s.stats().haloBytesLastTick
            return
        s.pdTickBegin()
        for _ in range(o.timeSubsteps):
            self.halo_bytes += self.halo_x([self.q, self.prev, self.vel])
            s.pdSubstepBegin()
            for _ in range(o.iterations):
                s.pdIteration()
                self.halo_bytes += self.halo_x([self.q])
            s.pdSubstepEnd()
        s.pdTickEnd()

    def projections_last_tick(self):
        """Owned static projections + contacts whose point node is owned, per iteration, times iterations."""
        o = self.solver.getOptions()
        nt, nf = self.solver.countOwnedContacts() if self.world > 1 else (self.solver.stats().triCollisions, self.solver.stats().staticCollisions)
        return o.timeSubsteps * o.iterations * (self.owned_static + nt + nf)

    # ---- global views (tests, repartition) ----
    def _gather(self, local_rows):
        """local_rows: [n_local, k] numpy of this rank; returns the [n_global, k] array assembled from every rank's owned rows."""
        n_global = int(self.plan.node_off[-1])
        out = np.zeros((n_global, local_rows.shape[1]), local_rows.dtype)
        mine = (self.l2g[self.owned], local_rows[self.owned])
        if self.world == 1 or self.dist is None:
            out[mine[0]] = mine[1]
            return out
        parts = [None] * self.world
        self.dist.all_gather_object(parts, mine, group=self.group)
        for ix, rows in parts:
            out[ix] = rows
        return out

    def gather_state(self):
        s = self.solver
        return self._gather(s.positions), self._gather(s.prevPositions), self._gather(s.velocities)

    def body_extents_x(self):
        """Current x extents of every body of the scene (owned bodies from the device, all-gathered)."""
        torch = self.torch
        nb_local = len(self.plan.local[self.rank])
        lo = torch.full((nb_local,), float("inf"), device=self.q.device).scatter_reduce(0, self.body_of_node, self.q[:, 0], "amin")
        hi = torch.full((nb_local,), float("-inf"), device=self.q.device).scatter_reduce(0, self.body_of_node, self.q[:, 0], "amax")
        mine = (self.owned_bodies, lo[self.owned_body_local].cpu().numpy(), hi[self.owned_body_local].cpu().numpy())
        glo = np.zeros(self.plan.n_bodies); ghi = np.zeros(self.plan.n_bodies)
        parts = [mine]
        if self.world > 1 and self.dist is not None:
            parts = [None] * self.world
            self.dist.all_gather_object(parts, mine, group=self.group)
        for ids, a, b in parts:
            glo[ids] = a; ghi[ids] = b
        return glo, ghi

    def check_halo(self, repartition=True):
        """True if every rank still holds every body its owned bodies could touch; otherwise repartitions
        (collective: call on all ranks at the same tick)."""
        lo, hi = self.body_extents_x()
        if not self.plan.missing_ghosts(lo, hi):
            return True
        if repartition:
            state = self.gather_state()
            self._build(lo, hi, state)
            self.repartitions += 1
        return False


def tick_lockstep(ranks):
    """In-process emulation of the distributed tick for tests on a single GPU: `ranks` are SlabSolver objects built
    with the same specs and world = len(ranks) (dist=None), all living on one device.  Every phase runs on all
    ranks, then ghost rows are copied from their owners with the same send/recv lists the NCCL path uses."""
    def exchange(which):
        for r in ranks:
            for peer, rix in r.halo_x.recv.items():
                six = ranks[peer].halo_x.send[r.rank]
                for w in which:
                    getattr(r, w).index_copy_(0, rix, getattr(ranks[peer], w).index_select(0, six))
    o = ranks[0].solver.getOptions()
    for r in ranks:
        r.solver.pdTickBegin()
    for _ in range(o.timeSubsteps):
        exchange(("q", "prev", "vel"))
        for r in ranks:
            r.solver.pdSubstepBegin()
        for _ in range(o.iterations):
            for r in ranks:
                r.solver.pdIteration()
            exchange(("q",))
        for r in ranks:
            r.solver.pdSubstepEnd()
    for r in ranks:
        r.solver.pdTickEnd()


def gather_lockstep(ranks):
    """Global (pos, prev, vel) assembled from every emulated rank's owned rows."""
    n_global = int(ranks[0].plan.node_off[-1])
    out = [np.zeros((n_global, 3), np.float32) for _ in range(3)]
    for r in ranks:
        for k, a in enumerate((r.solver.positions, r.solver.prevPositions, r.solver.velocities)):
            out[k][r.l2g[r.owned]] = a[r.owned]
    return out
