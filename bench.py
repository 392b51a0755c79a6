#!/usr/bin/env python
"""bench.py — constraint projections/s of the PD solver loop on scene S3 (BASELINE.json configs[2]):
1M-tet soft-body stack (20 834 x createTetBox), tet strain + volume constraints, 10 local/global
iterations per substep, one B200 per rank.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one Solver::tick (one substep of 10 PD iterations + collision detection/response).
The scene is first rolled forward, untimed, to simulation tick --preroll (default 60: the columns have landed on the
floor and on each other, ~290 k point-triangle contacts), so whatever --warmup / --steps the caller passes the timed
window covers collisions and a real global solve; the free-fall regime (ticks 3..13, no contacts) is measured too and
reported under `free_fall`.  The reference arm rolls its sample forward to the same tick.
`value`  : whole-job projections/s with all state resident in HBM (tick loop without host readback),
           CUDA-event timed, max over ranks.
`e2e`    : the same ticks through the host-facing API with HOST buffers: every step uploads the node
           state (pos/prev/vel) from pinned host memory, ticks, and reads the state back.
`cpu_baseline` / --impl reference: the unmodified reference (oracle/_ref) on the host cores, on a
           bounded sample of the same scene (6x6 columns x 21 layers = 756 bodies).
`roofline`: the kernel with the largest share of the step (the mat-vec of the global solve), algorithmic
           bytes / CUDA-event duration of sampled launches; the other hot kernels and the
           local step + RHS pair of the north_star target are reported next to it.
Multi-GPU: weak scaling over x slabs (pies_b200/multigpu.py): `world` S3 stacks side by side, each rank owns one
           S3-sized slab plus a two-body ghost layer; ghost rows travel over NCCL point-to-point every substep and
           every PD iteration.  No other collective on the data path.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL_BODIES = 20834
SAMPLE = dict(bodies=756, nx=6, nz=6)  # reference-arm sample: same column height as the full scene


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, steps, warmup, quiet=False, threads=8):
    """Times the unmodified reference (oracle/_ref) on the host cores on the bounded sample, over the same tick window
    as the GPU arm (preroll + warmup untimed, then `steps`).  threads = SolverOptions::threadCount: 8 is the reference
    default (Solver.h:36) and what fixes the collision-list order on both arms; the reference is serial
    outside detection (Solver.cpp:269), so more threads only speed up the hash and the CCD."""
    from oracle import refapi
    from pies_b200 import scenes
    if not refapi.available():
        raise RuntimeError("oracle/_ref/libpies_ref.so missing (build it where /root/reference exists)")
    r = refapi.RefSolver(threadCount=threads, **scenes.S3_OPTIONS)
    scenes.build_s3(r, **SAMPLE)
    per_iter = 2 * 48 * SAMPLE["bodies"]
    t0 = time.time()
    r.tick()  # first tick: stiffness assembly + first factorisation (quadratic in the reference)
    first = time.time() - t0
    skip = max(0, args.preroll + warmup - 1)
    for _ in range(skip):
        r.tick()
    proj = 0
    t0 = time.time()
    for _ in range(steps):
        r.tick()
        proj += 10 * (per_iter + r.count("tri_collision") + r.count("static_collision"))
    dt = time.time() - t0
    return {"value": proj / dt, "seconds": dt, "first_tick_s": first, "cores": threads, "steps": steps,
            "contacts_last_tick": {"point_triangle": int(r.count("tri_collision")), "floor": int(r.count("static_collision"))},
            "sample": "%d of %d bodies (%dx%d columns x 21 layers), ticks %d..%d, threadCount %d" % (
                SAMPLE["bodies"], FULL_BODIES, SAMPLE["nx"], SAMPLE["nz"], skip + 1, skip + 1 + steps, threads)}


def run_other(args, local_rank):
    if os.environ.get("PIES_BENCH_VERBOSE"):
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["PIES_BENCH_VERBOSE"]), exit=False)
    """configs[1] (S2: 100 000-node distance chain, PBD, node-node collisions through the node hash) and configs[3]
    (S4: 15 625 shape-matching bodies of 4 x 8 x 8 particles = 4 M particles, hull triangles, 244 goal regions driven by a
    scripted transform, CCD + friction) on one GPU: the same JSON line, the byte model of SURVEY section 8(d) applied to
    the whole tick (72 B per distance projection, 44 B per shape / goal member, 136 B per point-triangle and 48 B per
    floor contact, 52 + 76 B per node and substep for the advection / velocity passes)."""
    import torch
    import pies_b200 as pb
    from pies_b200 import scenes
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    t0 = time.time()
    if args.workload == "s2":
        n_nodes = 100000 if args.bodies == FULL_BODIES else args.bodies
        s = pb.Solver(device=local_rank, **scenes.S2_OPTIONS)
        scenes.build_rope(s, n=n_nodes, shape="spiral", pinned=False)
        s.setTuning(pbdColourBatches=True)
        per_tick = None
        workload = ("S2: %d-node distance-constraint chain (flat coil, arms overlapping: node-node collisions from the first tick), PBD, "
                    "4 iterations, node-node response in colour batches (the reference's sequential order is a serial chain as long as "
                    "the visit list on this scene: the ordered executor, which the parity tests use, needs seconds per tick)") % n_nodes
        iterations, script = 4, None
    else:
        bodies = 15625 if args.bodies == FULL_BODIES else args.bodies
        per_side = max(1, int(round(bodies ** (1.0 / 3.0))))
        s = pb.Solver(device=local_rank, iterations=4)
        _, regions = scenes.build_s4(s, bodies=bodies, per_side=per_side, goal_bodies=min(244, bodies))
        workload = "S4: %d x createShapeMatchingBox(4, 8, 8) = %d particles, one 256-particle cluster per body, hull triangles, %d goal regions, PD, 4 iterations" % (
            bodies, 256 * bodies, min(244, bodies))
        iterations = 4
        script = (lambda t: s.updateFixedRegions(scenes.s4_region_script(regions, t))) if len(regions) else None
    s.setStream(stream.cuda_stream)
    s.setTuning(profilePhases=True)   # keeps the flags set above
    build_s = time.time() - t0
    print("[bench %s] scene built in %.1f s" % (args.workload, build_s), file=sys.stderr, flush=True)
    n = len(s.getVertices())
    tick_no = [0]

    def tick():
        tick_no[0] += 1
        if script is not None:
            script(tick_no[0])
        s.tick()
    # The reference's PBD is unstable on chains (its distance projection moves only node 0 of a link: a self-overlapping rope
    # diverges within a handful of ticks IN THE REFERENCE, tests/test_pbd_gpu.py; at 100 k nodes our run follows it: tick 6
    # already leaves the regime the coil was built for), so the rope is put back to its initial state every four ticks; the
    # 3.6 MB upload is inside the timed region.
    if args.workload == "s2":
        reset = [torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(3)]
        reset[0][:] = s.positions; reset[1][:] = s.positions; reset[2][:] = 0.0
        plain_tick = tick

        def tick():
            if tick_no[0] % 4 == 0 and tick_no[0]:
                s.setState(reset[0], reset[1], reset[2])
            plain_tick()
    for k in range(args.preroll + max(args.warmup, 3)):
        tw = time.time()
        tick()
        print("[bench %s] roll tick %d: %.1f ms" % (args.workload, k + 1, 1e3 * (time.time() - tw)), file=sys.stderr, flush=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    proj = launches = 0
    contacts = 0
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    ev0.record(stream)
    for k in range(args.steps):
        tw = time.time()
        tick()
        st = s.stats()
        proj += st.projectionsLastTick; launches += st.kernelLaunchesLastTick
        contacts += st.collisionProjections
        if os.environ.get("PIES_BENCH_VERBOSE"):
            print("[bench %s] timed tick %d: %.1f ms" % (args.workload, k + 1, 1e3 * (time.time() - tw)), file=sys.stderr, flush=True)
    ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    # end to end: the reference-facing readback (getVertices) after every tick
    t0 = time.time()
    e2e_proj = 0
    for _ in range(args.steps):
        tick()
        s.getVertices(copy=False)
        e2e_proj += s.stats().projectionsLastTick
    torch.cuda.synchronize()
    e2e_s = time.time() - t0
    st = s.stats()
    peak, peak_kind = load_peaks()
    if args.workload == "s2":
        per_proj = 72.0
        static_bytes = iterations * per_proj * (n - 1) + (52 + 76) * n
        contact_bytes = 72.0 * (contacts / args.steps)     # a node-node visit moves two nodes like a distance projection
    else:
        members = int(st.staticProjections)
        static_bytes = iterations * (44.0 * members + 40.0 * n) + (52 + 76) * n
        contact_bytes = iterations * (136.0 * st.triCollisions + 48.0 * st.staticCollisions)
    alg = static_bytes + contact_bytes
    ms = dev_ms / args.steps
    line = {"metric": "constraint projections/s", "value": proj / (dev_ms * 1e-3), "unit": "projections/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.workload == "s2" else "f32 (shape matching in f64 like the reference)",
            "data": "synthetic", "config": {"workload": workload, "nodes": n, "iterations": iterations, "substeps": 1,
                                            "parallelism": "1 GPU", "preroll_ticks": args.preroll, "scene_build_s": build_s},
            "substeps_per_s": args.steps / (dev_ms * 1e-3),
            "e2e": {"value": e2e_proj / e2e_s, "unit": "projections/s", "h2d_bytes_per_step": int(64 * len(regions)) if args.workload == "s4" else 0,
                    "d2h_bytes_per_step": int(36 * n), "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole tick (all kernels of the " + ("PBD" if args.workload == "s2" else "PD") + " loop)",
                         "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(alg), "avg_launch_ms": ms,
                         "launches_timed": args.steps, "share_of_step": 1.0, "traffic": None},
            "phase_ms_per_step": {"local": st.msLocal, "global": st.msGlobal, "detect": st.msDetect, "contact": st.msContact, "other": st.msOther},
            "contacts_last_tick": {"point_triangle": int(st.triCollisions), "floor": int(st.staticCollisions),
                                   "collision_projections": int(st.collisionProjections)},
            "cpu_baseline": None}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bodies", type=int, default=FULL_BODIES, help="debug: smaller S3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", type=float, default=4.0, help="ghost layer width along x (4.0 = two S3 columns)")
    ap.add_argument("--preroll", type=int, default=None, help="untimed ticks before warm-up (default: 60 for s3, 70 for s5 = the contact regime)")
    ap.add_argument("--workload", default="s3", choices=["s3", "s5", "s2", "s4"],
                    help="s3 (default, the metric's configuration; weak scaling for --gpus > 1); s5: BASELINE configs[4], "
                         "512 TetGen bodies of 16.5 k tets, STRONG scaling: the same scene cut into --gpus x-slabs; "
                         "s2: configs[1], the 100 k-node PBD rope; s4: configs[3], 4 M shape-matching particles (1 GPU each)")
    ap.add_argument("--big-tier", action="store_true", help="island tier 3 on (one 1024-thread CTA per island of up to 7 168 nodes)")
    args = ap.parse_args()
    if args.preroll is None:
        args.preroll = 70 if args.workload == "s5" else 60
    if args.workload == "s5" and args.bodies == FULL_BODIES:
        args.bodies = 512
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    config = {"workload": "S3: %d x createTetBox soft-body stack, PD, tet strain+volume, 10 iterations/substep" % args.bodies,
              "nodes": 27 * args.bodies, "tets": 48 * args.bodies, "static_projections_per_iteration": 96 * args.bodies,
              "iterations": 10, "substeps": 1, "parallelism": "replicas x%d" % world if world > 1 else "1 GPU",
              "l2": "working set > L2 (elements 80 MB + contributions 64 MB + CSR/preconditioner 120 MB per PD iteration)"}

    if args.workload in ("s2", "s4"):
        if args.impl == "reference" or world > 1:
            if rank == 0:
                print(json.dumps({"impl": args.impl, "unavailable": "workloads s2 / s4 run our arm on one GPU only"}))
            return 0
        if args.preroll is None or args.preroll in (60, 70):
            args.preroll = 0 if args.workload == "s2" else 20
        return run_other(args, local_rank)
    if args.workload == "s5":
        config = {"workload": "S5: %d x TetGen cube body (16 546 tets, 4 518 nodes, 6 912 boundary triangles each; one committed mesh "
                              "replicated), PD, tet strain+volume, 10 iterations/substep, dropped onto the floor and each other" % args.bodies,
                  "nodes": 4518 * args.bodies, "tets": 16546 * args.bodies, "static_projections_per_iteration": 2 * 16546 * args.bodies,
                  "iterations": 10, "substeps": 1, "parallelism": "1 GPU", "l2": "working set > L2"}
    if args.impl == "reference":
        if args.workload != "s3":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the metric's configuration (S3) only"}))
            return 0
        if rank != 0:
            return 0
        res = run_reference(args, args.steps, args.warmup, threads=8)
        cores = os.cpu_count() or 1
        extra = None
        if cores > 8:  # the same window with every host core in the detection phase (SURVEY section 8d asks for both)
            try:
                e = run_reference(args, min(args.steps, 5), args.warmup, threads=min(cores, 64))
                extra = {"value": e["value"], "cores": e["cores"], "steps": e["steps"]}
            except Exception as ex:
                extra = {"error": str(ex)}
        line = {"impl": "reference", "metric": "constraint projections/s", "value": res["value"], "unit": "projections/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config, workload=config["workload"] + " [CPU sample: " + res["sample"] + "]"),
                "cpu_baseline": {"value": res["value"], "unit": "projections/s", "cores": res["cores"], "kind": "reference",
                                 "sample": res["sample"], "first_tick_s": res["first_tick_s"], "host_cores": cores,
                                 "all_cores_run": extra},
                "contacts_last_tick": res["contacts_last_tick"],
                "e2e": {"value": res["value"], "unit": "projections/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import pies_b200 as pb
    from pies_b200 import scenes
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "s3":
        try:
            cpu = run_reference(args, steps=10, warmup=min(warmup, 5), quiet=True, threads=8)
        except Exception as e:  # the oracle is optional here; say so instead of inventing a number
            cpu = {"error": str(e)}

    # One dedicated CUDA stream per rank: the solver launches on it, NCCL orders its transfers against it, and
    # the timed region is bracketed by events recorded on it.
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    drv = None
    scaling = "weak"
    if args.workload == "s5":
        scaling = "strong"
        if world == 1:
            s = pb.Solver(device=local_rank, **scenes.S3_OPTIONS)
            scenes.build_s5_replicated(s, args.bodies)
            s.setStream(stream.cuda_stream)
            tick = s.tick
        else:
            # Strong scaling: the SAME scene on every world size, cut into x slabs of equal constraint count.  Bodies are 8
            # wide on a pitch of 10, so a ghost layer of 1.0 is empty while the bodies fall straight down; every rank checks
            # the layer before the timed window (and repartitions with ghosts if bodies have drifted into reach).
            from pies_b200 import multigpu
            pts, tets, faces = scenes.cube24_mesh()
            specs = [multigpu.tetmesh(pts + o, tets, faces) for o in scenes.s5_origins(args.bodies)]
            drv = multigpu.SlabSolver(specs, rank=rank, world=world, halo=1.0, device=local_rank, dist=dist, snap=1.0,
                                      **scenes.S3_OPTIONS)
            s = drv.solver
            tick = drv.tick
            config["parallelism"] = ("x-slabs x%d of the same %d-body scene (strong scaling), ghost layer 1.0, NCCL halo exchange from inside "
                                     "pies_b200_tick when bodies of different slabs come within reach" % (world, args.bodies))
    elif world == 1:
        s = pb.Solver(device=local_rank, **scenes.S3_OPTIONS)
        scenes.build_s3(s, args.bodies)
        s.setStream(stream.cuda_stream)
        tick = s.tick
        owned_static = 96 * args.bodies
    else:
        # Weak scaling: the scene is `world` S3 stacks side by side along x (32*world x 32 columns, same column
        # height), cut into x slabs of equal constraint count; each rank simulates its slab plus a two-body-deep
        # ghost layer and exchanges ghost rows with its neighbours every substep and every PD iteration.
        from pies_b200 import multigpu
        trans = scenes.s3_translations(args.bodies * world, nx=32 * world)
        specs = [multigpu.tetbox(t) for t in trans]
        drv = multigpu.SlabSolver(specs, rank=rank, world=world, halo=args.halo, device=local_rank, dist=dist, snap=0.5,
                                  **scenes.S3_OPTIONS)
        s = drv.solver
        tick = drv.tick
        owned_static = drv.owned_static
        config["parallelism"] = ("x-slabs x%d, %d-body S3 per rank, halo %.1f (two bodies deep), NCCL halo exchange per substep and per PD "
                                 "iteration from inside pies_b200_tick (ncclSend/ncclRecv, csrc/halo.cu)" % (world, args.bodies, args.halo))
        config["nodes"] *= world; config["tets"] *= world; config["static_projections_per_iteration"] *= world
    n = len(s.getVertices())
    s.setTuning(profilePhases=True, islandBigTier=bool(args.big_tier))

    def projections():
        if drv is not None:
            return drv.projections_last_tick()
        return s.stats().projectionsLastTick

    # free-fall regime first (extra key): ticks 3..13, no contacts, the global solve is one exact block solve per body
    free_fall = None
    done_ticks = 0
    if args.preroll >= 13:
        for _ in range(3):
            tick()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record(stream)
        fproj = 0
        fph = {"local": 0.0, "global": 0.0, "detect": 0.0, "contact": 0.0, "other": 0.0}
        for _ in range(10):
            tick()
            fst = s.stats()
            fproj += projections()
            fph["local"] += fst.msLocal; fph["global"] += fst.msGlobal; fph["detect"] += fst.msDetect
            fph["contact"] += fst.msContact; fph["other"] += fst.msOther
        f1.record(stream)
        barrier()
        fms = f0.elapsed_time(f1)
        free_fall = {"ticks": "3..13", "ms_per_step": fms / 10, "projections_per_s_rank0": fproj / (fms * 1e-3),
                     "phase_ms_per_step": {k: v / 10 for k, v in fph.items()}}
        done_ticks = 13
    for _ in range(max(0, args.preroll - done_ticks)):
        tick()
    for _ in range(warmup):
        tick()
    repartitioned = False
    if drv is not None:
        # untimed: every rank checks that its ghost layer still holds every body its own bodies can touch and the scene is
        # repartitioned if not, so the timed window below never runs on a stale partition
        repartitioned = not drv.check_halo(repartition=True)
        s = drv.solver
        s.setTuning(profilePhases=True, islandBigTier=bool(args.big_tier))
        if repartitioned:
            for _ in range(3):
                tick()
        n = len(s.getVertices())
    # snapshot so the device-resident and the end-to-end measurements replay the same ticks
    snap = [torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(3)]
    snap[0][:] = s.positions; snap[1][:] = s.prevPositions; snap[2][:] = s.velocities

    # ---- device-resident run: K ticks, state stays in HBM ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern = {k: [0.0, 0] for k in ("tet", "gather", "spmv", "update", "island")}
    island_row_iters = 0
    island_inv_floats = 0
    cap_hits = 0
    phases = {"local": 0.0, "global": 0.0, "detect": 0.0, "contact": 0.0, "other": 0.0, "halo": 0.0}
    proj = launches = pcg_iters = 0
    if drv is not None:
        drv.halo_bytes = 0
    barrier()
    sampler = ClockSampler(local_rank)
    t0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        tick()  # no getVertices() here: the state never leaves HBM
        st = s.stats()
        proj += projections()
        launches += st.kernelLaunchesLastTick
        pcg_iters += st.pcgIterationsLastTick
        kern["tet"][0] += st.msTetKernel; kern["tet"][1] += st.tetKernelLaunches
        kern["gather"][0] += st.msGatherKernel; kern["gather"][1] += st.gatherKernelLaunches
        kern["spmv"][0] += st.msSpmvKernel; kern["spmv"][1] += st.spmvKernelLaunches
        kern["update"][0] += st.msUpdateKernel; kern["update"][1] += st.updateKernelLaunches
        kern["island"][0] += st.msIslandKernels; kern["island"][1] += st.islandKernelLaunches
        island_row_iters += st.pcgIslandRowIterations
        island_inv_floats += int(st.islandInverseFloats)
        cap_hits += st.pcgCapHits
        phases["local"] += st.msLocal; phases["global"] += st.msGlobal; phases["detect"] += st.msDetect
        phases["contact"] += st.msContact; phases["other"] += st.msOther; phases["halo"] += st.msHalo
    ev1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.time() - t0)
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)   # CUDA events on the launching stream around the K ticks
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(proj), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = t.tolist()
    proj_all, launches_all = tot.tolist()
    halo_bytes = drv.halo_bytes if drv is not None else 0
    halo_ok = drv.check_halo(repartition=False) if drv is not None else True   # every contact partner still inside the ghost layer

    # ---- end-to-end run: same ticks, host buffers in and out every step ----
    s.getVertices(copy=False)        # untimed: the first call page-locks the host mirror (once per topology)
    s.setState(snap[0], snap[1], snap[2])
    hp, hv, hq = snap
    e2e_proj = 0
    e2e_parts = {"h2d_state": 0.0, "tick": 0.0, "d2h_vertices": 0.0, "d2h_state": 0.0}
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        ta = time.time()
        s.setState(hp, hv, hq)          # H2D: 3 x 12 B per node from pinned memory
        tb = time.time()
        tick()
        tc = time.time()
        s.getVertices(copy=False)       # the reference-facing readback: D2H 12 B per node into the Vertex mirror
        td = time.time()
        s.getState(hp, hv, hq)          # D2H of the full state into the same host buffers
        te = time.time()
        e2e_parts["h2d_state"] += tb - ta; e2e_parts["tick"] += tc - tb
        e2e_parts["d2h_vertices"] += td - tc; e2e_parts["d2h_state"] += te - td
        e2e_proj += projections()
    barrier()
    e2e_s = time.time() - t0
    e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ep = torch.tensor([float(e2e_proj), float(n)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        dist.all_reduce(ep, op=dist.ReduceOp.SUM)
    e2e_proj_all, n_all = ep.tolist()

    if rank == 0:
        peak, peak_kind = load_peaks()
        st = s.stats()
        nnz = int(st.systemNonZeros)
        total_phase = sum(phases.values()) or 1.0
        # Algorithmic bytes per launch (the split of the 176 B per tet-type projection between the
        # kernel that writes the contributions and the gather that re-reads them is stated in DESIGN.md section 4).
        local_proj = int(st.staticProjections)
        # One island solve has to move, once: the rows of S + C_t (8 B per non-zero), x in and out, b and the per-row
        # tables (76 B per node), and the inverses its lists apply (packed block inverse of every single-block island,
        # dense inverse of every island of the dense lists, packed block inverses of the CG lists) — everything else
        # (refinement rounds, CG iterations) runs out of registers / shared memory.  Average over the timed ticks.
        alg = {"tet": 112 * local_proj, "gather": 64 * local_proj + 40 * n, "spmv": 8 * nnz + 64 * n, "update": 72 * n,
               "island": int(8 * nnz + 76 * n + 4.0 * island_inv_floats / max(1, args.steps))}
        names = {"tet": "k_tet_elems (fused tet strain+volume projection: ids, Qinv, parameters in; 4 contributions out)",
                 "gather": "k_gather_rhs / k_gather_rhs_contacts (CSR gather of the contributions, with contacts also the collision and floor terms, into the right-hand side)",
                 "spmv": "k_pcg_spmv (A z over SELL-32 windows staged by cp.async, with p / Ap recurrences)",
                 "update": "k_pcg_update (x, r update + packed block-Jacobi apply; the preconditioner stream is not in the SURVEY model)",
                 "island": "island solve: k_island_direct + k_island_dense<*> + k_island_pcg<*> side by side (one global solve; bytes "
                           "= what a solve has to read once: 8 B per non-zero of S, 76 B per node of vectors and row tables, and "
                           "the block / dense inverses its lists apply; rounds and iterations run on chip)"}
        traffic = {}
        traffic_file = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    traffic = json.load(f)
            except Exception:
                traffic = {}

        def roof(k):
            ms, cnt = kern[k]
            avg = ms / max(1, cnt)
            ach = alg[k] / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
            launches_per_step = 10.0
            if k in ("spmv", "update"):
                launches_per_step = (pcg_iters / args.steps) if kern[k][1] else 0.0
            return {"bound": "hbm", "kernel": names[k], "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                    "frac": ach / peak, "algorithmic_bytes_per_launch": alg[k], "avg_launch_ms": avg, "launches_timed": cnt,
                    "share_of_step": avg * launches_per_step / (dev_ms / args.steps),
                    "traffic": traffic.get(k, {}).get("dram_bytes_per_launch")}
        roofs = {k: roof(k) for k in kern}
        dominant = max(roofs, key=lambda k: roofs[k]["share_of_step"])
        # the local step + RHS assembly as the north_star words its target: 176 B per projection + 40 B per node
        lr_ms = roofs["tet"]["avg_launch_ms"] + roofs["gather"]["avg_launch_ms"]
        lr_bytes = alg["tet"] + alg["gather"]
        line = {
            "metric": "constraint projections/s", "value": proj_all / (dev_ms_max * 1e-3), "unit": "projections/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "substeps_per_s": args.steps / (dev_ms_max * 1e-3),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": {"value": e2e_proj_all / e.item(), "unit": "projections/s", "h2d_bytes_per_step": int(36 * n_all),
                    "d2h_bytes_per_step": int(48 * n_all), "ms_per_step": 1e3 * e.item() / args.steps,
                    "breakdown_ms_rank0": {k: 1e3 * v / args.steps for k, v in e2e_parts.items()}},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": roofs[dominant],
            "roofline_other_kernels": {k: v for k, v in roofs.items() if k != dominant},
            "roofline_local_step_plus_rhs": {"achieved": lr_bytes / (lr_ms * 1e-3) / 1e9 if lr_ms > 0 else 0.0, "peak": peak, "unit": "GB/s",
                                             "frac": (lr_bytes / (lr_ms * 1e-3) / 1e9 / peak) if lr_ms > 0 else 0.0,
                                             "algorithmic_bytes_per_iteration": lr_bytes, "ms_per_iteration": lr_ms,
                                             "projections_per_s": local_proj / (lr_ms * 1e-3) if lr_ms > 0 else 0.0},
            "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "pcg_iterations_per_step": pcg_iters / args.steps,
            "pcg_cap_hits": int(cap_hits),
            "islands_last_tick": {"warp": int(st.islandsTier[0]), "dense_cta128_cta320": int(st.islandsTier[1]), "cta512": int(st.islandsTier[2]),
                                  "cta1024": int(st.islandsTier[3]), "grid_wide": int(st.islandsGlobal),
                                  "grid_wide_nodes": int(st.islandNodesGlobal)},
            "tick_window": [args.preroll + warmup, args.preroll + warmup + args.steps],
            "free_fall": free_fall,
            "contacts_last_tick": {"point_triangle": int(st.triCollisions), "floor": int(st.staticCollisions)},
        }
        if drv is not None:
            line["halo"] = {"bytes_per_step_rank0": halo_bytes / args.steps, "ghost_layer_still_sufficient": bool(halo_ok), "ghost_nodes_rank0": int((~drv.owned).sum()),
                            "owned_nodes_rank0": int(drv.owned.sum()), "exchanges_per_step": int(st.haloExchangesLastTick),
                            "inside_the_library": bool(drv.native_halo), "repartitioned_before_timing": bool(repartitioned)}
        if cpu is not None:
            if "error" in cpu:
                line["cpu_baseline"] = {"value": None, "unit": "projections/s", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "unavailable: " + cpu["error"]}
            else:
                line["cpu_baseline"] = {"value": cpu["value"], "unit": "projections/s", "cores": cpu["cores"], "kind": "reference",
                                        "sample": cpu["sample"], "first_tick_s": cpu["first_tick_s"]}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
