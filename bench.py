#!/usr/bin/env python
"""bench.py — constraint projections/s of the PD solver loop on scene S3 (BASELINE.json configs[2]):
1M-tet soft-body stack (20 834 x createTetBox), tet strain + volume constraints, 10 local/global
iterations per substep, one B200 per rank.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one Solver::tick (one substep of 10 PD iterations + collision detection/response).
`value`  : whole-job projections/s with all state resident in HBM (tick loop without host readback),
           CUDA-event timed, max over ranks.
`e2e`    : the same ticks through the host-facing API with HOST buffers: every step uploads the node
           state (pos/prev/vel) from pinned host memory, ticks, and reads the state back.
`roofline`: the fused tet strain+volume projection kernel, algorithmic bytes / measured time.
`cpu_baseline` / --impl reference: the unmodified reference (oracle/_ref) on the host cores, on a
           bounded sample of the same scene (6x6 columns x 21 layers = 756 bodies).
Multi-GPU (round 1): ranks run independent replicas of S3 (weak scaling, no data-path collective);
slab partitioning with NCCL halo exchange is not built yet.
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


def run_reference(args, steps, warmup, quiet=False):
    """Times the unmodified reference (oracle/_ref) on the host cores on the bounded sample."""
    from oracle import refapi
    from pies_b200 import scenes
    if not refapi.available():
        raise RuntimeError("oracle/_ref/libpies_ref.so missing (build it where /root/reference exists)")
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    r = refapi.RefSolver(threadCount=threads, **scenes.S3_OPTIONS)
    scenes.build_s3(r, **SAMPLE)
    per_iter = 2 * 48 * SAMPLE["bodies"]
    t0 = time.time()
    r.tick()  # first tick: stiffness assembly + first factorisation (quadratic in the reference)
    first = time.time() - t0
    for _ in range(max(0, warmup - 1)):
        r.tick()
    proj = 0
    t0 = time.time()
    for _ in range(steps):
        r.tick()
        proj += 10 * (per_iter + r.count("tri_collision") + r.count("static_collision"))
    dt = time.time() - t0
    return {"value": proj / dt, "seconds": dt, "first_tick_s": first, "cores": threads, "steps": steps,
            "sample": "%d of %d bodies (%dx%d columns x 21 layers), ticks %d..%d" % (
                SAMPLE["bodies"], FULL_BODIES, SAMPLE["nx"], SAMPLE["nz"], warmup, warmup + steps)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bodies", type=int, default=FULL_BODIES, help="debug: smaller S3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    config = {"workload": "S3: %d x createTetBox soft-body stack, PD, tet strain+volume, 10 iterations/substep" % args.bodies,
              "nodes": 27 * args.bodies, "tets": 48 * args.bodies, "static_projections_per_iteration": 96 * args.bodies,
              "iterations": 10, "substeps": 1, "parallelism": "replicas x%d" % world if world > 1 else "1 GPU",
              "l2": "working set > L2 (elements 80 MB + contributions 64 MB + CSR/preconditioner 120 MB per PD iteration)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        res = run_reference(args, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "constraint projections/s", "value": res["value"], "unit": "projections/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config, workload=config["workload"] + " [CPU sample: " + res["sample"] + "]"),
                "cpu_baseline": {"value": res["value"], "unit": "projections/s", "cores": res["cores"], "kind": "reference",
                                 "sample": res["sample"], "first_tick_s": res["first_tick_s"]},
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
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = run_reference(args, steps=12, warmup=min(warmup, 20), quiet=True)
        except Exception as e:  # the oracle is optional here; say so instead of inventing a number
            cpu = {"error": str(e)}

    s = pb.Solver(device=local_rank, **scenes.S3_OPTIONS)
    scenes.build_s3(s, args.bodies)
    n = len(s.getVertices())
    s.setTuning(profilePhases=True)
    for _ in range(warmup):
        s.tick()
    # snapshot so the device-resident and the end-to-end measurements replay the same ticks
    snap = [torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(3)]
    snap[0][:] = s.positions; snap[1][:] = s.prevPositions; snap[2][:] = s.velocities

    # ---- device-resident run: K ticks, state stays in HBM ----
    barrier()
    sampler = ClockSampler(local_rank)
    dev_ms = 0.0
    proj = launches = 0
    tet_ms = 0.0
    tet_launches = 0
    tet_bytes = 0.0
    phases = {"local": 0.0, "global": 0.0, "detect": 0.0, "contact": 0.0, "other": 0.0}
    t0 = time.time()
    for _ in range(args.steps):
        s.tick()  # no getVertices() here: the state never leaves HBM
        st = s.stats()
        dev_ms += st.msTick
        proj += st.projectionsLastTick
        launches += st.kernelLaunchesLastTick
        tet_ms += st.msTetKernel
        tet_launches += st.tetKernelLaunches
        phases["local"] += st.msLocal; phases["global"] += st.msGlobal; phases["detect"] += st.msDetect
        phases["contact"] += st.msContact; phases["other"] += st.msOther
    barrier()
    wall_ms = 1e3 * (time.time() - t0)
    clocks = sampler.stop()
    # msTick = CUDA events on the solver stream around each tick's kernels (the device time);
    # wall_ms = the host's view of the same loop.
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(proj), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = t.tolist()
    proj_all, launches_all = tot.tolist()

    # ---- end-to-end run: same ticks, host buffers in and out every step ----
    s.setState(snap[0], snap[1], snap[2])
    hp, hv, hq = snap
    e2e_proj = 0
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        s.setState(hp, hv, hq)          # H2D: 3 x 12 B per node from pinned memory
        s.tick()
        s.getVertices()                 # the reference-facing readback: D2H 12 B per node into the Vertex mirror
        hp[:] = s.positions; hv[:] = s.prevPositions; hq[:] = s.velocities  # D2H of the full state
        e2e_proj += s.stats().projectionsLastTick
    barrier()
    e2e_s = time.time() - t0
    e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ep = torch.tensor([float(e2e_proj)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        dist.all_reduce(ep, op=dist.ReduceOp.SUM)

    if rank == 0:
        peak, peak_kind = load_peaks()
        st = s.stats()
        per_launch_bytes = scenes.s3_algorithmic_bytes(0, 2 * 48 * args.bodies, 0, 0)  # tet-type projections only: 176 B each
        avg_tet_ms = tet_ms / max(1, tet_launches)
        achieved = per_launch_bytes / (avg_tet_ms * 1e-3) / 1e9 if avg_tet_ms > 0 else 0.0
        total_phase = sum(phases.values()) or 1.0
        line = {
            "metric": "constraint projections/s", "value": proj_all / (dev_ms_max * 1e-3), "unit": "projections/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "substeps_per_s": world * args.steps / (dev_ms_max * 1e-3),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": {"value": ep.item() / e.item(), "unit": "projections/s", "h2d_bytes_per_step": 36 * n,
                    "d2h_bytes_per_step": 48 * n, "ms_per_step": 1e3 * e.item() / args.steps},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_tet_elems (fused tet strain+volume projection)", "achieved": achieved,
                         "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": avg_tet_ms,
                         "launches_timed": tet_launches, "share_of_step": tet_ms / total_phase, "traffic": None},
            "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "pcg_iterations_last_tick": int(st.pcgIterationsLastTick),
            "contacts_last_tick": {"point_triangle": int(st.triCollisions), "floor": int(st.staticCollisions)},
        }
        traffic_file = os.path.join(ROOT, "profiles", "tet_kernel_traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    line["roofline"]["traffic"] = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                pass
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
