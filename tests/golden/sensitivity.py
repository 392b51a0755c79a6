"""Self-sensitivity of the UNMODIFIED reference on the scenes whose late-tick tolerance is wider than 1e-4 x diagonal.

For each scene the reference is run unperturbed and with every position perturbed by a uniform random offset of at most
`eps` = 1e-6 (about one fp32 ulp at these coordinates, 4..16) — once, at tick `at` ("once"), and before every tick
("every_tick", the model of an implementation whose arithmetic rounds differently in every step: ours solves the global
step with a CG instead of a Cholesky factorisation, so every tick's positions differ from the reference's in the last
bit or two).  The maximal position difference from the unperturbed run at the check ticks, over the seeds, is the noise
floor an equally valid fp32 evaluation cannot be expected to beat; tests/test_solver_gpu.py reads the committed numbers
(tests/golden/sensitivity.json) and holds our trajectories to max(1e-4 x diagonal, the every_tick floor).

Run where /root/reference exists:   make -C oracle ref && python tests/golden/sensitivity.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.refapi import RefSolver  # noqa: E402
from make_golden import quiet, two_box  # noqa: E402
from pies_b200 import scenes  # noqa: E402


def run(build, ticks, checks, eps=0.0, at=0, seed=0, every=False, **opts):
    with quiet():
        r = RefSolver(**opts)
        build(r)
    out = {}
    rng = np.random.default_rng(seed)
    for t in range(1, ticks + 1):
        if eps and (t == at + 1 or (every and t > at)):
            p = r.positions.copy()
            p += ((rng.random(p.shape) - 0.5) * 2 * eps).astype(np.float32)
            r.setState(p, None, None)
        r.tick()
        if t in checks:
            out[t] = r.positions.copy()
    return out


def eps_sweep(build, ticks, check, seeds=32, **opts):
    """How large a per-tick perturbation the scene tolerates before a threshold contact flips: for each eps the
    position difference at `check` in units of 1e-4 x diagonal (median, max, number of seeds above 1)."""
    base = run(build, ticks, [check], **opts)
    diag = float(np.linalg.norm(base[check].max(0) - base[check].min(0)))
    out = {"diag": diag, "tick": check, "seeds": seeds, "unit": "1e-4 x diagonal", "eps": {}}
    for eps in (1e-6, 2e-6, 4e-6, 1e-5):
        res = sorted(float(np.abs(base[check] - run(build, ticks, [check], eps=eps, at=0, seed=seed, every=True, **opts)[check]).max())
                     / (1e-4 * diag) for seed in range(seeds))
        out["eps"]["%g" % eps] = {"median": res[len(res) // 2], "max": res[-1], "above_1": sum(r > 1 for r in res),
                                  "largest": res[-6:]}
    return out


def floor(build, ticks, checks, at, **opts):
    base = run(build, ticks, checks, **opts)
    diag = float(np.linalg.norm(base[checks[0]].max(0) - base[checks[0]].min(0)))
    res = {"diag": diag, "eps": 1e-6, "first_perturbed_tick": at + 1, "seeds": 4, "once": {}, "every_tick": {}}
    for mode, every in (("once", False), ("every_tick", True)):
        for seed in range(4):
            b = run(build, ticks, checks, eps=1e-6, at=at, seed=seed, every=every, **opts)
            for t in checks:
                d = float(np.abs(base[t] - b[t]).max())
                res[mode][str(t)] = max(res[mode].get(str(t), 0.0), d)
    res["every_tick_over_1e-4_diag"] = {t: v / (1e-4 * diag) for t, v in res["every_tick"].items()}
    return res


def rebuild(r, meshes, off):
    """White-box rebuild of committed TetGen bodies translated by `off`: nodes, strain + volume constraints and boundary
    triangles exactly as Solver::addTriMeshVolume leaves them (PrimitiveUtilities.cpp:270-327) — at off = 0 the run is
    bit-identical to the fixture's (checked below)."""
    total_n = sum(len(p) for p, _, _ in meshes); total_t = sum(len(t) for _, t, _ in meshes)
    r.reserve(nodes=total_n, tets=total_t, vols=total_t, tris=sum(len(f) for _, _, f in meshes))
    base = 0
    for pts, tets, faces in meshes:
        for p in pts + np.asarray(off, np.float32):
            r.appendNode(p, (0, 0, 0), 0.1, 1.0)
        for t in tets:
            r.appendTet([int(x) + base for x in t], 1000.0, 0.8, 1.0)
        for t in tets:
            r.appendVolume([int(x) + base for x in t], 1000.0, 1.0, 1.0)
        for f in faces:
            r.appendTriangle(int(f[0]) + base, int(f[1]) + base, int(f[2]) + base)
        base += len(pts)


def translation_floor(fixture, mesh_keys, ticks, checks, offsets=((8.0, 0.0, 0.0), (0.0, 0.0, 8.0), (-4.0, 0.0, -4.0))):
    """The reference against ITSELF on the same bodies translated by about one body size (the floor does not move: only
    x and z).  Physics is translation invariant; the reference's fp32 sparse Cholesky of a stiffness matrix with sliver
    tets is not (its backward error acts on absolute coordinates), so this is the distance below which "agrees with the
    reference" stops meaning anything on that mesh."""
    g = np.load(os.path.join(HERE, fixture + ".npz"))
    meshes = [(g[a], g[b], g[c]) for a, b, c in mesh_keys]

    def run_off(off):
        r = RefSolver()
        rebuild(r, meshes, off)
        out = {}
        for t in range(1, ticks + 1):
            r.tick()
            if t in checks:
                out[t] = r.positions - np.asarray(off, np.float32)
        return out
    same = run_off((0.0, 0.0, 0.0))
    res = {"offsets": [list(o) for o in offsets], "rebuild_at_offset_0_max_abs_diff_from_fixture":
           max(float(np.abs(same[t] - g["pos%d" % t]).max()) for t in checks), "max_abs_diff": {}}
    diag = float(np.linalg.norm(g["pos%d" % checks[0]].max(0) - g["pos%d" % checks[0]].min(0)))
    res["diag"] = diag
    for off in offsets:
        o = run_off(off)
        for t in checks:
            d = float(np.abs(o[t] - g["pos%d" % t]).max())
            res["max_abs_diff"][str(t)] = max(res["max_abs_diff"].get(str(t), 0.0), d)
    res["over_1e-4_diag"] = {t: v / (1e-4 * diag) for t, v in res["max_abs_diff"].items()}
    return res


def s5_pair(r):
    scenes.add_tetgen_cube(r, n=24, origin=(0.0, 0.07, 0.0))
    scenes.add_tetgen_cube(r, n=24, origin=(1.3, 8.57, 0.9))


if __name__ == "__main__":
    out = {}
    path = os.path.join(HERE, "sensitivity.json")
    if len(sys.argv) > 1 and os.path.exists(path):   # e.g. `sensitivity.py s1_full s5_pair`: add / refresh only these
        out = json.load(open(path))
    todo = sys.argv[1:]
    jobs = {
        "two_box": lambda: floor(two_box, 40, [1, 10, 40], at=0, iterations=10),
        "tetgen_cube": lambda: floor(lambda r: scenes.add_tetgen_cube(r, n=6, origin=(0.0, 0.4, 0.0)), 60, [1, 10, 30, 60], at=0),
        "stack16": lambda: floor(lambda r: scenes.build_s3(r, bodies=16, nx=2, nz=2), 50, [1, 10, 40, 44, 50], at=0, **scenes.S3_OPTIONS),
        "s1_full": lambda: floor(lambda r: scenes.add_tetgen_cube(r, n=20, origin=(0.0, 3.07, 0.0)), 100, [1, 10, 60, 70, 80, 100], at=0),
        "s5_pair": lambda: floor(s5_pair, 60, [1, 10, 30, 40, 50, 60], at=0),
        "two_box_eps_sweep": lambda: eps_sweep(two_box, 40, 40, iterations=10),
        "s1_full_translation": lambda: translation_floor("s1_full", [("points", "tets", "faces")], 100, [1, 10, 60, 70, 80, 100]),
        "s5_pair_translation": lambda: translation_floor("s5_pair", [("points", "tets", "faces"), ("points2", "tets2", "faces2")],
                                                         60, [1, 10, 30, 40, 50, 60]),
        "tetgen_cube_translation": lambda: translation_floor("tetgen_cube", [("points", "tets", "faces")], 60, [1, 10, 30, 60]),
    }
    for name, job in jobs.items():
        if not todo or name in todo:
            out[name] = job()
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))
