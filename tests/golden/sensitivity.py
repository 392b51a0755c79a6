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


if __name__ == "__main__":
    out = {}
    out["two_box"] = floor(two_box, 40, [1, 10, 40], at=0, iterations=10)
    out["tetgen_cube"] = floor(lambda r: scenes.add_tetgen_cube(r, n=6, origin=(0.0, 0.4, 0.0)), 60, [1, 10, 30, 60], at=0)
    out["stack16"] = floor(lambda r: scenes.build_s3(r, bodies=16, nx=2, nz=2), 50, [1, 10, 40, 44, 50], at=0, **scenes.S3_OPTIONS)
    with open(os.path.join(HERE, "sensitivity.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))
