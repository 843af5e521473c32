#!/usr/bin/env python
"""Time the reference's own Python/NumPy force routines (CPU baseline item 1) in THIS container -
the reference tree does not travel to the GPU box, so these numbers are measured here and committed:

    python oracle/time_reference_python.py        ->  profiles/r1_reference_python_cpu.json

md_lj_module.force at N = 256, md_lj_ll_module.force at N = 864, 4,000 and 32,000 (median of 5 calls), and the loop of
md_nve_lj.py (force + hessian every step, lifted from the driver as in make_golden_dynamics.py) at N = 256.
"""
import json
import os
import platform
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_dynamics import REF, ref_module  # noqa: E402
from examples_b200.lattice import fcc_positions, perturb, ran_velocities  # noqa: E402


def med(fn, k=5):
    ts = []
    for _ in range(k):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def main():
    lj, ll = ref_module("md_lj_module"), ref_module("md_lj_ll_module")
    out = {"where": "build container (no GPU), %d cores visible, NumPy %s single-threaded, Python %s"
                    % (os.cpu_count(), np.__version__, platform.python_version()), "rows": []}
    for name, mod, nc in (("md_lj_module.force", lj, 4), ("md_lj_ll_module.force", ll, 6),
                          ("md_lj_ll_module.force", ll, 10), ("md_lj_ll_module.force", ll, 20)):
        n, box, r = fcc_positions(nc, 0.75)
        r = perturb(r, box, 0.05)
        t = med(lambda: mod.force(box, 2.5, r), 5 if nc < 20 else 3)
        out["rows"].append({"what": name, "n": n, "sec_per_call": t, "atoms_per_s": n / t})
        print(out["rows"][-1], flush=True)
    # md_nve_lj.py loop body at N = 256: kick, drift, force, kick + calc_variables' hessian (md_nve_lj.py:46,182-190)
    n, box, r = fcc_positions(4, 0.75)
    r = perturb(r, box, 0.05)
    v = ran_velocities(n, 1.0, seed=1)
    dt = 0.005
    total, f = lj.force(box, 2.5, r)
    t0 = time.perf_counter()
    for _ in range(100):
        v = v + 0.5 * dt * f
        r = r + dt * v / box
        r = r - np.rint(r)
        total, f = lj.force(box, 2.5, r)
        v = v + 0.5 * dt * f
        lj.hessian(box, 2.5, r, f)
    t = (time.perf_counter() - t0) / 100
    out["rows"].append({"what": "md_nve_lj.py step (force + hessian)", "n": n, "sec_per_call": t,
                        "atom_steps_per_s": n / t})
    print(out["rows"][-1], flush=True)
    path = os.path.join(ROOT, "profiles", "r1_reference_python_cpu.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
