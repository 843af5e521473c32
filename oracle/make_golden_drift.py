#!/usr/bin/env python
"""Generate tests/golden/reference_drift.npz: a 10^4-step conserved-energy trace from the reference's own
velocity-Verlet loop body (md_nve_lj.py:178-192, lifted from the parsed driver source and executed unmodified) with
the reference md_lj_module.force, N=256, rho=0.75, T~1.0, r_cut=2.5, dt=0.005 (the driver's defaults).
Run in the build container only (needs /root/reference); takes ~5 minutes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_dynamics import lift, ref_module  # noqa: E402

from examples_b200.lattice import fcc_positions, perturb, ran_velocities  # noqa: E402


def main():
    lj = ref_module("md_lj_module")
    _, step = lift("md_nve_lj.py", [])
    n, box, r = fcc_positions(4, 0.75)
    r = perturb(r, box, 0.05)
    v = ran_velocities(n, 2.0, seed=11)  # melts to T ~ 1
    ns = dict(np=np, n=n, box=box, r_cut=2.5, dt=0.005, r=r.copy(), v=v.copy(), force=lj.force)
    ns["total"], ns["f"] = lj.force(box, 2.5, ns["r"])
    nsteps = 10000
    rec = np.empty((nsteps, 3))
    for s in range(nsteps):
        exec(step, ns)
        rec[s] = [ns["total"].pot, np.sum(ns["v"] ** 2), ns["total"].cut]
        if s % 1000 == 0:
            print(s, rec[s], flush=True)
    path = os.path.join(ROOT, "tests", "golden", "reference_drift.npz")
    np.savez_compressed(path, box=np.float64(box), r0=r, v0=v, rec=rec)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
