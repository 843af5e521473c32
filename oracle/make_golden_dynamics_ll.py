#!/usr/bin/env python
"""Generate tests/golden/reference_dynamics_ll.npz: the BAOAB and Nose-Hoover-chain step loops of the reference
(bd_nvt_lj.py, md_nvt_lj.py), executed from their parsed source as in make_golden_dynamics.py, but with the force
routine of md_lj_ll_module.py (the GUIDE's "edit the import" variant, python_examples/GUIDE.md:288-289) at N = 864
(sc = 4): the thermostatted loops on the link-cell route.  Run in the build container only:

    python oracle/make_golden_dynamics_ll.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_dynamics import lift, ref_module  # noqa: E402
from examples_b200.lattice import fcc_positions, perturb, ran_velocities  # noqa: E402


def main():
    out = {}
    ll = ref_module("md_lj_ll_module")
    dt, nsteps = 0.005, 24
    n, box, r0 = fcc_positions(6, 0.75)
    r0 = perturb(r0, box, 0.05)

    fcode, step = lift("bd_nvt_lj.py", ["a_propagator", "b_propagator", "o_propagator"])
    v0 = ran_velocities(n, 1.0, seed=17)
    ns = dict(np=np, n=n, box=box, r_cut=2.5, r=r0.copy(), v=v0.copy(), dt=dt, temperature=1.0, gamma=1.0,
              force=ll.force)
    exec(fcode, ns)
    ns["total"], ns["f"] = ll.force(box, 2.5, ns["r"])
    drawn, real_randn = [], np.random.randn

    def recording_randn(*shape):
        x = real_randn(*shape)
        drawn.append(x.copy())
        return x

    np.random.seed(20240612)
    np.random.randn = recording_randn
    try:
        rec = np.empty((nsteps, 6))
        for s in range(nsteps):
            exec(step, ns)
            t = ns["total"]
            rec[s] = [t.cut, t.pot, t.vir, t.lap, np.sum(ns["v"] ** 2), np.sum(ns["f"] ** 2)]
    finally:
        np.random.randn = real_randn
    out.update(box=np.float64(box), r0=r0, bd_v0=v0, bd_noise=np.array(drawn), bd_rec=rec, bd_r1=ns["r"],
               bd_v1=ns["v"])
    print("bd_nvt_lj (cells) done", rec[-1], flush=True)

    fcode, step = lift("md_nvt_lj.py", ["u1_propagator", "u2_propagator", "u3_propagator", "u4_propagator"])
    v0 = ran_velocities(n, 1.0, seed=18)
    m, temperature, tau = 3, 1.0, 2.0
    g = 3 * (n - 1)
    q = np.full(m, temperature * tau ** 2, dtype=np.float64)
    q[0] = g * temperature * tau ** 2
    p_eta = np.random.default_rng(6).standard_normal(m) * np.sqrt(temperature) * np.sqrt(q)
    ns = dict(np=np, n=n, box=box, r_cut=2.5, r=r0.copy(), v=v0.copy(), dt=dt, temperature=temperature, m=m, g=g,
              q=q.copy(), eta=np.zeros(m), p_eta=p_eta.copy(), force=ll.force)
    exec(fcode, ns)
    ns["total"], ns["f"] = ll.force(box, 2.5, ns["r"])
    rec = np.empty((nsteps, 7))
    for s in range(nsteps):
        exec(step, ns)
        t = ns["total"]
        ext = np.sum(0.5 * ns["p_eta"] ** 2 / ns["q"]) + temperature * (g * ns["eta"][0] + np.sum(ns["eta"][1:]))
        rec[s] = [t.cut, t.pot, t.vir, t.lap, np.sum(ns["v"] ** 2), np.sum(ns["f"] ** 2), ext]
    out.update(nhc_v0=v0, nhc_q=q, nhc_p_eta0=p_eta, nhc_rec=rec, nhc_r1=ns["r"], nhc_v1=ns["v"],
               nhc_eta1=ns["eta"], nhc_p_eta1=ns["p_eta"], nhc_g=np.float64(g))
    print("md_nvt_lj (cells) done", rec[-1], flush=True)
    path = os.path.join(ROOT, "tests", "golden", "reference_dynamics_ll.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
