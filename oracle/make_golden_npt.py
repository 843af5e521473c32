#!/usr/bin/env python
"""Generate tests/golden/reference_npt.npz by EXECUTING the propagators and the step loop of the reference's
md_npt_lj.py (lifted from its parsed source, unmodified - see make_golden_dynamics.py).

Run in the build container only (needs /root/reference):

    python oracle/make_golden_npt.py

Three cases: N=864 whose shrinking box leaves the link-cell regime (sc 4 -> 3) during the run, N=256 at rho=0.75 (all-pairs route, 100 steps, the driver's default parameters), and N=2048 compressed to
box = 12.49 sigma so that the expanding box crosses box/r_cut = 5 within the run: the cell grid of the link-cell route
changes from sc=4 to sc=5 on the way (md_lj_ll_module.py:106).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_dynamics import REF, lift, ref_module  # noqa: E402,F401
from examples_b200.lattice import fcc_positions, perturb, ran_velocities  # noqa: E402


def run_case(tag, nc, box_override, nsteps, pressure, seed, out, lj):
    fcode, step = lift("md_npt_lj.py", ["u1_propagator", "u2_propagator", "u2p_propagator", "u3_propagator",
                                        "u4_propagator"])
    n, box, r = fcc_positions(nc, 0.75)
    if box_override:
        box = box_override
    r = perturb(r, box, 0.03)
    v = ran_velocities(n, 1.0, seed=seed)
    dt, m, temperature, tau, tau_baro, r_cut = 0.005, 3, 1.0, 2.0, 2.0, 2.5
    rng = np.random.default_rng(seed)
    g = 3 * (n - 1)                                                       # md_npt_lj.py:311
    q = np.full(m, temperature * tau ** 2, dtype=np.float64)
    q[0] = g * temperature * tau ** 2
    eta = np.zeros(m)
    p_eta = rng.standard_normal(m) * np.sqrt(temperature) * np.sqrt(q)
    q_baro = np.full(m, temperature * tau_baro ** 2, dtype=np.float64)
    eta_baro = np.zeros(m)
    p_eta_baro = rng.standard_normal(m) * np.sqrt(temperature) * np.sqrt(q_baro)
    w_eps = g * temperature * tau_baro ** 2
    p_eps = rng.standard_normal() * np.sqrt(temperature * w_eps)
    ns = dict(np=np, n=n, box=box, box0=box, eps=0.0, p_eps=p_eps, w_eps=w_eps, r_cut=r_cut, r=r.copy(), v=v.copy(),
              dt=dt, temperature=temperature, pressure=pressure, m=m, g=g, q=q.copy(), eta=eta.copy(),
              p_eta=p_eta.copy(), q_baro=q_baro.copy(), eta_baro=eta_baro.copy(), p_eta_baro=p_eta_baro.copy(),
              force=lj.force)
    exec(fcode, ns)
    ns["total"], ns["f"] = lj.force(box, r_cut, ns["r"])
    rec = np.empty((nsteps, 9))
    for s in range(nsteps):
        exec(step, ns)
        t = ns["total"]
        ext = (np.sum(0.5 * ns["p_eta"] ** 2 / ns["q"]) + np.sum(0.5 * ns["p_eta_baro"] ** 2 / ns["q_baro"])
               + 0.5 * ns["p_eps"] ** 2 / w_eps
               + temperature * (g * ns["eta"][0] + np.sum(ns["eta"][1:]) + np.sum(ns["eta_baro"])))
        rec[s] = [t.cut, t.pot, t.vir, t.lap, np.sum(ns["v"] ** 2), np.sum(ns["f"] ** 2), ns["box"], ext, ns["p_eps"]]
    out.update({tag + "_box0": np.float64(box), tag + "_r0": r, tag + "_v0": v, tag + "_q": q, tag + "_q_baro": q_baro,
                tag + "_p_eta0": p_eta, tag + "_p_eta_baro0": p_eta_baro, tag + "_p_eps0": np.float64(p_eps),
                tag + "_w_eps": np.float64(w_eps), tag + "_pressure": np.float64(pressure), tag + "_rec": rec,
                tag + "_r1": ns["r"], tag + "_v1": ns["v"], tag + "_eta1": ns["eta"], tag + "_p_eta1": ns["p_eta"],
                tag + "_eta_baro1": ns["eta_baro"], tag + "_p_eta_baro1": ns["p_eta_baro"],
                tag + "_eps1": np.float64(ns["eps"]), tag + "_p_eps1": np.float64(ns["p_eps"])})
    print(tag, "done: box", box, "->", ns["box"], "sc", int(box // r_cut), "->", int(ns["box"] // r_cut), rec[-1],
          flush=True)


def main():
    out = {}
    lj = ref_module("md_lj_module")
    run_case("npt256", 4, None, 100, 0.99, 21, out, lj)
    run_case("npt2048", 8, 12.49, 40, 0.99, 22, out, lj)
    # N=864 at rho=0.75 (box 10.48, sc=4) under a pressure of 8: the shrinking box drops below 4 cells per edge, the
    # force path switches from link cells to all pairs during the run
    run_case("npt864", 6, None, 80, 8.0, 23, out, lj)
    path = os.path.join(ROOT, "tests", "golden", "reference_npt.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; numpy", np.__version__)


if __name__ == "__main__":
    main()
