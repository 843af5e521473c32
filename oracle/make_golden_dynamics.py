#!/usr/bin/env python
"""Generate tests/golden/reference_dynamics.npz by EXECUTING the reference drivers' own propagators and step loops.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_dynamics.py

The drivers (bd_nvt_lj.py, md_nvt_lj.py, md_nvt_lj_le.py) are scripts, not importable modules, so their propagator
functions and the body of their `for stp in range(nstep)` loop are lifted out of the parsed source (ast) and executed,
unmodified, in a namespace holding the driver's globals.  Nothing is re-typed: what runs is the reference's code.
np.random.randn is wrapped to record the Gaussian numbers the Langevin driver draws, so the CUDA path can be fed the
same noise (bd_nvt_lj.py:127 uses the unseeded legacy generator).
"""
import ast
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ATLJ_REFERENCE", "/root/reference/python_examples")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)  # the propagators import numpy / scipy only, but keep the reference importable

from examples_b200.lattice import fcc_positions, perturb, ran_velocities  # noqa: E402


def ref_module(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def lift(driver, func_names):
    """-> (dict of compiled function-def code, compiled step-loop body) from the driver's source."""
    src = open(os.path.join(REF, driver)).read()
    tree = ast.parse(src)
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in func_names]
    assert len(funcs) == len(func_names), (driver, [f.name for f in funcs])
    loop = None
    for node in ast.walk(tree):
        if isinstance(node, ast.For) and isinstance(node.target, ast.Name) and node.target.id == "stp":
            loop = node
    assert loop is not None
    body = [st for st in loop.body
            if not (isinstance(st, ast.Expr) and isinstance(st.value, ast.Call)
                    and getattr(st.value.func, "id", "") == "blk_add")]
    fmod = ast.Module(body=funcs, type_ignores=[])
    bmod = ast.Module(body=body, type_ignores=[])
    ast.fix_missing_locations(fmod), ast.fix_missing_locations(bmod)
    return compile(fmod, driver, "exec"), compile(bmod, driver + ":step", "exec")


def main():
    out = {}
    lj, le = ref_module("md_lj_module"), ref_module("md_lj_le_module")
    dt = 0.005

    # ---- BAOAB Langevin, bd_nvt_lj.py, N=256 LJ -------------------------------------------------------------
    fcode, step = lift("bd_nvt_lj.py", ["a_propagator", "b_propagator", "o_propagator"])
    n, box, r = fcc_positions(4, 0.75)
    r = perturb(r, box, 0.05)
    v = ran_velocities(n, 1.0, seed=7)
    ns = dict(np=np, n=n, box=box, r_cut=2.5, r=r.copy(), v=v.copy(), dt=dt, temperature=1.0, gamma=1.0,
              force=lj.force)
    exec(fcode, ns)
    ns["total"], ns["f"] = lj.force(box, 2.5, ns["r"])
    nsteps = 100
    drawn = []
    real_randn = np.random.randn

    def recording_randn(*shape):
        x = real_randn(*shape)
        drawn.append(x.copy())
        return x

    np.random.seed(20240607)
    np.random.randn = recording_randn
    try:
        rec = np.empty((nsteps, 6))
        for s in range(nsteps):
            exec(step, ns)
            t = ns["total"]
            rec[s] = [t.cut, t.pot, t.vir, t.lap, np.sum(ns["v"] ** 2), np.sum(ns["f"] ** 2)]
    finally:
        np.random.randn = real_randn
    out.update(bd_box=np.float64(box), bd_r0=r, bd_v0=v, bd_noise=np.array(drawn), bd_rec=rec, bd_r1=ns["r"],
               bd_v1=ns["v"], bd_dt=np.float64(dt), bd_gamma=np.float64(1.0), bd_temperature=np.float64(1.0))
    print("bd_nvt_lj done", rec[-1], flush=True)

    # ---- Nose-Hoover chain, md_nvt_lj.py, N=256 LJ, M=3 --------------------------------------------------------
    fcode, step = lift("md_nvt_lj.py", ["u1_propagator", "u2_propagator", "u3_propagator", "u4_propagator"])
    n, box, r = fcc_positions(4, 0.75)
    r = perturb(r, box, 0.05)
    v = ran_velocities(n, 1.0, seed=8)
    m, temperature, tau = 3, 1.0, 2.0
    g = 3 * (n - 1)                                       # md_nvt_lj.py:244
    q = np.full(m, temperature * tau ** 2, dtype=np.float64)
    q[0] = g * temperature * tau ** 2
    eta = np.zeros(m)
    p_eta = np.random.default_rng(5).standard_normal(m) * np.sqrt(temperature) * np.sqrt(q)
    ns = dict(np=np, n=n, box=box, r_cut=2.5, r=r.copy(), v=v.copy(), dt=dt, temperature=temperature, m=m, g=g,
              q=q.copy(), eta=eta.copy(), p_eta=p_eta.copy(), force=lj.force)
    exec(fcode, ns)
    ns["total"], ns["f"] = lj.force(box, 2.5, ns["r"])
    rec = np.empty((nsteps, 7))
    for s in range(nsteps):
        exec(step, ns)
        t = ns["total"]
        ext = np.sum(0.5 * ns["p_eta"] ** 2 / ns["q"]) + temperature * (g * ns["eta"][0] + np.sum(ns["eta"][1:]))
        rec[s] = [t.cut, t.pot, t.vir, t.lap, np.sum(ns["v"] ** 2), np.sum(ns["f"] ** 2), ext]
    out.update(nhc_box=np.float64(box), nhc_r0=r, nhc_v0=v, nhc_q=q, nhc_eta0=eta, nhc_p_eta0=p_eta, nhc_rec=rec,
               nhc_r1=ns["r"], nhc_v1=ns["v"], nhc_eta1=ns["eta"], nhc_p_eta1=ns["p_eta"],
               nhc_temperature=np.float64(temperature), nhc_g=np.float64(g))
    print("md_nvt_lj done", rec[-1], flush=True)

    # ---- isokinetic SLLOD, md_nvt_lj_le.py, N=256 WCA, rho=0.8442 ---------------------------------------------
    fcode, step = lift("md_nvt_lj_le.py", ["a_propagator", "b1_propagator", "b2_propagator"])
    for tag, rate in (("le", 0.04), ("le_fast", 0.64)):
        n, box, r = fcc_positions(4, 0.8442)
        r = perturb(r, box, 0.05)
        v = ran_velocities(n, 0.722, seed=9)
        ns = dict(np=np, n=n, box=box, r=r.copy(), v=v.copy(), dt=dt, strain=0.0, strain_rate=rate, force=le.force)
        exec(fcode, ns)
        ns["total"], ns["f"] = le.force(box, ns["strain"], ns["r"])
        rec = np.empty((nsteps, 8))
        for s in range(nsteps):
            exec(step, ns)
            t = ns["total"]
            vv = ns["v"]
            rec[s] = [t.pot, t.vir, t.lap, t.pyx, np.sum(vv ** 2), np.sum(ns["f"] ** 2), np.sum(vv[:, 0] * vv[:, 1]),
                      ns["strain"]]
        out.update({tag + "_box": np.float64(box), tag + "_r0": r, tag + "_v0": v, tag + "_rec": rec,
                    tag + "_r1": ns["r"], tag + "_v1": ns["v"], tag + "_rate": np.float64(rate)})
        print("md_nvt_lj_le", rate, "done", rec[-1], flush=True)

    path = os.path.join(ROOT, "tests", "golden", "reference_dynamics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; numpy", np.__version__)


if __name__ == "__main__":
    main()
