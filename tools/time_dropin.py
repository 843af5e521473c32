#!/usr/bin/env python
"""The drop-in modules timed on the reference's own configurations, row for row what oracle/time_reference_python.py
measures for the reference's NumPy code (bench.py prints the two side by side):

  md_lj_module.force            N = 256                    all pairs, NumPy arrays in and out
  md_lj_ll_module.force         N = 864, 4,000, 32,000     link cells
  md_nve_lj.py                  the reference's UNMODIFIED driver script with examples_b200/dropin first on the path,
                                1 block x 2,000 steps at N = 256 (force + hessian every step, md_nve_lj.py:46,178-192)

    python tools/time_dropin.py          -> one JSON object on stdout (needs a CUDA device: there is no CPU path)
"""
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from examples_b200 import md_lj_ll_module, md_lj_module  # noqa: E402
from examples_b200.lattice import fcc_positions, perturb, ran_velocities  # noqa: E402

DROPIN = os.path.join(ROOT, "examples_b200", "dropin")


def staged_reference():
    """Directory of the reference's driver scripts (oracle/_ref, staged by oracle/make_ref.py), or None."""
    d = os.path.join(ROOT, "oracle", "_ref", "python_examples")
    return d if os.path.isfile(os.path.join(d, "md_nve_lj.py")) else None


_WRAP = r"""
import io, json, runpy, sys, time
script, nstep = sys.argv[1], int(sys.argv[2])
sys.argv = [script]
def run(ns):
    sys.stdin = io.StringIO('{"nblock":1,"nstep":%d}' % ns)
    out, old = io.StringIO(), sys.stdout
    sys.stdout = out
    t0 = time.perf_counter()
    try:
        runpy.run_path(script, run_name='__main__')
    finally:
        sys.stdout = old
    return time.perf_counter() - t0, out.getvalue()
run(10)                      # imports, CUDA context, buffers
a, _ = run(10)
b, text = run(10 + nstep)
print(json.dumps({"short": a, "long": b, "ok": "Program ends" in text and "CUDA force routine" in text}))
"""


def time_driver(ref, nstep=2000):
    """The unmodified script, run three times IN ONE PROCESS (runpy): the difference between a 10-step and a
    (10 + nstep)-step run is the time of nstep steps of its loop, free of interpreter and CUDA start-up."""
    n, box, r = fcc_positions(4, 0.75)
    r = perturb(r, box, 0.05)
    v = ran_velocities(n, 1.0, seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "cnf.inp"), "w") as fh:  # config_io_module.py:88-95
            fh.write("%15d\n%15.8f\n" % (n, box))
            np.savetxt(fh, np.concatenate((r * box, v), axis=1), fmt="%15.10f")
        # PYTHONSAFEPATH: the script's own directory must not shadow the drop-in (same as tests/test_reference_driver.py)
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([DROPIN, ROOT, ref]), PYTHONSAFEPATH="1")
        t0 = time.perf_counter()
        p = subprocess.run([sys.executable, "-c", _WRAP, os.path.join(ref, "md_nve_lj.py"), str(nstep)],
                           capture_output=True, text=True, cwd=tmp, env=env, timeout=600)
        wall = time.perf_counter() - t0
        if p.returncode != 0:
            raise RuntimeError("md_nve_lj.py on the drop-in failed: " + (p.stderr or p.stdout)[-400:])
        d = json.loads(p.stdout.strip().splitlines()[-1])
        if not d["ok"]:
            raise RuntimeError("md_nve_lj.py did not run on the drop-in module")
        return n, wall, (d["long"] - d["short"]) / nstep


def main():
    out = {"rows": []}
    for name, mod, nc, calls in (("md_lj_module.force", md_lj_module, 4, 20), ("md_lj_ll_module.force", md_lj_ll_module, 6, 20),
                                 ("md_lj_ll_module.force", md_lj_ll_module, 10, 20),
                                 ("md_lj_ll_module.force", md_lj_ll_module, 20, 20)):
        n, box, r = fcc_positions(nc, 0.75)
        r = perturb(r, box, 0.05)
        mod.force(box, 2.5, r)  # context, buffers
        ts = []
        for _ in range(calls):
            t0 = time.perf_counter()
            mod.force(box, 2.5, r)
            ts.append(time.perf_counter() - t0)
        t = statistics.median(ts)
        out["rows"].append({"what": name, "n": n, "calls": calls, "sec_per_call": t, "atoms_per_s": n / t})
    ref = staged_reference()
    if ref is None:
        out["rows"].append({"what": "md_nve_lj.py", "unavailable": "oracle/_ref is not staged"})
    else:
        try:
            n, wall, per_step = time_driver(ref)
            out["rows"].append({"what": "md_nve_lj.py (unmodified script on the drop-in, 1 block x 2,000 steps, force + "
                                        "hessian per step)", "n": n, "wall_s": wall, "sec_per_step": per_step,
                                "atom_steps_per_s": n / per_step})
        except Exception as e:  # noqa: BLE001
            out["rows"].append({"what": "md_nve_lj.py", "error": str(e)[:300]})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
