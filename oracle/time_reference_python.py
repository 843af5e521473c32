#!/usr/bin/env python
"""Time the reference's own Python/NumPy routines on THIS host's cores (CPU baseline item 1).

BASELINE INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg runs it as a subprocess; nothing in the product imports it).
The reference files come from oracle/_ref/python_examples (staged by oracle/make_ref.py, travels to the GPU box) or,
in the build container, from /root/reference/python_examples.  Rows:

  md_lj_module.force            N = 256                    (md_lj_module.py:68-149)
  md_lj_ll_module.force         N = 864, 4,000, 32,000     (md_lj_ll_module.py:69-224)
  md_nve_lj.py                  the UNMODIFIED driver script, 1 block x 100 steps at N = 256 (force + hessian every
                                step, md_nve_lj.py:46,178-192), wall time of the whole process and of its step loop

    python oracle/time_reference_python.py [--budget SECONDS] [--out FILE]     -> one JSON object on stdout
"""
import argparse
import importlib.util
import json
import os
import platform
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_ref import staged_dir  # noqa: E402
from examples_b200.lattice import fcc_positions, perturb, ran_velocities  # noqa: E402  (pure NumPy input generator)


def ref_dir():
    d = staged_dir()
    if d:
        return d, "oracle/_ref (staged copy of the reference, sha256-verified)"
    d = os.path.join(os.environ.get("ATLJ_REFERENCE", "/root/reference"), "python_examples")
    if os.path.isdir(d):
        return d, d
    return None, None


def ref_module(d, name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(d, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def write_cnf(path, n, box, r, v):
    """cnf.inp in the layout config_io_module.write_cnf_atoms produces (config_io_module.py:88-95)."""
    with open(path, "w") as fh:
        fh.write("%15d\n%15.8f\n" % (n, box))
        np.savetxt(fh, np.concatenate((r, v), axis=1), fmt="%15.10f")


def time_driver(d, nstep=100):
    """Run the unmodified md_nve_lj.py (reference modules only on the path) -> wall seconds, seconds per step."""
    n, box, r = fcc_positions(4, 0.75)
    r = perturb(r, box, 0.05)
    v = ran_velocities(n, 1.0, seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        write_cnf(os.path.join(tmp, "cnf.inp"), n, box, r * box, v)
        env = dict(os.environ, PYTHONPATH=d, OMP_NUM_THREADS="1")
        t = []
        for ns in (10, 10 + nstep):  # the difference removes start-up, imports and the two extra force calls
            t0 = time.perf_counter()
            p = subprocess.run([sys.executable, os.path.join(d, "md_nve_lj.py")], input='{"nblock":1,"nstep":%d}' % ns,
                               capture_output=True, text=True, cwd=tmp, env=env, timeout=600)
            t.append(time.perf_counter() - t0)
            if p.returncode != 0 or "Program ends" not in p.stdout:
                raise RuntimeError("md_nve_lj.py failed: " + p.stderr[-400:])
        return n, t[1], (t[1] - t[0]) / nstep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=float, default=60.0, help="approximate seconds of CPU work allowed")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    d, where = ref_dir()
    if d is None:
        out = {"unavailable": "oracle/_ref is not staged and /root/reference is absent (run oracle/make_ref.py in "
                              "the build container)"}
        print(json.dumps(out))
        return
    t_begin = time.perf_counter()
    lj, ll = ref_module(d, "md_lj_module"), ref_module(d, "md_lj_ll_module")
    out = {"source": where, "host": "%d cores visible, NumPy %s (these routines run single-threaded), Python %s"
                                    % (os.cpu_count(), np.__version__, platform.python_version()),
           "cores": 1, "rows": []}
    for name, mod, nc, calls in (("md_lj_module.force", lj, 4, 5), ("md_lj_ll_module.force", ll, 6, 5),
                                 ("md_lj_ll_module.force", ll, 10, 5), ("md_lj_ll_module.force", ll, 20, 3)):
        n, box, r = fcc_positions(nc, 0.75)
        r = perturb(r, box, 0.05)
        ts = []
        for k in range(calls):
            t0 = time.perf_counter()
            mod.force(box, 2.5, r)
            ts.append(time.perf_counter() - t0)
            # the N = 32,000 call takes ~8 s: stop early when the budget is nearly spent (at least one call is kept)
            if time.perf_counter() - t_begin > args.budget:
                break
        t = statistics.median(ts)
        out["rows"].append({"what": name, "n": n, "calls": len(ts), "sec_per_call": t, "atoms_per_s": n / t})
    try:
        n, wall, per_step = time_driver(d)
        out["rows"].append({"what": "md_nve_lj.py (unmodified script, 1 block x 100 steps, force + hessian per step)",
                            "n": n, "wall_s": wall, "sec_per_step": per_step, "atom_steps_per_s": n / per_step})
    except Exception as e:  # noqa: BLE001  (a baseline row must not take the bench down)
        out["rows"].append({"what": "md_nve_lj.py", "error": str(e)[:300]})
    out["seconds"] = time.perf_counter() - t_begin
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
