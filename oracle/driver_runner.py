"""Run the reference's UNMODIFIED driver scripts as subprocesses and compare their printed tables.  TEST INFRASTRUCTURE.

Used by oracle/make_golden_driver.py (reference modules on PYTHONPATH -> golden stdout, build container only) and by
tests/test_reference_driver.py (examples_b200/dropin FIRST on PYTHONPATH -> the same scripts run on the GPU path).

A case = (driver script, JSON parameters on stdin, cnf.inp, optional module alias).  The alias implements the reference's
own instruction for the link-cell route -- "the main program is edited to import routines from md_lj_ll_module.py"
(python_examples/GUIDE.md:288-289) -- without editing the script: a directory holding a one-line md_lj_module.py that
re-exports md_lj_ll_module is put in front of PYTHONPATH.
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (script, stdin JSON, nc of the fcc start, alias {module file: source})
CASES = {
    "nve_256": ("md_nve_lj.py", '{"nblock":2,"nstep":50}', 4, None),
    "nve_ll_864": ("md_nve_lj.py", '{"nblock":2,"nstep":25}', 6,
                   {"md_lj_module.py": "from md_lj_ll_module import *  # GUIDE.md:288-289\n"}),
    # md_nvt_lj.py:227,256 and md_npt_lj.py:289,320-331 call np.random.seed() (OS entropy) and draw the thermostat /
    # barostat momenta: SEEDED cases are started through runpy with np.random.seed pinned to a fixed value, so that
    # the unmodified script draws the same numbers in the golden run and in the test
    "nvt_256": ("md_nvt_lj.py", '{"nblock":2,"nstep":50}', 4, None),
    "npt_256": ("md_npt_lj.py", '{"nblock":2,"nstep":50}', 4, None),
    # strain_rate * dt * nstep must be an integer (md_nvt_lj_le.py:198-201): 4.0 * 0.005 * 50 = 1, so the strain crosses
    # the +-0.5 wrap and every shifted-stencil case of md_lj_llle_module.py:131-135 inside one short block
    "le_256": ("md_nvt_lj_le.py", '{"nblock":2,"nstep":50,"strain_rate":4.0}', 4, None),
    "le_ll_2048": ("md_nvt_lj_le.py", '{"nblock":2,"nstep":50,"strain_rate":4.0}', 8,
                   {"md_lj_le_module.py": "from md_lj_llle_module import *  # GUIDE.md:476-477\n"}),
}

SEEDED = {"nvt_256": 20170101, "npt_256": 20170102}
_RUNPY = ("import sys, runpy, numpy as np\n"
          "_seed = np.random.seed\n"
          "np.random.seed = lambda *a, **k: _seed(%d)\n"
          "sys.argv = [%r]\n"
          "runpy.run_path(%r, run_name='__main__')\n")

# lines that legitimately differ between two runs or two force back-ends: versions, wall clock, the banner line that
# names the routine ("Fast NumPy force routine" / "CUDA force routine (libatlj, sm_100a)")
SKIP = re.compile(r"^(Python:|NumPy:|Date:|Time:|CPU time:|Fast |Slow |CUDA force routine|Uses neighbour lists)")
NUM = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


def case_input(name):
    """Deterministic start configuration of a case: perturbed fcc lattice + seeded velocities -> n, box, r*box, v."""
    from examples_b200.lattice import fcc_positions, perturb, ran_velocities
    nc = CASES[name][2]
    n, box, r = fcc_positions(nc, 0.75)
    r = perturb(r, box, 0.05)
    v = ran_velocities(n, 1.0, seed=7)
    return n, box, r * box, v


def write_cnf(path, n, box, r, v):
    """The layout config_io_module.write_cnf_atoms produces (config_io_module.py:80-95)."""
    with open(path, "w") as fh:
        fh.write("%15d\n%15.8f\n" % (n, box))
        np.savetxt(fh, np.concatenate((r, v), axis=1), fmt="%15.10f")


def run_case(name, module_dirs, timeout=900):
    """Run the case's script from script_dir = module_dirs[-1] with PYTHONPATH = [alias dir] + module_dirs.
    -> (stdout, cnf.out text)"""
    script, params, _, alias = CASES[name]
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(GOLDEN, "driver_%s.inp" % name)) as src, open(os.path.join(tmp, "cnf.inp"), "w") as dst:
            dst.write(src.read())
        path = list(module_dirs)
        if alias:
            ad = os.path.join(tmp, "_alias")
            os.makedirs(ad)
            for fn, text in alias.items():
                with open(os.path.join(ad, fn), "w") as fh:
                    fh.write(text)
            path.insert(0, ad)
        # PYTHONSAFEPATH (Python >= 3.11, same as `python -P`): do not put the SCRIPT's directory in front of sys.path,
        # otherwise the reference modules next to the script would always win over PYTHONPATH
        env = dict(os.environ, PYTHONPATH=os.pathsep.join(path), PYTHONSAFEPATH="1")
        path_script = os.path.join(module_dirs[-1], script)
        cmd = [sys.executable, path_script]
        if name in SEEDED:
            cmd = [sys.executable, "-c", _RUNPY % (SEEDED[name], path_script, path_script)]
        p = subprocess.run(cmd, input=params, capture_output=True, text=True, cwd=tmp, env=env, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("%s exited %d:\n%s\n%s" % (script, p.returncode, p.stdout[-2000:], p.stderr[-2000:]))
        with open(os.path.join(tmp, "cnf.out")) as fh:
            out = fh.read()
        return p.stdout, out


def compare_stdout(got, want, abs_tol=2.5e-6, rel_tol=2e-5):
    """Line by line: same words; fixed-point numbers (%15.6f) equal within the print precision; numbers in e-notation
    (mean-squared deviations of conserved quantities: 1e-8 for the NVE energy, pure roundoff ~1e-15 for the
    thermostatted ones) within 0.5 % or 1e-12 absolute.
    -> number of numeric fields compared.  Raises AssertionError with the offending line."""
    gl = [l for l in got.splitlines() if not SKIP.match(l.strip())]
    wl = [l for l in want.splitlines() if not SKIP.match(l.strip())]
    assert len(gl) == len(wl), "line count %d != %d\n%s" % (len(gl), len(wl), got[-1500:])
    nnum = 0
    for a, b in zip(gl, wl):
        ta, tb = a.split(), b.split()
        assert len(ta) == len(tb), "fields differ:\n%s\n%s" % (a, b)
        for x, y in zip(ta, tb):
            if NUM.match(x) and NUM.match(y):
                fx, fy = float(x), float(y)
                if "e" in y.lower():
                    assert abs(fx - fy) <= 1e-12 + 5e-3 * abs(fy), "number differs:\n%s\n%s" % (a, b)
                else:
                    assert abs(fx - fy) <= abs_tol + rel_tol * abs(fy), "number differs:\n%s\n%s" % (a, b)
                nnum += 1
            else:
                assert x == y, "text differs:\n%s\n%s" % (a, b)
    return nnum


def compare_cnf(got, want, tol=5e-9):
    ga = np.loadtxt(got.splitlines(), skiprows=2)
    wa = np.loadtxt(want.splitlines(), skiprows=2)
    assert got.splitlines()[0].split() == want.splitlines()[0].split()
    assert abs(float(got.splitlines()[1]) - float(want.splitlines()[1])) < 1e-7
    assert ga.shape == wa.shape
    # positions are written as r*box after wrapping: an atom within tol of the box face may appear on either side
    d = ga - wa
    box = float(want.splitlines()[1])
    d[:, :3] -= box * np.rint(d[:, :3] / box)
    return float(np.max(np.abs(d)))
