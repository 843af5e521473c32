"""The reference's UNMODIFIED driver scripts, executed as the reference's users do (stdin JSON, cnf.inp in the working
directory), with examples_b200/dropin FIRST on PYTHONPATH: md_nve_lj.py (all pairs, and the link-cell route the way
GUIDE.md:288-289 prescribes), md_nvt_lj.py, md_npt_lj.py, md_nvt_lj_le.py (all pairs and md_lj_llle_module cells).

The scripts themselves come from oracle/_ref/python_examples (byte copies staged by oracle/make_ref.py; sha256-checked),
their expected output from tests/golden/driver_*.stdout, which oracle/make_golden_driver.py produced by running the same
scripts with the reference's own force modules in the build container.  Compared: every number of the printed block
table / initial / final values (print precision), and the cnf.out the script wrote (1e-8).
"""
import os

import pytest

from conftest import ROOT



def _load(name):
    # by file, not through sys.path: a path entry for oracle/ would shadow the `oracle` package with oracle/oracle.py
    import importlib.util
    spec = importlib.util.spec_from_file_location("_atlj_" + name, os.path.join(ROOT, "oracle", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


D = _load("driver_runner")
staged_dir = _load("make_ref").staged_dir

DROPIN = os.path.join(ROOT, "examples_b200", "dropin")


def test_driver_goldens_are_committed():
    for name in D.CASES:
        for ext in ("inp", "stdout", "out"):
            assert os.path.exists(os.path.join(D.GOLDEN, "driver_%s.%s" % (name, ext))), (name, ext)
        want = open(os.path.join(D.GOLDEN, "driver_%s.stdout" % name)).read()
        assert "Program ends" in want and "Run averages" in want
        assert D.compare_stdout(want, want) > 40  # the comparison really finds the numbers


def test_staged_reference_is_unmodified_when_present():
    """oracle/_ref is git-ignored and staged by build(); when it is there its sha256 manifest must verify."""
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "python_examples")):
        pytest.skip("oracle/_ref not staged (run oracle/make_ref.py in the build container)")
    assert staged_dir() is not None


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(D.CASES))
def test_unmodified_reference_driver_runs_on_the_dropin(name):
    ref = staged_dir()
    if ref is None:
        pytest.skip("oracle/_ref not staged: the reference scripts are not on this machine")
    stdout, cnf = D.run_case(name, [DROPIN, ref])
    assert "CUDA force routine (libatlj, sm_100a)" in stdout, "the drop-in module was not the one imported"
    want = open(os.path.join(D.GOLDEN, "driver_%s.stdout" % name)).read()
    assert D.compare_stdout(stdout, want) > 40
    err = D.compare_cnf(cnf, open(os.path.join(D.GOLDEN, "driver_%s.out" % name)).read())
    assert err < 1e-8, err
