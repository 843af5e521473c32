"""CPU: libatlj.so loads, exports every symbol declared in include/atlj.h, and refuses to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "atlj.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atlj_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("atlj_force_lj", "atlj_force_le", "atlj_hessian_lj", "atlj_cells", "atlj_sim_run_nve",
                 "atlj_sim_run_baoab", "atlj_sim_run_sllod", "atlj_sim_run_nhc"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from examples_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    assert _lib.lib().atlj_version() >= 100


def test_no_cpu_fallback_without_a_device():
    """On a machine without a GPU the product path must fail loudly, never compute on the CPU."""
    from examples_b200 import _lib, md_lj_ll_module
    from examples_b200.lattice import fcc_positions
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    n, box, r = fcc_positions(6)
    with pytest.raises(_lib.AtljError):
        md_lj_ll_module.force(box, 2.5, r)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "examples_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.lower().replace("no oracle", ""), os.path.join(dirpath, fn)


def test_dropin_modules_export_the_reference_names():
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "examples_b200", "dropin"))
    try:
        for name, names in (("md_lj_module", ("introduction", "conclusion", "force", "hessian", "PotentialType")),
                            ("md_lj_ll_module", ("introduction", "conclusion", "force", "hessian", "PotentialType")),
                            ("md_lj_le_module", ("introduction", "conclusion", "force", "PotentialType")),
                            ("md_lj_llle_module", ("introduction", "conclusion", "force", "PotentialType"))):
            sys.modules.pop(name, None)
            m = importlib.import_module(name)
            assert "examples_b200" in os.path.abspath(m.__file__)
            for k in names:
                assert hasattr(m, k), (name, k)
            sys.modules.pop(name, None)
    finally:
        sys.path.pop(0)
    from examples_b200.md_lj_module import PotentialType
    a = PotentialType(cut=1.0, pot=2.0, vir=3.0, lap=4.0, ovr=False) + PotentialType(1.0, 1.0, 1.0, 1.0, True)
    assert (a.cut, a.pot, a.vir, a.lap, a.ovr) == (2.0, 3.0, 4.0, 5.0, True)
    from examples_b200.md_lj_le_module import PotentialType as P2
    b = P2(pot=1.0, vir=2.0, lap=3.0, pyx=4.0, ovr=False) + P2(1.0, 1.0, 1.0, 1.0, False)
    assert (b.pot, b.vir, b.lap, b.pyx, b.ovr) == (2.0, 3.0, 4.0, 5.0, False)
