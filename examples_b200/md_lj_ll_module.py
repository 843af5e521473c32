"""Force routine for MD simulation, LJ atoms, using neighbour lists -- B200 drop-in for the reference's
python_examples/md_lj_ll_module.py (same names, arguments, return values and assertion messages).

The link-cell construction (md_lj_ll_module.py:106-120, link_list_module.f90:103-161) is a counting sort on the
GPU whose cell triplets and per-cell atom order are bit-identical to the reference's; the pair loop
(:122-165) is the tiled fp64 kernel of examples_b200/csrc/pair.cu.  No CPU fallback.
"""
import math

import numpy as np

from . import _lib
from .md_lj_module import PotentialType, _force, _hessian, conclusion  # noqa: F401  (same type as md_lj_module)

fast = True  # kept for API compatibility (md_lj_ll_module.py:28)


def introduction():
    """Prints out introductory statements at start of run (md_lj_ll_module.py:50-62)."""
    print('Lennard-Jones potential')
    print('Cut-and-shifted version for dynamics')
    print('Cut (but not shifted) version also calculated')
    print('Diameter, sigma = 1')
    print('Well depth, epsilon = 1')
    print('CUDA force routine (libatlj, sm_100a)')
    print('Uses neighbour lists')


def force(box, r_cut, r):
    """Takes in box, cutoff range, and coordinate array, and calculates forces and potentials etc."""
    sc = math.floor(box / r_cut)                      # Number of cells along box edge (md_lj_ll_module.py:106)
    assert sc >= 3, 'System is too small for cells'   # md_lj_ll_module.py:107
    rc, totals, f, ovr, _ = _force(box, r_cut, r, sc, 1)
    _lib.check(rc)                                    # 'Index error' (md_lj_ll_module.py:109)
    total = PotentialType(cut=float(totals[0]), pot=float(totals[1]), vir=float(totals[2]), lap=float(totals[3]),
                          ovr=ovr)
    return total, f


def hessian(box, r_cut, r, f):
    """Calculates Hessian function (for 1/N correction to config temp) (md_lj_ll_module.py:227-357)."""
    assert np.all(r.shape == f.shape), 'Dimension mismatch in hessian'
    sc = math.floor(box / r_cut)
    assert sc >= 3, 'System is too small for cells'
    rc, hes = _hessian(box, r_cut, r, f, sc, 1)
    _lib.check(rc)
    return hes


def cells(box, r_cut, r):
    """The integer work on its own: cell triplets c (N,3) int64, perm (atoms in cell order, ascending index
    inside a cell) and cell_start (sc^3+1), exactly as md_lj_ll_module.py:106-120 lays the atoms out."""
    sc = math.floor(box / r_cut)
    assert sc >= 3, 'System is too small for cells'
    r = _lib.as_f64(r)
    n = r.shape[0]
    c = np.empty((n, 3), dtype=np.int64)
    perm = np.empty(n, dtype=np.int32)
    cell_start = np.empty(sc ** 3 + 1, dtype=np.int32)
    _lib.check(_lib.lib().atlj_cells(r, n, sc, 0, 0.0, c, perm, cell_start))
    return c, perm, cell_start
