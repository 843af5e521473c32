"""Force routines for SMC simulation, Lennard-Jones atoms -- B200 drop-in for the reference's
python_examples/smc_lj_module.py (same names, arguments, return values and assertion messages).

force() is the routine of md_lj_module (smc_lj_module.force sums force_1 over i < j: the same pairs, the same
per-pair arithmetic and the same final factors, smc_lj_module.py:76-98,185-190); force_1() hands one atom and its
partners to libatlj.so.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from .md_lj_module import PotentialType, _host_scalars, conclusion
from .md_lj_module import force as _md_force

fast = True  # kept for API compatibility (smc_lj_module.py:29); the CUDA route is always used


def introduction():
    """Prints out introductory statements at start of run (smc_lj_module.py:58-69)."""
    print('Lennard-Jones potential')
    print('Cut-and-shifted version for SMC dynamics')
    print('Cut (but not shifted) version also calculated')
    print('Diameter, sigma = 1')
    print('Well depth, epsilon = 1')
    print('CUDA force routine (libatlj, sm_100a)')


def force(box, r_cut, r):
    """Takes in box, cutoff range, and coordinate array, and calculates forces and potentials etc
    (smc_lj_module.py:76).  On an overlap the reference stops summing and returns ovr = True with unusable totals."""
    n, d = r.shape
    assert d == 3, 'Dimension error for r in force'
    return _md_force(box, r_cut, r)


def force_1(ri, box, r_cut, r):
    """Takes in coordinates of an atom and calculates its interactions (smc_lj_module.py:100): -> partial
    (PotentialType) and f_partial (nj,3), the force on ri due to each atom of r."""
    nj, d = r.shape
    assert d == 3, 'Dimension error for r in force_1'
    assert ri.size == 3, 'Dimension error for ri in force_1'
    r_cut_box_sq, box_sq, pot_cut = _host_scalars(box, r_cut)
    ri = _lib.as_f64(ri).reshape(3)
    r = _lib.as_f64(r)
    f_partial = np.empty((nj, 3))
    tot = np.zeros(4)
    ovr = C.c_int(0)
    _lib.check(_lib.lib().atlj_force_1(ri.ctypes.data, r.ctypes.data, nj, float(box), r_cut_box_sq, box_sq, pot_cut,
                                       f_partial.ctypes.data, tot.ctypes.data, C.byref(ovr)))
    if ovr.value:
        return PotentialType(pot=0.0, cut=0.0, vir=0.0, lap=0.0, ovr=True), f_partial
    return PotentialType(cut=tot[0], pot=tot[1], vir=tot[2], lap=tot[3], ovr=False), f_partial
