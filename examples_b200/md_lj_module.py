"""Force routine for MD simulation, Lennard-Jones atoms -- B200 drop-in for the reference's
python_examples/md_lj_module.py (same names, arguments, return values and assertion messages).

force()/hessian() hand the NumPy arrays to libatlj.so (hand-written sm_100a CUDA) through ctypes.  The pair
set and every in-range / overlap decision are the reference's; sums differ only by fp64 summation order
(<= 1e-10 relative, tests/test_force_parity.py).  Large systems (floor(box/r_cut) >= 4) are served by the
link-cell kernel, small ones by the all-pairs kernel -- both produce the all-pairs result of
md_lj_module.py:95-115.  There is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np

from . import _lib

fast = True  # kept for API compatibility (md_lj_module.py:29); the CUDA route is always used


class PotentialType:
    """A composite variable for interactions (md_lj_module.py:31-48)."""

    def __init__(self, cut, pot, vir, lap, ovr):
        self.cut = cut  # the potential energy cut (but not shifted) at r_cut
        self.pot = pot  # the potential energy cut-and-shifted at r_cut
        self.vir = vir  # the virial
        self.lap = lap  # the Laplacian
        self.ovr = ovr  # a flag indicating overlap (i.e. pot too high to use)

    def __add__(self, other):
        return PotentialType(self.cut + other.cut, self.pot + other.pot, self.vir + other.vir,
                             self.lap + other.lap, self.ovr or other.ovr)


def introduction():
    """Prints out introductory statements at start of run (md_lj_module.py:50-61)."""
    print('Lennard-Jones potential')
    print('Cut-and-shifted version for dynamics')
    print('Cut (but not shifted) version also calculated')
    print('Diameter, sigma = 1')
    print('Well depth, epsilon = 1')
    print('CUDA force routine (libatlj, sm_100a)')


def conclusion():
    """Prints out concluding statements at end of run (md_lj_module.py:63-66)."""
    print('Program ends')


def _host_scalars(box, r_cut):
    # the reference's own expressions, evaluated in Python so the thresholds are bit-identical
    # (md_lj_module.py:80-88)
    r_cut_box = r_cut / box
    r_cut_box_sq = r_cut_box ** 2
    box_sq = box ** 2
    sr2 = 1.0 / r_cut ** 2
    sr6 = sr2 ** 3
    sr12 = sr6 ** 2
    pot_cut = sr12 - sr6
    return r_cut_box_sq, box_sq, pot_cut


def _force(box, r_cut, r, sc, check_index):
    r = _lib.as_f64(r)
    n = r.shape[0]
    r_cut_box_sq, box_sq, pot_cut = _host_scalars(box, r_cut)
    f = _lib.result_empty(r.shape)      # page-locked when large: the GPU writes it by DMA
    totals = np.empty(4, dtype=np.float64)
    ovr = C.c_int(0)
    npair = C.c_int64(0)
    rc = _lib.lib().atlj_force_lj(r, n, box, r_cut_box_sq, box_sq, pot_cut, sc, check_index, f, totals,
                                  C.byref(ovr), C.byref(npair))
    return rc, totals, f, bool(ovr.value), npair.value


def force(box, r_cut, r):
    """Takes in box, cutoff range, and coordinate array, and calculates forces and potentials etc."""
    n, d = r.shape
    assert d == 3, 'Dimension error in force'
    sc = math.floor(box / r_cut)
    rc, totals, f, ovr, _ = _force(box, r_cut, r, sc if sc >= 4 else 0, 0)
    if rc == _lib.ERR_INDEX:
        # a coordinate exactly on the +0.5 face cannot be binned (md_lj_ll_module.py:109); the all-pairs
        # reference has no such restriction, so use the all-pairs kernel for this call
        rc, totals, f, ovr, _ = _force(box, r_cut, r, 0, 0)
    _lib.check(rc)
    total = PotentialType(cut=float(totals[0]), pot=float(totals[1]), vir=float(totals[2]), lap=float(totals[3]),
                          ovr=ovr)
    return total, f


def _hessian(box, r_cut, r, f, sc, check_index):
    r = _lib.as_f64(r)
    f = _lib.as_f64(f)
    r_cut_box_sq, box_sq, _ = _host_scalars(box, r_cut)
    hes = C.c_double(0.0)
    rc = _lib.lib().atlj_hessian_lj(r, f, r.shape[0], box, r_cut_box_sq, box_sq, sc, check_index, C.byref(hes))
    return rc, hes.value


def hessian(box, r_cut, r, f):
    """Calculates Hessian function (for 1/N correction to config temp) (md_lj_module.py:151-213)."""
    n, d = r.shape
    assert d == 3, 'Dimension error in hessian'
    assert np.all(r.shape == f.shape), 'Dimension mismatch in hessian'
    sc = math.floor(box / r_cut)
    rc, hes = _hessian(box, r_cut, r, f, sc if sc >= 4 else 0, 0)
    if rc == _lib.ERR_INDEX:
        rc, hes = _hessian(box, r_cut, r, f, 0, 0)
    _lib.check(rc)
    return hes
