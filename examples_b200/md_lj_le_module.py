"""Force routine for MD simulation, Lennard-Jones atoms, Lees-Edwards boundaries -- B200 drop-in for the
reference's python_examples/md_lj_le_module.py (WCA potential, force(box, strain, r)).  No CPU fallback."""
import ctypes as C
import math

import numpy as np

from . import _lib

fast = True
r_cut_sq = 2.0 ** (1.0 / 3.0)
r_cut = r_cut_sq ** (1.0 / 2.0)  # Minimum and cutoff of WCA LJ potential (md_lj_le_module.py:30-31)


class PotentialType:
    """A composite variable for interactions (md_lj_le_module.py:33-50)."""

    def __init__(self, pot, vir, lap, pyx, ovr):
        self.pot = pot  # the potential energy cut-and-shifted at r_cut
        self.vir = vir  # the virial
        self.lap = lap  # the Laplacian
        self.pyx = pyx  # the off-diagonal virial part of the pressure tensor
        self.ovr = ovr  # a flag indicating overlap (i.e. pot too high to use)

    def __add__(self, other):
        return PotentialType(self.pot + other.pot, self.vir + other.vir, self.lap + other.lap,
                             self.pyx + other.pyx, self.ovr or other.ovr)


def introduction():
    """Prints out introductory statements at start of run (md_lj_le_module.py:52-62)."""
    print('WCA shifted Lennard-Jones potential')
    print('Diameter, sigma = 1')
    print('Well depth, epsilon = 1')
    print('Lees-Edwards boundaries')
    print('CUDA force routine (libatlj, sm_100a)')


def conclusion():
    """Prints out concluding statements at end of run."""
    print('Program ends')


def _force_le(box, strain, r, sc, check_index):
    r = _lib.as_f64(r)
    n = r.shape[0]
    r_cut_box = r_cut / box            # md_lj_le_module.py:82-84
    r_cut_box_sq = r_cut_box ** 2
    box_sq = box ** 2
    shift = math.floor(strain * sc) if sc > 0 else 0
    f = _lib.result_empty(r.shape)      # page-locked when large: the GPU writes it by DMA
    totals = np.empty(4, dtype=np.float64)
    ovr = C.c_int(0)
    npair = C.c_int64(0)
    rc = _lib.lib().atlj_force_le(r, n, box, r_cut_box_sq, box_sq, float(strain), sc, shift, check_index, f, totals,
                                  C.byref(ovr), C.byref(npair))
    return rc, totals, f, bool(ovr.value), npair.value


def force(box, strain, r):
    """Takes in box, strain, and coordinate array, and calculates forces and potentials etc."""
    n, d = r.shape
    assert d == 3, 'Dimension error in force'
    # all pairs in the reference; the cell kernel needs LE-wrapped coordinates, which this module does not
    # require of its caller, so only the all-pairs kernel is used here
    rc, totals, f, ovr, _ = _force_le(box, strain, r, 0, 0)
    _lib.check(rc)
    total = PotentialType(pot=float(totals[0]), vir=float(totals[1]), lap=float(totals[2]), pyx=float(totals[3]),
                          ovr=ovr)
    return total, f
