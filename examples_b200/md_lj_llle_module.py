"""Force routine for MD simulation, LJ atoms, Lees-Edwards boundaries, using neighbour lists -- B200 drop-in for
the reference's python_examples/md_lj_llle_module.py.  The sheared cell neighbourhood
(md_lj_llle_module.py:88-92,112,131-135) is enumerated on the device (csrc/common.cuh neighbour_cell).

Documented deviation: the reference double-counts image cells when sc == 3; this
module serves sc == 3 with the all-pairs kernel, i.e. it returns the CORRECT forces there.  No CPU fallback."""
import math

import numpy as np

from . import _lib
from .md_lj_le_module import PotentialType, _force_le, conclusion, r_cut, r_cut_sq  # noqa: F401

fast = True


def introduction():
    """Prints out introductory statements at start of run (md_lj_llle_module.py:52-63)."""
    print('WCA shifted Lennard-Jones potential')
    print('Diameter, sigma = 1')
    print('Well depth, epsilon = 1')
    print('Lees-Edwards boundaries')
    print('CUDA force routine (libatlj, sm_100a)')
    print('Uses neighbour lists')


def force(box, strain, r):
    """Takes in box, strain, and coordinate array, and calculates forces and potentials etc."""
    # The reference applies the Lees-Edwards correction to the CALLER's array in place before rebinding a wrapped
    # local copy (md_lj_llle_module.py:94-95); kept, so that side effect is identical.
    r[:, 0] = r[:, 0] - np.rint(r[:, 1]) * strain
    sc = math.floor(box / r_cut)                     # md_lj_llle_module.py:107
    assert sc >= 3, 'System is too small for cells'  # md_lj_llle_module.py:108
    rc, totals, f, ovr, _ = _force_le(box, strain, r, sc, 1)
    _lib.check(rc)
    total = PotentialType(pot=float(totals[0]), vir=float(totals[1]), lap=float(totals[2]), pyx=float(totals[3]),
                          ovr=ovr)
    return total, f
