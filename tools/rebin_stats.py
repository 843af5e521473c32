"""SURVEY.md 8f-2 (reuse of the neighbour structure across steps), measured: how much of the cell-sorted order survives
one velocity-Verlet step?  `python tools/rebin_stats.py [nc] [steps]` melts the bench system, then for every step
reports the fraction of atoms that changed CELL and the fraction whose SLOT in the cell-sorted arrays changed (an
incremental re-bin could leave the second kind alone).  The arrays are compact (no slack between cells), so one atom
that changes cell shifts every atom behind it up to the cell that receives it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import DT, R_CUT, make_system  # noqa: E402
from examples_b200.device import Simulation  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 80
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n, box, r, v = make_system(nc)
sim = Simulation(box, R_CUT, r, v)
sim.force()
sim.run_nve(DT, 2000, record=False)
sc = sim.sc


def cells_and_slots():
    rs, _, _, ids = sim.download_sorted()
    c = np.floor((rs + 0.5) * sc).astype(np.int64)
    cell = (c[:, 0] * sc + c[:, 1]) * sc + c[:, 2]
    slot = np.empty(n, dtype=np.int64)
    slot[ids] = np.arange(n)
    cell_of = np.empty(n, dtype=np.int64)
    cell_of[ids] = cell
    return cell_of, slot


c0, s0 = cells_and_slots()
for k in range(steps):
    sim.run_nve(DT, 1, record=False)
    c1, s1 = cells_and_slots()
    moved_cell = float(np.mean(c1 != c0))
    moved_slot = float(np.mean(s1 != s0))
    rows = int(np.unique(np.concatenate((c1[c1 != c0] // sc, c0[c1 != c0] // sc))).size)
    print("step %d: N=%d sc=%d  changed cell %.3f %% (%d atoms; %d of %d z rows touched)  changed slot %.2f %%"
          % (k + 1, n, sc, 100 * moved_cell, int(moved_cell * n), rows, sc * sc, 100 * moved_slot))
    c0, s0 = c1, s1
sim.close()
