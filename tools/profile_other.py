"""Step rates of the other BASELINE configurations: c3 (SLLOD, WCA, N=256,000, strain rate 0.01) and c4 (BAOAB, LJ,
N=1,000,188): `python tools/profile_other.py`."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from examples_b200.device import Simulation  # noqa: E402
from examples_b200.lattice import fcc_positions, ran_velocities  # noqa: E402


def run(tag, sim, fn, n, steps):
    sim.enable_timing(True)
    sim.timing()
    sim.timer_start()
    fn(steps)
    ms = sim.timer_stop()
    t = sim.timing()
    print("%s N=%d: %.4f ms/step = %.3g atom-steps/s  (pair %.4f cells %.4f integrate %.4f)"
          % (tag, n, ms / steps, n * steps / (ms * 1e-3), t["pair"] / steps, t["cells"] / steps, t["integrate"] / steps))


n, box, r = fcc_positions(40, 0.75)
r = r + np.random.default_rng(1).normal(0.0, 0.05 / box, size=r.shape)  # the perfect lattice has no WCA pair in range
r = np.ascontiguousarray(r - np.rint(r))
sim = Simulation(box, None, r, ran_velocities(n, 1.0), model="wca")
sim.force(strain=0.0)
state = {"strain": 0.0}


def sllod(k):
    rec, state["strain"] = sim.run_sllod(0.005, 0.01, state["strain"], k)
    return rec


sllod(400)
run("c3 SLLOD WCA", sim, sllod, n, 200)
sim.close()

n, box, r = fcc_positions(63, 0.75)
sim = Simulation(box, 2.5, r, ran_velocities(n, 2.0))
sim.force()
sim.run_baoab(0.005, 1.0, 1.0, 400, seed=3)
run("c4 BAOAB LJ", sim, lambda k: sim.run_baoab(0.005, 1.0, 1.0, k, seed=3), n, 100)
run("c4 NVE LJ", sim, lambda k: sim.run_nve(0.005, k), n, 100)
sim.close()
