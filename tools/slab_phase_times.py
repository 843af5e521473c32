"""Kernel categories of a slab rank against the single domain of the same size, on ONE GPU (ranks emulated in lockstep):
    python tools/slab_phase_times.py [nc_total] [world] [steps]
Shows what the slab variants of the kernels cost beyond their single-domain forms (peer stores are local here)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import DT, R_CUT, make_system  # noqa: E402
from examples_b200.device import Simulation  # noqa: E402
from examples_b200.slab import EmulatedSlabs  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 80
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
n, box, r, v = make_system(nc)
many = EmulatedSlabs(box, R_CUT, r, v, world)
many.force()
many.run_nve(DT, 100)
for s in many.ranks:
    s.enable_timing(True)
    s.timing()
many.run_nve(DT, steps)
for k, s in enumerate(many.ranks):
    t = s.timing()
    print("rank %d of %d (N_total=%d, owned %d): " % (k, world, n, s.natoms)
          + "  ".join("%s %.4f" % (a, b / steps) for a, b in t.items() if b > 0), flush=True)
many.close()
