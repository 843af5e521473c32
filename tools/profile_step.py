"""Short device-resident NVE run for ncu: `python tools/profile_step.py [nc] [melt] [steps]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import DT, R_CUT, make_system  # noqa: E402
from examples_b200.device import Simulation  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 80
melt = int(sys.argv[2]) if len(sys.argv) > 2 else 300
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n, box, r, v = make_system(nc)
sc = int(os.environ["ATLJ_SC"]) if "ATLJ_SC" in os.environ else None   # coarser cell grid (skin experiments)
sim = Simulation(box, R_CUT, r, v, sc=sc)
sim.force()
sim.run_nve(DT, melt, record=False)
sim.sync()
sim.timer_start()
sim.run_nve(DT, steps, record=False)
ms = sim.timer_stop()
sim.enable_timing(True)
sim.timing()
rec = sim.run_nve(DT, steps)
t = sim.timing()
print("sc=%d whole step %.4f ms (graph replay unless ATLJ_NO_GRAPH=1) " % (sim.sc, ms / steps), end="")
print("N=%d steps=%d pairs/atom=%.3f T=%.3f ms/step: pair %.4f cells %.4f integrate %.4f"
      % (n, steps, rec[-1, 8] / n, rec[-1, 4] / (3 * n - 3), t["pair"] / steps, t["cells"] / steps,
         t["integrate"] / steps))
sim.close()
