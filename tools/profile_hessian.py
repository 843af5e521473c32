import sys, os
sys.path.insert(0, "/root/repo")
from bench import DT, R_CUT, make_system
from examples_b200.device import Simulation
n, box, r, v = make_system(80)
sim = Simulation(box, R_CUT, r, v)
sim.force()
sim.run_nve(DT, 300, record=False)
for hes in (False, True):
    sim.run_nve(DT, 5, hessian=hes)
    sim.enable_timing(True); sim.timing()
    sim.timer_start()
    sim.run_nve(DT, 20, hessian=hes)
    ms = sim.timer_stop()
    t = sim.timing()
    print("hessian=%s: %.4f ms/step  pair %.4f other(hessian) %.4f cells %.4f integrate %.4f" % (hes, ms / 20, t["pair"] / 20, t["other"] / 20, t["cells"] / 20, t["integrate"] / 20))
