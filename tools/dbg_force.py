import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import liquid_like
from examples_b200.device import Simulation
n, box, r = liquid_like(8, 0.75, 0.05)
s = Simulation(box, 2.5, r)
print(s.force()[:4])
