"""Soak test at BASELINE size: `python tools/soak.py [nc] [steps]` (default nc=80: N = 2,048,000; 20,000 steps).
NVE after a 2,000-step melt; reports the drift and the fluctuation of the conserved energy per atom, checks that no
step raised the overlap flag or an error, and cross-checks the final configuration between the two independent pair
kernels (pair_tile2 with the fp16 filter vs the first tiled kernel with the fp32 filter, in a child process): in-range
pair counts must be identical, totals and forces agree to 1e-10."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import DT, R_CUT, make_system  # noqa: E402
from examples_b200.device import Simulation  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    d = np.load(sys.argv[2])
    sim = Simulation(float(d["box"]), R_CUT, d["r"])
    rec = sim.force()
    r, v, f = sim.download()
    np.savez(sys.argv[3], rec=rec, f=f)
    sys.exit(0)

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 80
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
n, box, r, v = make_system(nc)
sim = Simulation(box, R_CUT, r, v)
sim.force()
sim.run_nve(DT, 2000, record=False)
t0 = time.perf_counter()
recs = []
for k in range(0, steps, 2000):
    recs.append(sim.run_nve(DT, min(2000, steps - k)))
rec = np.concatenate(recs)
wall = time.perf_counter() - t0
e = (0.5 * rec[:, 4] + rec[:, 1]) / n
t = np.arange(len(e))
slope = np.polyfit(t, e, 1)[0]
print("N=%d steps=%d wall %.1f s (%.3g atom-steps/s incl. record download)" % (n, steps, wall, n * steps / wall))
print("conserved energy per atom: mean %.8f, rms fluctuation %.2e, drift %.2e per step (%.2e over the run)"
      % (e.mean(), e.std(), slope, slope * len(e)))
print("overlap flag raised in %d steps; in-range pairs per atom %.4f +- %.4f; T %.4f"
      % (int(rec[:, 7].sum()), rec[:, 8].mean() / n, rec[:, 8].std() / n, rec[:, 4].mean() / (3 * n - 3)))
rs, vs, fs = sim.download()
rec2 = sim.force()
rs2, vs2, fs2 = sim.download()
with tempfile.TemporaryDirectory() as d:
    a, b = os.path.join(d, "in.npz"), os.path.join(d, "out.npz")
    np.savez(a, r=rs2, box=np.float64(box))
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child", a, b], check=True,
                   env={**os.environ, "ATLJ_PAIR_KERNEL": "tile1"}, timeout=600)
    o = np.load(b)
dt = np.max(np.abs(o["rec"][:4] - rec2[:4]) / np.abs(rec2[:4]))
df = np.max(np.abs(o["f"] - fs2)) / np.max(np.abs(fs2))
print("tile2 (fp16 filter) vs tile1 (fp32 filter) on the final configuration: pairs %d vs %d, totals rel %.1e, forces "
      "rel %.1e" % (rec2[8], o["rec"][8], dt, df))
ok = rec[:, 7].sum() == 0 and rec2[8] == o["rec"][8] and dt < 1e-10 and df < 1e-10 and abs(slope * len(e)) < 1e-4
print("SOAK", "OK" if ok else "FAILED")
sim.close()
sys.exit(0 if ok else 1)
