"""cnf text path at BASELINE size: `python tools/profile_cnf.py [nc]` - N = 4*nc^3 atoms (default nc=80: 2,048,000),
r and v.  Times the GPU formatter / parser (host arrays in, text out, copies included) and the checkpoint of a
device-resident state, beside the reference's np.savetxt / np.loadtxt on a 100,000-row sample of the same data."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import R_CUT, make_system  # noqa: E402
from examples_b200 import _lib  # noqa: E402
from examples_b200.config_io_module import read_cnf_atoms, write_cnf_atoms, write_cnf_sim  # noqa: E402
from examples_b200.device import Simulation  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 80
n, box, r, v = make_system(nc)
sim = Simulation(box, R_CUT, r, v)
sim.force()
sim.run_nve(0.005, 20, record=False)
d = tempfile.mkdtemp()
p = os.path.join(d, "cnf.out")
rb = r * box


def best(fn, k=3):
    ts = []
    for _ in range(k):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


t_w = best(lambda: write_cnf_atoms(p, n, box, rb, v))
size = os.path.getsize(p)
t_r = best(lambda: read_cnf_atoms(p, with_v=True))
t_s = best(lambda: write_cnf_sim(p, sim))
body = np.empty(n * 6 * 16, dtype=np.uint8)
L = _lib.lib()
t_f = best(lambda: _lib.check(L.atlj_cnf_format(rb.ctypes.data, v.ctypes.data, n, 1.0, body.ctypes.data)))
rr, vv = np.empty((n, 3)), np.empty((n, 3))
t_p = best(lambda: _lib.check(L.atlj_cnf_parse(body.ctypes.data, n, 6, rr.ctypes.data, vv.ctypes.data)))
ns = 100000
ps = os.path.join(d, "cnf.ref")
rv = np.concatenate([rb[:ns], v[:ns]], axis=1)
t0 = time.perf_counter()
np.savetxt(ps, rv, header="{:15d}\n{:15.8f}".format(ns, box), fmt="%15.10f", comments="")
t_ws = time.perf_counter() - t0
t0 = time.perf_counter()
np.loadtxt(ps, skiprows=2)
t_rs = time.perf_counter() - t0
print("N=%d rows x 6 cols, %.1f MB of text" % (n, size / 1e6))
print("write_cnf_atoms (GPU format + file write): %.3f s = %.2f M atoms/s;  library call alone %.3f s (%.2f GB/s text)"
      % (t_w, n / t_w / 1e6, t_f, size / t_f / 1e9))
print("write_cnf_sim   (device state -> file)   : %.3f s = %.2f M atoms/s" % (t_s, n / t_s / 1e6))
print("read_cnf_atoms  (file read + GPU parse)  : %.3f s = %.2f M atoms/s;  library call alone %.3f s (%.2f GB/s text)"
      % (t_r, n / t_r / 1e6, t_p, size / t_p / 1e9))
print("reference np.savetxt on %d rows: %.3f s = %.3f M atoms/s (-> %.1f s at N);  np.loadtxt: %.3f s = %.3f M atoms/s "
      "(-> %.1f s at N)" % (ns, t_ws, ns / t_ws / 1e6, t_ws * n / ns, t_rs, ns / t_rs / 1e6, t_rs * n / ns))
sim.close()
