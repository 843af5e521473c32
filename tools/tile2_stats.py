"""Phase-2 iteration statistics of pair_tile2_kernel (debug build: `make -C examples_b200/csrc stats`):
    ATLJ_LIB=examples_b200/libatlj_stats.so python tools/tile2_stats.py [nc] [melt]
Per warp batch (32 atoms): executed iterations against the lower bounds of different walk schemes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ATLJ_LIB", os.path.join(ROOT, "examples_b200", "libatlj_stats.so"))
from bench import DT, R_CUT, make_system  # noqa: E402
from examples_b200 import _lib  # noqa: E402
from examples_b200.device import Simulation  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 80
melt = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n, box, r, v = make_system(nc)
sim = Simulation(box, R_CUT, r, v)
sim.force()
sim.run_nve(DT, melt, record=False)
L = _lib.lib()
out = (C.c_ulonglong * 8)()
L.atlj_debug_tile2_stats(out)
sim.force()
L.atlj_debug_tile2_stats(out)
it, pc, nw, emp, half, warps, mx, mxh = [int(x) for x in out]
print("N=%d warp batches=%d  (per batch:) executed iterations %.2f | max_lane[sum ceil(popc/2) + empty words] %.2f | "
      "max_lane ceil(pairs/2) %.2f | mean_lane pairs/2 %.2f" % (n, warps, it / warps, mx / warps, mxh / warps,
                                                               pc / warps / 64.0))
print("per lane: words %.2f, empty words %.2f, candidates surviving the filter %.2f, sum ceil(popc/2) %.2f"
      % (nw / warps / 32.0, emp / warps / 32.0, pc / warps / 32.0, half / warps / 32.0))
sim.close()
