"""Multi-GPU parity check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_slab_nccl.py
Every rank runs its slab through libatlj's NCCL path; rank 0 also runs the whole system on its own GPU and compares the
per-step records (global sums) and the final forces/positions gathered from all ranks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from conftest import liquid_like, rel_err  # noqa: E402
from examples_b200 import _lib  # noqa: E402
from examples_b200.device import Simulation  # noqa: E402
from examples_b200.lattice import ran_velocities  # noqa: E402
from examples_b200.slab import SlabSimulation  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_lib.check(_lib.lib().atlj_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n, box, r = liquid_like(nc, 0.75, 0.06)
v = ran_velocities(n, 1.5)
sim = SlabSimulation(box, 2.5, r, v, rank, world)
rec0 = sim.force()
rec = sim.run_nve(0.005, nsteps)
rs, vs, fs, ids = sim.download_sorted()
parts = [None] * world
dist.all_gather_object(parts, (rs, vs, fs, ids))
ok = True
if rank == 0:
    one = Simulation(box, 2.5, r, v)
    ref0 = one.force()
    ref = one.run_nve(0.005, nsteps)
    r1, v1, f1 = one.download()
    rm, vm, fm = np.empty_like(r1), np.empty_like(r1), np.empty_like(r1)
    seen = np.zeros(n, dtype=int)
    for a, b, c, i in parts:
        rm[i], vm[i], fm[i] = a, b, c
        seen[i] += 1
    errs = {"totals0": rel_err(rec0[:4], ref0[:4]), "traj": max(rel_err(rec[:, k], ref[:, k]) for k in range(6)),
            "r": rel_err(rm, r1), "v": rel_err(vm, v1), "f": rel_err(fm, f1)}
    ok = (np.all(seen == 1) and np.all(rec[:, 11] == n) and np.array_equal(rec[:, 8], ref[:, 8])
          and errs["totals0"] < 1e-10 and errs["traj"] < 1e-9 and errs["r"] < 1e-9 and errs["v"] < 1e-9
          and errs["f"] < 1e-8)
    print("slab NCCL check: world=%d N=%d steps=%d %s -> %s" % (world, n, nsteps, errs, "OK" if ok else "FAIL"), flush=True)
sim.close()

# ---- the thermostats on slabs: BAOAB (noise keyed by global atom id) and the Nose-Hoover chain (all-reduced sum v^2)
from examples_b200 import _lib as L  # noqa: E402
L.set_multi_max(0)
m, g = 3, 3 * (n - 1)
q = np.full(m, 4.0)
q[0] = g * 4.0
for what in ("baoab", "nhc"):
    sim = SlabSimulation(box, 2.5, r, v, rank, world)
    sim.force()
    eta, p_eta = np.zeros(m), np.array([0.3, -0.2, 0.1])
    rec = sim.run_baoab(0.005, 1.0, 1.0, nsteps, seed=5) if what == "baoab" else sim.run_nhc(0.005, 1.0, eta, p_eta, q, nsteps)
    parts = [None] * world
    dist.all_gather_object(parts, sim.download_sorted())
    if rank == 0:
        one = Simulation(box, 2.5, r, v)
        one.force()
        e1, p1 = np.zeros(m), np.array([0.3, -0.2, 0.1])
        ref = one.run_baoab(0.005, 1.0, 1.0, nsteps, seed=5) if what == "baoab" else one.run_nhc(0.005, 1.0, e1, p1, q, nsteps)
        r1, v1, f1 = one.download()
        one.close()
        rm, vm, fm = np.empty_like(r1), np.empty_like(r1), np.empty_like(r1)
        for a, b, c, i in parts:
            rm[i], vm[i], fm[i] = a, b, c
        errs = {"traj": max(rel_err(rec[:, k], ref[:, k]) for k in range(6)), "r": rel_err(rm, r1), "v": rel_err(vm, v1),
                "f": rel_err(fm, f1)}
        if what == "nhc":
            errs["eta"] = rel_err(eta, e1)
            errs["conserved"] = rel_err(rec[:, 10], ref[:, 10])
        good = all(x < 1e-8 for x in errs.values()) and np.array_equal(rec[:, 8], ref[:, 8])
        ok = ok and good
        print("slab NCCL check (%s): world=%d N=%d steps=%d %s -> %s" % (what, world, n, nsteps, errs, "OK" if good else "FAIL"),
              flush=True)
    sim.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
