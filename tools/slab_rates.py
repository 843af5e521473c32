"""Step rates of the slab integrators at the weak-scaling size of bench.py (2.05 M atoms per GPU), under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_rates.py [steps]
NVE (replayed graph), NVE with the hessian in the loop, BAOAB, Nose-Hoover chain: ms per step, max over ranks."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import DT, R_CUT, make_system, workload_nc  # noqa: E402
from examples_b200 import _lib  # noqa: E402
from examples_b200.slab import SlabSimulation  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_lib.check(_lib.lib().atlj_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n, box, r, v = make_system(workload_nc("c5", world))
sim = SlabSimulation(box, R_CUT, r, v, rank, world)
del r, v
sim.force()
sim.run_nve(DT, 500, record=False)


def timed(fn):
    fn(5)
    sim.sync()
    dist.barrier()
    sim.sync()
    sim.timer_start()
    fn(steps)
    ms = sim.timer_stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


m, g = 3, 3 * (n - 1)
q = np.full(m, 4.0)
q[0] = g * 4.0
eta, p_eta = np.zeros(m), np.zeros(m)
out = {"nve": timed(lambda k: sim.run_nve(DT, k, record=False)),
       "nve_hessian": timed(lambda k: sim.run_nve(DT, k, hessian=True)),
       "baoab": timed(lambda k: sim.run_baoab(DT, 1.0, 1.0, k, seed=3)),
       "nhc": timed(lambda k: sim.run_nhc(DT, 1.0, eta, p_eta, q, k))}
if rank == 0:
    print("slab rates: world=%d N=%d  " % (world, n) + "  ".join("%s %.4f ms/step (%.3g atom-steps/s)"
                                                                 % (k, ms, n / (ms * 1e-3)) for k, ms in out.items()),
          flush=True)
sim.close()
dist.destroy_process_group()
