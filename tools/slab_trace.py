"""Where a slab step spends its time, waits included (run under torchrun with ATLJ_MG_TRACE=1):
    ATLJ_MG_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/slab_trace.py [steps]
Every rank prints the mean time between the %globaltimer marks libatlj leaves between the launches of a step
(one-thread kernels, also inside the replayed graph): integration kernel, arrivals (waits for the neighbours'
migration messages), scan, scatter + sort + gather (sends the halo layers), pair kernel (waits for the halo stamps
when it reaches a boundary layer), and the whole step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import DT, NC_WEAK, R_CUT, make_system  # noqa: E402
from examples_b200 import _lib  # noqa: E402
from examples_b200.slab import SlabSimulation  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_lib.check(_lib.lib().atlj_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n, box, r, v = make_system(NC_WEAK[world])
sim = SlabSimulation(box, R_CUT, r, v, rank, world)
sim.force()
sim.run_nve(DT, 600, record=False)
dist.barrier()
sim.timer_start()
sim.run_nve(DT, steps, record=False)
ms = sim.timer_stop()
tr = np.zeros((steps, 8), dtype=np.uint64)
_lib.check(_lib.lib().atlj_sim_mg_trace(sim._h, steps, tr.ctypes.data))
t = tr[4:, :6].astype(np.int64)      # skip the directly launched first steps
d = np.diff(t, axis=1).mean(axis=0) / 1e3
step = np.diff(t[:, 0]).mean() / 1e3
names = ["integrate+migrate", "arrivals(wait)", "scan", "scatter+sort+gather+halo", "pair(+halo wait)"]
print("rank %d: step %.1f us (%.4f ms/step by events); " % (rank, step, ms / steps)
      + ", ".join("%s %.1f" % (nm, x) for nm, x in zip(names, d))
      + ", gap to next step %.1f" % (step - d.sum()), flush=True)
sim.close()
dist.destroy_process_group()
