"""Slab domain decomposition of the LJ hot path across the GPUs of one box (one rank per GPU).

No reference counterpart: the reference's only distributed program is replica exchange.  The cell grid is
the reference's (sc = floor(box/r_cut), md_lj_ll_module.py:106); rank k owns the contiguous x layers
[x_lo, x_hi) = partition(sc, world)[k] and keeps one halo layer on each side.  The per-step exchange (migration,
halo) runs inside libatlj on the sim's stream: the pack kernels write straight into the neighbours' receive buffers
over NVLink (csrc/mg.cu); the step records are all-reduced once per run with NCCL.  This module holds the host-side
logic: the partition, the initial scatter of a configuration, capacity sizing, the bootstrap over torch.distributed
(CUDA IPC handles of the receive buffers, NCCL unique id), and a single-GPU emulation of several ranks in one process
used by the GPU tests.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from .device import NREC, Simulation


def partition(sc, world):
    """x cell layers of each rank: contiguous, as even as possible, every rank at least one layer."""
    assert world >= 1 and sc >= world, "need at least one cell layer per rank"
    base, extra = divmod(sc, world)
    out, lo = [], 0
    for k in range(world):
        hi = lo + base + (1 if k < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def x_layer(r, sc):
    """Global x cell layer of every atom, with the reference's expression (md_lj_ll_module.py:88,108)."""
    x = r[:, 0] - np.rint(r[:, 0])
    return np.floor((x + 0.5) * sc).astype(np.int64)


def owned_mask(r, sc, x_lo, x_hi):
    c0 = x_layer(r, sc)
    return (c0 >= x_lo) & (c0 < x_hi)


def halo_mask(r, sc, x_lo, x_hi):
    """Atoms of the two layers adjacent to the slab (periodic)."""
    c0 = x_layer(r, sc)
    return (c0 == (x_lo - 1) % sc) | (c0 == x_hi % sc)


def capacities(n_total, sc, nx_own, slack=1.5):
    """Per-rank buffer sizes (atoms): owned + halo capacity, migration message, halo layer.  Every rank of a run
    must use the same numbers (the neighbours write into each other's arrays): pass the LARGEST nx_own of the
    partition."""
    per_layer = n_total / sc
    halo_cap = int(per_layer * slack) + 1024
    halo_cap += halo_cap & 1
    # migration message: in a liquid ~1e-3 of a layer crosses a face per step, but a run that starts from the lattice
    # can have a whole atomic plane sitting on a slab boundary, half of which crosses at once
    mig_cap = max(4096, int(per_layer * 0.6))
    cap = int(per_layer * nx_own * 1.25) + 2 * halo_cap + 2 * mig_cap + 1024
    return cap, mig_cap, halo_cap


class SlabRank(Simulation):
    """One rank's share of the system: a Simulation restricted to x layers [x_lo, x_hi) with halo/migration buffers."""

    def __init__(self, box, r_cut, r_all, v_all, rank, world, model="lj", stream=0, ipc=False):
        self.rank, self.world = rank, world
        n_total = r_all.shape[0]
        sc = math.floor(box / r_cut)
        assert sc >= 4, "the cell route needs sc >= 4"
        parts = partition(sc, world)
        self.x_lo, self.x_hi = parts[rank]
        r_all = np.ascontiguousarray(r_all - np.rint(r_all))
        mine = np.flatnonzero(owned_mask(r_all, sc, self.x_lo, self.x_hi)).astype(np.int32)
        cap, mig_cap, halo_cap = capacities(n_total, sc, max(hi - lo for lo, hi in parts))
        v = None if v_all is None else np.ascontiguousarray(v_all[mine])
        super().__init__(box, r_cut, np.ascontiguousarray(r_all[mine]), v, model=model, ids=mine, capacity=cap,
                         stream=stream, slab=(self.x_lo, self.x_hi), ipc=ipc)
        self.n_total = n_total
        L = _lib.lib()
        L.atlj_sim_set_total(self._h, n_total)
        _lib.check(L.atlj_sim_mg_enable(self._h, mig_cap, halo_cap))

    # ---- the three device phases (emulation) and their buffers ----------------------------------------------------
    def phase1(self, dt):
        _lib.check(_lib.lib().atlj_sim_mg_phase1(self._h, float(dt)))

    def phase2(self):
        _lib.check(_lib.lib().atlj_sim_mg_phase2(self._h))

    def phase3(self, dt):
        rec = np.zeros(NREC)
        _lib.check(_lib.lib().atlj_sim_mg_phase3(self._h, float(dt), rec.ctypes.data))
        return rec

    def set_peers(self, left, right):
        """Neighbours in the same process (emulation): they write into each other's arrays through plain pointers."""
        _lib.check(_lib.lib().atlj_sim_mg_set_peers(self._h, left._h, right._h))

    # ---- one process per GPU ----------------------------------------------------------------------------------------
    def ipc_handle(self):
        """CUDA IPC handles of this rank's receive arena, position buffers and cell table + the layout."""
        buf = C.create_string_buffer(int(_lib.lib().atlj_sim_mg_blob_bytes()))
        _lib.check(_lib.lib().atlj_sim_mg_ipc_blob(self._h, buf))
        return buf.raw

    def ipc_open(self, left_handle, right_handle):
        _lib.check(_lib.lib().atlj_sim_mg_ipc_open(self._h, left_handle, right_handle))

    def connect(self, unique_id):
        _lib.check(_lib.lib().atlj_sim_mg_connect(self._h, self.rank, self.world, unique_id))

    def run_nve(self, dt, nsteps, hessian=False, record=True):
        """Velocity-Verlet steps of the whole system (md_nve_lj.py:178-192); hessian=True adds the hessian sum of
        md_nve_lj.py:46 to every record (REC_HES), with the forces of the halo atoms fetched from the neighbours."""
        rec = np.zeros((nsteps, NREC)) if (record or hessian) else None
        fn = _lib.lib().atlj_sim_run_nve_mg_hess if hessian else _lib.lib().atlj_sim_run_nve_mg
        _lib.check(fn(self._h, float(dt), int(nsteps), None if rec is None else rec.ctypes.data))
        return rec

    @property
    def own_cap(self):
        """Owned atoms this rank has room for (the size of the host arrays nve_step_host works on)."""
        return int(_lib.lib().atlj_sim_mg_own_cap(self._h))

    def nve_step_host(self, dt, r, v, f, ids, n):
        """One velocity-Verlet step of the whole system with this rank's owned atoms in HOST arrays of own_cap rows
        (the first n are valid; device order); -> (record of global sums, new n)."""
        rec = np.zeros(NREC)
        nn = C.c_int64(int(n))
        for a in (r, v, f):
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape[0] >= self.own_cap
        assert ids.dtype == np.int32 and ids.shape[0] >= self.own_cap
        _lib.check(_lib.lib().atlj_sim_nve_step_host_mg(self._h, float(dt), r.ctypes.data, v.ctypes.data, f.ctypes.data,
                                                        ids.ctypes.data, C.byref(nn), rec.ctypes.data))
        return rec, int(nn.value)

    def force(self, strain=0.0, hessian=False):
        """Force evaluation on the current positions = one step with dt = 0 (kick and drift are exact no-ops)."""
        return self.run_nve(0.0, 1, hessian=hessian)[0]

    def phase3_force(self):
        _lib.check(_lib.lib().atlj_sim_mg_phase3_force(self._h))

    def phase4_hess(self, dt):
        rec = np.zeros(NREC)
        _lib.check(_lib.lib().atlj_sim_mg_phase4_hess(self._h, float(dt), rec.ctypes.data))
        return rec

    def run_baoab(self, dt, gamma, temperature, nsteps, seed=1, noise=None):
        """BAOAB Langevin steps on the slab (bd_nvt_lj.py:221-233); the device noise is keyed by the GLOBAL atom id, so
        the trajectory is that of the single-domain run with the same seed."""
        assert noise is None, "injected noise is a single-domain test facility"
        rec = np.zeros((nsteps, NREC))
        decay, amp = self._o_scalars(dt, gamma, temperature)
        _lib.check(_lib.lib().atlj_sim_run_baoab_mg(self._h, float(dt), decay, amp, int(seed), int(nsteps),
                                                    rec.ctypes.data))
        return rec

    def run_nhc(self, dt, temperature, eta, p_eta, q, nsteps, g=None):
        """Nose-Hoover chain steps on the slab (md_nvt_lj.py:270-288); eta, p_eta updated in place on every rank."""
        rec = np.zeros((nsteps, NREC))
        for a in (eta, p_eta, q):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        if g is None:
            g = 3 * (self.n_total - 1)
        _lib.check(_lib.lib().atlj_sim_run_nhc_mg(self._h, float(dt), float(temperature), float(g), len(q),
                                                  eta.ctypes.data, p_eta.ctypes.data, q.ctypes.data, int(nsteps),
                                                  rec.ctypes.data))
        return rec


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _lib.check(_lib.lib().atlj_nccl_unique_id(buf))
    return buf.raw


class SlabSimulation(SlabRank):
    """The rank of THIS process in a torch.distributed job (one process per GPU): bootstraps libatlj's own NCCL
    communicator by broadcasting the unique id through torch.distributed."""

    def __init__(self, box, r_cut, r_all, v_all, rank, world, model="lj", stream=0):
        import torch.distributed as dist
        super().__init__(box, r_cut, r_all, v_all, rank, world, model=model, stream=stream, ipc=True)
        handles = [None] * world
        dist.all_gather_object(handles, self.ipc_handle())   # map the neighbours' receive buffers (CUDA IPC)
        self.ipc_open(handles[(rank - 1) % world], handles[(rank + 1) % world])
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        self.connect(ids[0])
        dist.barrier()


class EmulatedSlabs:
    """`world` ranks on ONE GPU in one process, driven in lockstep: every rank finishes a phase before any rank
    starts the next, so the messages a phase waits for have always been published already.  Same kernels, same
    peer-memory writes and stamps as the multi-process path; the all-reduce is a host sum."""

    def __init__(self, box, r_cut, r, v, world, model="lj"):
        self.world = world
        self.ranks = [SlabRank(box, r_cut, r, v, k, world, model=model) for k in range(world)]
        for k, s in enumerate(self.ranks):
            s.set_peers(self.ranks[(k - 1) % world], self.ranks[(k + 1) % world])

    def step(self, dt, hessian=False):
        for s in self.ranks:
            s.phase1(dt)
        for s in self.ranks:
            s.phase2()
        if hessian:   # every rank sends its boundary forces before any rank's hessian kernel waits for them
            for s in self.ranks:
                s.phase3_force()
            rec = sum(s.phase4_hess(dt) for s in self.ranks)
        else:
            rec = sum(s.phase3(dt) for s in self.ranks)
        rec[7] = 1.0 if rec[7] > 0 else 0.0
        return rec

    def force(self, hessian=False):
        return self.step(0.0, hessian=hessian)

    def run_nve(self, dt, nsteps, hessian=False):
        return np.array([self.step(dt, hessian=hessian) for _ in range(nsteps)])

    def _finish(self, dt):
        for s in self.ranks:
            s.phase2()
        rec = sum(s.phase3(dt) for s in self.ranks)
        rec[7] = 1.0 if rec[7] > 0 else 0.0
        return rec

    def run_baoab(self, dt, gamma, temperature, nsteps, seed=1):
        decay, amp = Simulation._o_scalars(dt, gamma, temperature)
        out = []
        for k in range(nsteps):
            for s in self.ranks:
                _lib.check(_lib.lib().atlj_sim_mg_phase1_baoab(s._h, float(dt), decay, amp, int(seed), int(k == 0)))
            out.append(self._finish(dt))
        return np.array(out)

    def run_nhc(self, dt, temperature, eta, p_eta, q, nsteps, g=None):
        L = _lib.lib()
        n = self.ranks[0].n_total
        if g is None:
            g = 3 * (n - 1)
        m = len(q)
        ksq = sum(float(np.sum(s.download_sorted()[1] ** 2)) for s in self.ranks)   # global sum v^2 at the start
        for s in self.ranks:
            _lib.check(L.atlj_sim_mg_nhc_open(s._h, m, eta.ctypes.data, p_eta.ctypes.data, q.ctypes.data, ksq))
        out = []
        for k in range(nsteps):
            for s in self.ranks:
                _lib.check(L.atlj_sim_mg_phase1_nhc(s._h, float(dt), float(temperature), float(g), m, int(k == 0)))
            rec = self._finish(dt)
            recs = []
            for s in self.ranks:                                                     # every rank: same global record
                rk = rec.copy()
                _lib.check(L.atlj_sim_mg_nhc_close(s._h, float(dt), float(temperature), float(g), m, rk.ctypes.data))
                recs.append(rk)
            assert all(np.array_equal(recs[0], rk) for rk in recs)
            out.append(recs[0])
        for s in self.ranks:
            e, p = eta.copy(), p_eta.copy()
            _lib.check(L.atlj_sim_mg_nhc_finish(s._h, m, e.ctypes.data, p.ctypes.data))
        eta[:], p_eta[:] = e, p
        return np.array(out)

    def gather(self):
        """r, v, f of the whole system in the original atom order."""
        n = self.ranks[0].n_total
        r, v, f = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
        seen = np.zeros(n, dtype=np.int64)
        for s in self.ranks:
            rs, vs, fs, ids = s.download_sorted()
            r[ids], v[ids], f[ids] = rs, vs, fs
            seen[ids] += 1
        assert np.all(seen == 1), "every atom must be owned by exactly one rank"
        return r, v, f

    def close(self):
        for s in self.ranks:
            s.close()
