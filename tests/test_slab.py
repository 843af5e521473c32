"""Slab domain decomposition.

CPU (gloo, world_size 2): the host-side logic -- partition, ownership, halo selection, ring exchange -- checked by
computing the forces on each rank's owned atoms from owned + received halo atoms with the C oracle and comparing with
the single-domain oracle forces.
GPU: several ranks emulated on one device (device copies in place of NCCL; same kernels, same message formats) against
the single-domain CUDA path: totals, forces, a 60-step trajectory with atoms migrating between slabs."""
import math
import os

import numpy as np
import pytest

from conftest import ROOT, liquid_like, rel_err


def test_partition_covers_all_layers():
    from examples_b200.slab import partition
    for sc in (4, 5, 13, 55, 111):
        for world in (1, 2, 3, 4, 8):
            if sc < world:
                continue
            p = partition(sc, world)
            assert p[0][0] == 0 and p[-1][1] == sc
            assert all(a[1] == b[0] for a, b in zip(p, p[1:]))
            sizes = [hi - lo for lo, hi in p]
            assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, ROOT)
        from examples_b200.slab import halo_mask, owned_mask, partition, x_layer
        from oracle import oracle as O
        n, box, r = liquid_like(10, 0.75, 0.06)  # N = 4000, sc = 6
        sc = math.floor(box / 2.5)
        x_lo, x_hi = partition(sc, world)[rank]
        own = owned_mask(r, sc, x_lo, x_hi)
        ids = np.flatnonzero(own)
        # ring exchange of the boundary layers: first owned layer to the left, last owned layer to the right
        c0 = x_layer(r, sc)
        first, last = np.flatnonzero(c0 == x_lo), np.flatnonzero(c0 == x_hi - 1)
        left, right = (rank - 1) % world, (rank + 1) % world

        def sendrecv(payload, dst, src):
            size = torch.tensor([payload.size], dtype=torch.int64)
            got = torch.zeros(1, dtype=torch.int64)
            reqs = [dist.isend(size, dst), dist.irecv(got, src)]
            [q_.wait() for q_ in reqs]
            buf = torch.empty(int(got.item()), dtype=torch.float64)
            reqs = [dist.isend(torch.from_numpy(payload.ravel().copy()), dst), dist.irecv(buf, src)]
            [q_.wait() for q_ in reqs]
            return buf.numpy().reshape(-1, 4)

        pack = lambda idx: np.concatenate([r[idx], idx[:, None].astype(np.float64)], axis=1)
        from_right = sendrecv(pack(first), left, right)   # the right neighbour's first layer = my right halo
        from_left = sendrecv(pack(last), right, left)
        halo = np.concatenate([from_left, from_right])
        hid = halo[:, 3].astype(np.int64)
        want = np.flatnonzero(halo_mask(r, sc, x_lo, x_hi))
        assert sorted(set(hid.tolist())) == sorted(want.tolist()), "halo = the two adjacent layers"
        # forces on owned atoms from owned + halo (all pairs on the subset, minimum image in the full box)
        sub = np.concatenate([r[ids], halo[:, :3]])
        _, f_sub, _, _ = O.force(np.ascontiguousarray(sub), box, 2.5, cells=False)
        _, f_all, _, _ = O.force(r, box, 2.5, cells=True)
        err = rel_err(f_sub[:len(ids)], f_all[ids])
        cnt = torch.tensor([len(ids)], dtype=torch.int64)
        dist.all_reduce(cnt)
        q.put((rank, err, int(cnt.item()), n))
    finally:
        dist.destroy_process_group()


def test_slab_scheme_two_ranks_gloo(oracle):
    """world_size 2 on CPU: ownership is a partition of the atoms, one halo layer per side is sufficient, the ring
    exchange delivers exactly the adjacent layers."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(k, 2, port, q)) for k in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, err, total, n in res:
        assert err < 1e-10 and total == n


@pytest.mark.gpu
@pytest.mark.parametrize("nc,world", [(10, 2), (10, 3), (20, 4), (14, 8)])
def test_emulated_slabs_match_single_domain(nc, world):
    from examples_b200.device import Simulation
    from examples_b200.lattice import ran_velocities
    from examples_b200 import _lib
    from examples_b200.slab import EmulatedSlabs
    # bit-identical forces need the same summation order: the single domain must take the tiled kernel the slabs use,
    # not the small-system kernel this size would select (the pair_route fixture restores the setting)
    _lib.set_multi_max(0)
    n, box, r = liquid_like(nc, 0.75, 0.06)
    v = ran_velocities(n, 1.5)
    one = Simulation(box, 2.5, r, v)
    rec1 = one.force()
    many = EmulatedSlabs(box, 2.5, r, v, world)
    recm = many.force()
    assert recm[11] == n and recm[8] == rec1[8], "atoms conserved, in-range pair count exact"
    assert rel_err(recm[:4], rec1[:4]) < 1e-10 and recm[7] == rec1[7]
    r1, v1, f1 = one.download()
    rm, vm, fm = many.gather()
    assert np.array_equal(fm, f1) and np.array_equal(rm, r1), "per-atom sums run in the same order on a slab"
    # a trajectory: atoms cross slab boundaries; every step's scalars and the final state agree
    nsteps = 60
    t1 = one.run_nve(0.005, nsteps)
    tm = many.run_nve(0.005, nsteps)
    assert np.all(tm[:, 11] == n)
    for k in (0, 1, 2, 3, 4, 5):
        assert rel_err(tm[:, k], t1[:, k]) < 1e-9, k
    assert np.array_equal(tm[:, 8], t1[:, 8])
    r1, v1, f1 = one.download()
    rm, vm, fm = many.gather()
    assert np.array_equal(rm, r1) and np.array_equal(vm, v1) and np.array_equal(fm, f1), "bit-identical trajectory"
    owners0 = [set(s.download_sorted()[3].tolist()) for s in many.ranks]
    assert sum(len(o) for o in owners0) == n
    one.close(), many.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nc,world", [(10, 2), (10, 3), (14, 4)])
def test_emulated_slabs_hessian_in_the_loop(nc, world):
    """md_nve_lj.py:46 calls hessian(box, r_cut, r, f) every step (md_lj_ll_module.py:227-357): on slabs the pair terms of
    a boundary atom need the forces of its halo partners, which the neighbours store into this rank's force array after
    their pair kernel (second stamp).  The records' hessian column must match the single-domain loop (same per-pair
    arithmetic, different summation order across ranks), and the trajectory stays bit-identical."""
    from examples_b200 import _lib
    from examples_b200.device import REC_HES, Simulation
    from examples_b200.lattice import ran_velocities
    from examples_b200.slab import EmulatedSlabs
    _lib.set_multi_max(0)  # same (tiled) pair kernel on both sides
    n, box, r = liquid_like(nc, 0.75, 0.06)
    v = ran_velocities(n, 1.5)
    one = Simulation(box, 2.5, r, v)
    rec1 = one.force(hessian=True)
    many = EmulatedSlabs(box, 2.5, r, v, world)
    recm = many.force(hessian=True)
    assert rec1[REC_HES] != 0.0
    assert rel_err(recm[[0, 1, 2, 3, REC_HES]], rec1[[0, 1, 2, 3, REC_HES]]) < 1e-10 and recm[8] == rec1[8]
    # against the CPU oracle as well (not only kernel against kernel)
    from oracle import oracle as O
    _, fo, _, _ = O.force(np.ascontiguousarray(r - np.rint(r)), box, 2.5, cells=True)
    hes_o = O.hessian(np.ascontiguousarray(r - np.rint(r)), fo, box, 2.5, cells=True)
    assert abs(recm[REC_HES] - hes_o) <= 1e-10 * abs(hes_o)
    nsteps = 30
    t1 = one.run_nve(0.005, nsteps, hessian=True)
    tm = many.run_nve(0.005, nsteps, hessian=True)
    for k in (0, 1, 2, 3, 4, 5, REC_HES):
        assert rel_err(tm[:, k], t1[:, k]) < 1e-9, k
    assert np.array_equal(tm[:, 8], t1[:, 8]) and np.all(tm[:, 11] == n)
    r1, v1, f1 = one.download()
    rm, vm, fm = many.gather()
    assert np.array_equal(rm, r1) and np.array_equal(vm, v1) and np.array_equal(fm, f1)
    one.close(), many.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_emulated_slabs_baoab_and_nose_hoover(world):
    """SURVEY.md 8e widened: the thermostats on slabs.  BAOAB (bd_nvt_lj.py:221-233) draws its noise from Philox keyed
    by the GLOBAL atom id and the step, the Nose-Hoover chain (md_nvt_lj.py:270-288) is fed the all-reduced sum v^2:
    both must reproduce the single-domain trajectory - positions, velocities, forces bit for bit (per-atom sums run
    in the same order), the step records to summation order."""
    from examples_b200 import _lib
    from examples_b200.device import Simulation
    from examples_b200.lattice import ran_velocities
    from examples_b200.slab import EmulatedSlabs
    _lib.set_multi_max(0)  # same (tiled) pair kernel on both sides
    n, box, r = liquid_like(10, 0.75, 0.06)
    v = ran_velocities(n, 1.2)
    nsteps = 40
    # ---- BAOAB
    one = Simulation(box, 2.5, r, v)
    one.force()
    many = EmulatedSlabs(box, 2.5, r, v, world)
    many.force()
    t1 = one.run_baoab(0.005, 1.0, 1.0, nsteps, seed=11)
    tm = many.run_baoab(0.005, 1.0, 1.0, nsteps, seed=11)
    for k in range(6):
        assert rel_err(tm[:, k], t1[:, k]) < 1e-9, k
    assert np.array_equal(tm[:, 8], t1[:, 8]) and np.all(tm[:, 11] == n)
    r1, v1, f1 = one.download()
    rm, vm, fm = many.gather()
    assert np.array_equal(rm, r1) and np.array_equal(vm, v1) and np.array_equal(fm, f1)
    one.close(), many.close()
    # ---- Nose-Hoover chain (m = 3, the reference's defaults: md_nvt_lj.py:245-257)
    m, g = 3, 3 * (n - 1)
    q = np.full(m, 1.0 * 2.0 ** 2)
    q[0] = g * 1.0 * 2.0 ** 2
    e1, p1 = np.zeros(m), np.array([0.3, -0.2, 0.1])
    em, pm = e1.copy(), p1.copy()
    one = Simulation(box, 2.5, r, v)
    one.force()
    many = EmulatedSlabs(box, 2.5, r, v, world)
    many.force()
    t1 = one.run_nhc(0.005, 1.0, e1, p1, q, nsteps)
    tm = many.run_nhc(0.005, 1.0, em, pm, q, nsteps)
    for k in (0, 1, 2, 3, 4, 5, 10):
        assert rel_err(tm[:, k], t1[:, k]) < 1e-9, k
    assert np.array_equal(tm[:, 8], t1[:, 8])
    assert rel_err(em, e1) < 1e-10 and rel_err(pm, p1) < 1e-10
    r1, v1, f1 = one.download()
    rm, vm, fm = many.gather()
    assert rel_err(rm, r1) < 1e-9 and rel_err(vm, v1) < 1e-9 and rel_err(fm, f1) < 1e-8
    one.close(), many.close()
