#!/usr/bin/env python
"""bench.py -- LJ atom-steps/s of the device-resident NVE loop (cell build + pair force + velocity Verlet).

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            our arm (libatlj.so, sm_100a CUDA)
  python bench.py --impl reference --gpus N --steps K ...  the CPU restatement of the reference path on host cores

Workloads (BASELINE.json configs; --workload):
  c5 (default)  NVE LJ weak scaling, fcc nc=80 -> 2,048,000 atoms per GPU (nc=101/127/160 at 2/4/8 GPUs), rho=0.75,
                r_cut=2.5, dt=0.005, melted before timing.  This is the configuration the metric
                ("atom-steps/s at 1/2/4/8 B200") is quoted on; its state (~350 MB) exceeds the 126 MB L2.
  c2            md_nve_lj + md_lj_ll_module, N=32,000 (launch-latency bound on a B200).
  c4            N=1,000,188.
The default line also carries "other_workloads": short runs of c1 (N=256 all pairs), c2, c3 (SLLOD, WCA, N=256,000) and
c4 (BAOAB, N=1,000,188; plus the constant-NPT loop on the same system) for information.
A "step" is one velocity-Verlet step of md_nve_lj.py:178-192: kick, drift+wrap, cell build, pair force, kick, sums.

`value`    atoms * steps / device time of the device-resident loop (CUDA events on the launching stream, max over
           ranks, barrier + synchronize on both sides).
`e2e`      same step, but the state (r, v, f, id) lives in pinned HOST memory as it does in the reference driver:
           every step copies it to the device, steps, and copies r, v, f, id + the per-step record back.
`roofline` pair kernel: algorithmic flops (45 per in-range pair) / its CUDA-event time, against the
           fp64 DFMA peak measured in the same run (MEASURED_PEAKS.json has no fp64 figure).
           `roofline_hbm`: cell build + integration kernels, 152 B/atom-step, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`  the C restatement of the reference (oracle/, "port": gfortran is not in the image) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R_CUT, DT, DENSITY = 2.5, 0.005, 0.75
FLOP_PER_PAIR = 45.0
HBM_BYTES_NON_PAIR = 152.0       # 200 B/atom-step minus the pair kernel's 48
NC_WEAK = {1: 80, 2: 101, 4: 127, 8: 160}
T_INIT = 2.0                     # lattice start temperature: melts to a T ~ 1 liquid (reported in config)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].strip().lower() == "active":
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_system(nc, seed=20170101):
    from examples_b200.lattice import fcc_positions, ran_velocities
    n, box, r = fcc_positions(nc, DENSITY)
    v = ran_velocities(n, T_INIT, seed)
    return n, box, r, v


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the C restatement of md_nve_lj.py + md_lj_ll_module on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_config(wl, nc, world):
    """The descriptor both arms print (nothing run-dependent in it, so the two lines can be compared key by key)."""
    n = 4 * nc ** 3
    box = (n / DENSITY) ** (1.0 / 3.0)
    n_rank = n / world
    return {"workload": "%s: NVE LJ, N=%d (fcc nc=%d), rho=0.75, r_cut=2.5, dt=0.005, link cells sc=%d"
                        % (wl, n, nc, int(np.floor(box / R_CUT))),
            "atoms_per_gpu": n_rank,
            "l2": ("state %.0f MB per GPU exceeds the 126 MB L2" % (n_rank * 24 * 7 / 1e6)) if n_rank * 24 * 7 > 126e6
            else "state fits in L2 (device-resident loop)",
            "parallelism": "slab%d" % world if world > 1 else "single"}


def cpu_nve(nc, steps, warmup, melt_state=None, threads=None):
    """-> atom-steps/s, threads, n.  Liquid-like input: melt_state (r, v) if given, else a perturbed lattice."""
    from oracle import oracle as O
    O.build()
    # explicit team size: torch.distributed.run exports OMP_NUM_THREADS=1 to its children
    O.set_num_threads(threads if threads else host_cores())
    n, box, r, v = make_system(nc)
    if melt_state is not None:
        r, v = melt_state
        n = r.shape[0]
    else:
        rng = np.random.default_rng(1)
        r = r + rng.normal(0.0, 0.08 / box, size=r.shape)
        r = np.ascontiguousarray(r - np.rint(r))
        v = v * np.sqrt(1.0 / T_INIT)
    r, v = np.ascontiguousarray(r.copy()), np.ascontiguousarray(v.copy())
    _, f, _, _ = O.force(r, box, R_CUT, cells=True, omp=True)
    if warmup > 0:
        O.nve_steps(r, v, f, box, R_CUT, DT, warmup, cells=True, omp=True, record=False)
    t0 = time.perf_counter()
    O.nve_steps(r, v, f, box, R_CUT, DT, steps, cells=True, omp=True, record=False)
    dt = time.perf_counter() - t0
    return n * steps / dt, O.num_threads(True), n, dt


def workload_nc(wl, world):
    if wl == "c5":
        return NC_WEAK[world]
    return {"c2": 20, "c4": 63}[wl]


def run_reference(args, rank, world):
    if rank != 0:
        return
    # the SAME system as our arm at this N (2,048,000 atoms at 1 GPU ... 16,384,000 at 8): every step is one
    # velocity-Verlet step of the whole system on all host cores; K + W steps stay within minutes (0.4 s per step
    # and 2 M atoms on 16 threads).  Not melted (2,000 CPU steps would take a quarter of an hour): the lattice is
    # disordered with Gaussian displacements instead, which gives the liquid's pair count.
    nc = workload_nc(args.workload, world)
    value, threads, n, dt = cpu_nve(nc, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "LJ atom-steps/sec (rho=0.75, r_cut=2.5)", "value": value,
        "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic fcc lattice + Gaussian displacement (sigma 0.08), T=1 velocities",
        "config": workload_config(args.workload, nc, world),
        "cpu_baseline": {"value": value, "unit": "atom-steps/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of the whole system, N=%d (fcc nc=%d): velocity Verlet + link-cell "
                                   "force, C restatement of md_nve_lj.py/md_lj_ll_module.py (gcc -O2 -fopenmp, %d "
                                   "threads set explicitly; like md_lj_omp_module.f90 it keeps a private N x 3 "
                                   "force array per thread and allocates its cell arrays every call); gfortran "
                                   "absent so the Fortran twin cannot be built" % (args.steps, n, nc, threads)},
        "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def other_workloads():
    """Short device-resident runs of the other BASELINE.json configurations (information only, not the headline):
    c1 N=256 all pairs, c2 N=32,000 NVE, c3 N=256,000 WCA Lees-Edwards SLLOD at strain rate 0.01, c4 N=1,000,188
    BAOAB with the device Gaussian RNG.  -> {name: {atom_steps_per_s, ms_per_step, n}}"""
    from examples_b200.device import Simulation
    from examples_b200.lattice import fcc_positions, ran_velocities
    out = {}

    def timed(tag, sim, n, fn, steps):
        sim.sync()
        sim.timer_start()
        fn(steps)
        ms = sim.timer_stop()
        out[tag] = {"n": n, "ms_per_step": ms / steps, "atom_steps_per_s": n * steps / (ms * 1e-3)}

    for tag, nc, steps in (("c1_nve_n256_allpairs", 4, 2000), ("c2_nve_n32000", 20, 1000)):
        n, box, r, v = make_system(nc)
        sim = Simulation(box, R_CUT, r, v)
        sim.force()
        sim.run_nve(DT, 1000, record=False)
        timed(tag, sim, n, lambda k: sim.run_nve(DT, k, record=False), steps)
        sim.close()
    n, box, r = fcc_positions(40, DENSITY)
    r = r + np.random.default_rng(1).normal(0.0, 0.05 / box, size=r.shape)  # the perfect lattice has no WCA pair
    r = np.ascontiguousarray(r - np.rint(r))
    sim = Simulation(box, None, r, ran_velocities(n, 1.0), model="wca")
    sim.force(strain=0.0)
    st = {"strain": 0.0}

    def sllod(k):
        _, st["strain"] = sim.run_sllod(DT, 0.01, st["strain"], k)

    sllod(400)
    timed("c3_sllod_wca_n256000_rate0.01", sim, n, sllod, 200)
    sim.close()
    n, box, r, v = make_system(63)
    sim = Simulation(box, R_CUT, r, v)
    sim.force()
    sim.run_baoab(DT, 1.0, 1.0, 500, seed=3)
    timed("c4_baoab_n1000188", sim, n, lambda k: sim.run_baoab(DT, 1.0, 1.0, k, seed=3), 100)
    sim.run_nve(DT, 500, record=False)
    # constant-NPT on the same system (md_npt_lj.py defaults: P = 0.99, tau = tau_baro = 2): one host
    # synchronisation per step, box and cell grid re-derived every step
    m, g = 3, 3 * (n - 1)
    q = np.full(m, 4.0)
    q[0] = g * 4.0
    st = {"eps": 0.0, "p_eps": 0.0}
    ch = [np.zeros(m), np.zeros(m), q, np.zeros(m), np.zeros(m), np.full(m, 4.0)]

    def npt(k):
        _, st["eps"], st["p_eps"] = sim.run_npt(DT, 1.0, 0.99, ch[0], ch[1], ch[2], ch[3], ch[4], ch[5], g * 4.0, box,
                                                st["eps"], st["p_eps"], k)

    npt(50)
    timed("npt_n1000188", sim, n, npt, 100)
    sim.close()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# driver-visible parity records
# ---------------------------------------------------------------------------------------------------------------------
def slab_parity(rank, world, dist, nc=20, nsteps=100, host_steps=5, hess_steps=3):
    """The REAL multi-process slab path (CUDA IPC peer buffers, stamps, NCCL all-reduce) against the single-domain run
    of the same system on rank 0's GPU: N = 32,000 (BASELINE config 2), 100 velocity-Verlet steps.  Every rank takes
    part; rank 0 gets the record, the others None.  Forces are per-atom sums in a fixed order, so r, v, f must agree
    to the last bit; the totals are reduced in a different order (<= 1e-12)."""
    from examples_b200.device import Simulation
    from examples_b200.lattice import fcc_positions, ran_velocities
    from examples_b200.slab import SlabSimulation
    n, box, r = fcc_positions(nc, DENSITY)
    r = r + np.random.default_rng(1).normal(0.0, 0.06 / box, size=r.shape)
    r = np.ascontiguousarray(r - np.rint(r))
    v = ran_velocities(n, 1.5)
    sim = SlabSimulation(box, R_CUT, r, v, rank, world)
    rec0 = sim.force()
    rec = sim.run_nve(DT, nsteps)
    # ... and host_steps more through the host-buffer step (the e2e arm's entry point at N > 1): same trajectory
    rs, vs, fs, ids = sim.download_sorted()
    n_own, cap = ids.shape[0], sim.own_cap
    R, V, F = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 3))
    I = np.zeros(cap, dtype=np.int32)
    R[:n_own], V[:n_own], F[:n_own], I[:n_own] = rs, vs, fs, ids
    rec_h = np.zeros((host_steps, rec.shape[1]))
    for k in range(host_steps):
        rec_h[k], n_own = sim.nve_step_host(DT, R, V, F, I, n_own)
    # ... and hess_steps with the hessian in the loop (forces of the halo atoms fetched from the neighbours)
    rec_x = sim.run_nve(DT, hess_steps, hessian=True)
    parts = [None] * world
    dist.all_gather_object(parts, sim.download_sorted())
    sim.close()
    if rank != 0:
        return None
    from examples_b200 import _lib
    old = _lib.set_multi_max(0)  # the same (tiled) kernel as on the slabs: bit-identical per-atom sums are the claim
    one = Simulation(box, R_CUT, r, v)
    ref0 = one.force()
    ref = one.run_nve(DT, nsteps)
    ref_h = one.run_nve(DT, host_steps)
    ref_x = one.run_nve(DT, hess_steps, hessian=True)
    r1, v1, f1 = one.download()
    one.close()
    _lib.set_multi_max(old)
    rm, vm, fm = np.empty_like(r1), np.empty_like(r1), np.empty_like(r1)
    seen = np.zeros(n, dtype=np.int64)
    for a, b, c, i in parts:
        rm[i], vm[i], fm[i] = a, b, c
        seen[i] += 1
    tot = max(float(np.max(np.abs(rec[:, k] - ref[:, k]) / np.maximum(np.abs(ref[:, k]), 1e-300))) for k in range(6))
    tot = max(tot, float(np.max(np.abs(rec0[:4] - ref0[:4]) / np.abs(ref0[:4]))))
    tot = max([tot] + [float(np.max(np.abs(rec_h[:, k] - ref_h[:, k]) / np.maximum(np.abs(ref_h[:, k]), 1e-300)))
                       for k in range(6)])
    tot = max([tot] + [float(np.max(np.abs(rec_x[:, k] - ref_x[:, k]) / np.maximum(np.abs(ref_x[:, k]), 1e-300)))
                       for k in range(6)])
    hes = float(np.max(np.abs(rec_x[:, 6] - ref_x[:, 6]) / np.abs(ref_x[:, 6])))
    return {"world": world, "n": n, "steps": nsteps, "host_buffer_steps": host_steps, "hessian_steps": hess_steps,
            "max_rel_diff_hessian": hes,
            "against": "single-domain run of the same system on rank 0 (%d device-resident steps, %d steps "
                       "through atlj_sim_nve_step_host_mg, %d steps with the hessian in the loop)"
                       % (nsteps, host_steps, hess_steps),
            "every_atom_owned_once": bool(np.all(seen == 1)), "owned_count_conserved": bool(np.all(rec[:, 11] == n)),
            "pair_counts_equal": bool(np.array_equal(rec[:, 8], ref[:, 8]) and rec0[8] == ref0[8]
                                      and np.array_equal(rec_h[:, 8], ref_h[:, 8])
                                      and np.array_equal(rec_x[:, 8], ref_x[:, 8])),
            "max_abs_diff_r": float(np.max(np.abs(rm - r1))), "max_abs_diff_v": float(np.max(np.abs(vm - v1))),
            "max_abs_diff_f": float(np.max(np.abs(fm - f1))), "max_rel_diff_step_records": tot}


def melted_parity(sim, box, n):
    """The state the bench times (melted liquid), checked against the CPU restatement of the reference
    (oracle/, OpenMP): forces, totals, in-range pair count, overlap flag.  oracle/ is the checker here, nothing else."""
    from oracle import oracle as O
    O.build()
    O.set_num_threads(host_cores())
    rec = sim.force()
    r, _, f = sim.download()
    t0 = time.perf_counter()
    tot, fo, ovr, npair = O.force(r, box, R_CUT, cells=True, omp=True)
    sec = time.perf_counter() - t0
    return {"n": n, "against": "oracle/atlj_oracle.c (C restatement of md_lj_ll_module.py:69-224), %d threads, %.2f s"
                               % (O.num_threads(True), sec),
            "pair_counts_equal": bool(int(rec[8]) == int(npair)), "pairs": int(npair),
            "overlap_equal": bool((rec[7] != 0.0) == bool(ovr)),
            "max_rel_diff_totals": float(np.max(np.abs(rec[:4] - tot) / np.abs(tot))),
            "max_abs_diff_f": float(np.max(np.abs(f - fo))),
            "max_rel_diff_f": float(np.max(np.abs(f - fo)) / np.max(np.abs(fo)))}


def reference_python_rows(budget=40.0):
    """The reference's own NumPy routines and its unmodified md_nve_lj.py timed on THIS host (subprocess, so that
    nothing of the reference is imported into the bench process) -> dict for cpu_baseline["reference_python"]."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "time_reference_python.py"), "--budget",
                            str(budget)], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="1"))
        return json.loads(p.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"unavailable": "time_reference_python.py failed: %s" % str(e)[:200]}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from examples_b200 import _lib
    from examples_b200.device import REC_NPAIR, REC_T1, REC_VSQ, Simulation
    if _lib.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libatlj has no CPU path")
    _lib.check(_lib.lib().atlj_set_device(local_rank))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = args.workload
    nc = workload_nc(wl, world)
    parity = slab_parity(rank, world, dist) if world > 1 else None
    n, box, r, v = make_system(nc)

    if world > 1:
        from examples_b200.slab import SlabSimulation
        sim = SlabSimulation(box, R_CUT, r, v, rank, world)
    else:
        sim = Simulation(box, R_CUT, r, v)
    del r, v
    sim.force()
    t_melt0 = time.perf_counter()
    done = 0
    while done < args.melt:  # un-timed: melt the lattice so the pair count is liquid-like
        k = min(500, args.melt - done)
        sim.run_nve(DT, k, record=False)
        done += k
    sim.sync()
    t_melt = time.perf_counter() - t_melt0

    def barrier():
        sim.sync()
        if dist is not None:
            dist.barrier()
        sim.sync()

    # ---- device-resident loop ---------------------------------------------------------------------------------------
    sim.run_nve(DT, args.warmup, record=False)
    sampler = ClockSampler(local_rank)
    fp64_peak = _lib.measure_fp64_peak() if rank == 0 else 0.0
    barrier()
    if rank == 0:
        sampler.start()
    l0 = _lib.kernel_launches()
    sim.timer_start()
    rec = sim.run_nve(DT, args.steps, record=True)
    ms = sim.timer_stop()
    l1 = _lib.kernel_launches()
    barrier()
    # second pass of the same K steps with a CUDA-event bracket around every kernel category (the headline pass above
    # replays the step as a CUDA graph and is timed as a whole): kernel_ms_per_step, roofline
    sim.enable_timing(True)
    sim.timing()
    sim.timer_start()
    sim.run_nve(DT, args.steps, record=False)
    ms_events = sim.timer_stop()
    cat = sim.timing()
    sim.enable_timing(False)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_total = n
    value = n_total * args.steps / (ms * 1e-3)
    pairs_per_step = float(np.mean(rec[:, REC_NPAIR]))
    temp = float(np.mean(rec[:, REC_VSQ])) / (3 * n_total - 3)
    epot = float(np.mean(rec[:, REC_T1])) / n_total
    ovr = bool(rec[:, 7].any())

    # ---- end to end through host buffers ----------------------------------------------------------------------------
    e2e = None
    if world == 1:
        from examples_b200._lib import PinnedArray
        rs, vs, fs, ids = sim.download_sorted()
        pr, pv, pf = PinnedArray((n, 3)), PinnedArray((n, 3)), PinnedArray((n, 3))
        pi = PinnedArray((n,), np.int32)
        pr.array[:], pv.array[:], pf.array[:], pi.array[:] = rs, vs, fs, ids
        ke = max(3, min(args.steps, 50))
        for _ in range(3):
            sim.nve_step_host(DT, pr.array, pv.array, pf.array, pi.array)
        sim.sync()
        t0 = time.perf_counter()
        sim.timer_start()
        for _ in range(ke):
            sim.nve_step_host(DT, pr.array, pv.array, pf.array, pi.array)
        ms_e = sim.timer_stop()
        wall_e = time.perf_counter() - t0
        ms_e = max(ms_e, wall_e * 1e3)  # host-side time counts too
        e2e = {"value": n * ke / (ms_e * 1e-3), "unit": "atom-steps/s", "steps": ke,
               "h2d_bytes_per_step": 3 * 24 * n + 4 * n, "d2h_bytes_per_step": 3 * 24 * n + 4 * n + 8 * _lib.NREC,
               "api": "atlj_sim_nve_step_host: r,v,f,id in pinned host memory each step"}
        # the drop-in force(box, r_cut, r) call with NumPy (pageable) arrays
        from examples_b200 import md_lj_ll_module
        rr = np.ascontiguousarray(pr.array.copy())
        md_lj_ll_module.force(box, R_CUT, rr)
        t0 = time.perf_counter()
        kf = 5
        for _ in range(kf):
            md_lj_ll_module.force(box, R_CUT, rr)
        sec = (time.perf_counter() - t0) / kf
        e2e["force_call_atoms_per_s"] = n / sec
        e2e["force_call"] = {"api": "md_lj_ll_module.force(box, r_cut, r): pageable NumPy r in (host threads copy it "
                                    "chunk-wise into a page-locked ring, DMA per chunk), f returned in a page-locked "
                                    "NumPy array written by DMA", "ms_per_call": sec * 1e3,
                             "pcie_gb_per_s": 2 * 24 * n / sec / 1e9, "bytes_per_call": 2 * 24 * n}
        for a in (pr, pv, pf, pi):
            a.free()
        # the drop-in on the reference's own configurations (same rows as cpu_baseline.reference_python)
        if not args.no_cpu:
            try:
                p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_dropin.py")], capture_output=True,
                                   text=True, timeout=600)
                e2e["dropin_on_reference_configs"] = json.loads(p.stdout.strip().splitlines()[-1])["rows"]
            except Exception as e:  # noqa: BLE001
                e2e["dropin_on_reference_configs"] = {"unavailable": str(e)[:200]}

    if world > 1:
        # every rank's owned atoms live in page-locked host arrays; each step copies them in, runs one exchange step
        # of the whole system and copies them out again (atoms migrate: the owned count travels with the arrays)
        from examples_b200._lib import PinnedArray
        rs, vs, fs, ids = sim.download_sorted()
        n_own, cap = ids.shape[0], sim.own_cap
        pr, pv, pf = PinnedArray((cap, 3)), PinnedArray((cap, 3)), PinnedArray((cap, 3))
        pi = PinnedArray((cap,), np.int32)
        pr.array[:n_own], pv.array[:n_own], pf.array[:n_own], pi.array[:n_own] = rs, vs, fs, ids
        del rs, vs, fs, ids
        ke = max(3, min(args.steps, 50))
        for _ in range(3):
            _, n_own = sim.nve_step_host(DT, pr.array, pv.array, pf.array, pi.array, n_own)
        barrier()
        t0 = time.perf_counter()
        sim.timer_start()
        for _ in range(ke):
            _, n_own = sim.nve_step_host(DT, pr.array, pv.array, pf.array, pi.array, n_own)
        ms_e = sim.timer_stop()
        ms_e = max(ms_e, (time.perf_counter() - t0) * 1e3)  # host-side time counts too
        barrier()
        import torch
        t = torch.tensor([ms_e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e = float(t.item())
        e2e = {"value": n_total * ke / (ms_e * 1e-3), "unit": "atom-steps/s", "steps": ke,
               "h2d_bytes_per_step": (3 * 24 + 4) * n_total, "d2h_bytes_per_step": (3 * 24 + 4) * n_total + 8 * _lib.NREC,
               "api": "atlj_sim_nve_step_host_mg on every rank: the rank's owned r,v,f,id in pinned host memory each "
                      "step (bytes are the sum over the %d ranks)" % world}
        for a in (pr, pv, pf, pi):
            a.free()

    parity_m = melted_parity(sim, box, n) if world == 1 and not args.no_cpu else None

    other = None
    if world == 1 and not args.no_other:
        other = other_workloads()

    if rank != 0:
        return
    hbm_peak, hbm_src = peaks()
    pair_ms = cat["pair"] / args.steps
    other_ms = (cat["cells"] + cat["integrate"]) / args.steps
    n_rank = n_total / world
    flops = FLOP_PER_PAIR * pairs_per_step / world
    ach = flops / (pair_ms * 1e-3) / 1e12 if pair_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r2_pair_traffic.json")
    if os.path.exists(tp) and wl == "c5" and world == 1:  # dram bytes per launch from the committed ncu capture
        td = json.load(open(tp))
        traffic = td["dram_bytes_read"] + td["dram_bytes_write"]
    roofline = {"kernel": "pair_tile2_kernel", "bound": "fp64", "achieved": ach, "peak": fp64_peak / 1e12,
                "unit": "TFLOP/s", "frac": ach / (fp64_peak / 1e12) if fp64_peak > 0 else None, "traffic": traffic,
                "peak_source": "register-resident DFMA probe (atlj_measure_fp64_peak) in this run; "
                               "MEASURED_PEAKS.json has no fp64 figure",
                "ms_per_launch": pair_ms, "flop_per_launch": flops,
                "share_of_step": pair_ms / (ms_events / args.steps),
                "timed_in": "second pass of the same %d steps, every kernel category bracketed by CUDA events on the "
                            "launching stream (%.4f ms/step; the headline pass replays the step as a CUDA graph)"
                            % (args.steps, ms_events / args.steps)}
    ach_h = HBM_BYTES_NON_PAIR * n_rank / (other_ms * 1e-3) / 1e9 if other_ms > 0 else 0.0
    roofline_hbm = {"kernel": "kdk_assign (integration + cell assignment) + scan + scatter + rank/gather: the step "
                              "without the pair kernel and its item table",
                    "bound": "hbm", "achieved": ach_h,
                    "peak": hbm_peak, "unit": "GB/s", "frac": ach_h / hbm_peak, "traffic": None,
                    "peak_source": hbm_src, "ms_per_step": other_ms, "bytes_per_step": HBM_BYTES_NON_PAIR * n_rank}

    cpu = None
    if world == 1 and not args.no_cpu:
        # bounded sample: 25 steps of the SAME system (about 12 s on 16 threads)
        t0 = time.perf_counter()
        cv, threads, cn, cdt = cpu_nve(nc, 25, 2)
        cpu = {"value": cv, "unit": "atom-steps/s", "cores": threads, "kind": "port",
               "sample": "25 velocity-Verlet steps of the whole system, N=%d (fcc nc=%d + Gaussian displacement), C "
                         "restatement of the reference with OpenMP over cells (%d threads set explicitly), %.1f s"
                         % (cn, nc, threads, time.perf_counter() - t0),
               # the reference's own NumPy path on the same host cores; gfortran is not in the
               # image, so the Fortran twin (item 3) cannot be built
               "reference_python": reference_python_rows()}

    line = {
        "metric": "LJ atom-steps/sec (rho=0.75, r_cut=2.5)", "value": value, "unit": "atom-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic: fcc lattice (initialize.py formula), Gaussian velocities seed 20170101 at T=%.1f, "
                "melted %d steps before timing" % (T_INIT, args.melt),
        "config": workload_config(wl, nc, world),
        "state": {"temperature": temp, "pot_per_atom": epot, "pairs_per_atom": pairs_per_step / n_total,
                  "overlap": ovr, "melt_s": t_melt},
        "parity": parity, "parity_melted": parity_m,
        "e2e": e2e, "gpu_launches": int(l1 - l0), "clocks": clocks, "roofline": roofline,
        "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "other_workloads": other,
        "kernel_ms_per_step": dict({k: cat[k] / args.steps for k in cat},
                                   # second (event-bracketed, directly launched) pass: what the brackets do not cover -
                                   # launch gaps, event records, and on a slab the skew between the ranks
                                   unaccounted=ms_events / args.steps - sum(cat.values()) / args.steps,
                                   whole_step_event_pass=ms_events / args.steps),
    }
    emit(line)
    sim.close()
    if dist is not None:
        dist.destroy_process_group()


_JSON_FD = None


def quiet_stdout():
    """Libraries below us write to stdout on their own (NCCL prints its version line there): point file descriptor 1
    at stderr for the run and keep the real stdout for the ONE JSON line."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c2", "c4"])
    ap.add_argument("--melt", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the short runs of the other BASELINE configs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
