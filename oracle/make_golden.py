#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz by IMPORTING the unmodified reference modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

The reference has no golden vectors of its own, so these outputs of
python_examples/md_lj_module.py, md_lj_ll_module.py, md_lj_le_module.py and
md_lj_llle_module.py on deterministic configurations are what pins both the C oracle and the
CUDA path.  Inputs are stored next to the outputs, so tests never re-derive them.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ATLJ_REFERENCE", "/root/reference/python_examples")
sys.path.insert(0, ROOT)

from examples_b200.lattice import fcc_positions, perturb, ran_velocities  # noqa: E402


def ref_module(name):
    """Import a reference module by file path so the drop-in of the same name is never picked up."""
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def lj_totals(t):
    return np.array([t.cut, t.pot, t.vir, t.lap])


def le_totals(t):
    return np.array([t.pot, t.vir, t.lap, t.pyx])


def main():
    import importlib.util  # noqa: F401
    lj, ll = ref_module("md_lj_module"), ref_module("md_lj_ll_module")
    le, llle = ref_module("md_lj_le_module"), ref_module("md_lj_llle_module")
    out = {}
    r_cut = 2.5

    # ---- LJ, rho=0.75, r_cut=2.5 (SURVEY.md Appendix B configurations) -----------------
    for nc, amps in ((4, (0.0, 0.1)), (6, (0.0, 0.1)), (10, (0.0, 0.1))):
        n, box, r0 = fcc_positions(nc, 0.75)
        for amp in amps:
            tag = "lj_n%d_a%g" % (n, amp)
            r = perturb(r0, box, amp)
            out[tag + "_box"] = np.float64(box)
            out[tag + "_r"] = r
            if n <= 864:
                t, f = lj.force(box, r_cut, r.copy())
                out[tag + "_ap_totals"], out[tag + "_ap_f"], out[tag + "_ap_ovr"] = lj_totals(t), f, np.bool_(t.ovr)
                out[tag + "_ap_hes"] = np.float64(lj.hessian(box, r_cut, r.copy(), f))
            if n >= 864:
                t, f = ll.force(box, r_cut, r.copy())
                out[tag + "_ll_totals"], out[tag + "_ll_f"], out[tag + "_ll_ovr"] = lj_totals(t), f, np.bool_(t.ovr)
                out[tag + "_ll_hes"] = np.float64(ll.hessian(box, r_cut, r.copy(), f))
                # integer work: cell triplets and per-cell atom lists as md_lj_ll_module.py:106-120 builds them
                sc = int(np.floor(box / r_cut))
                rw = r - np.rint(r)
                c = np.floor((rw + 0.5) * sc).astype(np.int_)
                ids = np.arange(n)
                lists = []
                from itertools import product
                for ci in product(range(sc), repeat=3):
                    lists.append(ids[np.all(c == ci, axis=1)])
                out[tag + "_sc"] = np.int64(sc)
                out[tag + "_c"] = c
                out[tag + "_perm"] = np.concatenate(lists).astype(np.int32)
                out[tag + "_cell_start"] = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int32)
        print("LJ nc=%d done" % nc, flush=True)

    # neighbour cell ids exactly as the reference computes them (ravel_multi_index, mode='wrap')
    d14 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [-1, 1, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 1], [1, 0, 1],
                    [-1, -1, 1], [0, -1, 1], [1, -1, 1], [-1, 1, 1], [0, 1, 1], [1, 1, 1]])
    d17 = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 1], [-1, 0, 1], [0, 0, 1], [1, 1, -1], [1, 1, 0], [1, 1, 1],
                    [0, 1, -1], [0, 1, 0], [0, 1, 1], [-1, 1, -1], [-1, 1, 0], [-1, 1, 1], [-2, 1, -1], [-2, 1, 0],
                    [-2, 1, 1]])
    for sc in (3, 4, 7):
        nb = np.empty((sc ** 3, 14), dtype=np.int32)
        for ci1 in range(sc ** 3):
            ci = np.unravel_index(ci1, (sc, sc, sc))
            for k, dj in enumerate(d14):
                nb[ci1, k] = np.ravel_multi_index(ci + dj, (sc, sc, sc), mode="wrap")
        out["nb14_sc%d" % sc] = nb
    for sc, shift in ((5, 0), (5, 2), (8, -3), (8, 3)):
        nb = np.full((sc ** 3, 17), -1, dtype=np.int32)
        for ci1 in range(sc ** 3):
            ci = np.unravel_index(ci1, (sc, sc, sc))
            if ci[1] == sc - 1:
                dd = d17.copy()
                dd[5:, 0] = d17[5:, 0] - shift
            else:
                dd = d17[:-3, :].copy()
            for k, dj in enumerate(dd):
                nb[ci1, k] = np.ravel_multi_index(ci + dj, (sc, sc, sc), mode="wrap")
        out["nb17_sc%d_shift%d" % (sc, shift)] = nb

    # ---- LE / WCA, rho=0.8442, amp=0.1 ---------------------------------------------------
    for nc in (4, 6):
        n, box, r0 = fcc_positions(nc, 0.8442)
        for strain in (0.0, 0.13, -0.37, 0.499):
            tag = "le_n%d_s%g" % (n, strain)
            r = perturb(r0, box, 0.1)
            r[:, 0] = r[:, 0] - np.rint(r[:, 1]) * strain
            r = r - np.rint(r)
            out[tag + "_box"] = np.float64(box)
            out[tag + "_r"] = r
            t, f = le.force(box, strain, r.copy())
            out[tag + "_ap_totals"], out[tag + "_ap_f"], out[tag + "_ap_ovr"] = le_totals(t), f, np.bool_(t.ovr)
            t, f = llle.force(box, strain, r.copy())
            out[tag + "_ll_totals"], out[tag + "_ll_f"], out[tag + "_ll_ovr"] = le_totals(t), f, np.bool_(t.ovr)
        print("LE nc=%d done" % nc, flush=True)

    # ---- NVE trajectory through the reference's own loop body (md_nve_lj.py:178-192) ----------
    for nc, mod, nsteps, tag in ((4, lj, 200, "nve_n256_ap"), (6, ll, 40, "nve_n864_ll")):
        n, box, r = fcc_positions(nc, 0.75)
        v = ran_velocities(n, 1.0, seed=20170101)
        dt = 0.005
        out[tag + "_box"], out[tag + "_r0"], out[tag + "_v0"] = np.float64(box), r.copy(), v.copy()
        total, f = mod.force(box, r_cut, r)
        rec = np.empty((nsteps, 7))
        for s in range(nsteps):
            v = v + 0.5 * dt * f
            r = r + dt * v / box
            r = r - np.rint(r)
            total, f = mod.force(box, r_cut, r)
            assert not total.ovr
            v = v + 0.5 * dt * f
            rec[s] = [total.cut, total.pot, total.vir, total.lap, np.sum(v ** 2), np.sum(f ** 2),
                      mod.hessian(box, r_cut, r, f)]
        out[tag + "_rec"], out[tag + "_r1"], out[tag + "_v1"] = rec, r, v
        print(tag, "done", flush=True)

    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays; numpy", np.__version__)


if __name__ == "__main__":
    main()
