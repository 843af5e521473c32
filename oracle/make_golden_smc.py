#!/usr/bin/env python
"""Generate tests/golden/reference_smc.npz with the reference's smc_lj_module (run in the build container only):
force_1 for several atoms of an N = 256 configuration against all the others, one call that ends in an overlap, and the
module's force() on the whole configuration.

    python oracle/make_golden_smc.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_dynamics import ref_module  # noqa: E402
from examples_b200.lattice import fcc_positions, perturb  # noqa: E402


def main():
    smc = ref_module("smc_lj_module")
    n, box, r = fcc_positions(4, 0.75)
    r = perturb(r, box, 0.06)
    out = dict(box=np.float64(box), r=r)
    idx = np.array([0, 17, 101, 255])
    tots, fps = [], []
    for i in idx:
        rj = np.delete(r, i, 0)
        partial, fp = smc.force_1(r[i, :], box, 2.5, rj)
        assert not partial.ovr
        tots.append([partial.cut, partial.pot, partial.vir, partial.lap])
        fps.append(fp)
    out.update(idx=idx, f1_totals=np.array(tots), f1_f=np.array(fps))
    ri = r[3] + np.array([0.0, 0.0, 0.0])
    rj = np.delete(r, 3, 0)
    rj[10] = ri + np.array([0.6 / box, 0.0, 0.0])          # a partner 0.6 sigma away: overlap
    partial, fp = smc.force_1(ri, box, 2.5, rj)
    assert partial.ovr and not fp.any() and partial.pot == 0.0
    out.update(ovr_ri=ri, ovr_rj=rj)
    total, f = smc.force(box, 2.5, r)
    out.update(force_totals=np.array([total.cut, total.pot, total.vir, total.lap]), force_f=f)
    path = os.path.join(ROOT, "tests", "golden", "reference_smc.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
