#!/usr/bin/env python
"""Generate the cnf fixtures under tests/golden/ with the reference's OWN config_io_module (run in the build container
only; needs /root/reference):

    python oracle/make_golden_cnf.py

cnf_golden_r.inp / cnf_golden_rv.inp are written by the reference's write_cnf_atoms from the arrays saved in
cnf_golden.npz; the npz also holds what the reference's read_cnf_atoms reads back from those files.  The values
include the awkward cases of '%15.10f': exact ties at the 11th decimal (odd multiples of 2^-11), negative zero and
negatives that round to zero, the largest values a 15-character field holds, subnormals.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ATLJ_REFERENCE", "/root/reference/python_examples")
sys.path.insert(0, REF)
from config_io_module import read_cnf_atoms, write_cnf_atoms  # noqa: E402


def main():
    rng = np.random.default_rng(20240611)
    n = 4000
    special = np.array([
        0.0, -0.0, 1.0 / 2048, 3.0 / 2048, -5.0 / 2048, 7.0 / 2048 + 3.0, 1e-11, -1e-11, 5e-11, -5e-11, 4.9999999999e-11,
        0.1, -0.1, 0.29999999999999999, 1.0 / 3.0, 9999.99999999994, -999.99999999994, 9999.0, -999.0, 123.4567890125,
        5e-324, -5e-324, 2.2250738585072014e-308, 0.99999999995, 0.999999999949999, 1.00000000005, 139.78864371789,
        -69.894321858945, 2.5, 1e-10, 1.5e-10, 2.5e-10, -2.5e-10, 0.00000000015, 999.99999999995, 4096.00000000005,
    ])
    special = np.concatenate([special, (2 * rng.integers(0, 2 ** 20, 200) + 1) / 2048.0])   # more exact ties
    r = rng.uniform(-70.0, 70.0, size=(n, 3))
    r.ravel()[:special.size] = special
    r.ravel()[special.size:2 * special.size] = -special[::-1] * 0.05
    v = rng.standard_normal((n, 3)) * 1.3
    v.ravel()[:special.size] = special[::-1] * 1e-3
    box = 139.78864371789039
    gold = os.path.join(ROOT, "tests", "golden")
    write_cnf_atoms(os.path.join(gold, "cnf_golden_r.inp"), n, box, r)
    write_cnf_atoms(os.path.join(gold, "cnf_golden_rv.inp"), n, box, r, v)
    n1, box1, r1 = read_cnf_atoms(os.path.join(gold, "cnf_golden_r.inp"))
    n2, box2, r2, v2 = read_cnf_atoms(os.path.join(gold, "cnf_golden_rv.inp"), with_v=True)
    assert n1 == n2 == n and np.array_equal(r1, r2)
    np.savez_compressed(os.path.join(gold, "cnf_golden.npz"), r=r, v=v, box=np.float64(box), box_read=np.float64(box1),
                        r_read=r1, v_read=v2)
    for f in ("cnf_golden_r.inp", "cnf_golden_rv.inp", "cnf_golden.npz"):
        print(f, os.path.getsize(os.path.join(gold, f)), "bytes")


if __name__ == "__main__":
    main()
