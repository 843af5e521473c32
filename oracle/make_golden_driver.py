#!/usr/bin/env python
"""Generate tests/golden/driver_<case>.{inp,stdout,out}: the start configuration, the printed output and the final
cnf.out of the reference's UNMODIFIED driver scripts run with the reference's own force modules.

Build container only (needs the reference, staged under oracle/_ref by oracle/make_ref.py):

    python oracle/make_ref.py && python oracle/make_golden_driver.py [case ...]

tests/test_reference_driver.py runs the same scripts with examples_b200/dropin first on PYTHONPATH and compares.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import driver_runner as D  # noqa: E402
from make_ref import staged_dir  # noqa: E402


def main():
    ref = staged_dir()
    assert ref, "run oracle/make_ref.py first"
    for name in (sys.argv[1:] or D.CASES):
        n, box, r, v = D.case_input(name)
        D.write_cnf(os.path.join(D.GOLDEN, "driver_%s.inp" % name), n, box, r, v)
        stdout, cnf = D.run_case(name, [ref])
        open(os.path.join(D.GOLDEN, "driver_%s.stdout" % name), "w").write(stdout)
        open(os.path.join(D.GOLDEN, "driver_%s.out" % name), "w").write(cnf)
        print("%-12s N=%d  %d lines of output" % (name, n, len(stdout.splitlines())), flush=True)


if __name__ == "__main__":
    main()
