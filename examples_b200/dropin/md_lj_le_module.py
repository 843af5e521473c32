"""Drop-in shim: put this directory in front of the reference's python_examples/ on PYTHONPATH and the
unmodified drivers (md_nve_lj.py, md_nvt_lj.py, bd_nvt_lj.py, md_nvt_lj_le.py ...) import the B200 path."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from examples_b200.md_lj_le_module import *  # noqa: E402,F401,F403
from examples_b200.md_lj_le_module import PotentialType, conclusion, force, introduction  # noqa: E402,F401
