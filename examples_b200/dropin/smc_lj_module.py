"""Drop-in shim: put this directory in front of the reference's python_examples/ on PYTHONPATH and the unmodified
smc_nvt_lj.py imports the B200 path (force, force_1)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from examples_b200.smc_lj_module import *  # noqa: E402,F401,F403
from examples_b200.smc_lj_module import PotentialType, conclusion, force, force_1, introduction  # noqa: E402,F401
