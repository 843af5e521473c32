"""Drop-in shim for the reference's config_io_module: read_cnf_atoms / write_cnf_atoms go through the GPU text path
(examples_b200/config_io_module.py); every other name (read_cnf_mols, write_cnf_mols) resolves to the reference's own
module of the same name further along sys.path, so programs outside the LJ path keep working."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(os.path.dirname(_here))
if _root not in sys.path:
    sys.path.insert(0, _root)
from examples_b200.config_io_module import read_cnf_atoms, write_cnf_atoms, write_cnf_sim  # noqa: E402,F401

_ref = None


def __getattr__(name):
    global _ref
    if _ref is None:
        for d in sys.path:
            p = os.path.join(d or ".", "config_io_module.py")
            if os.path.isfile(p) and os.path.abspath(os.path.dirname(p)) != _here:
                spec = importlib.util.spec_from_file_location("_reference_config_io_module", p)
                _ref = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(_ref)
                break
        else:
            raise AttributeError("config_io_module.%s: the reference's config_io_module.py is not on sys.path" % name)
    return getattr(_ref, name)
