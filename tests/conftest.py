import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WCA_R_CUT = (2.0 ** (1.0 / 3.0)) ** (1.0 / 2.0)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_generate_tests(metafunc):
    """Every GPU test runs twice: with the small-system pair kernel (eight lanes per atom; forced on up to 1e5 atoms
    here, the default is 16,384) and with the routes large systems take (tiled / direct / all-pairs kernels), so that each kernel keeps
    its parity coverage at the sizes the oracle and the goldens can check."""
    if metafunc.definition.get_closest_marker("gpu") is not None:
        metafunc.fixturenames.append("pair_route")
        metafunc.parametrize("pair_route", ["small", "tiled"], indirect=True)


@pytest.fixture
def pair_route(request):
    from examples_b200 import _lib
    limit = 100000 if request.param == "small" else 0
    old = _lib.set_multi_max(limit)
    old_env = os.environ.get("ATLJ_MULTI_MAX")
    os.environ["ATLJ_MULTI_MAX"] = str(limit)      # subprocesses (the unmodified reference drivers)
    yield request.param
    _lib.set_multi_max(old)
    if old_env is None:
        os.environ.pop("ATLJ_MULTI_MAX", None)
    else:
        os.environ["ATLJ_MULTI_MAX"] = old_env


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference Python modules (oracle/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale


def liquid_like(nc, density=0.75, sigma=0.05, seed=1):
    """fcc lattice with Gaussian displacements (sigma in units of sigma_LJ): a disordered, overlap-free
    configuration at any size without a melt."""
    from examples_b200.lattice import fcc_positions
    n, box, r = fcc_positions(nc, density)
    rng = np.random.default_rng(seed)
    r = r + rng.normal(0.0, sigma / box, size=r.shape)
    return n, box, np.ascontiguousarray(r - np.rint(r))


@pytest.fixture(scope="session")
def golden_dyn():
    """Trajectories produced by executing the reference drivers' own propagators (oracle/make_golden_dynamics.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_dynamics.npz"))
