"""Synthetic inputs for the LJ hot path: fcc lattice and Maxwell-Boltzmann velocities.

Same formulas as the reference's configuration generator (python_examples/initialize.py:
fcc_positions :43-63, box from density :546, ran_velocities :138-142), vectorised so that a
16 M-atom lattice takes seconds, and seeded (np.random.default_rng) so runs are reproducible;
the reference uses the unseeded legacy generator (initialize.py:465) -- a documented deviation.
"""
import numpy as np

_BASIS = np.array([[0.25, 0.25, 0.25], [0.25, 0.75, 0.75], [0.75, 0.75, 0.25], [0.75, 0.25, 0.75]])


def fcc_box(n, density):
    """Box length for n atoms at the given density (initialize.py:546)."""
    return (n / density) ** (1.0 / 3.0)


def fcc_positions(nc, density=0.75):
    """-> n, box, r (n,3) in box=1 units, atom i = 4*((ix*nc+iy)*nc+iz)+a  (initialize.py:52-63)."""
    n = 4 * nc ** 3
    box = fcc_box(n, density)
    cell = box / nc
    box2 = box / 2.0
    g = np.arange(nc, dtype=np.float64)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    base = np.stack([ix, iy, iz], axis=-1).reshape(-1, 1, 3)  # (nc^3,1,3), C order = product(range(nc),repeat=3)
    r = (_BASIS[None, :, :] + base) * cell - box2             # same op order as initialize.py:57-59
    r = r.reshape(n, 3) / box
    return n, box, np.ascontiguousarray(r)


def perturb(r, box, amp):
    """Deterministic displacement used by the parity tests:
    r[i,k] += (amp/box)*sin(3i+k+1), then wrap."""
    n = r.shape[0]
    i = np.arange(n, dtype=np.float64)[:, None]
    k = np.arange(3, dtype=np.float64)[None, :]
    r = r + (amp / box) * np.sin(3.0 * i + k + 1.0)
    return r - np.rint(r)


def ran_velocities(n, temperature=1.0, seed=20170101):
    """Gaussian velocities, zero total momentum, exact kinetic temperature (initialize.py:138-142)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 3))
    v = v - np.sum(v, axis=0) / n
    factor = np.sqrt((3 * n - 3) * temperature / np.sum(v ** 2))
    return factor * v
