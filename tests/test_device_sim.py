"""GPU: the device-resident integrators against the reference's own loop (golden) and the C oracle."""
import numpy as np
import pytest

from conftest import liquid_like, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["nve_n256_ap", "nve_n864_ll"])
def test_nve_trajectory_golden(golden, tag):
    """md_nve_lj.py:178-192 run unmodified with the reference force module (golden) vs the device loop.
    Step-by-step agreement of every per-step scalar to 1e-9 relative over 200 / 40 steps, final r, v too."""
    from examples_b200.device import Simulation
    box = float(golden[tag + "_box"])
    rec_g = golden[tag + "_rec"]
    sim = Simulation(box, 2.5, golden[tag + "_r0"], golden[tag + "_v0"])
    rec0 = sim.force()
    rec = sim.run_nve(0.005, rec_g.shape[0], hessian=True)
    for k in range(6):
        assert rel_err(rec[:, k], rec_g[:, k]) < 1e-9, k
    assert rel_err(rec[:, 6], rec_g[:, 6]) < 1e-9  # hessian
    assert not rec[:, 7].any()
    r, v, f = sim.download()
    assert rel_err(r, golden[tag + "_r1"]) < 1e-9
    assert rel_err(v, golden[tag + "_v1"]) < 1e-9
    sim.close()


def test_nve_32k_vs_oracle_and_energy_conservation(oracle):
    """BASELINE config 2 (N=32,000): 20 steps against the C oracle step by step, then 400 steps of energy
    conservation (conserved-energy MSD per atom small, no drift)."""
    from examples_b200.device import Simulation
    from examples_b200.lattice import ran_velocities
    n, box, r = liquid_like(20, 0.75, 0.05)
    v = ran_velocities(n, 1.0)
    sim = Simulation(box, 2.5, r, v)
    sim.force()
    ro, vo = r.copy(), v.copy()
    _, fo, _, _ = oracle.force(ro, box, 2.5, cells=True)
    rec_o, _ = oracle.nve_steps(ro, vo, fo, box, 2.5, 0.005, 20, cells=True)
    rec = sim.run_nve(0.005, 20)
    for k in range(6):
        assert rel_err(rec[:, k], rec_o[:, k]) < 1e-9, k
    rg, vg, fg = sim.download()
    assert rel_err(rg, ro) < 1e-9 and rel_err(vg, vo) < 1e-9 and rel_err(fg, fo) < 1e-8
    rec = sim.run_nve(0.005, 400)
    e = (0.5 * rec[:, 4] + rec[:, 1]) / n
    assert np.std(e) < 2e-4 and abs(e[-1] - e[0]) < 5e-4
    sim.close()
