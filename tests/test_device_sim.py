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


def test_baoab_trajectory_golden(golden_dyn):
    """bd_nvt_lj.py:221-233 executed from the reference source with recorded np.random.randn draws (golden) vs the
    device loop fed the same noise: every per-step scalar to 1e-9 relative over 100 steps, final r, v too."""
    from examples_b200.device import Simulation
    g = golden_dyn
    box, rec_g = float(g["bd_box"]), g["bd_rec"]
    sim = Simulation(box, 2.5, g["bd_r0"], g["bd_v0"])
    sim.force()
    rec = sim.run_baoab(float(g["bd_dt"]), float(g["bd_gamma"]), float(g["bd_temperature"]), rec_g.shape[0],
                        noise=g["bd_noise"])
    for k in range(6):
        assert rel_err(rec[:, k], rec_g[:, k]) < 1e-9, k
    r, v, f = sim.download()
    assert rel_err(r, g["bd_r1"]) < 1e-9 and rel_err(v, g["bd_v1"]) < 1e-9
    sim.close()


def test_baoab_device_rng_statistics():
    """Philox + Box-Muller on the device: with a huge friction one O step replaces v by sqrt(T) * N(0,1)."""
    from examples_b200.device import Simulation
    n, box, r = liquid_like(20, 0.75, 0.03)
    temperature = 1.7
    sim = Simulation(box, 2.5, r, np.zeros((n, 3)))
    sim.force()
    sim.run_baoab(1e-6, 1e9, temperature, 1, seed=11)
    _, v1, f, _ = sim.download_sorted()
    v1 = v1 - 0.5e-6 * f                                     # undo the final half kick
    x = v1.ravel() / np.sqrt(temperature)
    nn = x.size
    assert abs(x.mean()) < 5 / np.sqrt(nn)
    assert abs(x.var() - 1.0) < 5 * np.sqrt(2.0 / nn)
    assert abs(np.mean(x ** 4) - 3.0) < 5 * np.sqrt(96.0 / nn)
    c = np.corrcoef(v1.T)
    assert np.max(np.abs(c - np.eye(3))) < 5 / np.sqrt(n)
    sim.run_baoab(1e-6, 1e9, temperature, 1, seed=11)       # next step: fresh numbers (the counter advanced)
    _, v2, f2, _ = sim.download_sorted()
    assert abs(np.corrcoef(v1.ravel(), (v2 - 0.5e-6 * f2).ravel())[0, 1]) < 5 / np.sqrt(nn)
    sim.close()
    sim2 = Simulation(box, 2.5, r, np.zeros((n, 3)))        # same seed, same step counter => same numbers
    sim2.force()
    sim2.run_baoab(1e-6, 1e9, temperature, 1, seed=11)
    r2, v3, f3 = sim2.download()
    sim3 = Simulation(box, 2.5, r, np.zeros((n, 3)))
    sim3.force()
    sim3.run_baoab(1e-6, 1e9, temperature, 1, seed=12)
    _, v4, _ = sim3.download()
    assert not np.array_equal(v3, v4)
    sim2.close(), sim3.close()


def test_nhc_trajectory_golden(golden_dyn):
    """md_nvt_lj.py:270-288 executed from the reference source (golden) vs the device loop, 100 steps, M = 3."""
    from examples_b200.device import Simulation
    g = golden_dyn
    box, rec_g = float(g["nhc_box"]), g["nhc_rec"]
    sim = Simulation(box, 2.5, g["nhc_r0"], g["nhc_v0"])
    sim.force()
    eta, p_eta, q = g["nhc_eta0"].copy(), g["nhc_p_eta0"].copy(), g["nhc_q"].copy()
    rec = sim.run_nhc(0.005, float(g["nhc_temperature"]), eta, p_eta, q, rec_g.shape[0], g=float(g["nhc_g"]))
    for k in range(6):
        assert rel_err(rec[:, k], rec_g[:, k]) < 1e-9, k
    assert rel_err(rec[:, 10], rec_g[:, 6]) < 1e-9            # thermostat part of the conserved energy
    assert rel_err(eta, g["nhc_eta1"]) < 1e-9 and rel_err(p_eta, g["nhc_p_eta1"]) < 1e-9
    r, v, f = sim.download()
    assert rel_err(r, g["nhc_r1"]) < 1e-9 and rel_err(v, g["nhc_v1"]) < 1e-9
    # conserved quantity (md_nvt_lj.py:81): energy + thermostat terms
    cons = 0.5 * rec[:, 4] + rec[:, 1] + rec[:, 10]
    assert np.std(cons) / 256 < 5e-4  # fluctuates at O(dt^2), no drift
    sim.close()


@pytest.mark.parametrize("tag", ["le", "le_fast"])
def test_sllod_trajectory_golden(golden_dyn, tag):
    """md_nvt_lj_le.py:226-240 executed from the reference source (golden) vs the device loop (WCA, Lees-Edwards
    link cells, sc = 5), 100 steps at strain rates 0.04 and 0.64."""
    from examples_b200.device import Simulation
    g = golden_dyn
    box, rec_g = float(g[tag + "_box"]), g[tag + "_rec"]
    sim = Simulation(box, None, g[tag + "_r0"], g[tag + "_v0"], model="wca")
    sim.force(strain=0.0)
    rec, strain = sim.run_sllod(0.005, float(g[tag + "_rate"]), 0.0, rec_g.shape[0])
    for k in range(4):
        assert rel_err(rec[:, k], rec_g[:, k]) < 1e-9, k      # pot, vir, lap, pyx
    assert rel_err(rec[:, 4], rec_g[:, 4]) < 1e-9              # sum v^2 (isokinetic: constant)
    assert rel_err(rec[:, 5], rec_g[:, 5]) < 1e-9              # sum f^2
    assert rel_err(rec[:, 9], rec_g[:, 6]) < 1e-9              # sum vx*vy
    assert np.max(np.abs(rec[:, 10] - rec_g[:, 7])) < 1e-14 and abs(strain - rec_g[-1, 7]) < 1e-14
    assert np.ptp(rec[:, 4]) < 1e-9 * rec[0, 4]
    r, v, f = sim.download()
    assert rel_err(r, g[tag + "_r1"]) < 1e-9 and rel_err(v, g[tag + "_v1"]) < 1e-9
    sim.close()
