"""CPU: the C oracle against the golden vectors generated from the reference Python modules.
Tolerances: integer work bit-exact; fp64 sums <= 1e-10 relative (summation order differs); on the perfect
lattice (amp=0) forces cancel to ~1e-13 absolute, so forces there are compared absolutely."""
import numpy as np
import pytest

from conftest import WCA_R_CUT, rel_err

TOL = 1e-10
LJ_TAGS = ["lj_n256_a0", "lj_n256_a0.1", "lj_n864_a0", "lj_n864_a0.1", "lj_n4000_a0", "lj_n4000_a0.1"]


@pytest.mark.parametrize("tag", LJ_TAGS)
def test_lj_force_and_hessian(golden, oracle, tag):
    r, box = golden[tag + "_r"], float(golden[tag + "_box"])
    for kind, cells in (("ap", False), ("ll", True)):
        if tag + "_" + kind + "_totals" not in golden:
            continue
        t, f, ovr, npair = oracle.force(r, box, 2.5, cells=cells)
        gt, gf = golden[tag + "_" + kind + "_totals"], golden[tag + "_" + kind + "_f"]
        assert np.max(np.abs(t - gt) / np.abs(gt)) < TOL
        if np.max(np.abs(gf)) > 1e-6:
            assert rel_err(f, gf) < TOL
        else:
            assert np.max(np.abs(f - gf)) < 1e-9
        assert ovr == bool(golden[tag + "_" + kind + "_ovr"])
        h = oracle.hessian(r, gf, box, 2.5, cells=cells)
        gh = float(golden[tag + "_" + kind + "_hes"])
        assert abs(h - gh) <= TOL * abs(gh)


@pytest.mark.parametrize("tag", [t for t in LJ_TAGS if "n256" not in t])
def test_cells_bit_exact(golden, oracle, tag):
    r = golden[tag + "_r"]
    sc = int(golden[tag + "_sc"])
    rw, c, perm, cs = oracle.cells(r, sc)
    assert np.array_equal(c, golden[tag + "_c"])
    assert np.array_equal(perm, golden[tag + "_perm"])
    assert np.array_equal(cs, golden[tag + "_cell_start"])
    assert np.array_equal(rw, r - np.rint(r))


def test_neighbour_cells_bit_exact(golden, oracle):
    for sc in (3, 4, 7):
        nb = golden["nb14_sc%d" % sc]
        for c in range(sc ** 3):
            assert np.array_equal(oracle.neighbours(sc, c), nb[c])
    for sc, shift in ((5, 0), (5, 2), (8, -3), (8, 3)):
        nb = golden["nb17_sc%d_shift%d" % (sc, shift)]
        for c in range(sc ** 3):
            o = oracle.neighbours(sc, c, le=True, shift=shift)
            assert np.array_equal(o, nb[c][:len(o)])
            assert len(o) == 17 or nb[c][14] == -1


@pytest.mark.parametrize("n", [256, 864])
@pytest.mark.parametrize("strain", [0.0, 0.13, -0.37, 0.499])
def test_le_force(golden, oracle, n, strain):
    tag = "le_n%d_s%g" % (n, strain)
    r, box = golden[tag + "_r"], float(golden[tag + "_box"])
    for kind, cells in (("ap", False), ("ll", True)):
        t, f, ovr, _ = oracle.force(r, box, WCA_R_CUT, cells=cells, le=True, strain=strain)
        assert np.max(np.abs(t - golden[tag + "_" + kind + "_totals"]) / np.abs(golden[tag + "_" + kind + "_totals"])) < TOL
        assert rel_err(f, golden[tag + "_" + kind + "_f"]) < TOL
        assert ovr == bool(golden[tag + "_" + kind + "_ovr"])


def test_known_answers_survey_appendix_b(golden):
    """The golden file reproduces the known-answer values recorded in SURVEY.md Appendix B."""
    t = golden["lj_n256_a0.1_ap_totals"]
    assert np.allclose(t, [-1370.5275859513945, -1267.7148549034605, -1051.4041117662784, 89670.3966567883], rtol=1e-13)
    assert abs(float(golden["lj_n256_a0.1_ap_hes"]) - 86032141.89611648) < 1e-5
    t = golden["lj_n4000_a0.1_ll_totals"]
    assert np.allclose(t, [-21038.5217314108, -19387.660270725948, -14072.201469209294, 1589716.543164366], rtol=1e-13)
    t = golden["le_n864_s0.13_ap_totals"]
    assert np.allclose(t, [1251.9249815191486, 8151.080939866314, 912932.7598035531, 4824.218109774356], rtol=1e-13)


@pytest.mark.parametrize("tag,cells", [("nve_n256_ap", False), ("nve_n864_ll", True)])
def test_nve_trajectory(golden, oracle, tag, cells):
    """Velocity-Verlet loop of md_nve_lj.py:178-192: oracle vs the reference's own loop, step by step.
    Tolerance 1e-9 relative over <= 200 steps (chaotic growth of rounding differences)."""
    box = float(golden[tag + "_box"])
    r, v = golden[tag + "_r0"].copy(), golden[tag + "_v0"].copy()
    rec_g = golden[tag + "_rec"]
    _, f, _, _ = oracle.force(r, box, 2.5, cells=cells)
    rec, ovr = oracle.nve_steps(r, v, f, box, 2.5, 0.005, rec_g.shape[0], cells=cells)
    assert not ovr
    for k in range(6):
        assert rel_err(rec[:, k], rec_g[:, k]) < 1e-9
    assert rel_err(r, golden[tag + "_r1"]) < 1e-9
    assert rel_err(v, golden[tag + "_v1"]) < 1e-9


def test_lattice_matches_reference_inputs(golden):
    from examples_b200.lattice import fcc_positions, perturb
    n, box, r = fcc_positions(6, 0.75)
    assert n == 864 and box == float(golden["lj_n864_a0_box"])
    assert np.array_equal(perturb(r, box, 0.0), golden["lj_n864_a0_r"])
    assert np.array_equal(perturb(r, box, 0.1), golden["lj_n864_a0.1_r"])
