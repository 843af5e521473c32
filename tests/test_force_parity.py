"""GPU: the CUDA path (through the drop-in modules -> ctypes -> C ABI) against the golden vectors generated
from the reference Python modules, and against the C oracle at sizes the reference cannot reach.

Bars (BASELINE.json north_star): integer cell / neighbour work bit-exact; in-range pair count and overlap flag
exact; energies, virials, Laplacian, forces <= 1e-10 relative (fp64, summation order differs)."""
import os

import numpy as np
import pytest

from conftest import WCA_R_CUT, liquid_like, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-10
LJ_TAGS = ["lj_n256_a0", "lj_n256_a0.1", "lj_n864_a0", "lj_n864_a0.1", "lj_n4000_a0", "lj_n4000_a0.1"]


def check_lj(total, f, gt, gf, govr):
    t = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(t - gt) / np.abs(gt)) < TOL
    if np.max(np.abs(gf)) > 1e-6:
        assert rel_err(f, gf) < TOL
    else:  # perfect lattice: forces cancel
        assert np.max(np.abs(f - gf)) < 1e-9
    assert total.ovr == govr


@pytest.mark.parametrize("tag", LJ_TAGS)
def test_lj_force_golden(golden, tag):
    from examples_b200 import md_lj_ll_module, md_lj_module
    r, box = golden[tag + "_r"], float(golden[tag + "_box"])
    for kind, mod in (("ap", md_lj_module), ("ll", md_lj_ll_module)):
        if tag + "_" + kind + "_totals" not in golden:
            continue
        r_in = r.copy()
        total, f = mod.force(box, 2.5, r_in)
        assert np.array_equal(r_in, r), "caller's r must not be modified (SURVEY.md 8b)"
        assert f.shape == r.shape and f.dtype == np.float64
        check_lj(total, f, golden[tag + "_" + kind + "_totals"], golden[tag + "_" + kind + "_f"],
                 bool(golden[tag + "_" + kind + "_ovr"]))
        hes = mod.hessian(box, 2.5, r, golden[tag + "_" + kind + "_f"])
        gh = float(golden[tag + "_" + kind + "_hes"])
        assert abs(hes - gh) <= TOL * abs(gh)
    # the all-pairs module on a system large enough for cells goes through the cell kernel: same answer
    if "n4000" in tag:
        total, f = md_lj_module.force(box, 2.5, r)
        check_lj(total, f, golden[tag + "_ll_totals"], golden[tag + "_ll_f"], bool(golden[tag + "_ll_ovr"]))


@pytest.mark.parametrize("tag", [t for t in LJ_TAGS if "n256" not in t])
def test_cells_bit_exact_golden(golden, tag):
    from examples_b200 import md_lj_ll_module
    r, box = golden[tag + "_r"], float(golden[tag + "_box"])
    c, perm, cell_start = md_lj_ll_module.cells(box, 2.5, r)
    assert np.array_equal(c, golden[tag + "_c"])
    assert np.array_equal(perm, golden[tag + "_perm"])
    assert np.array_equal(cell_start, golden[tag + "_cell_start"])


@pytest.mark.parametrize("n", [256, 864])
@pytest.mark.parametrize("strain", [0.0, 0.13, -0.37, 0.499])
def test_le_force_golden(golden, n, strain):
    from examples_b200 import md_lj_le_module, md_lj_llle_module
    tag = "le_n%d_s%g" % (n, strain)
    r, box = golden[tag + "_r"], float(golden[tag + "_box"])
    for kind, mod in (("ap", md_lj_le_module), ("ll", md_lj_llle_module)):
        total, f = mod.force(box, strain, r.copy())
        t = np.array([total.pot, total.vir, total.lap, total.pyx])
        gt = golden[tag + "_" + kind + "_totals"]
        assert np.max(np.abs(t - gt) / np.abs(gt)) < TOL, (kind, t, gt)
        assert rel_err(f, golden[tag + "_" + kind + "_f"]) < TOL
        assert total.ovr == bool(golden[tag + "_" + kind + "_ovr"])


def test_neighbour_cells_match_reference_stencils(golden):
    """Full stencil of the kernel == reference half stencil (golden) plus its mirror image, as sets, per cell."""
    from examples_b200 import _lib
    L = _lib.lib()
    cases = [(sc, 0, 0, golden["nb14_sc%d" % sc]) for sc in (4, 7)]
    cases += [(sc, 1, shift, golden["nb17_sc%d_shift%d" % (sc, shift)]) for sc, shift in ((5, 0), (5, 2), (8, -3), (8, 3))]
    for sc, le, shift, half in cases:
        nb = np.empty((sc ** 3, 30), dtype=np.int32)
        cnt = np.empty(sc ** 3, dtype=np.int32)
        _lib.check(L.atlj_neighbour_cells(sc, le, shift, nb, cnt))
        want = [set() for _ in range(sc ** 3)]
        for ci in range(sc ** 3):
            for cj in half[ci]:
                if cj >= 0:
                    want[ci].add(int(cj))
                    want[int(cj)].add(ci)
        for ci in range(sc ** 3):
            got = [int(x) for x in nb[ci] if x >= 0]
            assert len(got) == len(set(got)) == cnt[ci], "duplicate neighbour cell"
            assert set(got) == want[ci], (sc, le, shift, ci)


@pytest.mark.parametrize("nc,sigma", [(20, 0.02), (20, 0.08), (13, 0.05)])
def test_lj_cells_vs_oracle_32k(oracle, nc, sigma):
    """BASELINE config 2 size (N=32,000, sc=13) and a non-commensurate size, against the C oracle."""
    from examples_b200 import md_lj_ll_module
    n, box, r = liquid_like(nc, 0.75, sigma)
    t, f, ovr, npair = oracle.force(r, box, 2.5, cells=True)
    total, fg = md_lj_ll_module.force(box, 2.5, r)
    tg = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(tg - t) / np.abs(t)) < TOL
    assert rel_err(fg, f) < TOL
    assert total.ovr == ovr
    from examples_b200.md_lj_module import _force
    import math
    rc, _, _, _, npair_g = _force(box, 2.5, r, math.floor(box / 2.5), 1)
    assert rc == 0 and npair_g == npair, "in-range pair count must be exact"
    hes = md_lj_ll_module.hessian(box, 2.5, r, f)
    ho = oracle.hessian(r, f, box, 2.5, cells=True)
    assert abs(hes - ho) <= TOL * abs(ho)
    c, perm, cs = md_lj_ll_module.cells(box, 2.5, r)
    rw, co, permo, cso = oracle.cells(r, math.floor(box / 2.5))
    assert np.array_equal(c, co) and np.array_equal(perm, permo) and np.array_equal(cs, cso)


@pytest.mark.parametrize("strain", [0.0, 0.2501, -0.4999])
def test_le_cells_vs_oracle_tiny_cells(oracle, strain):
    """WCA cells hold ~1 atom each (BASELINE config 3 geometry, smaller N)."""
    from examples_b200 import md_lj_llle_module
    n, box, r = liquid_like(16, 0.8442, 0.05)  # N=16,384, sc=23
    r[:, 0] = r[:, 0] - np.rint(r[:, 1]) * strain
    r = r - np.rint(r)
    t, f, ovr, npair = oracle.force(r, box, WCA_R_CUT, cells=True, le=True, strain=strain)
    total, fg = md_lj_llle_module.force(box, strain, r.copy())
    tg = np.array([total.pot, total.vir, total.lap, total.pyx])
    assert np.max(np.abs(tg - t) / np.maximum(np.abs(t), 1e-300)) < TOL
    assert rel_err(fg, f) < TOL
    assert total.ovr == ovr


def test_overlap_flag_and_crowded_cell(oracle):
    """Two atoms nearly on top of each other set ovr; a crowded cell exceeds the per-cell fast paths."""
    from examples_b200 import md_lj_ll_module
    n, box, r = liquid_like(6, 0.75, 0.02)
    r[1] = r[0] + np.array([0.7 / box, 0.0, 0.0])
    r = r - np.rint(r)
    total, f = md_lj_ll_module.force(box, 2.5, r)
    t, fo, ovr, _ = oracle.force(r, box, 2.5, cells=True)
    assert ovr and total.ovr
    assert rel_err(f, fo) < TOL
    # 700 extra atoms on a tiny lattice inside one cell: per-cell sort uses the heap path, the pair kernel's
    # candidate set exceeds its shared-memory chunk
    rng = np.random.default_rng(3)
    g = np.stack(np.meshgrid(*[np.arange(9)] * 3, indexing="ij"), -1).reshape(-1, 3)[:700]
    extra = (g * 0.9 + 0.3) / box - 0.5 + rng.normal(0, 1e-3, (700, 3))
    r2 = np.ascontiguousarray(np.concatenate([r[2:], extra]))
    rng.shuffle(r2)
    total, f = md_lj_ll_module.force(box, 2.5, r2)
    t, fo, ovr, npair = oracle.force(r2, box, 2.5, cells=True)
    tg = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(tg - t) / np.abs(t)) < TOL
    assert rel_err(f, fo) < TOL
    c, perm, cs = md_lj_ll_module.cells(box, 2.5, r2)
    _, co, permo, cso = oracle.cells(r2, 4)
    assert np.array_equal(perm, permo) and np.array_equal(cs, cso)


def test_reference_assertions():
    from examples_b200 import md_lj_ll_module, md_lj_llle_module, md_lj_module
    from examples_b200.lattice import fcc_positions
    n, box, r = fcc_positions(4)  # N=256: sc=2
    with pytest.raises(AssertionError, match="System is too small for cells"):
        md_lj_ll_module.force(box, 2.5, r)
    with pytest.raises(AssertionError, match="Dimension error in force"):
        md_lj_module.force(box, 2.5, np.zeros((10, 2)))
    n, box, r = fcc_positions(6)
    total, f = md_lj_module.force(box, 2.5, r)
    with pytest.raises(AssertionError, match="Dimension mismatch in hessian"):
        md_lj_module.hessian(box, 2.5, r, f[:-1])
    r2 = r.copy()
    r2[0, 0] = 0.5  # rint(0.5)=0 keeps it at +0.5 -> cell index sc
    with pytest.raises(AssertionError, match="Index error"):
        md_lj_ll_module.force(box, 2.5, r2)
    r2[0, 0] = -0.5
    md_lj_ll_module.force(box, 2.5, r2)
    # the all-pairs module has no such restriction
    r2[0, 0] = 0.5
    n, box, r = fcc_positions(10)
    r[0, 0] = 0.5
    md_lj_module.force(box, 2.5, r)
    with pytest.raises(AssertionError, match="System is too small for cells"):
        md_lj_llle_module.force(2.0, 0.0, np.zeros((4, 3)))


def test_sc3_served_by_all_pairs(oracle):
    """sc == 3 (N=500): every cell's stencil is the whole box; the cell module must still agree."""
    from examples_b200 import md_lj_ll_module
    from examples_b200.lattice import fcc_positions, perturb
    n, box, r = fcc_positions(5)
    r = perturb(r, box, 0.1)
    assert int(np.floor(box / 2.5)) == 3
    total, f = md_lj_ll_module.force(box, 2.5, r)
    t, fo, ovr, _ = oracle.force(r, box, 2.5, cells=True)
    tg = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(tg - t) / np.abs(t)) < TOL and rel_err(f, fo) < TOL


def test_size_independent_properties_large():
    """N = 256,000: Newton's third law (sum f = 0), invariance of the totals under a rigid translation and under
    a permutation of the atom order, and determinism (bitwise identical repeat)."""
    from examples_b200 import md_lj_ll_module
    n, box, r = liquid_like(40, 0.75, 0.06)
    total, f = md_lj_ll_module.force(box, 2.5, r)
    assert not total.ovr
    assert np.max(np.abs(f.sum(axis=0))) < 1e-8 * np.max(np.abs(f))
    total2, f2 = md_lj_ll_module.force(box, 2.5, r)
    assert np.array_equal(f, f2) and total.pot == total2.pot and total.lap == total2.lap
    perm = np.random.default_rng(0).permutation(n)
    total3, f3 = md_lj_ll_module.force(box, 2.5, np.ascontiguousarray(r[perm]))
    assert abs(total3.pot - total.pot) < 1e-11 * abs(total.pot)
    assert rel_err(f3, f[perm]) < 1e-11
    rs = r + np.array([0.123, -0.377, 0.25])
    rs = rs - np.rint(rs)
    total4, f4 = md_lj_ll_module.force(box, 2.5, rs)
    for a, b in ((total4.pot, total.pot), (total4.vir, total.vir), (total4.lap, total.lap)):
        assert abs(a - b) < 1e-9 * abs(b)  # translation changes rounding of r, hence a few borderline pairs


def test_sparse_gas_and_tiny_systems(oracle):
    """Ragged inputs: a dilute gas (most cells empty, chunks span many cells), strongly non-uniform density (a slab of
    liquid in an otherwise empty box), and systems of one or two atoms."""
    from examples_b200 import md_lj_ll_module, md_lj_module
    rng = np.random.default_rng(11)
    # dilute gas: N = 600 in a box of 30 sigma -> sc = 12, 1728 cells, 0.35 atoms per cell
    box = 30.0
    r = rng.uniform(-0.5, 0.5, size=(600, 3))
    total, f = md_lj_ll_module.force(box, 2.5, r)
    t, fo, ovr, npair = oracle.force(r, box, 2.5, cells=True)
    tg = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(tg - t) / np.maximum(np.abs(t), 1e-300)) < TOL and total.ovr == ovr
    assert rel_err(f, fo) < TOL
    # liquid slab: all atoms in a quarter of the box along y
    n, b2, rl = liquid_like(12, 0.75, 0.05)
    rl[:, 1] = rl[:, 1] * 0.25
    b2 = b2 * 1.0
    total, f = md_lj_ll_module.force(b2, 2.5, rl)
    t, fo, ovr, npair = oracle.force(rl, b2, 2.5, cells=True)
    tg = np.array([total.cut, total.pot, total.vir, total.lap])
    assert total.ovr == ovr and rel_err(f, fo) < TOL
    assert np.max(np.abs(tg - t) / np.abs(t)) < TOL
    # two atoms, one atom
    r2 = np.array([[0.0, 0.0, 0.0], [0.1, 0.0, 0.0]])
    total, f = md_lj_module.force(10.0, 2.5, r2)       # separation 1.0 sigma: pot = 4*(1-1) - shift
    t, fo, ovr, npair = oracle.force(r2, 10.0, 2.5, cells=False)
    assert npair == 1 and abs(total.pot - t[1]) <= 1e-12 * abs(t[1]) and rel_err(f, fo) < TOL
    total, f = md_lj_module.force(10.0, 2.5, r2[:1])
    assert (total.cut, total.pot, total.vir, total.lap, total.ovr) == (0.0, 0.0, 0.0, 0.0, False) and not f.any()
    total, f = md_lj_ll_module.force(10.0, 2.5, r2)     # same through the cell route (sc = 4)
    assert abs(total.pot - t[1]) <= 1e-12 * abs(t[1]) and rel_err(f, fo) < TOL


@pytest.mark.parametrize("n", [600, 1400])
def test_direct_route_force_and_hessian(oracle, n):
    """Grids with fewer than 3 atoms per cell take the direct per-atom kernel (csrc/pair.cu): LJ force, exact pair
    count and hessian against the oracle; n = 1400 (74 candidates per atom) overflows the per-atom candidate list."""
    import math
    from examples_b200 import md_lj_ll_module
    from examples_b200.md_lj_module import _force
    box = 20.0                                            # sc = 8, 512 cells
    rng = np.random.default_rng(n)
    g = np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3)
    r = (g[rng.permutation(len(g))[:n]] + 0.5) / 12.0 - 0.5 + rng.normal(0.0, 0.01, (n, 3))   # no overlaps
    r = np.ascontiguousarray(r - np.rint(r))
    t, fo, ovr, npair = oracle.force(r, box, 2.5, cells=True)
    total, f = md_lj_ll_module.force(box, 2.5, r)
    tg = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(tg - t) / np.abs(t)) < TOL and total.ovr == ovr
    assert rel_err(f, fo) < TOL
    rc, _, _, _, npair_g = _force(box, 2.5, r, math.floor(box / 2.5), 1)
    assert rc == 0 and npair_g == npair
    hes = md_lj_ll_module.hessian(box, 2.5, r, f)
    ho = oracle.hessian(r, f, box, 2.5, cells=True)
    assert abs(hes - ho) <= TOL * abs(ho)


def test_tiled_wca_route_matches_direct(oracle):
    """The tiled WCA/LE kernel (csrc/pair_tile.cu) is only chosen for >= 3 atoms per cell; ATLJ_PAIR_KERNEL=tile1 forces
    it at the usual density so that it stays covered.  Runs in a child process because the switch is read at
    sim creation from the environment."""
    import subprocess
    import sys
    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "from examples_b200 import md_lj_llle_module\n"
        "from examples_b200.lattice import fcc_positions\n"
        "n, box, r = fcc_positions(10, 0.8442)\n"
        "r = r + np.random.default_rng(5).normal(0, 0.05 / box, r.shape)\n"
        "r[:, 0] -= np.rint(r[:, 1]) * 0.31\n"
        "r = np.ascontiguousarray(r - np.rint(r))\n"
        "t, f = md_lj_llle_module.force(box, 0.31, r.copy())\n"
        "np.save(sys.argv[1], np.concatenate([[t.pot, t.vir, t.lap, t.pyx], f.ravel()]))\n"
    ) % ROOT
    import os
    import tempfile
    out = {}
    for tag, env in (("direct", {}), ("tile1", {"ATLJ_PAIR_KERNEL": "tile1"})):
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "o.npy")
            subprocess.run([sys.executable, "-c", code, p], check=True, env={**os.environ, **env}, timeout=300)
            out[tag] = np.load(p)
    a, b = out["direct"], out["tile1"]
    assert np.max(np.abs(a[:4] - b[:4]) / np.abs(a[:4])) < TOL
    assert rel_err(a[4:].reshape(-1, 3), b[4:].reshape(-1, 3)) < TOL
    from examples_b200.lattice import fcc_positions
    n, box, r = fcc_positions(10, 0.8442)
    r = r + np.random.default_rng(5).normal(0, 0.05 / box, r.shape)
    r[:, 0] -= np.rint(r[:, 1]) * 0.31
    r = np.ascontiguousarray(r - np.rint(r))
    t, fo, ovr, npair = oracle.force(r, box, WCA_R_CUT, cells=True, le=True, strain=0.31)
    assert np.max(np.abs(a[:4] - t) / np.abs(t)) < TOL and rel_err(a[4:].reshape(-1, 3), fo) < TOL


@pytest.mark.parametrize("n,box", [(2200, 20.0), (5000, 23.0), (900, 10.5)])
def test_thin_cells_through_the_tiled_kernel(oracle, n, box):
    """3-6 atoms per cell: the tiled LJ kernel with chunks that span many cells (up to the whole z extent of a row
    pair, staged planes wrapping around the box), sc = 8, 9 and 4."""
    import math
    from examples_b200 import md_lj_ll_module
    from examples_b200.device import Simulation
    from examples_b200.md_lj_module import _force
    rng = np.random.default_rng(n)
    m = math.ceil(n ** (1.0 / 3.0)) + 1
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    r = (g[rng.permutation(len(g))[:n]] + 0.5) / m - 0.5 + rng.normal(0.0, 0.12 / m, (n, 3))
    r = np.ascontiguousarray(r - np.rint(r))
    sc = math.floor(box / 2.5)
    assert n >= 3 * sc ** 3, "meant for the tiled kernel, not the direct one"
    t, fo, ovr, npair = oracle.force(r, box, 2.5, cells=True)
    total, f = md_lj_ll_module.force(box, 2.5, r)
    tg = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(tg - t) / np.abs(t)) < TOL and total.ovr == ovr
    assert rel_err(f, fo) < TOL
    rc, _, _, _, npair_g = _force(box, 2.5, r, sc, 1)
    assert rc == 0 and npair_g == npair
    sim = Simulation(box, 2.5, r)
    rec = sim.force(hessian=True)
    rs, vs, fs, ids = sim.download_sorted()
    ho = oracle.hessian(rs, fs, box, 2.5, cells=True)
    assert abs(rec[6] - ho) <= TOL * abs(ho) and rec[8] == npair
    sim.close()


@pytest.mark.parametrize("r_cut,box", [(30.0, 125.0), (42.0, 172.0)])
def test_long_cutoff_dilute_system(oracle, r_cut, box):
    """Cells of 31 and 43 sigma: the fp16 filter of the tiled LJ kernel works on chunk-centred coordinates up to ~100
    sigma (first case); beyond 40 sigma per cell the first tiled kernel takes over (second case)."""
    import math
    from examples_b200 import md_lj_ll_module
    from examples_b200.md_lj_module import _force
    sc = math.floor(box / r_cut)
    assert sc == 4
    n = 420
    rng = np.random.default_rng(int(r_cut))
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3)
    r = (g[rng.permutation(len(g))[:n]] + 0.5) / 8.0 - 0.5 + rng.normal(0.0, 0.02, (n, 3))   # spacing >> sigma
    r = np.ascontiguousarray(r - np.rint(r))
    t, fo, ovr, npair = oracle.force(r, box, r_cut, cells=True)
    total, f = md_lj_ll_module.force(box, r_cut, r)
    tg = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(tg - t) / np.abs(t)) < TOL and total.ovr == ovr
    assert rel_err(f, fo) < TOL
    rc, _, _, _, npair_g = _force(box, r_cut, r, sc, 1)
    assert rc == 0 and npair_g == npair


@pytest.mark.parametrize("tag,nc,model", [("c3", 40, "wca"), ("c4", 63, "lj"), ("c5", 80, "lj")])
def test_baseline_sizes_against_the_oracle(oracle, tag, nc, model):
    """The BASELINE.json configurations at FULL size against the C oracle (OpenMP over cells on the host cores):
    c3 N = 256,000 WCA with Lees-Edwards cells at strain 0.3, c4 N = 1,000,188 and c5 N = 2,048,000 LJ liquids.
    Totals and forces to 1e-10, in-range pair count and overlap flag exact."""
    from examples_b200.device import Simulation
    n, box, r = liquid_like(nc, 0.75, 0.07)
    if model == "wca":
        strain = 0.3
        r[:, 0] = r[:, 0] - np.rint(r[:, 1]) * strain
        r = np.ascontiguousarray(r - np.rint(r))
        t, fo, ovr, npair = oracle.force(r, box, WCA_R_CUT, cells=True, le=True, strain=strain, omp=True)
        sim = Simulation(box, None, r, model="wca")
        rec = sim.force(strain=strain)
    else:
        t, fo, ovr, npair = oracle.force(r, box, 2.5, cells=True, omp=True)
        sim = Simulation(box, 2.5, r)
        rec = sim.force()
    rs, vs, fs, ids = sim.download_sorted()
    f = np.empty_like(fs)
    f[ids] = fs
    assert rec[8] == npair and bool(rec[7]) == ovr
    assert np.max(np.abs(rec[:4] - t) / np.abs(t)) < TOL
    assert rel_err(f, fo) < TOL
    sim.close()


def test_smc_force_1_golden():
    """smc_lj_module.force_1 / force (the single-atom and multi-atom moves of smc_nvt_lj.py) against values produced
    by the reference module (oracle/make_golden_smc.py): per-partner forces and totals to 1e-10, the overlap contract
    (ovr = True, everything zero, smc_lj_module.py:146-149), assertion messages."""
    from examples_b200 import smc_lj_module
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_smc.npz"))
    box, r = float(g["box"]), g["r"]
    for k, i in enumerate(g["idx"]):
        rj = np.delete(r, i, 0)
        partial, fp = smc_lj_module.force_1(r[i, :], box, 2.5, rj)
        t = np.array([partial.cut, partial.pot, partial.vir, partial.lap])
        assert not partial.ovr and np.max(np.abs(t - g["f1_totals"][k]) / np.abs(g["f1_totals"][k])) < TOL
        assert fp.shape == rj.shape and rel_err(fp, g["f1_f"][k]) < TOL
        assert np.array_equal(fp == 0.0, g["f1_f"][k] == 0.0), "the same partners are out of range"
    partial, fp = smc_lj_module.force_1(g["ovr_ri"], box, 2.5, g["ovr_rj"])
    assert partial.ovr and not fp.any() and (partial.pot, partial.cut, partial.vir, partial.lap) == (0.0, 0.0, 0.0, 0.0)
    total, f = smc_lj_module.force(box, 2.5, r)
    t = np.array([total.cut, total.pot, total.vir, total.lap])
    assert np.max(np.abs(t - g["force_totals"]) / np.abs(g["force_totals"])) < TOL and rel_err(f, g["force_f"]) < TOL
    with pytest.raises(AssertionError, match="Dimension error for r in force_1"):
        smc_lj_module.force_1(r[0], box, 2.5, r[:, :2])
    with pytest.raises(AssertionError, match="Dimension error for ri in force_1"):
        smc_lj_module.force_1(r[0, :2], box, 2.5, r)
