"""Configuration-file text path (csrc/cnf.cu, examples_b200/config_io_module.py) against fixtures written and read
back by the reference's own config_io_module (oracle/make_golden_cnf.py).  Bar: byte-identical files, bit-identical
values."""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")


def _py_body(rv):
    """What np.savetxt(fmt='%15.10f') writes, restated with Python's own % operator (config_io_module.py:88-95)."""
    return "".join(" ".join("%15.10f" % x for x in row) + "\n" for row in rv).encode("ascii")


def test_fixture_is_the_documented_format():
    """CPU: the reference-written fixture is header + n*cols fixed 16-byte fields, and np.loadtxt's values are the
    correctly rounded decimals - the two facts the GPU kernels are built on."""
    g = np.load(os.path.join(GOLD, "cnf_golden.npz"))
    raw = open(os.path.join(GOLD, "cnf_golden_rv.inp"), "rb").read()
    n = g["r"].shape[0]
    head = ("{:15d}\n{:15.8f}".format(n, float(g["box"])) + "\n").encode("ascii")
    assert raw.startswith(head) and len(raw) == len(head) + n * 6 * 16
    assert raw[len(head):] == _py_body(np.concatenate([g["r"], g["v"]], axis=1))
    txt = raw[len(head):].decode("ascii").split()
    ints = np.array([int(t.replace(".", "").replace("-", "")) for t in txt], dtype=np.float64)
    sign = np.array([-1.0 if t.startswith("-") else 1.0 for t in txt])
    back = (sign * (ints / 1e10)).reshape(n, 6)            # one exact division = the correctly rounded value
    assert np.array_equal(back[:, :3], g["r_read"]) and np.array_equal(back[:, 3:], g["v_read"])
    assert np.array_equal(np.signbit(back[:, :3]), np.signbit(g["r_read"]))


@pytest.mark.gpu
def test_write_matches_reference_bytes(tmp_path):
    from examples_b200.config_io_module import write_cnf_atoms
    g = np.load(os.path.join(GOLD, "cnf_golden.npz"))
    n, box = g["r"].shape[0], float(g["box"])
    p = str(tmp_path / "cnf.out")
    write_cnf_atoms(p, n, box, g["r"])
    assert open(p, "rb").read() == open(os.path.join(GOLD, "cnf_golden_r.inp"), "rb").read()
    write_cnf_atoms(p, n, box, g["r"], g["v"])
    assert open(p, "rb").read() == open(os.path.join(GOLD, "cnf_golden_rv.inp"), "rb").read()
    with pytest.raises(AssertionError, match="r shape mismatch"):
        write_cnf_atoms(p, n + 1, box, g["r"])
    with pytest.raises(AssertionError, match="Wrong number of arguments"):
        write_cnf_atoms(p, n, box)


@pytest.mark.gpu
def test_read_matches_reference_values():
    from examples_b200.config_io_module import read_cnf_atoms
    g = np.load(os.path.join(GOLD, "cnf_golden.npz"))
    n, box, r = read_cnf_atoms(os.path.join(GOLD, "cnf_golden_r.inp"))
    assert n == g["r"].shape[0] and box == float(g["box_read"])
    assert np.array_equal(r, g["r_read"]) and np.array_equal(np.signbit(r), np.signbit(g["r_read"]))
    n, box, r, v = read_cnf_atoms(os.path.join(GOLD, "cnf_golden_rv.inp"), with_v=True)
    assert np.array_equal(r, g["r_read"]) and np.array_equal(v, g["v_read"])
    n, box, r = read_cnf_atoms(os.path.join(GOLD, "cnf_golden_rv.inp"))          # extra columns ignored, as loadtxt
    assert np.array_equal(r, g["r_read"])
    with pytest.raises(AssertionError, match="cols less than 6"):
        read_cnf_atoms(os.path.join(GOLD, "cnf_golden_r.inp"), with_v=True)


@pytest.mark.gpu
def test_random_values_against_python_formatting(tmp_path):
    """300,000 values over twelve decades, both signs, plus every odd multiple of 2^-11 below 4 (exact ties)."""
    from examples_b200.config_io_module import read_cnf_atoms, write_cnf_atoms
    rng = np.random.default_rng(3)
    x = rng.uniform(-1.0, 1.0, 300000) * 10.0 ** rng.integers(-12, 3, 300000)
    x[:4096] = (2 * np.arange(4096) + 1) / 2048.0 * np.where(np.arange(4096) % 3 == 0, -1.0, 1.0)
    r = x.reshape(-1, 3)
    p = str(tmp_path / "cnf.rnd")
    write_cnf_atoms(p, r.shape[0], 10.0, r)
    raw = open(p, "rb").read()
    assert raw[32:] == _py_body(r)
    n, box, rr = read_cnf_atoms(p)
    want = np.loadtxt(p, skiprows=2)
    assert np.array_equal(rr, want) and np.array_equal(np.signbit(rr), np.signbit(want))
    assert np.max(np.abs(rr - r)) <= 0.5e-10 * (1 + 1e-6)   # ties sit exactly on the half


@pytest.mark.gpu
def test_loud_failures(tmp_path):
    from examples_b200 import _lib
    from examples_b200.config_io_module import read_cnf_atoms, write_cnf_atoms
    p = str(tmp_path / "cnf.bad")
    for bad in (1.0e4, -1.0e3, np.nan, np.inf):
        r = np.zeros((4, 3))
        r[2, 1] = bad
        with pytest.raises(_lib.AtljError, match="does not fit"):
            write_cnf_atoms(p, 4, 1.0, r)
    r = np.zeros((4, 3))
    r[0, 0] = 9999.99999999994            # the widest positive field
    write_cnf_atoms(p, 4, 1.0, r)
    assert read_cnf_atoms(p)[2][0, 0] == 9999.9999999999
    raw = bytearray(open(p, "rb").read())
    raw[32 + 16 * 4 + 9] = ord("e")       # exponent notation inside a field
    open(p, "wb").write(bytes(raw))
    # right size but not the fixed notation: the GPU parser declines, the reference's own reader semantics apply
    got = read_cnf_atoms(p)[2]
    assert np.array_equal(got, np.loadtxt(p, skiprows=2)[:, :3])


def _fortran_style(path, n, box, r, v=None, eol="\n", trailing_blank=False):
    """A cnf file as the Fortran programs write it (config_io_module.f90: `(*(f15.10))`, no separator column)."""
    with open(path, "w", newline="") as fh:
        fh.write("%15d%s%15.8f%s" % (n, eol, box, eol))
        rows = r if v is None else np.concatenate((r, v), axis=1)
        for row in rows:
            fh.write("".join("%15.10f" % x for x in row) + eol)
        if trailing_blank:
            fh.write(eol)


@pytest.mark.parametrize("with_v", [False, True])
@pytest.mark.parametrize("eol,blank", [("\n", False), ("\r\n", False), ("\n", True)])
def test_reads_every_layout_the_reference_reader_accepts(tmp_path, with_v, eol, blank):
    """a cnf.inp from the Fortran `initialize` (15-character fields), CRLF line ends or a trailing blank
    line must still be readable, with the values np.loadtxt (config_io_module.py:37) returns.  Host text path: no GPU."""
    from examples_b200.config_io_module import read_cnf_atoms
    rng = np.random.default_rng(5)
    n, box = 37, 7.5
    r, v = rng.uniform(-3.7, 3.7, (n, 3)), rng.normal(0, 1, (n, 3))
    p = str(tmp_path / "cnf.f90")
    _fortran_style(p, n, box, r, v if with_v else None, eol, blank)
    out = read_cnf_atoms(p, with_v=with_v)
    want = np.loadtxt(p, skiprows=2)
    assert out[0] == n and out[1] == box
    assert np.array_equal(out[2], want[:, :3])
    if with_v:
        assert np.array_equal(out[3], want[:, 3:6])
    else:
        with pytest.raises(AssertionError, match="cols less than 6"):
            read_cnf_atoms(p, with_v=True)


def test_reads_exponent_notation(tmp_path):
    from examples_b200.config_io_module import read_cnf_atoms
    p = str(tmp_path / "cnf.exp")
    r = np.array([[1.5e-3, -2.25, 3.0], [0.0, 1e1, -1e-7]])
    with open(p, "w") as fh:
        fh.write("2\n3.0\n")
        np.savetxt(fh, r, fmt="%.8e")
    n, box, got = read_cnf_atoms(p)
    assert n == 2 and box == 3.0 and np.array_equal(got, np.loadtxt(p, skiprows=2))


@pytest.mark.gpu
def test_checkpoint_of_device_state(tmp_path):
    """write_cnf_sim == the driver's write_cnf_atoms(filename, n, box, r*box, v) on the downloaded state
    (md_nve_lj.py:196), in the original atom order, after steps that re-sorted the atoms."""
    from conftest import liquid_like
    from examples_b200.config_io_module import read_cnf_atoms, write_cnf_atoms, write_cnf_sim
    from examples_b200.device import Simulation
    from examples_b200.lattice import ran_velocities
    n, box, r = liquid_like(8, 0.75, 0.05)
    sim = Simulation(box, 2.5, r, ran_velocities(n, 1.0))
    sim.force()
    sim.run_nve(0.005, 25)
    p1, p2 = str(tmp_path / "a"), str(tmp_path / "b")
    write_cnf_sim(p1, sim)
    rd, vd, fd = sim.download()
    write_cnf_atoms(p2, n, box, rd * box, vd)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    nn, bb, rr, vv = read_cnf_atoms(p1, with_v=True)
    assert nn == n and abs(bb - box) < 1e-8 and np.max(np.abs(rr - rd * box)) <= 0.5e-10 * (1 + 1e-9)
    write_cnf_sim(p1, sim, with_v=False)
    assert np.array_equal(read_cnf_atoms(p1)[2], rr)
    sim.close()
