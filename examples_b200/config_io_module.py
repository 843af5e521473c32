"""read_cnf_atoms / write_cnf_atoms of the reference's config_io_module.py (:30-48, :74-95) for multi-million-atom
configurations: same file format, same signatures, same assertion messages; the text body is formatted / parsed on
the GPU (csrc/cnf.cu), byte-for-byte what np.savetxt(fmt='%15.10f') writes and bit-for-bit what np.loadtxt returns.

np.savetxt / np.loadtxt cost minutes at 16 M atoms (one Python-level format call per row); block-end checkpoints
(md_nve_lj.py:194-196) would dominate a device-resident run.

The GPU parser is the fast path for the fixed layout write_cnf_atoms produces (n*cols fields of 16 bytes).  Any
other layout the reference's reader accepts - files written by the Fortran programs with `(*(f15.10))` (15-character
fields, no separator column), CRLF line ends, a trailing blank line, exponent notation, other widths - is read the
way config_io_module.py:37 reads it, with np.loadtxt on the host: this is text input, not the force path, and a
cnf.inp from the Fortran `initialize` must still start md_nve_lj.py.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

FIELD = 16                      # '%15.10f' + one separator


def _header(nn, box):
    return ("{:15d}\n{:15.8f}".format(nn, box) + "\n").encode("ascii")     # config_io_module.py:82 + savetxt's newline


def write_cnf_atoms(filename, nn, box, *args):
    """Write out atomic configuration (config_io_module.py:74)."""
    nargs = len(args)
    assert nargs == 1 or nargs == 2, "{}{:d}".format('Wrong number of arguments ', nargs)
    n, d = args[0].shape
    assert n == nn, "r shape mismatch {:d}{:d}".format(n, nn)
    assert d == 3, "r shape mismatch {:d}".format(d)
    r = _lib.as_f64(args[0])
    v = None
    if nargs == 2:
        n, d = args[1].shape
        assert n == nn, "v shape mismatch {:d}{:d}".format(n, nn)
        assert d == 3, "v shape mismatch {:d}".format(d)
        v = _lib.as_f64(args[1])
    cols = 3 if v is None else 6
    body = np.empty(nn * cols * FIELD, dtype=np.uint8)
    _lib.check(_lib.lib().atlj_cnf_format(r.ctypes.data, None if v is None else v.ctypes.data, nn, 1.0,
                                          body.ctypes.data))
    with open(filename, "wb") as f:
        f.write(_header(nn, box))
        body.tofile(f)


def write_cnf_sim(filename, sim, with_v=True):
    """Checkpoint a device-resident Simulation: write_cnf_atoms(filename, n, box, r*box, v) of md_nve_lj.py:196
    without bringing r, v to the host as doubles - the un-sort, the scaling by box and the formatting are one kernel."""
    nn, cols = sim.n_total, 6 if with_v else 3
    body = np.empty(nn * cols * FIELD, dtype=np.uint8)
    _lib.check(_lib.lib().atlj_sim_cnf_format(sim._h, 1 if with_v else 0, body.ctypes.data))
    with open(filename, "wb") as f:
        f.write(_header(nn, sim.box))
        body.tofile(f)


def _read_generic(filename, with_v):
    """config_io_module.py:30-48 as written there (np.loadtxt): every layout the reference's reader accepts."""
    with open(filename, "r") as f:
        n = int(f.readline())
        box = float(f.readline())
    rv = np.loadtxt(filename, skiprows=2, ndmin=2)
    rows, cols = rv.shape
    assert rows == n, "{:d}{}{:d}".format(rows, ' rows not equal to ', n)
    assert cols >= 3, "{:d}{}".format(cols, ' cols less than 3')
    r = rv[:, 0:3].astype(np.float64)
    if with_v:
        assert cols >= 6, "{:d}{}".format(cols, ' cols less than 6')
        return n, box, r, rv[:, 3:6].astype(np.float64)
    return n, box, r


def read_cnf_atoms(filename, with_v=False):
    """Read in atomic configuration (config_io_module.py:30)."""
    with open(filename, "rb") as f:
        n = int(f.readline())        # Number of atoms
        box = float(f.readline())    # Simulation box length (assumed cubic)
        start = f.tell()
        size = os.fstat(f.fileno()).st_size - start
        fixed = n > 0 and size > 0 and size % (n * FIELD) == 0
        cols = size // (n * FIELD) if fixed else 0
        body = np.fromfile(f, dtype=np.uint8, count=size) if fixed and cols >= 3 else None
    if body is not None:
        assert not with_v or cols >= 6, "{:d}{}".format(cols, ' cols less than 6')
        r = np.empty((n, 3))
        v = np.empty((n, 3)) if with_v else None
        rc = _lib.lib().atlj_cnf_parse(body.ctypes.data, n, int(cols), r.ctypes.data,
                                       None if v is None else v.ctypes.data)
        if rc != _lib.ERR_ARG:       # ERR_ARG: right size, but not the fixed '%15.10f ' notation -> generic reader
            _lib.check(rc)
            if with_v:
                return n, box, r, v
            return n, box, r
    return _read_generic(filename, with_v)
