"""Device-resident simulation: positions, velocities and forces stay in HBM between steps.

Replaces the per-step loop bodies of the reference drivers -- md_nve_lj.py:178-192 (velocity Verlet),
bd_nvt_lj.py:221-233 (BAOAB), md_nvt_lj.py:270-288 (Nose-Hoover chain), md_nvt_lj_le.py:226-240 (SLLOD) --
including the force call inside them.  Each run_* call returns one record per step with exactly the scalars the
drivers' calc_variables() need (SURVEY.md section 5 "Metrics"), so averages_module can be fed unchanged.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import MODEL_LJ, MODEL_WCA, NREC

# record columns (include/atlj.h)
REC_T0, REC_T1, REC_T2, REC_T3, REC_VSQ, REC_FSQ, REC_HES, REC_OVR, REC_NPAIR, REC_VXVY, REC_AUX, REC_NOWN = range(12)

WCA_R_CUT = (2.0 ** (1.0 / 3.0)) ** (1.0 / 2.0)  # md_lj_le_module.py:30-31


def _ptr(a):
    return None if a is None else a.ctypes.data


class Simulation:
    """One GPU's worth of atoms.  `model` is 'lj' (cut-and-shifted LJ, totals cut,pot,vir,lap) or 'wca'
    (WCA + Lees-Edwards, totals pot,vir,lap,pyx)."""

    def __init__(self, box, r_cut, r, v=None, model="lj", ids=None, capacity=None, stream=0, slab=None, sc=None,
                 ipc=False):
        L = _lib.lib()
        self.model = MODEL_LJ if model == "lj" else MODEL_WCA
        if self.model == MODEL_WCA:
            r_cut = WCA_R_CUT
        self.box, self.r_cut = float(box), float(r_cut)
        r = _lib.as_f64(r)
        n = r.shape[0]
        assert r.ndim == 2 and r.shape[1] == 3, 'Dimension error in force'
        self.n_total = n
        # host scalars with the reference's expressions (md_lj_module.py:80-88)
        r_cut_box = self.r_cut / self.box
        self.r_cut_box_sq = r_cut_box ** 2
        self.box_sq = self.box ** 2
        sr2 = 1.0 / self.r_cut ** 2
        sr6 = sr2 ** 3
        self.pot_cut = sr6 ** 2 - sr6
        self.sc = math.floor(self.box / self.r_cut)
        if sc is not None:      # a coarser grid than the reference's finds the same pairs (cells wider than r_cut)
            assert 1 <= sc <= self.sc, "cells must not be narrower than r_cut"
            self.sc = int(sc)
        x_lo, x_hi = (0, max(self.sc, 1)) if slab is None else slab
        cap = int(capacity if capacity is not None else n + n // 8 + 1024)
        self._h = L.atlj_sim_create_ex(cap, C.c_void_p(stream), 1 if ipc else 0)   # 1 = ATLJ_SIM_IPC
        if not self._h:
            raise _lib.AtljError("atlj_sim_create failed: " + L.atlj_last_error().decode())
        _lib.check(L.atlj_sim_configure(self._h, self.model, self.box, self.r_cut_box_sq, self.box_sq, self.pot_cut,
                                        max(self.sc, 1), x_lo, x_hi))
        self.upload(r, v, ids)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().atlj_sim_destroy(h)
            except Exception:  # interpreter shutdown: the module globals are already gone
                pass

    __del__ = close

    # ---- host <-> device ---------------------------------------------------------------------------------------
    def upload(self, r, v=None, ids=None):
        r = _lib.as_f64(r)
        v = None if v is None else _lib.as_f64(v)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int32)
        _lib.check(_lib.lib().atlj_sim_upload(self._h, _ptr(r), _ptr(v), _ptr(ids), r.shape[0]))

    @property
    def natoms(self):
        return int(_lib.lib().atlj_sim_natoms(self._h))

    def download_sorted(self):
        """r, v, f, ids in the order held on the device (cell order)."""
        n = self.natoms
        r, v, f = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
        ids = np.empty(n, dtype=np.int32)
        nout = C.c_int64(0)
        _lib.check(_lib.lib().atlj_sim_download(self._h, _ptr(r), _ptr(v), _ptr(f), _ptr(ids), C.byref(nout)))
        return r, v, f, ids

    def download(self):
        """r, v, f in the caller's original atom order (single-GPU: ids are 0..n-1)."""
        r, v, f, ids = self.download_sorted()
        out = []
        for a in (r, v, f):
            b = np.empty_like(a)
            b[ids] = a
            out.append(b)
        return tuple(out)

    # ---- dynamics ---------------------------------------------------------------------------------------------
    def force(self, strain=0.0, hessian=False):
        rec = np.zeros(NREC)
        _lib.check(_lib.lib().atlj_sim_force(self._h, float(strain), int(hessian), _ptr(rec)))
        return rec

    def run_nve(self, dt, nsteps, hessian=False, record=True):
        rec = np.zeros((nsteps, NREC)) if record else None
        _lib.check(_lib.lib().atlj_sim_run_nve(self._h, float(dt), int(nsteps), int(hessian), _ptr(rec)))
        return rec

    @staticmethod
    def _o_scalars(dt, gamma, temperature):
        # the two scalars of o_propagator, with the reference's own expressions (bd_nvt_lj.py:124-127)
        x = gamma * dt
        c = -np.expm1(-2 * x)
        c = np.sqrt(c)
        return float(np.exp(-x)), float(c * np.sqrt(temperature))

    def run_baoab(self, dt, gamma, temperature, nsteps, seed=1, noise=None):
        """BAOAB Langevin steps (bd_nvt_lj.py:221-233).  noise=None uses the device Philox generator; an array
        (nsteps, n, 3) indexed by original atom number replaces np.random.randn of bd_nvt_lj.py:127 (tests)."""
        rec = np.zeros((nsteps, NREC))
        decay, amp = self._o_scalars(dt, gamma, temperature)
        if noise is None:
            _lib.check(_lib.lib().atlj_sim_run_baoab(self._h, float(dt), decay, amp, int(seed), int(nsteps),
                                                     _ptr(rec)))
        else:
            noise = _lib.as_f64(noise)
            assert noise.shape == (nsteps, self.n_total, 3)
            _lib.check(_lib.lib().atlj_sim_run_baoab_noise(self._h, float(dt), decay, amp, _ptr(noise), int(nsteps),
                                                           _ptr(rec)))
        return rec

    def run_nhc(self, dt, temperature, eta, p_eta, q, nsteps, g=None):
        """Nose-Hoover chain steps (md_nvt_lj.py:270-288).  eta, p_eta (updated in place) and q are the length-m
        chain arrays of md_nvt_lj.py:245-257; g = 3*(n-1) as in md_nvt_lj.py:244 unless given."""
        rec = np.zeros((nsteps, NREC))
        for a in (eta, p_eta, q):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        if g is None:
            g = 3 * (self.n_total - 1)
        _lib.check(_lib.lib().atlj_sim_run_nhc(self._h, float(dt), float(temperature), float(g), len(q), _ptr(eta),
                                               _ptr(p_eta), _ptr(q), int(nsteps), _ptr(rec)))
        return rec

    def run_npt(self, dt, temperature, pressure, eta, p_eta, q, eta_baro, p_eta_baro, q_baro, w_eps, box0, eps,
                p_eps, nsteps, g=None):
        """Constant-NPT steps (md_npt_lj.py:344-366).  The chain arrays are updated in place; returns
        (records, eps, p_eps).  The box follows box0*exp(eps) (md_npt_lj.py:126): self.box, self.sc and the host
        scalars are those of the last step on return; rec[:, 9] holds the box length after every step."""
        rec = np.zeros((nsteps, NREC))
        for a in (eta, p_eta, q, eta_baro, p_eta_baro, q_baro):
            assert a.dtype == np.float64 and a.flags.c_contiguous and len(a) == len(q)
        if g is None:
            g = 3 * (self.n_total - 1)
        e, pe = C.c_double(float(eps)), C.c_double(float(p_eps))
        _lib.check(_lib.lib().atlj_sim_run_npt(self._h, float(dt), float(temperature), float(pressure), self.r_cut,
                                               float(g), len(q), _ptr(eta), _ptr(p_eta), _ptr(q), _ptr(eta_baro),
                                               _ptr(p_eta_baro), _ptr(q_baro), float(w_eps), float(box0),
                                               C.byref(e), C.byref(pe), int(nsteps), _ptr(rec)))
        if nsteps > 0:
            self.box = float(box0) * math.exp(e.value)
            self.box_sq = self.box ** 2
            self.r_cut_box_sq = (self.r_cut / self.box) ** 2
            self.sc = math.floor(self.box / self.r_cut)
        return rec, e.value, pe.value

    def run_sllod(self, dt, strain_rate, strain, nsteps):
        """-> records, strain after the last step (md_nvt_lj_le.py:226-240)."""
        rec = np.zeros((nsteps, NREC))
        s = C.c_double(float(strain))
        _lib.check(_lib.lib().atlj_sim_run_sllod(self._h, float(dt), float(strain_rate), C.byref(s), int(nsteps),
                                                 _ptr(rec)))
        return rec, s.value

    def nve_step_host(self, dt, r, v, f, ids):
        """One velocity-Verlet step with the state in HOST arrays (updated in place, cell order on return)."""
        rec = np.zeros(NREC)
        _lib.check(_lib.lib().atlj_sim_nve_step_host(self._h, float(dt), _ptr(r), _ptr(v), _ptr(f), _ptr(ids),
                                                     r.shape[0], _ptr(rec)))
        return rec

    def timer_start(self):
        _lib.check(_lib.lib().atlj_sim_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double(0.0)
        _lib.check(_lib.lib().atlj_sim_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def sync(self):
        _lib.check(_lib.lib().atlj_sim_sync(self._h))

    # ---- measurement ------------------------------------------------------------------------------------------
    def enable_timing(self, on=True):
        _lib.check(_lib.lib().atlj_sim_enable_timing(self._h, int(on)))

    def timing(self):
        """Accumulated device milliseconds {pair, cells, integrate, other} since the last call (CUDA events on
        the launching stream)."""
        ms = np.zeros(8)
        _lib.check(_lib.lib().atlj_sim_timing(self._h, ms))
        return dict(pair=ms[0], cells=ms[1], integrate=ms[2], other=ms[3], mig_exchange=ms[4], halo_pack=ms[5],
                    halo_exchange=ms[6], pair_boundary=ms[7])
