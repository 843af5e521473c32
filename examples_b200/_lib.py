"""ctypes binding of libatlj.so (include/atlj.h).  No CPU fallback: if the library is missing or no
CUDA device is usable, every compute call raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ATLJ_LIB", os.path.join(_HERE, "libatlj.so"))  # ATLJ_LIB: A/B runs of another build

OK, ERR_DIM, ERR_SMALL, ERR_INDEX, ERR_MISMATCH, ERR_CUDA, ERR_ARG = 0, -1, -2, -3, -4, -100, -101
MODEL_LJ, MODEL_WCA = 0, 1
NREC = 12

# the reference's assertion messages (SURVEY.md 8b "Errors")
_ASSERTS = {
    ERR_DIM: "Dimension error in force",            # md_lj_module.py:77
    ERR_SMALL: "System is too small for cells",     # md_lj_ll_module.py:107
    ERR_INDEX: "Index error",                       # md_lj_ll_module.py:109
    ERR_MISMATCH: "Dimension mismatch in hessian",  # md_lj_module.py:163
}

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_lib = None


class AtljError(RuntimeError):
    pass


def lib():
    """Load libatlj.so (built by __graft_entry__.build() / make -C examples_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AtljError("%s not found: build it with `make -C examples_b200/csrc` (nvcc, sm_100a). "
                        "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.atlj_version.restype = C.c_int
    L.atlj_last_error.restype = C.c_char_p
    L.atlj_device_count.restype = C.c_int
    L.atlj_set_device.argtypes = [C.c_int]
    L.atlj_kernel_launches.restype = C.c_int64
    L.atlj_set_multi_max.restype = C.c_int64
    L.atlj_set_multi_max.argtypes = [C.c_int64]
    L.atlj_force_lj.argtypes = [_dp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    L.atlj_force_le.argtypes = [_dp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                C.c_int, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    L.atlj_hessian_lj.argtypes = [_dp, _dp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                  C.POINTER(C.c_double)]
    L.atlj_cells.argtypes = [_dp, C.c_int64, C.c_int, C.c_int, C.c_double, _i64p, _i32p, _i32p]
    L.atlj_neighbour_cells.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p]
    L.atlj_sim_create.restype = _vp
    L.atlj_sim_create.argtypes = [C.c_int64, _vp]
    L.atlj_sim_create_ex.restype = _vp
    L.atlj_sim_create_ex.argtypes = [C.c_int64, _vp, C.c_int]
    L.atlj_sim_destroy.restype = None
    L.atlj_sim_destroy.argtypes = [_vp]
    L.atlj_sim_configure.argtypes = [_vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                     C.c_int]
    L.atlj_sim_upload.argtypes = [_vp, _vp, _vp, _vp, C.c_int64]
    L.atlj_sim_download.argtypes = [_vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64)]
    L.atlj_sim_natoms.restype = C.c_int64
    L.atlj_sim_natoms.argtypes = [_vp]
    L.atlj_sim_force.argtypes = [_vp, C.c_double, C.c_int, _vp]
    L.atlj_sim_run_nve.argtypes = [_vp, C.c_double, C.c_int, C.c_int, _vp]
    L.atlj_sim_run_baoab.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, _vp]
    L.atlj_sim_run_baoab_noise.argtypes = [_vp, C.c_double, C.c_double, C.c_double, _vp, C.c_int, _vp]
    L.atlj_sim_run_nhc.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_int, _vp, _vp, _vp, C.c_int, _vp]
    L.atlj_sim_run_npt.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, _vp, _vp,
                                   _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.c_int, _vp]
    L.atlj_sim_run_sllod.argtypes = [_vp, C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int, _vp]
    L.atlj_force_1.argtypes = [_vp, _vp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _vp, _vp,
                               C.POINTER(C.c_int)]
    L.atlj_cnf_format.argtypes = [_vp, _vp, C.c_int64, C.c_double, _vp]
    L.atlj_cnf_parse.argtypes = [_vp, C.c_int64, C.c_int, _vp, _vp]
    L.atlj_sim_cnf_format.argtypes = [_vp, C.c_int, _vp]
    L.atlj_sim_mg_enable.argtypes = [_vp, C.c_int64, C.c_int64]
    L.atlj_sim_set_total.restype = None
    L.atlj_sim_set_total.argtypes = [_vp, C.c_int64]
    L.atlj_nccl_unique_id.argtypes = [C.c_char_p]
    L.atlj_sim_mg_connect.argtypes = [_vp, C.c_int, C.c_int, C.c_char_p]
    L.atlj_sim_run_nve_mg.argtypes = [_vp, C.c_double, C.c_int, _vp]
    L.atlj_sim_run_nve_mg_hess.argtypes = [_vp, C.c_double, C.c_int, _vp]
    L.atlj_sim_mg_phase3_force.argtypes = [_vp]
    L.atlj_sim_mg_phase4_hess.argtypes = [_vp, C.c_double, _vp]
    L.atlj_sim_mg_phase1.argtypes = [_vp, C.c_double]
    L.atlj_sim_mg_phase2.argtypes = [_vp]
    L.atlj_sim_mg_phase3.argtypes = [_vp, C.c_double, _vp]
    L.atlj_sim_mg_set_peers.argtypes = [_vp, _vp, _vp]
    L.atlj_sim_mg_trace.argtypes = [_vp, C.c_int, _vp]
    L.atlj_sim_run_baoab_mg.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, _vp]
    L.atlj_sim_run_nhc_mg.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_int, _vp, _vp, _vp, C.c_int, _vp]
    L.atlj_sim_mg_phase1_baoab.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int]
    L.atlj_sim_mg_nhc_open.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_double]
    L.atlj_sim_mg_phase1_nhc.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    L.atlj_sim_mg_nhc_close.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_int, _vp]
    L.atlj_sim_mg_nhc_finish.argtypes = [_vp, C.c_int, _vp, _vp]
    L.atlj_sim_mg_blob_bytes.restype = C.c_int64
    L.atlj_sim_mg_ipc_blob.argtypes = [_vp, C.c_char_p]
    L.atlj_sim_mg_ipc_open.argtypes = [_vp, C.c_char_p, C.c_char_p]
    L.atlj_sim_nve_step_host.argtypes = [_vp, C.c_double, _vp, _vp, _vp, _vp, C.c_int64, _vp]
    L.atlj_sim_nve_step_host_mg.argtypes = [_vp, C.c_double, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64), _vp]
    L.atlj_sim_mg_own_cap.argtypes = [_vp]
    L.atlj_sim_mg_own_cap.restype = C.c_int64
    L.atlj_host_alloc.restype = _vp
    L.atlj_host_alloc.argtypes = [C.c_int64]
    L.atlj_host_free.restype = None
    L.atlj_host_free.argtypes = [_vp]
    L.atlj_sim_timer_start.argtypes = [_vp]
    L.atlj_sim_timer_stop.argtypes = [_vp, C.POINTER(C.c_double)]
    L.atlj_sim_sync.argtypes = [_vp]
    L.atlj_measure_fp64_peak.argtypes = [C.POINTER(C.c_double)]
    L.atlj_sim_enable_timing.argtypes = [_vp, C.c_int]
    L.atlj_sim_timing.argtypes = [_vp, _dp]
    _lib = L
    return L


def check(rc):
    """Map a library status to the reference's behaviour: AssertionError with the reference's text for its
    own assertion conditions, AtljError otherwise."""
    if rc == OK:
        return
    if rc in _ASSERTS:
        raise AssertionError(_ASSERTS[rc])
    msg = lib().atlj_last_error().decode("utf-8", "replace")
    raise AtljError("libatlj error %d: %s" % (rc, msg))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    return lib().atlj_device_count()


def set_multi_max(natoms):
    """Systems of at most `natoms` atoms use the eight-lanes-per-atom pair kernel (0: never).  -> previous value."""
    return int(lib().atlj_set_multi_max(int(natoms)))


def kernel_launches():
    return int(lib().atlj_kernel_launches())


def measure_fp64_peak():
    v = C.c_double(0.0)
    check(lib().atlj_measure_fp64_peak(C.byref(v)))
    return v.value


class _PinnedBlock:
    """One page-locked block handed out as the memory of a NumPy array (buffer protocol, PEP 688).  NumPy keeps a
    reference to the exporter for as long as the array or any view of it lives; when the last one is gone the block
    goes back to its pool."""

    def __init__(self, pool, ptr, nbytes):
        self.pool, self.ptr, self.nbytes = pool, ptr, nbytes
        self._c = (C.c_char * max(nbytes, 1)).from_address(ptr)

    def __buffer__(self, flags):
        return memoryview(self._c)

    def __del__(self):
        try:
            self.pool._give_back(self.ptr, self.nbytes)
        except Exception:  # interpreter shutdown
            pass


class PinnedPool:
    """Result arrays in page-locked host memory: force() returns `f` as an ordinary ndarray whose memory the GPU writes
    by DMA (no staging copy, no page faults of a fresh np.empty).  A block is reused once the caller has dropped the
    array (`total, f = force(...)` in the drivers' loops rebinds f every step, md_nve_lj.py:187)."""

    KEEP = 3          # free blocks kept per size
    MIN_BYTES = 1 << 20   # smaller results are plain NumPy arrays

    def __init__(self, alloc=None, free=None):
        self._alloc = alloc or (lambda n: lib().atlj_host_alloc(n))
        self._free = free or (lambda p: lib().atlj_host_free(p))
        self._idle = {}

    def empty(self, shape, dtype=np.float64):
        count = int(np.prod(shape))
        nbytes = count * np.dtype(dtype).itemsize
        if nbytes < self.MIN_BYTES:
            return np.empty(shape, dtype=dtype)
        idle = self._idle.get(nbytes)
        ptr = idle.pop() if idle else self._alloc(nbytes)
        if not ptr:
            return np.empty(shape, dtype=dtype)      # page-locked memory exhausted: pageable result, staged copy
        blk = _PinnedBlock(self, ptr, nbytes)
        return np.frombuffer(blk, dtype=dtype, count=count).reshape(shape)

    def _give_back(self, ptr, nbytes):
        idle = self._idle.setdefault(nbytes, [])
        if len(idle) < self.KEEP:
            idle.append(ptr)
        else:
            self._free(ptr)


_result_pool = None


def result_empty(shape, dtype=np.float64):
    """An uninitialised result array for a drop-in call (page-locked when large)."""
    global _result_pool
    if _result_pool is None:
        _result_pool = PinnedPool()
    return _result_pool.empty(shape, dtype)


class PinnedArray:
    """A NumPy view of page-locked host memory (cudaHostAlloc) so that H2D/D2H copies are asynchronous DMA."""

    def __init__(self, shape, dtype=np.float64):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().atlj_host_alloc(self.nbytes)
        if not self.ptr:
            raise AtljError("pinned allocation failed: " + lib().atlj_last_error().decode())
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().atlj_host_free(self.ptr)
            self.ptr = None
