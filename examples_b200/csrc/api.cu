// C ABI of libatlj.so (include/atlj.h): device-resident simulation object + stateless host-array entry points.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "hostpipe.cuh"
#include "integrate.cuh"
#include "sim.cuh"

namespace atlj {

static thread_local char g_err[512] = "";
int64_t g_launches = 0;
static int pdl_default() {
  const char *e = getenv("ATLJ_PDL");  // bit mask of the kernels launched with the attribute (A/B runs)
  return e ? atoi(e) : (PDL_SCAN | PDL_SCATTER | PDL_GATHER | PDL_PAIR);
}
int g_pdl = pdl_default();
int g_pdl_off = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_ok() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no usable CUDA device (%s); libatlj has no CPU path", e == cudaSuccess ? "device count 0"
                                                                                       : cudaGetErrorString(e));
    return ATLJ_ERR_CUDA;
  }
  return ATLJ_OK;
}

// ---- timing -------------------------------------------------------------------------------------------------------
void Timer::begin(atlj_sim *s, int cat, cudaStream_t st, bool other_stream) {
  if (!s->timing) return;
  cudaEvent_t e0, e1;
  if (s->ev_free.size() >= 2) {
    e0 = s->ev_free.back(), s->ev_free.pop_back();
    e1 = s->ev_free.back(), s->ev_free.pop_back();
  } else {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
  }
  cudaEventRecord(e0, other_stream ? st : s->st);
  (other_stream ? s->ev_open2 : s->ev_open).push_back({e0, e1, cat});
}
void Timer::end(atlj_sim *s, cudaStream_t st, bool other_stream) {
  if (!s->timing) return;
  auto &open = other_stream ? s->ev_open2 : s->ev_open;
  auto &t = open.back();
  cudaEventRecord(t.e1, other_stream ? st : s->st);
  s->ev_done.push_back(t);
  open.pop_back();
}

// ---- allocation ---------------------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(T **p, size_t count) {
  ATLJ_CUDA(cudaMalloc((void **)p, sizeof(T) * (count ? count : 1)));
  return ATLJ_OK;
}

// An allocation that cudaIpcGetMemHandle can export on its own: at least 8 MB and a multiple of 2 MB (small cudaMalloc
// blocks are carved out of shared 2 MB pages and the handle would name the whole page); the last 256 bytes carry a
// signature the importing process verifies.
static const unsigned long long IPC_MAGIC = 0x61746c6a6d670002ull;  // "atljmg" 2
static int dev_alloc_ipc(void **p, size_t bytes, size_t *alloc_bytes, cudaStream_t st) {
  size_t want = bytes + 256;
  if (want < (8u << 20)) want = 8u << 20;
  want = (want + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  ATLJ_CUDA(cudaMalloc(p, want));
  ATLJ_CUDA(cudaMemsetAsync(*p, 0, want, st));
  ATLJ_CUDA(cudaMemcpyAsync((unsigned char *)*p + want - 256, &IPC_MAGIC, sizeof(IPC_MAGIC), cudaMemcpyHostToDevice, st));
  ATLJ_CUDA(cudaStreamSynchronize(st));
  *alloc_bytes = want;
  return ATLJ_OK;
}

static int ensure_tile_tab(atlj_sim *s, int ints) {
  if (ints <= s->tile_tab_ints) return ATLJ_OK;
  cudaFree(s->tile_tab);
  s->tile_tab = nullptr;
  int rc = dev_alloc(&s->tile_tab, (size_t)ints);
  if (rc) return rc;
  s->tile_tab_ints = ints;
  return ATLJ_OK;
}

static int ensure_cells(atlj_sim *s, int ncell) {
  if (ncell <= s->ncell_alloc) return ATLJ_OK;
  cudaFree(s->cb.cell_count);
  cudaFree(s->cb.cell_start);
  s->cb.cell_count = s->cb.cell_start = nullptr;
  int rc;
  if ((rc = dev_alloc(&s->cb.cell_count, (size_t)ncell))) return rc;
  // padded table: one end sentinel per x layer (Geom::cs_idx); sized for the worst padding (sc = 1)
  if (s->ipc_safe) {
    if ((rc = dev_alloc_ipc((void **)&s->cb.cell_start, sizeof(int) * (2 * (size_t)ncell + 2), &s->cs_alloc_bytes, s->st)))
      return rc;
  } else if ((rc = dev_alloc(&s->cb.cell_start, 2 * (size_t)ncell + 2))) {
    return rc;
  }
  ATLJ_CUDA(cudaMemsetAsync(s->cb.cell_count, 0, sizeof(int) * (size_t)ncell, s->st));
  ATLJ_CUDA(cudaMemsetAsync(s->cb.cell_start, 0, sizeof(int) * (2 * (size_t)ncell + 2), s->st));
  s->ncell_alloc = ncell;
  return ATLJ_OK;
}

static int ensure_partials(atlj_sim *s, int rows) {
  if (rows <= s->partials_rows) return ATLJ_OK;
  cudaFree(s->partials);
  s->partials = nullptr;
  int rc = dev_alloc(&s->partials, (size_t)rows * 8);
  if (rc) return rc;
  s->partials_rows = rows;
  return ATLJ_OK;
}

int sim_ensure_rec(atlj_sim *s, int64_t nrec) {
  if (nrec <= s->rec_cap) return ATLJ_OK;
  cudaFree(s->rec_dev);
  s->rec_dev = nullptr;
  int rc = dev_alloc(&s->rec_dev, (size_t)nrec * ATLJ_NREC);
  if (rc) return rc;
  s->rec_cap = nrec;
  return ATLJ_OK;
}

// The second-generation LJ kernel keeps fp16 copies of chunk-centred coordinates (sigma units): its chunks must stay
// within the fp16 range, which a cell wider than 40 sigma (a cutoff beyond 32 sigma) could break - such systems and the
// WCA model go through the first tiled kernel.
bool sim_tile2(const atlj_sim *s) {
  return s->pair_kernel == 0 && s->p.model == ATLJ_MODEL_LJ && s->p.box <= 40.0 * s->g.sc;
}

// mean atoms per cell of the whole system (all ranks)
double sim_per_cell(const atlj_sim *s) {
  const double cells = (double)s->g.sc * s->g.sc * s->g.sc;
  const long long nt = s->n_total > 0 ? s->n_total : s->n;
  return cells > 0 ? (double)nt / cells : 0.0;
}

static int ensure_masks(atlj_sim *s) {
  if (s->masks) return ATLJ_OK;
  const size_t bytes = pair_tile2_mask_bytes(s->cap);
  cudaError_t e = cudaMalloc((void **)&s->masks, bytes);
  if (e != cudaSuccess) {
    s->masks = nullptr;
    set_error("cudaMalloc(%zu) for the mask words -> %s", bytes, cudaGetErrorString(e));
    return ATLJ_ERR_CUDA;
  }
  return ATLJ_OK;
}

int sim_ensure_masks(atlj_sim *s) { return ensure_masks(s); }

// cells of ~1 atom (WCA cut-off): the direct kernel; ATLJ_PAIR_KERNEL=tile1 keeps the tiled one for A/B runs
static bool sim_direct(const atlj_sim *s) {
  const long long cells = s->g.n_own_cells();
  return s->use_cells && s->pair_kernel == 0 && !s->mg.on && cells > 0 && s->n < 3 * cells;
}

// small systems (the tiled kernels cannot fill the GPU): eight lanes per atom, exact arithmetic.  ATLJ_MULTI_MAX moves
// the threshold for A/B runs (0 switches the kernel off).
static int64_t g_multi_max = -1;  // atoms; -1: not initialised yet
static int64_t multi_max() {
  if (g_multi_max < 0) {
    const char *mm = getenv("ATLJ_MULTI_MAX");
    g_multi_max = mm ? atoll(mm) : 16384;
    if (g_multi_max < 0) g_multi_max = 0;
  }
  return g_multi_max;
}
static bool sim_multi(const atlj_sim *s) {
  return s->pair_kernel == 0 && !s->mg.on && s->n <= multi_max();
}

// ---- one force evaluation on the current positions ----------------------------------------------------------------------
// Cell route: wrap+assign -> scan -> scatter/sort/gather (atoms re-ordered into cell order, buffers swapped) -> pair
// kernel -> finalize.  All-pairs route (sc < 4): pair_all on the current order.
int sim_force(atlj_sim *s, double strain, bool with_hessian, double *rec_dev, bool check_index_allpairs,
              cudaEvent_t after_sort) {
  int rc;
  const int n = (int)s->n;
  Phys p = s->p;
  p.strain = strain;
  p.shift = (p.model == ATLJ_MODEL_WCA) ? (int)floor(strain * s->g.sc) : 0;  // md_lj_llle_module.py:112
  ATLJ_CUDA(cudaMemsetAsync(rec_dev, 0, sizeof(double) * ATLJ_NREC, s->st));
  if (s->use_cells) {
    const int cur = s->cur, alt = 1 - s->cur;
    Timer::begin(s, 1);
    // the Lees-Edwards x pre-correction (md_lj_llle_module.py:94) is applied by the caller (shim / A propagator)
    if ((rc = launch_cell_assign(s->st, s->g, s->r[cur], s->r[cur], n, false, strain, s->cb, nullptr))) return rc;
    if ((rc = launch_cell_scan(s->st, s->g, s->cb, 0))) return rc;
    if ((rc = launch_cell_sort_gather(s->st, s->g, s->cb, n, 0, s->id[cur], s->r[cur], s->v[cur], s->r[alt], s->v[alt],
                                      s->id[alt])))
      return rc;
    Timer::end(s);
    s->cur = alt;
    if (after_sort) ATLJ_CUDA(cudaEventRecord(after_sort, s->st));
    PairArgs a;
    a.r = s->r[alt];
    a.fin = nullptr;
    a.cell_start = s->cb.cell_start;
    a.cell_count = s->cb.cell_count;
    a.f = s->f;
    a.partials = s->partials;
    a.tile_tab = s->tile_tab;
    a.masks = nullptr;
    a.natoms = (int)s->n;
    a.per_cell = sim_per_cell(s);
    a.g = s->g;
    a.p = p;
    Timer::begin(s, 0);
    int nb;
    const bool multi = sim_multi(s);
    const bool direct = !multi && sim_direct(s);
    const bool tile2 = !multi && !direct && sim_tile2(s);
    // hessian after the tile2 force kernel: that kernel keeps its mask words, the hessian kernel walks them
    const bool hess2 = with_hessian && tile2;
    if (hess2) {
      if ((rc = ensure_masks(s))) return rc;
      a.masks = s->masks;
    }
    if (multi)
      nb = launch_pair_multi(s->st, a, false, n, true);
    else if (direct)
      nb = launch_pair_direct(s->st, a, false, n);
    else if (tile2)
      nb = launch_pair_tile2(s->st, a, s->cb.err + 2, hess2 ? 1 : 0);
    else
      nb = launch_pair_tile(s->st, a, false);
    Timer::end(s);
    if (nb < 0) return nb;
    s->fin = FinTail();
    if (s->defer_fin && !with_hessian) {  // the caller's next kernel finishes the totals (FinTail, pair_math.cuh)
      s->fin.partials = s->partials, s->fin.nb = nb, s->fin.model = p.model, s->fin.rec = rec_dev;
      s->fin.reset_counter = s->cb.err + 2, s->fin.rows_dev = tile2 ? s->cb.err + 6 : nullptr;
    } else if ((rc = launch_finalize(s->st, s->partials, nb, p.model, false, rec_dev, s->cb.err + 2, nullptr, 0,
                                     tile2 ? s->cb.err + 6 : nullptr))) {
      return rc;
    }
    s->defer_fin = false;
    if (with_hessian) {
      a.fin = s->f;
      a.f = nullptr;
      Timer::begin(s, 3);
      if (hess2)
        nb = launch_pair_tile2(s->st, a, s->cb.err + 2, 2);
      else if (multi)
        nb = launch_pair_multi(s->st, a, true, n, true);
      else
        nb = direct ? launch_pair_direct(s->st, a, true, n) : launch_pair_tile(s->st, a, true);
      Timer::end(s);
      if (nb < 0) return nb;
      if ((rc = launch_finalize(s->st, s->partials, nb, p.model, true, rec_dev, hess2 ? s->cb.err + 2 : nullptr, nullptr,
                                0, hess2 ? s->cb.err + 6 : nullptr)))
        return rc;
    }
  } else {
    if (after_sort) ATLJ_CUDA(cudaEventRecord(after_sort, s->st));
    if (check_index_allpairs) {  // honour md_lj_ll_module.py:107-109 even though sc < 4 is served by all pairs
      Geom g = s->g;
      if ((rc = launch_cell_assign(s->st, g, s->r[s->cur], nullptr, n, false, strain, s->cb, nullptr))) return rc;
    }
    int nb = 0;
    const bool multi = sim_multi(s);
    PairArgs a;
    a.r = s->r[s->cur], a.fin = nullptr, a.cell_start = nullptr, a.cell_count = nullptr, a.f = s->f;
    a.partials = s->partials, a.tile_tab = nullptr, a.masks = nullptr, a.natoms = n, a.per_cell = 0.0, a.g = s->g, a.p = p;
    Timer::begin(s, 0);
    if (multi) {
      if ((nb = launch_pair_multi(s->st, a, false, n, false)) < 0) return nb;
    } else if ((rc = launch_pair_all(s->st, s->r[s->cur], nullptr, n, p, s->f, s->partials, false, &nb))) {
      return rc;
    }
    Timer::end(s);
    if ((rc = launch_finalize(s->st, s->partials, nb, p.model, false, rec_dev))) return rc;
    if (with_hessian) {
      a.fin = s->f, a.f = nullptr;
      if (multi) {
        if ((nb = launch_pair_multi(s->st, a, true, n, false)) < 0) return nb;
      } else if ((rc = launch_pair_all(s->st, s->r[s->cur], s->f, n, p, nullptr, s->partials, true, &nb))) {
        return rc;
      }
      if ((rc = launch_finalize(s->st, s->partials, nb, p.model, true, rec_dev))) return rc;
    }
  }
  s->defer_fin = false;
  s->forces_valid = true;
  return ATLJ_OK;
}

// The integrators start with a half kick from s->f (md_nve_lj.py:168 computes the forces before the loop): after an
// upload without a force call the forces are evaluated here instead of silently integrating with f = 0.
int sim_ensure_forces(atlj_sim *s, double strain) {
  if (s->forces_valid || s->mg.on) return ATLJ_OK;  // a slab's first evaluation is its dt = 0 step
  int rc = sim_ensure_rec(s, 1);
  if (rc) return rc;
  return sim_force(s, strain, false, s->rec_dev, false);
}

// read back and clear the device error flags; maps to the reference's assertion codes
int sim_check_flags(atlj_sim *s) {
  int h[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // [0] index error, [1] capacity, [2..4] work / done counters, [5] peer timeout
  ATLJ_CUDA(cudaMemcpyAsync(h, s->cb.err, sizeof(h), cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  if (h[0] || h[1] || h[5]) ATLJ_CUDA(cudaMemsetAsync(s->cb.err, 0, sizeof(h), s->st));
  if (h[5]) {
    set_error("timed out waiting for a neighbour rank's message (slab exchange; ATLJ_PEER_TIMEOUT_MS): every kernel "
              "queued after the timeout was skipped, the state is that of the last complete step - upload it again "
              "before continuing");
    return ATLJ_ERR_CUDA;
  }
  if (h[0]) {
    set_error("Index error");
    return ATLJ_ERR_INDEX;
  }
  if (h[1]) {
    set_error("buffer capacity exceeded (migration / halo)");
    return ATLJ_ERR_ARG;
  }
  return ATLJ_OK;
}

int sim_copy_rec(atlj_sim *s, double *rec_host, int nsteps) {
  if (rec_host)
    ATLJ_CUDA(cudaMemcpyAsync(rec_host, s->rec_dev, sizeof(double) * ATLJ_NREC * (size_t)nsteps,
                              cudaMemcpyDeviceToHost, s->st));
  return sim_check_flags(s);
}

}  // namespace atlj

using namespace atlj;

// =====================================================================================================================
extern "C" {

int atlj_version(void) { return 100; }
const char *atlj_last_error(void) { return g_err; }
int64_t atlj_kernel_launches(void) { return g_launches; }

int64_t atlj_set_multi_max(int64_t natoms) {
  const int64_t old = multi_max();
  g_multi_max = natoms < 0 ? 0 : natoms;
  return old;
}

int atlj_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int atlj_set_device(int device) {
  int rc = device_ok();
  if (rc) return rc;
  ATLJ_CUDA(cudaSetDevice(device));
  return ATLJ_OK;
}

atlj_sim *atlj_sim_create(int64_t capacity, void *stream) { return atlj_sim_create_ex(capacity, stream, 0); }

atlj_sim *atlj_sim_create_ex(int64_t capacity, void *stream, int flags) {
  if (device_ok() != ATLJ_OK) return nullptr;
  if (capacity < 1) capacity = 1;
  if (capacity > 0x7fffffffll / 3 - 4096) {  // kernels index components with 3*n in 32-bit integers
    set_error("capacity %lld exceeds %lld atoms per GPU", (long long)capacity, 0x7fffffffll / 3 - 4096);
    return nullptr;
  }
  atlj_sim *s = new atlj_sim();
  s->st = (cudaStream_t)stream;
  // A/B runs: ATLJ_PAIR_KERNEL=tile1 routes LJ forces through the first tiled kernel as well
  const char *op = getenv("ATLJ_PAIR_KERNEL");
  s->pair_kernel = (op && strcmp(op, "tile1") == 0) ? 1 : 0;
  const char *uf = getenv("ATLJ_UNFUSED");  // A/B runs: the unfused 12-launch velocity-Verlet step
  s->unfused = uf && uf[0] == '1';
  const char *ng = getenv("ATLJ_NO_GRAPH");  // A/B runs: every step launched kernel by kernel
  s->no_graph = ng && ng[0] == '1';
  s->cap = capacity;
  s->ipc_safe = (flags & ATLJ_SIM_IPC) != 0;
  size_t c = (size_t)capacity;
  bool ok = true;
  for (int k = 0; k < 2 && ok; k++) {
    if (s->ipc_safe)
      ok = dev_alloc_ipc((void **)&s->r[k], sizeof(double) * 3 * c, &s->r_alloc_bytes, s->st) == ATLJ_OK;
    else
      ok = dev_alloc(&s->r[k], 3 * c) == ATLJ_OK;
    ok = ok && dev_alloc(&s->v[k], 3 * c) == ATLJ_OK && dev_alloc(&s->id[k], c) == ATLJ_OK;
  }
  if (ok && s->ipc_safe)  // (slab ranks read the forces of their halo atoms from the neighbours: hessian)
    ok = dev_alloc_ipc((void **)&s->f, sizeof(double) * 3 * c, &s->f_alloc_bytes, s->st) == ATLJ_OK;
  else
    ok = ok && dev_alloc(&s->f, 3 * c) == ATLJ_OK;
  ok = ok && dev_alloc(&s->cb.cellid, c) == ATLJ_OK &&
       dev_alloc(&s->cb.rank, c) == ATLJ_OK && dev_alloc(&s->cb.keys, c) == ATLJ_OK &&
       dev_alloc(&s->cb.block_sums, 8192) == ATLJ_OK && dev_alloc(&s->cb.err, 8) == ATLJ_OK &&
       dev_alloc(&s->scal, 64) == ATLJ_OK && dev_alloc(&s->vf_partials, 4 * 4096) == ATLJ_OK;
  if (ok) ok = ensure_partials(s, pair_grid_blocks()) == ATLJ_OK;
  if (ok) ok = cudaMemsetAsync(s->cb.err, 0, 8 * sizeof(int), s->st) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(s->cb.block_sums, 0, 8192 * sizeof(int), s->st) == cudaSuccess;  // scan tile states
  if (ok) ok = cudaMemsetAsync(s->scal, 0, 64 * sizeof(double), s->st) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(s->f, 0, 3 * c * sizeof(double), s->st) == cudaSuccess;
  if (ok) ok = sim_ensure_rec(s, 64) == ATLJ_OK;
  if (!ok) {
    atlj_sim_destroy(s);
    return nullptr;
  }
  return s;
}

void atlj_sim_destroy(atlj_sim *s) {
  if (!s) return;
  cudaStreamSynchronize(s->st);
  sim_mg_destroy(s);
  for (int k = 0; k < 2; k++) cudaFree(s->r[k]), cudaFree(s->v[k]), cudaFree(s->id[k]);
  cudaFree(s->f), cudaFree(s->cb.cellid), cudaFree(s->cb.rank), cudaFree(s->cb.keys), cudaFree(s->cb.block_sums);
  cudaFree(s->cb.err), cudaFree(s->cb.cell_count), cudaFree(s->cb.cell_start), cudaFree(s->partials);
  cudaFree(s->scal), cudaFree(s->rec_dev), cudaFree(s->noise_dev), cudaFree(s->vf_partials), cudaFree(s->tile_tab), cudaFree(s->masks);
  if (s->sw0) cudaEventDestroy(s->sw0), cudaEventDestroy(s->sw1);
  if (s->st_copy) cudaStreamDestroy(s->st_copy), cudaEventDestroy(s->ev_sorted), cudaEventDestroy(s->ev_copied);
  if (s->nve_graph) cudaGraphExecDestroy(s->nve_graph);
  if (s->st_cap) cudaStreamDestroy(s->st_cap), cudaStreamDestroy(s->st_side), cudaEventDestroy(s->ev_fork), cudaEventDestroy(s->ev_join);
  for (auto e : s->ev_free) cudaEventDestroy(e);
  for (auto &t : s->ev_done) cudaEventDestroy(t.e0), cudaEventDestroy(t.e1);
  for (auto &t : s->ev_open) cudaEventDestroy(t.e0), cudaEventDestroy(t.e1);
  for (auto &t : s->ev_open2) cudaEventDestroy(t.e0), cudaEventDestroy(t.e1);
  delete s;
}

int atlj_sim_configure(atlj_sim *s, int model, double box, double r_cut_box_sq, double box_sq, double pot_cut, int sc,
                       int x_lo, int x_hi) {
  if (!s) return ATLJ_ERR_ARG;
  if (model != ATLJ_MODEL_LJ && model != ATLJ_MODEL_WCA) {
    set_error("unknown model %d", model);
    return ATLJ_ERR_ARG;
  }
  s->p.model = model;
  s->p.box = box;
  s->p.box_sq = box_sq;
  s->p.r_cut_box_sq = r_cut_box_sq;
  s->p.pot_cut = pot_cut;
  s->p.strain = 0.0;
  s->p.shift = 0;
  s->use_cells = sc >= 4;
  if (sc < 1) sc = 1;
  if (sc > 1024) {
    set_error("sc = %d exceeds 1024 cells per edge", sc);
    return ATLJ_ERR_ARG;
  }
  if (x_lo < 0 || x_hi > sc || x_hi <= x_lo) {
    set_error("bad slab [%d,%d) for sc=%d", x_lo, x_hi, sc);
    return ATLJ_ERR_ARG;
  }
  Geom g;
  g.sc = sc;
  g.x_lo = x_lo;
  g.nx_own = x_hi - x_lo;
  g.h = (g.nx_own == sc) ? 0 : 1;
  g.nxl = g.nx_own + 2 * g.h;
  s->g = g;
  s->forces_valid = false;
  // prefilter threshold in cell units with a 1e-4 relative margin (error bound of the fp32 path ~1.5e-6)
  double rc_cell = sqrt(r_cut_box_sq) * sc;
  s->p.rc2_cell = (float)(rc_cell * rc_cell * (1.0 + 1e-4));
  s->configured = true;
  int rows = pair_grid_blocks();
  if (pair_tile_grid(g) > rows) rows = pair_tile_grid(g);
  if (pair_tile2_items(g) > rows) rows = pair_tile2_items(g);
  if (pair_direct_blocks((int)s->cap) > rows) rows = pair_direct_blocks((int)s->cap);
  {
    const int64_t m = s->cap < multi_max() ? s->cap : multi_max();
    if (m > 0 && pair_multi_blocks((int)m) > rows) rows = pair_multi_blocks((int)m);
  }
  int rc = ensure_partials(s, rows);
  if (rc) return rc;
  if ((rc = ensure_tile_tab(s, pair_tile2_tab_ints(g)))) return rc;
  return ensure_cells(s, g.ncell());
}

int atlj_sim_upload(atlj_sim *s, const double *r_host, const double *v_host, const int32_t *id_host, int64_t n) {
  if (!s || !r_host || n < 0 || n > s->cap) {
    set_error("upload: bad arguments (n=%lld, capacity=%lld)", (long long)n, s ? (long long)s->cap : -1ll);
    return ATLJ_ERR_ARG;
  }
  s->n = n;
  s->n_total = n;
  s->cur = 0;
  size_t b = sizeof(double) * 3 * (size_t)n;
  ATLJ_CUDA(cudaMemcpyAsync(s->r[0], r_host, b, cudaMemcpyHostToDevice, s->st));
  if (v_host)
    ATLJ_CUDA(cudaMemcpyAsync(s->v[0], v_host, b, cudaMemcpyHostToDevice, s->st));
  else
    ATLJ_CUDA(cudaMemsetAsync(s->v[0], 0, b, s->st));
  if (id_host)
    ATLJ_CUDA(cudaMemcpyAsync(s->id[0], id_host, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s->st));
  else {
    int rc = launch_iota(s->st, (int)n, s->id[0]);
    if (rc) return rc;
  }
  ATLJ_CUDA(cudaMemsetAsync(s->f, 0, b, s->st));
  s->forces_valid = false;  // the run_* entry points evaluate the forces first when they are not current
  s->mg.sorted = false;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));  // the host arrays may be reused by the caller
  return ATLJ_OK;
}

int64_t atlj_sim_natoms(atlj_sim *s) { return s ? s->n : -1; }

int atlj_sim_download(atlj_sim *s, double *r_host, double *v_host, double *f_host, int32_t *id_host, int64_t *n_out) {
  if (!s) return ATLJ_ERR_ARG;
  size_t b = sizeof(double) * 3 * (size_t)s->n;
  if (r_host) ATLJ_CUDA(cudaMemcpyAsync(r_host, s->r[s->cur], b, cudaMemcpyDeviceToHost, s->st));
  if (v_host) ATLJ_CUDA(cudaMemcpyAsync(v_host, s->v[s->cur], b, cudaMemcpyDeviceToHost, s->st));
  if (f_host) ATLJ_CUDA(cudaMemcpyAsync(f_host, s->f, b, cudaMemcpyDeviceToHost, s->st));
  if (id_host)
    ATLJ_CUDA(cudaMemcpyAsync(id_host, s->id[s->cur], sizeof(int) * (size_t)s->n, cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  if (n_out) *n_out = s->n;
  return ATLJ_OK;
}

int atlj_sim_force(atlj_sim *s, double strain, int with_hessian, double *rec_host) {
  if (!s || !s->configured) return ATLJ_ERR_ARG;
  int rc = sim_force(s, strain, with_hessian != 0, s->rec_dev, true);
  if (rc) return rc;
  if ((rc = launch_vf_sums(s->st, 3 * s->n, s->v[s->cur], s->f, s->partials, s->rec_dev + 4))) return rc;
  return sim_copy_rec(s, rec_host, 1);
}

}  // extern "C" (internal helpers of atlj_sim_run_nve follow)

// ---- the fused velocity-Verlet loop ------------------------------------------------------------------------------------
// Between two pair kernels ONE pass does the trailing half kick of the finished step (with its sums), the leading half
// kick, the drift, the wrap and the cell assignment, and finishes the record of the step that just ended (two-level
// fixed-order reduction of the pair kernel's per-item rows and of sum v^2, sum f^2 inside that kernel).  Per step:
// kdk_assign -> scan -> scatter -> sort+gather (which leaves cell_count zeroed for the next step) || item table -> pair
// kernel: 6 launches, 5 on the critical path, none of whose arguments depends on the step number when `replay` is set
// (the record pointer follows a device counter), so TWO steps (the double-buffer parity returns) are captured once as a
// CUDA graph and replayed.  Results are identical to the unfused sequence up to the order of the final sums.
//   rec: replay ? base of the records : record of the step that just ended (null at the first step)
static int nve_fused_step(atlj_sim *s, cudaStream_t st, cudaStream_t side, double dt, bool first_kick, double *rec,
                          int nb_prev, bool replay, int *nb_out) {
  int rc;
  const int n = (int)s->n;
  Phys p = s->p;
  p.strain = 0.0, p.shift = 0;
  const bool multi = sim_multi(s);
  const bool direct = s->use_cells && !multi && sim_direct(s);
  const bool tile2 = s->use_cells && !multi && !direct && sim_tile2(s);
  const int cur = s->cur, alt = s->use_cells ? 1 - cur : cur;
  KdkLoop loop;
  loop.step_ctr = replay ? s->cb.err + 7 : nullptr;
  loop.tab0 = tile2 ? s->tile_tab : nullptr;
  loop.counts_zeroed = true;
  loop.no_assign = !s->use_cells;
  loop.tail_in_scan = s->use_cells;  // the scan launch below finishes the record of the step that just ended
  Timer::begin(s, 2);
  const int nvf = launch_kdk_assign(st, s->g, n, s->r[cur], s->v[cur], s->f, 0.5 * dt, dt, p.box, first_kick, s->cb,
                                    s->vf_partials, rec, s->partials, nb_prev, p.model, nullptr,
                                    tile2 ? s->cb.err + 6 : nullptr, &loop);
  if (nvf < 0) return nvf;
  Timer::end(s);
  if (!s->use_cells) {
    // all-pairs route (sc < 4, the reference's N = 256 case): two launches per step
    PairArgs a;
    a.r = s->r[cur], a.fin = nullptr, a.cell_start = nullptr, a.cell_count = nullptr, a.f = s->f;
    a.partials = s->partials, a.tile_tab = nullptr, a.masks = nullptr, a.natoms = n, a.per_cell = 0.0, a.g = s->g, a.p = p;
    int nb = 0;
    Timer::begin(s, 0);
    if (multi) {
      if ((nb = launch_pair_multi(st, a, false, n, false)) < 0) return nb;
    } else if ((rc = launch_pair_all(st, s->r[cur], nullptr, n, p, s->f, s->partials, false, &nb))) {
      return rc;
    }
    Timer::end(s);
    *nb_out = nb;
    return ATLJ_OK;
  }
  Timer::begin(s, 1);
  RecTail tail;
  if (first_kick) {
    tail.rows = s->vf_partials, tail.nrows = nvf, tail.rec = rec, tail.step_ctr = loop.step_ctr, tail.model = p.model;
    tail.has_pair = 1, tail.err = s->cb.err;
  }
  if ((rc = launch_cell_scan(st, s->g, s->cb, 0, first_kick ? &tail : nullptr))) return rc;
  PairArgs a;
  a.r = s->r[alt], a.fin = nullptr, a.cell_start = s->cb.cell_start, a.cell_count = s->cb.cell_count;
  a.f = s->f, a.partials = s->partials, a.tile_tab = s->tile_tab, a.natoms = (int)s->n, a.g = s->g, a.p = p;
  a.per_cell = sim_per_cell(s);
  a.masks = nullptr;
  const bool fork = tile2 && side != nullptr;
  if (fork) {  // the item table needs cell_start only: beside the sort + gather
    ATLJ_CUDA(cudaEventRecord(s->ev_fork, st));
    ATLJ_CUDA(cudaStreamWaitEvent(side, s->ev_fork, 0));
    if ((rc = launch_pair_tile2(side, a, s->cb.err + 2, 0, 1, true)) < 0) return rc;
    ATLJ_CUDA(cudaEventRecord(s->ev_join, side));
  }
  if ((rc = launch_cell_sort_gather(st, s->g, s->cb, n, 0, s->id[cur], s->r[cur], s->v[cur], s->r[alt], s->v[alt],
                                    s->id[alt], nullptr, nullptr, true)))
    return rc;
  Timer::end(s);
  s->cur = alt;
  if (fork) ATLJ_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
  Timer::begin(s, 0);
  int nb;
  if (multi)
    nb = launch_pair_multi(st, a, false, n, true);
  else if (direct)
    nb = launch_pair_direct(st, a, false, n);
  else if (tile2)
    nb = launch_pair_tile2(st, a, s->cb.err + 2, 0, fork ? 2 : 0, true);
  else
    nb = launch_pair_tile(st, a, false);
  Timer::end(s);
  if (nb < 0) return nb;
  *nb_out = nb;
  return ATLJ_OK;
}

// everything the captured launches bake in
struct NveGraphKey {
  int64_t n, n_total;
  int kind, cur, sc, model, nb_prev, multi;
  double dt, box, box_sq, r_cut_box_sq, pot_cut;
  void *r0, *r1, *f, *rec, *partials, *tab, *cell_start;
};

// kind: 0 = two steps of the fused link-cell loop, 1 = one step of the all-pairs loop, 2 = the same with the hessian
template <typename Body>
static int nve_graph_prepare(atlj_sim *s, int kind, double dt, int nb_prev, Body body) {
  NveGraphKey key;
  memset(&key, 0, sizeof(key));
  key.n = s->n, key.n_total = s->n_total, key.kind = kind, key.cur = s->cur, key.sc = s->g.sc, key.model = s->p.model;
  key.nb_prev = nb_prev, key.multi = sim_multi(s) ? 1 : 0;
  key.dt = dt, key.box = s->p.box, key.box_sq = s->p.box_sq, key.r_cut_box_sq = s->p.r_cut_box_sq, key.pot_cut = s->p.pot_cut;
  key.r0 = s->r[0], key.r1 = s->r[1], key.f = s->f, key.rec = s->rec_dev, key.partials = s->partials, key.tab = s->tile_tab;
  key.cell_start = s->cb.cell_start;
  static_assert(sizeof(NveGraphKey) <= sizeof(s->nve_graph_key), "key storage");
  if (s->nve_graph && memcmp(&key, s->nve_graph_key, sizeof(key)) == 0) return ATLJ_OK;
  if (s->nve_graph) cudaGraphExecDestroy(s->nve_graph), s->nve_graph = nullptr;
  if (!s->st_cap) {
    ATLJ_CUDA(cudaStreamCreateWithFlags(&s->st_cap, cudaStreamNonBlocking));
    ATLJ_CUDA(cudaStreamCreateWithFlags(&s->st_side, cudaStreamNonBlocking));
    ATLJ_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    ATLJ_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
  }
  const int64_t l0 = g_launches;
  const int cur0 = s->cur;
  ATLJ_CUDA(cudaStreamBeginCapture(s->st_cap, cudaStreamCaptureModeThreadLocal));
  g_pdl_off = (s->n_total > 0 ? s->n_total : s->n) > 44000;  // (common.cuh: programmatic launches inside a graph)
  int rc = body(s->st_cap, s->st_side);
  g_pdl_off = 0;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(s->st_cap, &graph);
  s->nve_graph_launches = (int)(g_launches - l0);
  g_launches = l0;  // captured, not launched
  if (rc == ATLJ_OK && (e != cudaSuccess || s->cur != cur0)) {
    set_error("CUDA graph capture of the NVE step failed: %s", cudaGetErrorString(e));
    rc = ATLJ_ERR_CUDA;
  }
  s->cur = cur0;
  if (rc == ATLJ_OK) {
    e = cudaGraphInstantiate(&s->nve_graph, graph, 0);
    if (e != cudaSuccess) {
      set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e));
      s->nve_graph = nullptr;
      rc = ATLJ_ERR_CUDA;
    }
  }
  if (getenv("ATLJ_DEBUG_GRAPH")) {
    size_t nn = 0;
    if (graph) cudaGraphGetNodes(graph, nullptr, &nn);
    fprintf(stderr, "atlj: graph capture rc=%d (%s), %zu nodes, pdl mask %d: %s\n", rc, cudaGetErrorString(e), nn, g_pdl,
            rc == ATLJ_OK ? "replayed" : atlj_last_error());
  }
  if (graph) cudaGraphDestroy(graph);
  if (rc == ATLJ_OK) memcpy(s->nve_graph_key, &key, sizeof(key));
  else cudaGetLastError();
  return rc;
}

// one velocity-Verlet step of the all-pairs route (sc < 4: md_nve_lj.py:178-192 with md_lj_module.force, and the
// hessian of calc_variables, md_nve_lj.py:46): kick + drift + wrap -> pair kernel -> totals [-> hessian -> its sum] ->
// kick + sums.  ctr (replayed launches): rec is the base of the records and the row follows the device counter.
static int nve_allpairs_step(atlj_sim *s, cudaStream_t st, double dt, bool with_hessian, double *rec, int *ctr) {
  int rc, nb = 0;
  const int n = (int)s->n;
  const long long n3 = 3 * s->n;
  Phys p = s->p;
  p.strain = 0.0, p.shift = 0;
  Timer::begin(s, 2);
  if ((rc = launch_kick_drift(st, n3, s->r[s->cur], s->v[s->cur], s->f, 0.5 * dt, dt, p.box, nullptr))) return rc;
  Timer::end(s);
  const bool multi = sim_multi(s);
  PairArgs a;
  a.r = s->r[s->cur], a.fin = nullptr, a.cell_start = nullptr, a.cell_count = nullptr, a.f = s->f;
  a.partials = s->partials, a.tile_tab = nullptr, a.masks = nullptr, a.natoms = n, a.per_cell = 0.0, a.g = s->g, a.p = p;
  Timer::begin(s, 0);
  if (multi) {
    if ((nb = launch_pair_multi(st, a, false, n, false)) < 0) return nb;
  } else if ((rc = launch_pair_all(st, s->r[s->cur], nullptr, n, p, s->f, s->partials, false, &nb))) {
    return rc;
  }
  Timer::end(s);
  if ((rc = launch_finalize(st, s->partials, nb, p.model, false, rec, nullptr, nullptr, 0, nullptr, ctr))) return rc;
  if (with_hessian) {
    a.fin = s->f, a.f = nullptr;
    Timer::begin(s, 3);
    if (multi) {
      if ((nb = launch_pair_multi(st, a, true, n, false)) < 0) return nb;
    } else if ((rc = launch_pair_all(st, s->r[s->cur], s->f, n, p, nullptr, s->partials, true, &nb))) {
      return rc;
    }
    Timer::end(s);
    if ((rc = launch_finalize(st, s->partials, nb, p.model, true, rec, nullptr, nullptr, 0, nullptr, ctr))) return rc;
  }
  Timer::begin(s, 2);
  rc = launch_kick_sums(st, n3, s->v[s->cur], s->f, 0.5 * dt, s->vf_partials, rec + 4, s->cb.err + 4, nullptr, ctr);
  Timer::end(s);
  return rc;
}

extern "C" {

int atlj_sim_run_nve(atlj_sim *s, double dt, int nsteps, int with_hessian, double *rec_host) {
  if (!s || !s->configured || nsteps < 0) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc || (rc = sim_ensure_forces(s, 0.0))) return rc;
  const long long n3 = 3 * s->n;
  const double half_dt = 0.5 * dt;  // md_nve_lj.py:182: 0.5 * dt evaluated first
  if (!with_hessian && nsteps > 0 && !s->unfused) {
    const int n = (int)s->n;
    const bool tile2 = s->use_cells && !sim_multi(s) && !sim_direct(s) && sim_tile2(s);
    ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC * (size_t)nsteps, s->st));
    if (s->use_cells) ATLJ_CUDA(cudaMemsetAsync(s->cb.cell_count, 0, sizeof(int) * (size_t)s->g.ncell(), s->st));
    ATLJ_CUDA(cudaMemsetAsync(s->cb.err + 7, 0, sizeof(int), s->st));  // step counter of the replayed launches
    int nb_prev = 0, k = 0;
    // step 0 (no finished step to close), launched directly; it also warms the kernels' attributes up
    if ((rc = nve_fused_step(s, s->st, nullptr, dt, false, nullptr, 0, false, &nb_prev))) return rc;
    k = 1;
    // pairs of steps as one replayed graph (not while the per-kernel event timing is on)
    // (the all-pairs route keeps its atom order, so the captured pair of steps is simply the same step twice)
    auto two_steps = [&](cudaStream_t cap, cudaStream_t side) -> int {
      int rc2 = ATLJ_OK, nb = nb_prev;
      for (int j = 0; j < 2 && rc2 == ATLJ_OK; j++) rc2 = nve_fused_step(s, cap, side, dt, true, s->rec_dev, nb, true, &nb);
      return rc2 == ATLJ_OK && nb != nb_prev ? ATLJ_ERR_ARG : rc2;
    };
    if (!s->timing && !s->no_graph && nsteps - k >= 2 && nve_graph_prepare(s, 0, dt, nb_prev, two_steps) == ATLJ_OK) {
      for (; nsteps - k >= 2; k += 2) {
        ATLJ_CUDA(cudaGraphLaunch(s->nve_graph, s->st));
        g_launches += s->nve_graph_launches;
      }
    }
    for (; k < nsteps; k++)
      if ((rc = nve_fused_step(s, s->st, nullptr, dt, true, s->rec_dev + (size_t)(k - 1) * ATLJ_NREC, nb_prev, false,
                               &nb_prev)))
        return rc;
    double *rec_last = s->rec_dev + (size_t)(nsteps - 1) * ATLJ_NREC;
    Timer::begin(s, 2);
    if ((rc = launch_finalize(s->st, s->partials, nb_prev, s->p.model, false, rec_last, s->cb.err + 2, nullptr, 0,
                              tile2 ? s->cb.err + 6 : nullptr)))
      return rc;
    if ((rc = launch_kick_sums(s->st, n3, s->v[s->cur], s->f, half_dt, s->vf_partials, rec_last + 4, s->cb.err + 4))) return rc;
    Timer::end(s);
    (void)n;
    s->step_count += nsteps;
    return sim_copy_rec(s, rec_host, nsteps);
  }
  if (!s->use_cells && nsteps > 0 && !s->unfused) {
    // all-pairs route (the reference's N = 256 case): launch-latency bound, so one step is captured and replayed
    ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC * (size_t)nsteps, s->st));
    ATLJ_CUDA(cudaMemsetAsync(s->cb.err + 7, 0, sizeof(int), s->st));
    int *ctr = s->cb.err + 7;
    int k = 0;
    if ((rc = nve_allpairs_step(s, s->st, dt, with_hessian != 0, s->rec_dev, ctr))) return rc;  // warms the kernels up
    k = 1;
    auto one_step = [&](cudaStream_t cap, cudaStream_t) -> int {
      return nve_allpairs_step(s, cap, dt, with_hessian != 0, s->rec_dev, ctr);
    };
    if (!s->timing && !s->no_graph && nsteps - k >= 2 &&
        nve_graph_prepare(s, with_hessian ? 2 : 1, dt, 0, one_step) == ATLJ_OK) {
      for (; k < nsteps; k++) {
        ATLJ_CUDA(cudaGraphLaunch(s->nve_graph, s->st));
        g_launches += s->nve_graph_launches;
      }
    }
    for (; k < nsteps; k++)
      if ((rc = nve_allpairs_step(s, s->st, dt, with_hessian != 0, s->rec_dev, ctr))) return rc;
    s->step_count += nsteps;
    return sim_copy_rec(s, rec_host, nsteps);
  }
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    Timer::begin(s, 2);
    if ((rc = launch_kick_drift(s->st, n3, s->r[s->cur], s->v[s->cur], s->f, half_dt, dt, s->p.box, nullptr)))
      return rc;
    Timer::end(s);
    if ((rc = sim_force(s, 0.0, with_hessian != 0, rec, false))) return rc;
    Timer::begin(s, 2);
    if ((rc = launch_kick_sums(s->st, n3, s->v[s->cur], s->f, half_dt, s->partials, rec + 4, s->cb.err + 4))) return rc;
    Timer::end(s);
  }
  s->step_count += nsteps;
  return sim_copy_rec(s, rec_host, nsteps);
}

int atlj_sim_nve_step_host(atlj_sim *s, double dt, double *r_host, double *v_host, double *f_host, int32_t *id_host,
                           int64_t n, double *rec_host) {
  if (!s || !s->configured || !r_host || !v_host || !f_host || !id_host || n < 1 || n > s->cap) {
    set_error("nve_step_host: bad arguments");
    return ATLJ_ERR_ARG;
  }
  int rc;
  s->n = n;
  s->n_total = n;
  s->cur = 0;
  const size_t b = sizeof(double) * 3 * (size_t)n;
  ATLJ_CUDA(cudaMemcpyAsync(s->r[0], r_host, b, cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(s->v[0], v_host, b, cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(s->f, f_host, b, cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(s->id[0], id_host, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s->st));
  const long long n3 = 3 * n;
  const double half_dt = 0.5 * dt;
  if (!s->st_copy) {
    ATLJ_CUDA(cudaStreamCreateWithFlags(&s->st_copy, cudaStreamNonBlocking));
    ATLJ_CUDA(cudaEventCreateWithFlags(&s->ev_sorted, cudaEventDisableTiming));
    ATLJ_CUDA(cudaEventCreateWithFlags(&s->ev_copied, cudaEventDisableTiming));
  }
  if ((rc = launch_kick_drift(s->st, n3, s->r[0], s->v[0], s->f, half_dt, dt, s->p.box, nullptr))) return rc;
  if ((rc = sim_force(s, 0.0, false, s->rec_dev, false, s->ev_sorted))) return rc;
  // positions and ids are final once the cell sort is done: they travel home on the copy stream while the pair
  // kernel runs; v and f follow on the main stream
  ATLJ_CUDA(cudaStreamWaitEvent(s->st_copy, s->ev_sorted, 0));
  ATLJ_CUDA(cudaMemcpyAsync(r_host, s->r[s->cur], b, cudaMemcpyDeviceToHost, s->st_copy));
  ATLJ_CUDA(cudaMemcpyAsync(id_host, s->id[s->cur], sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->st_copy));
  ATLJ_CUDA(cudaEventRecord(s->ev_copied, s->st_copy));
  if ((rc = launch_kick_sums(s->st, n3, s->v[s->cur], s->f, half_dt, s->partials, s->rec_dev + 4, s->cb.err + 4))) return rc;
  ATLJ_CUDA(cudaMemcpyAsync(v_host, s->v[s->cur], b, cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(f_host, s->f, b, cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaStreamWaitEvent(s->st, s->ev_copied, 0));
  return sim_copy_rec(s, rec_host, 1);  // synchronises
}

void *atlj_host_alloc(int64_t bytes) {
  if (device_ok() != ATLJ_OK || bytes < 0) return nullptr;
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    set_error("cudaHostAlloc(%lld) -> %s", (long long)bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}
void atlj_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int atlj_sim_timer_start(atlj_sim *s) {
  if (!s) return ATLJ_ERR_ARG;
  if (!s->sw0) {
    ATLJ_CUDA(cudaEventCreate(&s->sw0));
    ATLJ_CUDA(cudaEventCreate(&s->sw1));
  }
  ATLJ_CUDA(cudaEventRecord(s->sw0, s->st));
  return ATLJ_OK;
}
int atlj_sim_timer_stop(atlj_sim *s, double *ms) {
  if (!s || !s->sw0 || !ms) return ATLJ_ERR_ARG;
  ATLJ_CUDA(cudaEventRecord(s->sw1, s->st));
  ATLJ_CUDA(cudaEventSynchronize(s->sw1));
  float m = 0.f;
  ATLJ_CUDA(cudaEventElapsedTime(&m, s->sw0, s->sw1));
  *ms = m;
  return ATLJ_OK;
}
int atlj_sim_sync(atlj_sim *s) {
  if (!s) return ATLJ_ERR_ARG;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  return ATLJ_OK;
}

int atlj_sim_enable_timing(atlj_sim *s, int on) {
  if (!s) return ATLJ_ERR_ARG;
  s->timing = on != 0;
  return ATLJ_OK;
}

int atlj_sim_timing(atlj_sim *s, double ms[8]) {
  if (!s) return ATLJ_ERR_ARG;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  for (int k = 0; k < 8; k++) ms[k] = 0.0;
  for (auto &t : s->ev_done) {
    float m = 0.f;
    if (cudaEventElapsedTime(&m, t.e0, t.e1) == cudaSuccess) ms[t.cat & 7] += m;
    s->ev_free.push_back(t.e0);
    s->ev_free.push_back(t.e1);
  }
  s->ev_done.clear();
  return ATLJ_OK;
}

int atlj_measure_fp64_peak(double *flops_per_s) {
  int rc = device_ok();
  if (rc) return rc;
  double *scratch = nullptr;
  ATLJ_CUDA(cudaMalloc((void **)&scratch, 64));
  rc = measure_fp64_peak(0, scratch, flops_per_s);
  cudaFree(scratch);
  return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// stateless drop-in entry points: a cached simulation object sized to the largest n seen so far
// ---------------------------------------------------------------------------------------------------------------------
static atlj_sim *g_host_sim = nullptr;

static atlj_sim *host_sim(int64_t n) {
  if (g_host_sim && g_host_sim->cap >= n) return g_host_sim;
  if (g_host_sim) atlj_sim_destroy(g_host_sim);
  g_host_sim = atlj_sim_create(n + n / 8 + 1024, nullptr);
  return g_host_sim;
}

static int force_host(int model, const double *r_host, int64_t n, double box, double rcbsq, double bsq, double pot_cut,
                      double strain, int sc, int check_index, double *f_host, double totals[4], int *ovr,
                      int64_t *npair) {
  int rc = device_ok();
  if (rc) return rc;
  if (!r_host || !f_host || !totals || !ovr || n < 1) {
    set_error("force: bad arguments");
    return ATLJ_ERR_ARG;
  }
  if (check_index && sc < 3) return ATLJ_ERR_SMALL;  // md_lj_ll_module.py:107
  atlj_sim *s = host_sim(n);
  if (!s) return ATLJ_ERR_CUDA;
  if ((rc = atlj_sim_configure(s, model, box, rcbsq, bsq, pot_cut, sc > 0 ? sc : 1, 0, sc > 0 ? sc : 1))) return rc;
  // positions in (chunked through the page-locked ring when the caller's array is pageable; velocities are not part of
  // this call: the slot is zeroed because the cell sort carries it along)
  const size_t b = sizeof(double) * 3 * (size_t)n;
  s->n = n, s->n_total = n, s->cur = 0, s->forces_valid = false;
  ATLJ_CUDA(cudaMemsetAsync(s->v[0], 0, b, s->st));
  if ((rc = launch_iota(s->st, (int)n, s->id[0]))) return rc;
  if ((rc = hostpipe_h2d(s->st, s->r[0], r_host, b))) return rc;
  if ((rc = sim_force(s, strain, false, s->rec_dev, check_index != 0))) return rc;
  double *src = s->f;
  if (s->use_cells) {  // back to the caller's atom order
    src = s->r[1 - s->cur];
    if ((rc = launch_unsort(s->st, (int)n, s->id[s->cur], s->f, src))) return rc;
  }
  double rec[ATLJ_NREC];
  ATLJ_CUDA(cudaMemcpyAsync(rec, s->rec_dev, sizeof(rec), cudaMemcpyDeviceToHost, s->st));
  if ((rc = hostpipe_d2h(s->st, f_host, src, b))) return rc;  // synchronises
  if ((rc = sim_check_flags(s))) return rc;
  for (int k = 0; k < 4; k++) totals[k] = rec[k];
  *ovr = rec[7] != 0.0;
  if (npair) *npair = (int64_t)rec[8];
  return ATLJ_OK;
}

int atlj_force_lj(const double *r_host, int64_t n, double box, double r_cut_box_sq, double box_sq, double pot_cut,
                  int sc, int check_index, double *f_host, double totals[4], int *ovr, int64_t *npair) {
  return force_host(ATLJ_MODEL_LJ, r_host, n, box, r_cut_box_sq, box_sq, pot_cut, 0.0, sc, check_index, f_host, totals,
                    ovr, npair);
}

int atlj_force_1(const double *ri_host, const double *r_host, int64_t nj, double box, double r_cut_box_sq,
                 double box_sq, double pot_cut, double *f_partial_host, double totals[4], int *ovr) {
  int rc = device_ok();
  if (rc) return rc;
  if (!ri_host || !f_partial_host || !totals || !ovr || nj < 0 || (nj > 0 && !r_host)) {
    set_error("force_1: bad arguments");
    return ATLJ_ERR_ARG;
  }
  for (int k = 0; k < 4; k++) totals[k] = 0.0;
  *ovr = 0;
  if (nj == 0) return ATLJ_OK;
  atlj_sim *s = host_sim(nj);
  if (!s) return ATLJ_ERR_CUDA;
  if ((rc = atlj_sim_configure(s, ATLJ_MODEL_LJ, box, r_cut_box_sq, box_sq, pot_cut, 1, 0, 1))) return rc;
  const size_t b = sizeof(double) * 3 * (size_t)nj;
  ATLJ_CUDA(cudaMemcpyAsync(s->r[0], r_host, b, cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC, s->st));
  if ((rc = ensure_partials(s, (int)((nj + 127) / 128) + 1))) return rc;
  const int nb = launch_force_1(s->st, ri_host, s->r[0], (int)nj, s->p, s->f, s->partials);
  if (nb < 0) return nb;
  if ((rc = launch_finalize(s->st, s->partials, nb, ATLJ_MODEL_LJ, false, s->rec_dev))) return rc;
  ATLJ_CUDA(cudaMemcpyAsync(f_partial_host, s->f, b, cudaMemcpyDeviceToHost, s->st));
  double rec[ATLJ_NREC];
  if ((rc = sim_copy_rec(s, rec, 1))) return rc;
  *ovr = rec[7] != 0.0;
  if (*ovr) {  // smc_lj_module.py:146-149: nothing is usable after an overlap
    memset(f_partial_host, 0, b);
    return ATLJ_OK;
  }
  for (int k = 0; k < 4; k++) totals[k] = rec[k];
  return ATLJ_OK;
}

int atlj_force_le(const double *r_host, int64_t n, double box, double r_cut_box_sq, double box_sq, double strain,
                  int sc, int shift, int check_index, double *f_host, double totals[4], int *ovr, int64_t *npair) {
  // the library derives the shift itself (md_lj_llle_module.py:112); the argument states the caller's value and must
  // agree, otherwise the two sides disagree about the sheared stencil
  if (sc >= 4 && shift != (int)floor(strain * sc)) {
    set_error("force_le: shift %d is not floor(strain * sc) = %d (md_lj_llle_module.py:112)", shift,
              (int)floor(strain * sc));
    return ATLJ_ERR_ARG;
  }
  return force_host(ATLJ_MODEL_WCA, r_host, n, box, r_cut_box_sq, box_sq, 0.0, strain, sc, check_index, f_host, totals,
                    ovr, npair);
}

int atlj_hessian_lj(const double *r_host, const double *f_host, int64_t n, double box, double r_cut_box_sq,
                    double box_sq, int sc, int check_index, double *hes) {
  int rc = device_ok();
  if (rc) return rc;
  if (!r_host || !f_host || !hes || n < 1) {
    set_error("hessian: bad arguments");
    return ATLJ_ERR_ARG;
  }
  if (check_index && sc < 3) return ATLJ_ERR_SMALL;
  atlj_sim *s = host_sim(n);
  if (!s) return ATLJ_ERR_CUDA;
  if ((rc = atlj_sim_configure(s, ATLJ_MODEL_LJ, box, r_cut_box_sq, box_sq, 0.0, sc > 0 ? sc : 1, 0, sc > 0 ? sc : 1)))
    return rc;
  // the caller's forces ride through the sort in the velocity slot
  if ((rc = atlj_sim_upload(s, r_host, f_host, nullptr, n))) return rc;
  Phys p = s->p;
  int nb = 0;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC, s->st));
  if (s->use_cells) {
    const int cur = s->cur, alt = 1 - cur;
    if ((rc = launch_cell_assign(s->st, s->g, s->r[cur], s->r[cur], (int)n, false, 0.0, s->cb, nullptr))) return rc;
    if ((rc = launch_cell_scan(s->st, s->g, s->cb, 0))) return rc;
    if ((rc = launch_cell_sort_gather(s->st, s->g, s->cb, (int)n, 0, s->id[cur], s->r[cur], s->v[cur], s->r[alt],
                                      s->v[alt], s->id[alt])))
      return rc;
    s->cur = alt;
    PairArgs a;
    a.r = s->r[alt];
    a.fin = s->v[alt];
    a.cell_start = s->cb.cell_start;
    a.cell_count = s->cb.cell_count;
    a.f = nullptr;
    a.partials = s->partials;
    a.tile_tab = s->tile_tab;
    a.masks = nullptr;
    a.natoms = (int)s->n;
    a.per_cell = sim_per_cell(s);
    a.g = s->g;
    a.p = p;
    nb = sim_multi(s) ? launch_pair_multi(s->st, a, true, (int)n, true)
                      : (sim_direct(s) ? launch_pair_direct(s->st, a, true, (int)n) : launch_pair_tile(s->st, a, true));
    if (nb < 0) return nb;
  } else {
    if (check_index)
      if ((rc = launch_cell_assign(s->st, s->g, s->r[s->cur], nullptr, (int)n, false, 0.0, s->cb, nullptr))) return rc;
    if (sim_multi(s)) {
      PairArgs a;
      a.r = s->r[s->cur], a.fin = s->v[s->cur], a.cell_start = nullptr, a.cell_count = nullptr, a.f = nullptr;
      a.partials = s->partials, a.tile_tab = nullptr, a.masks = nullptr, a.natoms = (int)n, a.per_cell = 0.0;
      a.g = s->g, a.p = p;
      if ((nb = launch_pair_multi(s->st, a, true, (int)n, false)) < 0) return nb;
    } else if ((rc = launch_pair_all(s->st, s->r[s->cur], s->v[s->cur], (int)n, p, nullptr, s->partials, true, &nb))) {
      return rc;
    }
  }
  if ((rc = launch_finalize(s->st, s->partials, nb, p.model, true, s->rec_dev))) return rc;
  double rec[ATLJ_NREC];
  if ((rc = sim_copy_rec(s, rec, 1))) return rc;
  *hes = rec[6];
  return ATLJ_OK;
}

int atlj_cells(const double *r_host, int64_t n, int sc, int le, double strain, int64_t *c_host, int32_t *perm_host,
               int32_t *cell_start_host) {
  int rc = device_ok();
  if (rc) return rc;
  if (!r_host || n < 1) return ATLJ_ERR_ARG;
  if (sc < 3) return ATLJ_ERR_SMALL;
  atlj_sim *s = host_sim(n);
  if (!s) return ATLJ_ERR_CUDA;
  if ((rc = atlj_sim_configure(s, le ? ATLJ_MODEL_WCA : ATLJ_MODEL_LJ, 1.0, 1.0, 1.0, 0.0, sc, 0, sc))) return rc;
  if ((rc = atlj_sim_upload(s, r_host, nullptr, nullptr, n))) return rc;
  long long *c_dev = nullptr;
  ATLJ_CUDA(cudaMalloc((void **)&c_dev, sizeof(long long) * 3 * (size_t)n));
  const int cur = s->cur, alt = 1 - cur;
  rc = launch_cell_assign(s->st, s->g, s->r[cur], s->r[cur], (int)n, le != 0, strain, s->cb, c_dev);
  if (!rc) rc = launch_cell_scan(s->st, s->g, s->cb, 0);
  if (!rc)
    rc = launch_cell_sort_gather(s->st, s->g, s->cb, (int)n, 0, s->id[cur], s->r[cur], s->v[cur], s->r[alt], s->v[alt],
                                 s->id[alt]);
  if (!rc) {
    s->cur = alt;
    cudaError_t e = cudaSuccess;
    if (c_host)
      e = cudaMemcpyAsync(c_host, c_dev, sizeof(long long) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess && perm_host)
      e = cudaMemcpyAsync(perm_host, s->id[alt], sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess && cell_start_host) {  // drop the per-layer sentinels: plain CSR of sc^3 + 1 entries
      const size_t sc2 = (size_t)sc * sc;
      e = cudaMemcpy2DAsync(cell_start_host, sizeof(int) * sc2, s->cb.cell_start, sizeof(int) * (sc2 + 1),
                            sizeof(int) * sc2, (size_t)sc, cudaMemcpyDeviceToHost, s->st);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(cell_start_host + sc2 * sc, s->cb.cell_start + s->g.cs_size() - 1, sizeof(int),
                            cudaMemcpyDeviceToHost, s->st);
    }
    if (e != cudaSuccess) {
      set_error("atlj_cells copy: %s", cudaGetErrorString(e));
      rc = ATLJ_ERR_CUDA;
    }
  }
  int rc2 = sim_check_flags(s);  // synchronises
  cudaFree(c_dev);
  return rc ? rc : rc2;
}

__global__ void neighbour_table_kernel(Geom g, int le, int shift, int *__restrict__ nb, int *__restrict__ cnt) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.ncell()) return;
  int sc = g.sc;
  int lx = c / (sc * sc), cy = (c / sc) % sc, cz = c % sc;
  int m = 0;
  for (int k = 0; k < MAX_NB; k++) {
    int nc = neighbour_cell(g, le != 0, shift, lx, cy, cz, k);
    nb[(size_t)c * MAX_NB + k] = nc;
    if (nc >= 0) m++;
  }
  cnt[c] = m;
}

int atlj_neighbour_cells(int sc, int le, int shift, int32_t *nb_host, int32_t *count_host) {
  int rc = device_ok();
  if (rc) return rc;
  if (sc < 3 || sc > 256 || !nb_host || !count_host) return ATLJ_ERR_ARG;
  Geom g;
  g.sc = sc, g.x_lo = 0, g.nx_own = sc, g.h = 0, g.nxl = sc;
  int nc = g.ncell();
  int *nb = nullptr, *cnt = nullptr;
  ATLJ_CUDA(cudaMalloc((void **)&nb, sizeof(int) * (size_t)nc * MAX_NB));
  ATLJ_CUDA(cudaMalloc((void **)&cnt, sizeof(int) * (size_t)nc));
  neighbour_table_kernel<<<(nc + 127) / 128, 128>>>(g, le, shift, nb, cnt);
  ATLJ_LAUNCHED();
  ATLJ_CUDA(cudaMemcpy(nb_host, nb, sizeof(int) * (size_t)nc * MAX_NB, cudaMemcpyDeviceToHost));
  ATLJ_CUDA(cudaMemcpy(count_host, cnt, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost));
  cudaFree(nb);
  cudaFree(cnt);
  return ATLJ_OK;
}

}  // extern "C"
