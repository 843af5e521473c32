// Slab domain decomposition across the GPUs of one box: migration + halo exchange + all-reduce.
//
// The reference has no counterpart (its only distributed code is replica exchange, mc_nvt_lj_re.f90); the yardstick of
// this path is the single-domain result on the same configuration.  Each rank owns a contiguous range of x cell
// layers [x_lo, x_hi) of the reference's cell grid (md_lj_ll_module.py:106-109) and keeps one halo layer on each side.
// One velocity-Verlet step (md_nve_lj.py:178-192) is
//   phase 1  kick + drift + wrap on the owned atoms; atoms that left the slab are packed (r, v, id) for the neighbour
//            -> exchange 1 (migration, tens of atoms per face)
//   phase 2  arrivals appended, owned atoms binned and sorted (same counting sort as the single-GPU path), first and
//            last owned layer packed: per-cell counts + positions, already in cell order
//            -> exchange 2 (halo, one layer of positions per face: ~1.5e5 atoms = 3.5 MB at 16 M atoms / 8 GPUs)
//   phase 3  halo layers unpacked behind the owned atoms ([owned | left halo | right halo]; the per-layer sentinels of
//            the padded cell_start make that legal), full-stencil pair kernel on the owned rows (no force return
//            traffic), kick, local sums -> all-reduce of the 12-double step record.
// Transport: NCCL point-to-point send/recv on the sim's stream (NVLink 5 / NVSwitch), bound at run time with dlopen so
// that libatlj has no link-time dependency on a particular NCCL build (torch's bundled 2.28 or the system 2.27).
// The phases are also exported one by one so that the exchange can be emulated with device copies on a single GPU
// (tests/test_slab.py) -- kernels of different ranks never wait on each other, so that emulation is safe.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"
#include "integrate.cuh"
#include "sim.cuh"

namespace atlj {

constexpr int HDR_INTS = 16;  // 64-byte header of every message; [0] = number of atoms

__host__ __device__ static inline size_t halo_counts_ints(int sc) { return (((size_t)sc * sc + 15) / 16) * 16; }

// ---- phase 1: pack the atoms that left the slab ------------------------------------------------------------------
__global__ void __launch_bounds__(256) migrate_pack_kernel(Geom g, int n, const int *__restrict__ n_dev,
                                                           const double *__restrict__ r,
                                                           const double *__restrict__ v, const int *__restrict__ id,
                                                           unsigned char *send_left, unsigned char *send_right,
                                                           int cap, int *__restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(n, *n_dev)) return;
  double x = r[3 * (size_t)i];
  x = __dsub_rn(x, rint(x));
  const long long c0 = (long long)floor(__dmul_rn(__dadd_rn(x, 0.5), (double)g.sc));  // md_lj_ll_module.py:108
  if (c0 >= g.x_lo && c0 < g.x_lo + g.nx_own) return;
  const int left = g.x_lo == 0 ? g.sc - 1 : g.x_lo - 1, right = g.x_lo + g.nx_own == g.sc ? 0 : g.x_lo + g.nx_own;
  unsigned char *buf;
  if (c0 == left)
    buf = send_left;
  else if (c0 == right)
    buf = send_right;
  else {
    atomicOr(&err[0], 1);  // moved more than one layer in a step, or a coordinate on the +0.5 face ('Index error')
    return;
  }
  const int slot = atomicAdd(reinterpret_cast<int *>(buf), 1);
  if (slot >= cap) {
    atomicOr(&err[1], 1);
    return;
  }
  double *d = reinterpret_cast<double *>(buf + sizeof(int) * HDR_INTS) + 7 * (size_t)slot;
#pragma unroll
  for (int k = 0; k < 3; k++) d[k] = r[3 * (size_t)i + k], d[3 + k] = v[3 * (size_t)i + k];
  d[6] = (double)id[i];
}

// ---- phase 2: append the arrivals behind the atoms already held ----------------------------------------------------
__global__ void __launch_bounds__(256) append_kernel(const unsigned char *__restrict__ recv_left,
                                                     const unsigned char *__restrict__ recv_right, int cap,
                                                     double *__restrict__ r, double *__restrict__ v,
                                                     int *__restrict__ id, const int *__restrict__ n_old_dev,
                                                     long long capacity, int *__restrict__ n_after,
                                                     int *__restrict__ err) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_old = *n_old_dev;
  const int cl = min(*reinterpret_cast<const int *>(recv_left), cap);
  const int cr = min(*reinterpret_cast<const int *>(recv_right), cap);
  if (t == 0) {
    *n_after = n_old + cl + cr;
    if ((long long)n_old + cl + cr > capacity) atomicOr(&err[1], 1);
  }
  if (t >= 2 * cap) return;
  const int which = t >= cap, k = t - which * cap;
  if (k >= (which ? cr : cl)) return;
  const long long dst = (long long)n_old + (which ? cl : 0) + k;
  if (dst >= capacity) return;
  const double *s = reinterpret_cast<const double *>((which ? recv_right : recv_left) + sizeof(int) * HDR_INTS) + 7 * (size_t)k;
#pragma unroll
  for (int c = 0; c < 3; c++) r[3 * dst + c] = s[c], v[3 * dst + c] = s[3 + c];
  id[dst] = (int)s[6];
}

// one owned layer -> message: header, per-cell counts, positions (already contiguous and in cell order)
__global__ void __launch_bounds__(256) halo_pack_kernel(Geom g, int lx, const int *__restrict__ cs,
                                                        const double *__restrict__ r_sorted, unsigned char *buf,
                                                        int cap, int *__restrict__ err) {
  const int sc2 = g.sc * g.sc;
  const int *row = cs + (size_t)lx * (sc2 + 1);
  int *hdr = reinterpret_cast<int *>(buf);
  int *cnt = hdr + HDR_INTS;
  double *pos = reinterpret_cast<double *>(cnt + halo_counts_ints(g.sc));
  const int start = row[0];
  int total = row[sc2] - start;
  if (total > cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&err[1], 1);
    total = cap;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) hdr[0] = total;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long c = gt; c < sc2; c += stride) cnt[c] = row[c + 1] - row[c];
  for (long long e = gt; e < 3ll * total; e += stride) pos[e] = r_sorted[3 * (size_t)start + e];
}

// ---- phase 3: a received layer becomes halo layer lx of the local grid ---------------------------------------------
// cell_start of the layer = base + exclusive scan of the per-cell counts (one CTA; sc^2 <= 1024 * 64)
__global__ void __launch_bounds__(1024) halo_cs_kernel(Geom g, int lx, const unsigned char *__restrict__ buf,
                                                       const unsigned char *__restrict__ before,
                                                       const int *__restrict__ n_own_dev, int *__restrict__ cs,
                                                       long long capacity, int cap, int *__restrict__ err) {
  __shared__ int wsum[32];
  const int n_own = *n_own_dev;
  const int sc2 = g.sc * g.sc;
  const int *hdr = reinterpret_cast<const int *>(buf);
  const int *cnt = hdr + HDR_INTS;
  int base = n_own + (before ? min(*reinterpret_cast<const int *>(before), cap) : 0);
  const int per = (sc2 + 1023) / 1024;
  const int c0 = threadIdx.x * per;
  int sum = 0;
  for (int k = 0; k < per; k++)
    if (c0 + k < sc2) sum += cnt[c0 + k];
  int incl = sum;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    int s = wsum[lane], si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    wsum[lane] = si - s;
  }
  __syncthreads();
  int off = base + wsum[w] + incl - sum;
  int *row = cs + (size_t)lx * (sc2 + 1);
  for (int k = 0; k < per; k++) {
    if (c0 + k < sc2) {
      row[c0 + k] = off;
      off += cnt[c0 + k];
      if (c0 + k == sc2 - 1) {
        row[sc2] = off;
        if (off > capacity || off - base != min(hdr[0], cap)) atomicOr(&err[1], 1);
      }
    }
  }
}

__global__ void __launch_bounds__(256) halo_copy_kernel(Geom g, const unsigned char *__restrict__ buf,
                                                        const unsigned char *__restrict__ before,
                                                        const int *__restrict__ n_own_dev, int cap,
                                                        long long capacity, double *__restrict__ r_sorted) {
  const int n_own = *n_own_dev;
  const int *hdr = reinterpret_cast<const int *>(buf);
  const double *pos = reinterpret_cast<const double *>(hdr + HDR_INTS + halo_counts_ints(g.sc));
  const long long base = n_own + (before ? min(*reinterpret_cast<const int *>(before), cap) : 0);
  long long total = min(hdr[0], cap);
  if (base + total > capacity) total = capacity - base > 0 ? capacity - base : 0;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long e = gt; e < 3 * total; e += stride) r_sorted[3 * base + e] = pos[e];
}

__global__ void set_nown_kernel(double *rec, const int *n_own) { rec[11] = (double)*n_own; }
// after the sort the number of owned atoms is the end sentinel of the last owned layer
__global__ void copy_int_kernel(int *dst, const int *src) { *dst = *src; }

// ---- NCCL, bound at run time ----------------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return ATLJ_OK;
  const char *names[] = {getenv("ATLJ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *nm : names) {
    if (!nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("cannot load NCCL (libnccl.so.2): %s", dlerror());
    return ATLJ_ERR_CUDA;
  }
#define ATLJ_SYM(field, name)                                              \
  do {                                                                     \
    *(void **)(&g_nccl.field) = dlsym(h, name);                            \
    if (!g_nccl.field) {                                                   \
      set_error("NCCL symbol %s not found", name);                         \
      return ATLJ_ERR_CUDA;                                                \
    }                                                                      \
  } while (0)
  ATLJ_SYM(GetUniqueId, "ncclGetUniqueId");
  ATLJ_SYM(CommInitRank, "ncclCommInitRank");
  ATLJ_SYM(CommDestroy, "ncclCommDestroy");
  ATLJ_SYM(Send, "ncclSend");
  ATLJ_SYM(Recv, "ncclRecv");
  ATLJ_SYM(AllReduce, "ncclAllReduce");
  ATLJ_SYM(GroupStart, "ncclGroupStart");
  ATLJ_SYM(GroupEnd, "ncclGroupEnd");
  ATLJ_SYM(GetErrorString, "ncclGetErrorString");
#undef ATLJ_SYM
  g_nccl.handle = h;
  return ATLJ_OK;
}

#define ATLJ_NCCL(call)                                                                                 \
  do {                                                                                                  \
    ncclResult_t r_ = (call);                                                                           \
    if (r_ != ncclSuccess) {                                                                            \
      atlj::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, atlj::g_nccl.GetErrorString(r_));    \
      return ATLJ_ERR_CUDA;                                                                             \
    }                                                                                                   \
  } while (0)

// my "to the left" message lands in the left neighbour's "from the right" buffer and vice versa.  With two ranks
// both neighbours are the same peer; sends and receives between one pair match in order, hence recv right first.
static int exchange(atlj_sim *s, cudaStream_t st, unsigned char *const send[2], unsigned char *const recv[2],
                    size_t bytes) {
  MgState &m = s->mg;
  ncclComm_t comm = (ncclComm_t)m.comm;
  ATLJ_NCCL(g_nccl.GroupStart());
  ATLJ_NCCL(g_nccl.Send(send[0], bytes, ncclChar, m.left, comm, st));
  ATLJ_NCCL(g_nccl.Send(send[1], bytes, ncclChar, m.right, comm, st));
  ATLJ_NCCL(g_nccl.Recv(recv[1], bytes, ncclChar, m.right, comm, st));
  ATLJ_NCCL(g_nccl.Recv(recv[0], bytes, ncclChar, m.left, comm, st));
  ATLJ_NCCL(g_nccl.GroupEnd());
  return ATLJ_OK;
}

// ---- the three phases ---------------------------------------------------------------------------------------------------
// The number of owned atoms changes every step and lives on the device (cnt_dev[1]); kernels are launched over the
// capacity and read it, so a step needs no host synchronisation.  The host copy s->n is refreshed at the end of a run.
static int phase1(atlj_sim *s, double dt) {
  int rc;
  MgState &m = s->mg;
  const int nup = (int)s->cap;
  const int *n_dev = m.cnt_dev + 1;
  Timer::begin(s, 2);
  if ((rc = launch_kick_drift(s->st, 3ll * nup, s->r[s->cur], s->v[s->cur], s->f, 0.5 * dt, dt, s->p.box, nullptr,
                              n_dev)))
    return rc;
  Timer::end(s);
  Timer::begin(s, 3);
  ATLJ_CUDA(cudaMemsetAsync(m.mig_send[0], 0, sizeof(int) * HDR_INTS, s->st));
  ATLJ_CUDA(cudaMemsetAsync(m.mig_send[1], 0, sizeof(int) * HDR_INTS, s->st));
  migrate_pack_kernel<<<(nup + 255) / 256, 256, 0, s->st>>>(s->g, nup, n_dev, s->r[s->cur], s->v[s->cur],
                                                            s->id[s->cur], m.mig_send[0], m.mig_send[1],
                                                            (int)m.mig_cap, s->cb.err);
  ATLJ_LAUNCHED();
  Timer::end(s);
  return ATLJ_OK;
}

// phase 2 in three pieces so that the two exchanges can hide behind independent work:
//   2a  cell assignment of the atoms already held (leavers are skipped)          -- overlaps the migration exchange
//   2b  arrivals appended and assigned, scan, scatter, sort+gather of the blocks of cells that cover the first and
//       last owned layer, halo messages packed
//   2c  sort+gather of the remaining (interior) blocks                            -- overlaps the halo exchange
static int phase2a(atlj_sim *s) {
  MgState &m = s->mg;
  Timer::begin(s, 1);
  int rc = launch_cell_assign(s->st, s->g, s->r[s->cur], s->r[s->cur], (int)s->cap, false, 0.0, s->cb, nullptr,
                              m.cnt_dev + 1, true);
  Timer::end(s);
  return rc;
}

static int phase2b(atlj_sim *s) {
  int rc;
  MgState &m = s->mg;
  const int cap = (int)m.mig_cap, nup = (int)s->cap;
  const int cur = s->cur, alt = 1 - cur;
  Timer::begin(s, 3);
  append_kernel<<<(2 * cap + 255) / 256, 256, 0, s->st>>>(m.mig_recv[0], m.mig_recv[1], cap, s->r[cur], s->v[cur],
                                                         s->id[cur], m.cnt_dev + 1, (long long)s->cap, m.cnt_dev,
                                                         s->cb.err);
  ATLJ_LAUNCHED();
  if ((rc = launch_cell_assign(s->st, s->g, s->r[cur], s->r[cur], 2 * cap, false, 0.0, s->cb, nullptr, m.cnt_dev, true,
                               m.cnt_dev + 1)))
    return rc;
  Timer::end(s);
  Timer::begin(s, 1);
  if ((rc = launch_cell_scan(s->st, s->g, s->cb, 0))) return rc;
  if ((rc = launch_cell_sort_gather(s->st, s->g, s->cb, nup, 0, s->id[cur], s->r[cur], s->v[cur], s->r[alt], s->v[alt],
                                    s->id[alt], m.cnt_dev, 1)))
    return rc;
  Timer::end(s);
  Timer::begin(s, 5);
  const int sc2 = s->g.sc * s->g.sc;
  copy_int_kernel<<<1, 1, 0, s->st>>>(m.cnt_dev + 1,
                                      s->cb.cell_start + (size_t)(s->g.h + s->g.nx_own - 1) * (sc2 + 1) + sc2);
  ATLJ_LAUNCHED();
  const int blocks = 148 * 2;
  halo_pack_kernel<<<blocks, 256, 0, s->st>>>(s->g, s->g.h, s->cb.cell_start, s->r[alt], m.halo_send[0],
                                              (int)m.halo_cap, s->cb.err);
  ATLJ_LAUNCHED();
  halo_pack_kernel<<<blocks, 256, 0, s->st>>>(s->g, s->g.h + s->g.nx_own - 1, s->cb.cell_start, s->r[alt],
                                              m.halo_send[1], (int)m.halo_cap, s->cb.err);
  ATLJ_LAUNCHED();
  Timer::end(s);
  return ATLJ_OK;
}

static int phase2c(atlj_sim *s) {
  const int cur = s->cur, alt = 1 - cur;
  Timer::begin(s, 1);
  int rc = launch_cell_sort_gather(s->st, s->g, s->cb, (int)s->cap, 0, s->id[cur], s->r[cur], s->v[cur], s->r[alt],
                                   s->v[alt], s->id[alt], s->mg.cnt_dev, 2);
  Timer::end(s);
  s->cur = alt;
  return rc;
}

// the received layers become the halo layers 0 and nxl-1 of the local grid, behind the owned atoms
static int unpack_halos(atlj_sim *s, cudaStream_t st) {
  MgState &m = s->mg;
  const Geom &g = s->g;
  const int *n_dev = m.cnt_dev + 1;
  halo_cs_kernel<<<1, 1024, 0, st>>>(g, 0, m.halo_recv[0], nullptr, n_dev, s->cb.cell_start, (long long)s->cap,
                                     (int)m.halo_cap, s->cb.err);
  ATLJ_LAUNCHED();
  halo_cs_kernel<<<1, 1024, 0, st>>>(g, g.nxl - 1, m.halo_recv[1], m.halo_recv[0], n_dev, s->cb.cell_start,
                                     (long long)s->cap, (int)m.halo_cap, s->cb.err);
  ATLJ_LAUNCHED();
  halo_copy_kernel<<<148, 256, 0, st>>>(g, m.halo_recv[0], nullptr, n_dev, (int)m.halo_cap, (long long)s->cap,
                                        s->r[1 - s->cur]);
  ATLJ_LAUNCHED();
  halo_copy_kernel<<<148, 256, 0, st>>>(g, m.halo_recv[1], m.halo_recv[0], n_dev, (int)m.halo_cap, (long long)s->cap,
                                        s->r[1 - s->cur]);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

// emulation path: phase 2c has already swapped the buffers, so "alt" of unpack_halos is the current one
static int unpack_halos_current(atlj_sim *s) {
  s->cur = 1 - s->cur;
  int rc = unpack_halos(s, s->st);
  s->cur = 1 - s->cur;
  return rc;
}

static PairArgs pair_args(atlj_sim *s) {
  PairArgs a;
  a.r = s->r[s->cur], a.fin = nullptr, a.cell_start = s->cb.cell_start, a.cell_count = s->cb.cell_count;
  a.f = s->f, a.partials = s->partials, a.g = s->g, a.p = s->p;
  a.p.strain = 0.0, a.p.shift = 0;
  return a;
}

// pair kernel over the owned rows (the halos are in place) + kick + local sums
static int phase3(atlj_sim *s, double dt, double *rec) {
  int rc;
  MgState &m = s->mg;
  ATLJ_CUDA(cudaMemsetAsync(rec, 0, sizeof(double) * ATLJ_NREC, s->st));
  const PairArgs a = pair_args(s);
  Timer::begin(s, 0);
  const int nb = (s->pair_kernel == 0 && a.p.model == ATLJ_MODEL_LJ) ? launch_pair_tile2(s->st, a, s->cb.err + 2)
                                                                   : launch_pair_tile(s->st, a, false);
  Timer::end(s);
  if (nb < 0) return nb;
  if ((rc = launch_finalize(s->st, s->partials, nb, a.p.model, false, rec, s->cb.err + 2))) return rc;
  Timer::begin(s, 2);
  if ((rc = launch_kick_sums(s->st, 3ll * s->cap, s->v[s->cur], s->f, 0.5 * dt, s->vf_partials, rec + 4, m.cnt_dev + 1)))
    return rc;
  set_nown_kernel<<<1, 1, 0, s->st>>>(rec, m.cnt_dev + 1);
  ATLJ_LAUNCHED();
  Timer::end(s);
  return ATLJ_OK;
}

// host copy of the owned count <-> device
static int push_n(atlj_sim *s) {
  s->mg.cnt_host[0] = (int)s->n;
  ATLJ_CUDA(cudaMemcpyAsync(s->mg.cnt_dev + 1, s->mg.cnt_host, sizeof(int), cudaMemcpyHostToDevice, s->st));
  return ATLJ_OK;
}
static int pull_n(atlj_sim *s) {
  ATLJ_CUDA(cudaMemcpyAsync(s->mg.cnt_host, s->mg.cnt_dev + 1, sizeof(int), cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  s->n = s->mg.cnt_host[0];
  return ATLJ_OK;
}

static void mg_free(atlj_sim *s) {
  MgState &m = s->mg;
  for (int k = 0; k < 2; k++) {
    cudaFree(m.mig_send[k]), cudaFree(m.mig_recv[k]), cudaFree(m.halo_send[k]), cudaFree(m.halo_recv[k]);
    m.mig_send[k] = m.mig_recv[k] = m.halo_send[k] = m.halo_recv[k] = nullptr;
  }
  cudaFree(m.cnt_dev);
  m.cnt_dev = nullptr;
  if (m.cnt_host) cudaFreeHost(m.cnt_host);
  m.cnt_host = nullptr;
  if (m.st2) cudaStreamDestroy(m.st2), cudaEventDestroy(m.e1), cudaEventDestroy(m.e2);
  m.st2 = nullptr;
  if (m.comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)m.comm);
  m.comm = nullptr;
  m.on = false;
}
void sim_mg_destroy(atlj_sim *s) { mg_free(s); }

}  // namespace atlj

using namespace atlj;

extern "C" {

int atlj_sim_mg_enable(atlj_sim *s, int64_t mig_cap, int64_t halo_cap) {
  if (!s || !s->configured || mig_cap < 1 || halo_cap < 1 || s->g.h != 1) {
    set_error("mg_enable: configure the sim with a proper sub-range of x layers first");
    return ATLJ_ERR_ARG;
  }
  if ((long long)s->g.sc * s->g.sc > 1024ll * 64) {
    set_error("mg_enable: sc = %d too large for the one-CTA halo scan", s->g.sc);
    return ATLJ_ERR_ARG;
  }
  mg_free(s);
  MgState &m = s->mg;
  m.mig_cap = mig_cap, m.halo_cap = halo_cap;
  m.mig_bytes = sizeof(int) * HDR_INTS + sizeof(double) * 7 * (size_t)mig_cap;
  m.halo_bytes = sizeof(int) * (HDR_INTS + halo_counts_ints(s->g.sc)) + sizeof(double) * 3 * (size_t)halo_cap;
  for (int k = 0; k < 2; k++) {
    ATLJ_CUDA(cudaMalloc((void **)&m.mig_send[k], m.mig_bytes));
    ATLJ_CUDA(cudaMalloc((void **)&m.mig_recv[k], m.mig_bytes));
    ATLJ_CUDA(cudaMalloc((void **)&m.halo_send[k], m.halo_bytes));
    ATLJ_CUDA(cudaMalloc((void **)&m.halo_recv[k], m.halo_bytes));
    ATLJ_CUDA(cudaMemsetAsync(m.mig_send[k], 0, m.mig_bytes, s->st));
    ATLJ_CUDA(cudaMemsetAsync(m.mig_recv[k], 0, m.mig_bytes, s->st));
    ATLJ_CUDA(cudaMemsetAsync(m.halo_send[k], 0, m.halo_bytes, s->st));
    ATLJ_CUDA(cudaMemsetAsync(m.halo_recv[k], 0, m.halo_bytes, s->st));
  }
  ATLJ_CUDA(cudaMalloc((void **)&m.cnt_dev, 16 * sizeof(int)));
  ATLJ_CUDA(cudaMemsetAsync(m.cnt_dev, 0, 16 * sizeof(int), s->st));
  ATLJ_CUDA(cudaHostAlloc((void **)&m.cnt_host, 16 * sizeof(int), cudaHostAllocDefault));
  {
    int lo = 0, hi = 0;
    ATLJ_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    ATLJ_CUDA(cudaStreamCreateWithPriority(&m.st2, cudaStreamNonBlocking, hi));  // exchanges get SM slots first
  }
  ATLJ_CUDA(cudaEventCreateWithFlags(&m.e1, cudaEventDisableTiming));
  ATLJ_CUDA(cudaEventCreateWithFlags(&m.e2, cudaEventDisableTiming));
  m.on = true;
  return push_n(s);
}

/* which: 0/1 migration send left/right, 2/3 migration recv from left/right, 4/5 halo send left/right,
 * 6/7 halo recv from left/right */
void *atlj_sim_mg_buffer(atlj_sim *s, int which) {
  if (!s || !s->mg.on || which < 0 || which > 7) return nullptr;
  MgState &m = s->mg;
  unsigned char *t[8] = {m.mig_send[0], m.mig_send[1], m.mig_recv[0], m.mig_recv[1],
                         m.halo_send[0], m.halo_send[1], m.halo_recv[0], m.halo_recv[1]};
  return t[which];
}
int64_t atlj_sim_mg_buffer_bytes(atlj_sim *s, int which) {
  if (!s || !s->mg.on || which < 0 || which > 7) return -1;
  return (int64_t)(which < 4 ? s->mg.mig_bytes : s->mg.halo_bytes);
}
int atlj_device_copy(void *dst, const void *src, int64_t bytes) {
  ATLJ_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice));
  return ATLJ_OK;
}
void atlj_sim_set_total(atlj_sim *s, int64_t n_total) {
  if (s) s->n_total = n_total;
}

int atlj_sim_mg_phase1(atlj_sim *s, double dt) {
  if (!s || !s->mg.on) return ATLJ_ERR_ARG;
  int rc = push_n(s);
  if (rc || (rc = phase1(s, dt))) return rc;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  return ATLJ_OK;
}
int atlj_sim_mg_phase2(atlj_sim *s) {
  if (!s || !s->mg.on) return ATLJ_ERR_ARG;
  int rc = phase2a(s);
  if (rc || (rc = phase2b(s)) || (rc = phase2c(s)) || (rc = pull_n(s))) return rc;
  return sim_check_flags(s);
}
int atlj_sim_mg_phase3(atlj_sim *s, double dt, double *rec_host) {
  if (!s || !s->mg.on) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, 1);
  if (rc) return rc;
  if ((rc = unpack_halos_current(s)) || (rc = phase3(s, dt, s->rec_dev))) return rc;
  return sim_copy_rec(s, rec_host, 1);
}

int atlj_nccl_unique_id(char *out128) {
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  ATLJ_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id) < 128 ? sizeof(id) : 128);
  return ATLJ_OK;
}

int atlj_sim_mg_connect(atlj_sim *s, int rank, int world, const char *id128) {
  if (!s || !s->mg.on || world < 2 || rank < 0 || rank >= world || !id128) return ATLJ_ERR_ARG;
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id) < 128 ? sizeof(id) : 128);
  ncclComm_t comm;
  ATLJ_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  MgState &m = s->mg;
  m.comm = comm, m.rank = rank, m.world = world;
  m.left = (rank + world - 1) % world, m.right = (rank + 1) % world;
  return ATLJ_OK;
}

int atlj_sim_run_nve_mg(atlj_sim *s, double dt, int nsteps, double *rec_host) {
  if (!s || !s->mg.on || !s->mg.comm || nsteps < 0) {
    set_error("run_nve_mg: enable and connect the slab first");
    return ATLJ_ERR_ARG;
  }
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc) return rc;
  MgState &m = s->mg;
  if ((rc = push_n(s))) return rc;
  // Per step, main stream | second (high-priority) stream:
  //   kick, drift, pack leavers        |
  //   assign the atoms already held    | migration exchange
  //   append + assign arrivals, scan, sort boundary blocks, pack halos
  //   sort interior blocks             | halo exchange, unpack
  //   pair kernel, finalize, kick      |
  // The NCCL kernels need a few free SM slots, which the saturating pair kernel never leaves, so the exchanges are
  // hidden behind the short-lived cell-build kernels instead (measured: behind the pair kernel they simply wait).
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    if ((rc = phase1(s, dt))) return rc;
    ATLJ_CUDA(cudaEventRecord(m.e1, s->st));
    ATLJ_CUDA(cudaStreamWaitEvent(m.st2, m.e1, 0));
    Timer::begin(s, 4, m.st2, true);
    if ((rc = exchange(s, m.st2, m.mig_send, m.mig_recv, m.mig_bytes))) return rc;
    Timer::end(s, m.st2, true);
    ATLJ_CUDA(cudaEventRecord(m.e2, m.st2));
    if ((rc = phase2a(s))) return rc;
    ATLJ_CUDA(cudaStreamWaitEvent(s->st, m.e2, 0));
    if ((rc = phase2b(s))) return rc;
    ATLJ_CUDA(cudaEventRecord(m.e1, s->st));
    ATLJ_CUDA(cudaStreamWaitEvent(m.st2, m.e1, 0));
    Timer::begin(s, 6, m.st2, true);
    if ((rc = exchange(s, m.st2, m.halo_send, m.halo_recv, m.halo_bytes))) return rc;
    if ((rc = unpack_halos(s, m.st2))) return rc;  // writes behind the owned atoms of r[alt]; 2c swaps cur
    Timer::end(s, m.st2, true);
    ATLJ_CUDA(cudaEventRecord(m.e2, m.st2));
    if ((rc = phase2c(s))) return rc;
    ATLJ_CUDA(cudaStreamWaitEvent(s->st, m.e2, 0));
    if ((rc = phase3(s, dt, rec))) return rc;
  }
  // the records are read on the host only now: ONE all-reduce for all steps
  if (nsteps > 0) {
    Timer::begin(s, 3);
    ATLJ_NCCL(g_nccl.AllReduce(s->rec_dev, s->rec_dev, (size_t)nsteps * ATLJ_NREC, ncclDouble, ncclSum,
                               (ncclComm_t)m.comm, s->st));
    Timer::end(s);
  }
  s->step_count += nsteps;
  if ((rc = pull_n(s))) return rc;
  rc = sim_copy_rec(s, rec_host, nsteps);
  if (rec_host)
    for (int k = 0; k < nsteps; k++)
      if (rec_host[(size_t)k * ATLJ_NREC + 7] > 0.0) rec_host[(size_t)k * ATLJ_NREC + 7] = 1.0;
  return rc;
}

}  // extern "C"
