// Slab domain decomposition across the GPUs of one box: migration + halo exchange + all-reduce.
//
// The reference has no counterpart (its only distributed code is replica exchange, mc_nvt_lj_re.f90); the yardstick of
// this path is the single-domain result on the same configuration.  Each rank owns a contiguous range of x cell
// layers [x_lo, x_hi) of the reference's cell grid (md_lj_ll_module.py:106-109) and keeps one halo layer on each side.
// One velocity-Verlet step (md_nve_lj.py:178-192) is
//   phase 1  kick + drift + wrap on the owned atoms; atoms that left the slab are packed (r, v, id) STRAIGHT INTO THE
//            NEIGHBOUR'S receive buffer over NVLink (tens of atoms per face)
//   phase 2  cell assignment of the atoms already held; arrivals appended and assigned; scan; counting sort; the
//            first and last owned layer are packed -- per-cell counts + positions, already in cell order -- again
//            straight into the neighbours' receive buffers (one layer per face: ~6e4 atoms = 1.4 MB at 2 M atoms/GPU)
//   phase 3  halo layers unpacked behind the owned atoms ([owned | left halo | right halo]; the per-layer sentinels of
//            the padded cell_start make that legal), full-stencil pair kernel on the owned rows (no force return
//            traffic), kick, local sums.  The 12-double step records are all-reduced ONCE per run (NCCL).
//
// Transport.  The exchange is fused into the pack kernels: peer memory is mapped once (CUDA IPC between the one-process-
// per-GPU ranks, plain pointers when several ranks are emulated in one process) and the pack kernels store over
// NVLink 5 / NVSwitch into the neighbour's buffer; the CTA that finishes last publishes the message with a system-
// scope release of a step stamp in the buffer's header, and the consumer's stream holds a one-thread kernel that
// acquires that stamp.  No copy, no SM-hungry communication kernel, no host synchronisation per step.
// Receive buffers are double-buffered by step parity; that is sufficient flow control because a rank cannot start
// step s before it has consumed its neighbours' step s-1 messages, which they sent only after consuming the step s-2
// messages that used the same parity.  (An NCCL send/recv version of this exchange measured 35 us latency per
// exchange and could not overlap with the saturating pair kernel: profiles/r1_slab_history.md.)
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"
#include "integrate.cuh"
#include "sim.cuh"

namespace atlj {

constexpr int HDR_INTS = 16;  // 64-byte header of every message: [0] = number of atoms, [1] = step stamp (ready flag)

__host__ __device__ static inline size_t halo_counts_ints(int sc) { return (((size_t)sc * sc + 15) / 16) * 16; }

__device__ __forceinline__ void publish(int *hdr, int count, int stamp) {
  hdr[0] = count;
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(hdr + 1), "r"(stamp) : "memory");
}

// the consumer side: spins until the producer (another GPU) has published the message of this step
// (gives up after ~4e9 clocks, about 2 s, and raises err[5]: a neighbour that died must not hang this GPU)
__global__ void wait_stamp_kernel(const int *hdr_a, const int *hdr_b, int stamp, int *err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long t0 = clock64();
  for (const int *h : {hdr_a, hdr_b}) {
    int v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(h + 1) : "memory");
      if (v == stamp) break;
      if (clock64() - t0 > 4000000000ll) {
        atomicOr(&err[5], 1);
        return;
      }
    }
  }
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
  const double *s =
      reinterpret_cast<const double *>((which ? recv_right : recv_left) + sizeof(int) * HDR_INTS) + 7 * (size_t)k;
#pragma unroll
  for (int c = 0; c < 3; c++) r[3 * dst + c] = s[c], v[3 * dst + c] = s[3 + c];
  id[dst] = (int)s[6];
}

// first / last owned layer -> message in the left / right neighbour's buffer (blockIdx.y = direction): header,
// per-cell counts, positions (already contiguous and in cell order).  done[dir] counts finished CTAs; the last one
// publishes.  Thread (0,0,0) also moves the new owned count (end sentinel of the last owned layer) to n_own.
__global__ void __launch_bounds__(256) halo_pack_kernel(Geom g, const int *__restrict__ cs,
                                                        const double *__restrict__ r_sorted, unsigned char *to_left,
                                                        unsigned char *to_right, int cap, int stamp,
                                                        int *__restrict__ done, int *__restrict__ err,
                                                        int *__restrict__ n_own, double *__restrict__ rec) {
  const int sc2 = g.sc * g.sc;
  const int dir = blockIdx.y;
  const int lx = dir == 0 ? g.h : g.h + g.nx_own - 1;
  const int *row = cs + (size_t)lx * (sc2 + 1);
  if (dir == 1 && blockIdx.x == 0 && threadIdx.x == 0) {
    *n_own = row[sc2];
    if (rec) rec[11] = (double)row[sc2];
  }
  int *hdr = reinterpret_cast<int *>(dir ? to_right : to_left);
  int *cnt = hdr + HDR_INTS;
  double *pos = reinterpret_cast<double *>(cnt + halo_counts_ints(g.sc));
  const int start = row[0];
  int total = row[sc2] - start;
  if (total > cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&err[1], 1);
    total = cap;
  }
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long c = gt; c < sc2; c += stride) cnt[c] = row[c + 1] - row[c];
  // 16-byte stores towards the neighbour (the destination is 64-byte aligned; the source only 8-byte aligned)
  const double *src = r_sorted + 3 * (size_t)start;
  const long long n2 = (3ll * total) >> 1;
  double2 *pos2 = reinterpret_cast<double2 *>(pos);
  for (long long e = gt; e < n2; e += stride) pos2[e] = make_double2(src[2 * e], src[2 * e + 1]);
  if (gt == 0 && ((3ll * total) & 1)) pos[3ll * total - 1] = src[3ll * total - 1];
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&done[dir], 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    publish(hdr, total, stamp);
    done[dir] = 0;
  }
}

// ---- phase 3: a received layer becomes halo layer lx of the local grid ---------------------------------------------
// cell_start of the layer = base + exclusive scan of the per-cell counts (one CTA; sc^2 <= 1024 * 64)
__global__ void __launch_bounds__(1024) halo_cs_kernel(Geom g, const unsigned char *__restrict__ from_left,
                                                       const unsigned char *__restrict__ from_right,
                                                       const int *__restrict__ n_own_dev, int *__restrict__ cs,
                                                       long long capacity, int cap, int *__restrict__ err) {
  __shared__ int wsum[32];
  // block 0: the left neighbour's layer -> local layer 0, behind the owned atoms; block 1: the right neighbour's
  // layer -> local layer nxl-1, behind the left halo
  const int lx = blockIdx.x == 0 ? 0 : g.nxl - 1;
  const unsigned char *buf = blockIdx.x == 0 ? from_left : from_right;
  const unsigned char *before = blockIdx.x == 0 ? nullptr : from_left;
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

__global__ void __launch_bounds__(256) halo_copy_kernel(Geom g, const unsigned char *__restrict__ from_left,
                                                        const unsigned char *__restrict__ from_right,
                                                        const int *__restrict__ n_own_dev, int cap,
                                                        long long capacity, double *__restrict__ r_sorted) {
  const unsigned char *buf = blockIdx.y == 0 ? from_left : from_right;
  const unsigned char *before = blockIdx.y == 0 ? nullptr : from_left;
  const int n_own = *n_own_dev;
  const int *hdr = reinterpret_cast<const int *>(buf);
  const double *pos = reinterpret_cast<const double *>(hdr + HDR_INTS + halo_counts_ints(g.sc));
  const long long base = n_own + (before ? min(*reinterpret_cast<const int *>(before), cap) : 0);
  long long total = min(hdr[0], cap);
  if (base + total > capacity) total = capacity - base > 0 ? capacity - base : 0;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long e = gt; e < 3 * total; e += stride) r_sorted[3 * base + e] = pos[e];
}


// ---- NCCL (only the per-run all-reduce of the step records), bound at run time -------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
#ifndef ATLJ_HALO_PACK_CTAS
#define ATLJ_HALO_PACK_CTAS 296
#endif
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
  ATLJ_SYM(AllReduce, "ncclAllReduce");
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

// ---- receive arena: [kind: migration, halo][from: left, right][parity] ------------------------------------------------
static inline size_t arena_off(const MgState &m, int kind, int from, int parity) {
  const size_t mig = m.mig_bytes, halo = m.halo_bytes;
  return kind == 0 ? (size_t)(2 * from + parity) * mig : 4 * mig + (size_t)(2 * from + parity) * halo;
}
static inline unsigned char *my_recv(const MgState &m, int kind, int from, int parity) {
  return m.arena + arena_off(m, kind, from, parity);
}
// where my message towards `dir` (0 = left, 1 = right) lands: the neighbour's "from the other side" buffer
static inline unsigned char *peer_recv(const MgState &m, int kind, int dir, int parity) {
  return m.peer_arena[dir] + arena_off(m, kind, 1 - dir, parity);
}

// device-side row count of the tiled LJ kernel's partials (null for the other pair kernels)
static const int *tile2_rows(atlj_sim *s) {
  return sim_tile2(s) ? s->cb.err + 6 : nullptr;
}

// ---- the phases -----------------------------------------------------------------------------------------------------------
// The number of owned atoms changes every step and lives on the device (cnt_dev[1]); kernels are launched over the
// capacity and read it, so a step needs no host synchronisation.  The host copy s->n is refreshed at the end of a run.

// phase 1: ONE kernel: [trailing half kick + record of the finished step] + leading half kick + drift + wrap + cell
// assignment of the atoms that stay + leavers packed into the neighbours' buffers and published
static int phase1(atlj_sim *s, double dt, bool first_kick, double *rec_prev, int nb_prev) {
  MgState &m = s->mg;
  const int par = (int)(m.step & 1);
  SlabArgs sl;
  sl.n_dev = m.cnt_dev + 1, sl.id = s->id[s->cur];
  sl.to_left = peer_recv(m, 0, 0, par), sl.to_right = peer_recv(m, 0, 1, par);
  sl.mig_cap = (int)m.mig_cap, sl.stamp = (int)(m.step + 1), sl.mig_cnt = m.cnt_dev + 4;
  Timer::begin(s, 2);
  const int nb = launch_kdk_assign(s->st, s->g, (int)s->cap, s->r[s->cur], s->v[s->cur], s->f, 0.5 * dt, dt, s->p.box,
                                   first_kick, s->cb, s->vf_partials, rec_prev, s->partials, nb_prev, s->p.model, &sl,
                                   tile2_rows(s));
  Timer::end(s);
  return nb < 0 ? nb : ATLJ_OK;
}

// phase 2: arrivals appended and assigned, scan, counting sort, boundary layers packed into the neighbours' buffers
static int phase2(atlj_sim *s, double *rec) {
  int rc;
  MgState &m = s->mg;
  const int cap = (int)m.mig_cap, nup = (int)s->cap, par = (int)(m.step & 1), stamp = (int)(m.step + 1);
  const int cur = s->cur, alt = 1 - cur;
  Timer::begin(s, 4);
  const unsigned char *ml = my_recv(m, 0, 0, par), *mr = my_recv(m, 0, 1, par);
  wait_stamp_kernel<<<1, 32, 0, s->st>>>(reinterpret_cast<const int *>(ml), reinterpret_cast<const int *>(mr), stamp,
                                         s->cb.err);
  ATLJ_LAUNCHED();
  Timer::end(s);
  Timer::begin(s, 3);
  append_kernel<<<(2 * cap + 255) / 256, 256, 0, s->st>>>(ml, mr, cap, s->r[cur], s->v[cur], s->id[cur], m.cnt_dev + 1,
                                                         (long long)s->cap, m.cnt_dev, s->cb.err);
  ATLJ_LAUNCHED();
  if ((rc = launch_cell_assign(s->st, s->g, s->r[cur], s->r[cur], 2 * cap, false, 0.0, s->cb, nullptr, m.cnt_dev, true,
                               m.cnt_dev + 1)))
    return rc;
  Timer::end(s);
  Timer::begin(s, 1);
  if ((rc = launch_cell_scan(s->st, s->g, s->cb, 0))) return rc;
  if ((rc = launch_cell_sort_gather(s->st, s->g, s->cb, nup, 0, s->id[cur], s->r[cur], s->v[cur], s->r[alt], s->v[alt],
                                    s->id[alt], m.cnt_dev)))
    return rc;
  Timer::end(s);
  s->cur = alt;
  Timer::begin(s, 5);
  halo_pack_kernel<<<dim3(ATLJ_HALO_PACK_CTAS, 2), 256, 0, s->st>>>(s->g, s->cb.cell_start, s->r[alt], peer_recv(m, 1, 0, par),
                                                    peer_recv(m, 1, 1, par), (int)m.halo_cap, stamp, m.cnt_dev + 8,
                                                    s->cb.err, m.cnt_dev + 1, rec);
  ATLJ_LAUNCHED();
  Timer::end(s);
  return ATLJ_OK;
}

static PairArgs pair_args(atlj_sim *s) {
  PairArgs a;
  a.r = s->r[s->cur], a.fin = nullptr, a.cell_start = s->cb.cell_start, a.cell_count = s->cb.cell_count;
  a.f = s->f, a.partials = s->partials, a.tile_tab = s->tile_tab, a.natoms = (int)s->n, a.g = s->g, a.p = s->p;
  a.per_cell = sim_per_cell(s);
  a.masks = nullptr;
  a.p.strain = 0.0, a.p.shift = 0;
  return a;
}

// phase 3: the received layers become the halo layers 0 and nxl-1 of the local grid, behind the owned atoms; pair
// kernel over the owned rows.  -> rows of partials written (the record is finished by the next phase 1 or by tail())
static int phase3(atlj_sim *s) {
  MgState &m = s->mg;
  const Geom &g = s->g;
  const int par = (int)(m.step & 1), stamp = (int)(m.step + 1);
  const int *n_dev = m.cnt_dev + 1;
  const unsigned char *hl = my_recv(m, 1, 0, par), *hr = my_recv(m, 1, 1, par);
  Timer::begin(s, 6);
  wait_stamp_kernel<<<1, 32, 0, s->st>>>(reinterpret_cast<const int *>(hl), reinterpret_cast<const int *>(hr), stamp,
                                         s->cb.err);
  ATLJ_LAUNCHED();
  Timer::end(s);
  Timer::begin(s, 3);
  halo_cs_kernel<<<2, 1024, 0, s->st>>>(g, hl, hr, n_dev, s->cb.cell_start, (long long)s->cap, (int)m.halo_cap,
                                        s->cb.err);
  ATLJ_LAUNCHED();
  halo_copy_kernel<<<dim3(444, 2), 256, 0, s->st>>>(g, hl, hr, n_dev, (int)m.halo_cap, (long long)s->cap, s->r[s->cur]);
  ATLJ_LAUNCHED();
  Timer::end(s);
  const PairArgs a = pair_args(s);
  Timer::begin(s, 0);
  const int nb = sim_tile2(s) ? launch_pair_tile2(s->st, a, s->cb.err + 2)
                                                                   : launch_pair_tile(s->st, a, false);
  Timer::end(s);
  m.step++;
  return nb;
}

// the record of the last step of a run: pair totals, trailing half kick, sum v^2, sum f^2
static int tail(atlj_sim *s, double dt, double *rec, int nb) {
  int rc;
  if ((rc = launch_finalize(s->st, s->partials, nb, s->p.model, false, rec, s->cb.err + 2, nullptr, 0, tile2_rows(s))))
    return rc;
  Timer::begin(s, 2);
  rc = launch_kick_sums(s->st, 3ll * s->cap, s->v[s->cur], s->f, 0.5 * dt, s->vf_partials, rec + 4, s->cb.err + 4, s->mg.cnt_dev + 1);
  Timer::end(s);
  return rc;
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
  for (int d = 0; d < 2; d++) {
    if (m.peer_ipc[d]) cudaIpcCloseMemHandle(m.peer_arena[d]);
    m.peer_arena[d] = nullptr, m.peer_ipc[d] = false;
  }
  cudaFree(m.arena);
  m.arena = nullptr;
  cudaFree(m.cnt_dev);
  m.cnt_dev = nullptr;
  if (m.cnt_host) cudaFreeHost(m.cnt_host);
  m.cnt_host = nullptr;
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
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  m.mig_bytes = up(sizeof(int) * HDR_INTS + sizeof(double) * 7 * (size_t)mig_cap);
  m.halo_bytes = up(sizeof(int) * (HDR_INTS + halo_counts_ints(s->g.sc)) + sizeof(double) * 3 * (size_t)halo_cap);
  // 256 B tail holds a magic word the neighbours verify after mapping.  At least 8 MB and a multiple of 2 MB, so that
  // the arena is an allocation of its own: cudaIpcGetMemHandle exports the whole underlying allocation, and small
  // cudaMalloc blocks are carved out of shared 2 MB pages.
  m.arena_bytes = 4 * m.mig_bytes + 4 * m.halo_bytes + 256;
  const size_t alloc = ((m.arena_bytes > (8u << 20) ? m.arena_bytes : (8u << 20)) + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  ATLJ_CUDA(cudaMalloc((void **)&m.arena, alloc));
  ATLJ_CUDA(cudaMemsetAsync(m.arena, 0, alloc, s->st));
  {
    const unsigned long long magic = 0x61746c6a6d670001ull;  // "atljmg" 1
    ATLJ_CUDA(cudaMemcpyAsync(m.arena + m.arena_bytes - 256, &magic, sizeof(magic), cudaMemcpyHostToDevice, s->st));
    ATLJ_CUDA(cudaStreamSynchronize(s->st));
  }
  ATLJ_CUDA(cudaMalloc((void **)&m.cnt_dev, 16 * sizeof(int)));
  ATLJ_CUDA(cudaMemsetAsync(m.cnt_dev, 0, 16 * sizeof(int), s->st));
  ATLJ_CUDA(cudaHostAlloc((void **)&m.cnt_host, 16 * sizeof(int), cudaHostAllocDefault));
  m.step = 0;
  m.on = true;
  int rc = push_n(s);
  if (rc) return rc;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  return ATLJ_OK;
}

void atlj_sim_set_total(atlj_sim *s, int64_t n_total) {
  if (s) s->n_total = n_total;
}

/* receive arena of this rank: device pointer (same-process peers) and size */
void *atlj_sim_mg_arena(atlj_sim *s) { return (s && s->mg.on) ? s->mg.arena : nullptr; }
int64_t atlj_sim_mg_arena_bytes(atlj_sim *s) { return (s && s->mg.on) ? (int64_t)s->mg.arena_bytes : -1; }

/* peers living in the same process (several ranks emulated on one GPU): plain device pointers */
int atlj_sim_mg_set_peers(atlj_sim *s, void *left_arena, void *right_arena) {
  if (!s || !s->mg.on || !left_arena || !right_arena) return ATLJ_ERR_ARG;
  s->mg.peer_arena[0] = (unsigned char *)left_arena, s->mg.peer_arena[1] = (unsigned char *)right_arena;
  s->mg.peer_ipc[0] = s->mg.peer_ipc[1] = false;
  return ATLJ_OK;
}

/* peers in other processes (one process per GPU): CUDA IPC handle of the arena, 64 bytes */
int atlj_sim_mg_ipc_handle(atlj_sim *s, char *out64) {
  if (!s || !s->mg.on || !out64) return ATLJ_ERR_ARG;
  cudaIpcMemHandle_t h;
  ATLJ_CUDA(cudaIpcGetMemHandle(&h, s->mg.arena));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, 64);
  return ATLJ_OK;
}
int atlj_sim_mg_ipc_open(atlj_sim *s, const char *left64, const char *right64) {
  if (!s || !s->mg.on || !left64 || !right64) return ATLJ_ERR_ARG;
  MgState &m = s->mg;
  const bool same = memcmp(left64, right64, 64) == 0;  // two ranks: both neighbours are the one peer
  for (int d = 0; d < 2; d++) {
    if (d == 1 && same) {
      m.peer_arena[1] = m.peer_arena[0], m.peer_ipc[1] = false;
      break;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, d == 0 ? left64 : right64, 64);
    void *p = nullptr;
    ATLJ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    m.peer_arena[d] = (unsigned char *)p, m.peer_ipc[d] = true;
    unsigned long long magic = 0;  // the mapping must show the neighbour's arena, not some other part of its memory
    ATLJ_CUDA(cudaMemcpy(&magic, m.peer_arena[d] + m.arena_bytes - 256, sizeof(magic), cudaMemcpyDeviceToHost));
    if (magic != 0x61746c6a6d670001ull) {
      set_error("mg_ipc_open: the mapped peer buffer does not carry the arena signature (IPC mapping offset)");
      return ATLJ_ERR_CUDA;
    }
  }
  return ATLJ_OK;
}

int atlj_sim_mg_phase1(atlj_sim *s, double dt) {
  if (!s || !s->mg.on || !s->mg.peer_arena[0]) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, 1);
  if (rc || (rc = push_n(s))) return rc;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC, s->st));
  if ((rc = phase1(s, dt, false, nullptr, 0))) return rc;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  return ATLJ_OK;
}
int atlj_sim_mg_phase2(atlj_sim *s) {
  if (!s || !s->mg.on) return ATLJ_ERR_ARG;
  int rc = phase2(s, s->rec_dev);
  if (rc || (rc = pull_n(s))) return rc;
  return sim_check_flags(s);
}
int atlj_sim_mg_phase3(atlj_sim *s, double dt, double *rec_host) {
  if (!s || !s->mg.on) return ATLJ_ERR_ARG;
  const int nb = phase3(s);
  if (nb < 0) return nb;
  int rc = tail(s, dt, s->rec_dev, nb);
  if (rc) return rc;
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
  return ATLJ_OK;
}

int atlj_sim_run_nve_mg(atlj_sim *s, double dt, int nsteps, double *rec_host) {
  if (!s || !s->mg.on || !s->mg.comm || !s->mg.peer_arena[0] || !s->mg.peer_arena[1] || nsteps < 0) {
    set_error("run_nve_mg: enable the slab, map the neighbours' buffers and connect first");
    return ATLJ_ERR_ARG;
  }
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc) return rc;
  MgState &m = s->mg;
  if ((rc = push_n(s))) return rc;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC * (size_t)nsteps, s->st));
  int nb = 0;
  double *rec_prev = nullptr;
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    if ((rc = phase1(s, dt, k > 0, rec_prev, nb)) || (rc = phase2(s, rec))) return rc;
    if ((nb = phase3(s)) < 0) return nb;
    rec_prev = rec;
  }
  if (nsteps > 0 && (rc = tail(s, dt, rec_prev, nb))) return rc;
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
