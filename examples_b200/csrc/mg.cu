// Slab domain decomposition across the GPUs of one box: migration + halo exchange + all-reduce.
//
// The reference has no counterpart (its only distributed code is replica exchange, mc_nvt_lj_re.f90); the yardstick of
// this path is the single-domain result on the same configuration.  Each rank owns a contiguous range of x cell
// layers [x_lo, x_hi) of the reference's cell grid (md_lj_ll_module.py:106-109) and keeps one halo layer on each side.
// One velocity-Verlet step (md_nve_lj.py:178-192) is
//   phase 1  kick + drift + wrap on the owned atoms; atoms that left the slab are packed (r, v, id) STRAIGHT INTO THE
//            NEIGHBOUR'S receive buffer over NVLink (tens of atoms per face)
//   phase 2  ONE kernel waits for the neighbours' migration messages, appends the arrivals and assigns their cells;
//            scan; counting sort, whose gather kernel takes the first and last owned layer first and stores their
//            positions and their cell table a second time, STRAIGHT INTO THE NEIGHBOURS' POSITION ARRAYS AND CELL TABLES
//            (their halo regions: one layer per face, ~6e4 atoms = 1.4 MB at 2 M atoms/GPU).  No pack kernel, and
//            nothing is unpacked on the receiving side.
//   phase 3  the pair kernel starts at once on the rows that need no halo (interior x layers first) and only the CTAs
//            that reach a boundary layer acquire the neighbours' stamps: the exchange is hidden behind ~85 % of the
//            pair kernel.  Full stencil on the owned rows (no force return traffic).  The 12-double step records are
//            all-reduced ONCE per run (NCCL).
//
// Transport.  Peer memory is mapped once (CUDA IPC between the one-process-per-GPU ranks, plain pointers when several
// ranks are emulated in one process); kernels store over NVLink 5 / NVSwitch into it, the CTA that finishes last
// publishes a message with a system-scope release of a step stamp, consumers acquire that stamp.  No copy engine, no
// SM-hungry communication kernel, no host synchronisation per step.
// Flow control.  Migration buffers are double-buffered by step parity; the halo regions live in the double-buffered
// position arrays (r[0], r[1] alternate every step on every rank).  A rank cannot start step s+1's halo writes before
// it has consumed its neighbours' step s+1 migration messages, which they send only after their step-s pair kernel --
// the last reader of the step-s halo data and of the (single) halo rows of the cell table -- has finished.
// (An NCCL send/recv version of this exchange measured 35 us latency per exchange and could not overlap with the
// saturating pair kernel: profiles/r1_slab_history.md.)
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"
#include "integrate.cuh"
#include "sim.cuh"

namespace atlj {

constexpr int HDR_INTS = 16;  // 64-byte header of every message: [0] = number of atoms, [1] = step stamp (ready flag)

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// spins until the producer (another GPU) has published stamp in hdr[1]; false after timeout_ns (a neighbour that died
// must not hang this GPU): the caller raises err[5] and every later kernel of the run becomes a no-op
__device__ __forceinline__ bool acquire_stamp(const int *hdr, int stamp, unsigned long long timeout_ns) {
  const unsigned long long t0 = global_ns();
  for (;;) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(hdr + 1) : "memory");
    if (v == stamp) return true;
    if (global_ns() - t0 > timeout_ns) return false;
  }
}

__device__ __forceinline__ void publish(int *hdr, int count, int stamp) {
  hdr[0] = count;
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(hdr + 1), "r"(stamp) : "memory");
}

// Hessian on slabs (md_nve_lj.py:46 calls hessian(box, r_cut, r, f) every step): the pair terms of a boundary atom need
// the forces of its halo partners.  After the pair kernel the forces of the first / last owned layer are stored into the
// neighbours' force arrays at the halo offsets (where their positions went during the sort) and a second stamp
// (header [2]) is published; the hessian kernel acquires it when it reaches a boundary layer, like the force kernel
// acquires [1].  No double buffering: a neighbour can only send step s+1's forces after it has received this rank's
// step s+1 positions, which are sent after this rank's step-s hessian kernel - the last reader - in stream order.
__global__ void __launch_bounds__(256) halo_force_send_kernel(Geom g, const int *__restrict__ cell_start,
                                                              const double *__restrict__ f, double *dst_left,
                                                              double *dst_right, int cap, int *hdr_left,
                                                              int *hdr_right, const int *xstep, int *done,
                                                              const int *err) {
  if (err[5]) return;  // a neighbour's message timed out earlier in this run
  const int sc2 = g.sc * g.sc;
  const int *rowF = cell_start + (size_t)g.h * (sc2 + 1), *rowL = cell_start + (size_t)(g.h + g.nx_own - 1) * (sc2 + 1);
  const int F0 = rowF[0], F1 = min(rowF[sc2], F0 + cap), L0 = rowL[0], L1 = min(rowL[sc2], L0 + cap);
  const long long nF = 3ll * (F1 - F0), nL = 3ll * (L1 - L0);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nF + nL; t += (long long)gridDim.x * blockDim.x) {
    if (t < nF)
      dst_left[t] = f[3 * (size_t)F0 + t];
    else
      dst_right[t - nF] = f[3 * (size_t)L0 + (t - nF)];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      const int stamp = *xstep;
      __threadfence_system();
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(hdr_left + 2), "r"(stamp) : "memory");
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(hdr_right + 2), "r"(stamp) : "memory");
      *done = 0;
    }
  }
}

// The migration messages of a step: the integration kernel has stored the leavers (r, v, id) straight into the
// neighbours' receive buffers with plain stores and counted them; this one-warp kernel, ordered after it by the kernel
// boundary, writes the counts into the headers, fences ONCE at system scope and publishes the stamps.
__global__ void mig_publish_kernel(SlabArgs sl, const int *err) {
  if (threadIdx.x != 0 || err[5]) return;
  const int stamp = *sl.xstep;  // already moved on by the integration kernel: its value is this step's stamp
  const bool odd = (stamp - 1) & 1;
  int *hl = reinterpret_cast<int *>(odd ? sl.to_left1 : sl.to_left0);
  int *hr = reinterpret_cast<int *>(odd ? sl.to_right1 : sl.to_right0);
  hl[0] = min(sl.mig_cnt[0], sl.mig_cap), hr[0] = min(sl.mig_cnt[1], sl.mig_cap);
  sl.mig_cnt[0] = sl.mig_cnt[1] = sl.mig_cnt[2] = 0;
  __threadfence_system();
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(hl + 1), "r"(stamp) : "memory");
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(hr + 1), "r"(stamp) : "memory");
}

// consumer side for the pair kernels that cannot wait themselves (first tiled kernel): one thread acquires both stamps
__global__ void wait_stamp_kernel(const int *hdr_a, const int *hdr_b, const int *xstep, unsigned long long timeout_ns,
                                  int *err) {
  if (threadIdx.x != 0 || blockIdx.x != 0 || err[5]) return;
  const int stamp = *xstep;
  if (!acquire_stamp(hdr_a, stamp, timeout_ns) || !acquire_stamp(hdr_b, stamp, timeout_ns)) atomicOr(&err[5], 1);
}

// ---- optional trace (ATLJ_MG_TRACE=1): one-thread kernels between the launches of a step write %globaltimer into
// trace[step % TRACE_STEPS][slot]; tools/slab_trace.py prints where a slab step spends its time (waits included) even
// when the step is replayed as a graph.  slot 0 = before the integration kernel, 1 = after it, 2 = after the arrivals,
// 3 = after the scan, 4 = after the sort + gather (halo sent), 5 = after the pair kernel.
constexpr int TRACE_STEPS = 4096, TRACE_SLOTS = 8;
__global__ void trace_kernel(unsigned long long *trace, const int *xstep, int slot) {
  const int step = *xstep - (slot == 0 ? 0 : 1);
  trace[(size_t)(step & (TRACE_STEPS - 1)) * TRACE_SLOTS + slot] = global_ns();
}
static int trace_mark(atlj_sim *s, cudaStream_t st, int slot) {
  if (!s->mg.trace) return ATLJ_OK;
  trace_kernel<<<1, 1, 0, st>>>(s->mg.trace, s->mg.cnt_dev + 10, slot);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

// ---- phase 2: wait for the migration messages, append the arrivals behind the atoms already held, assign their cells --
// (one launch: every CTA acquires the two stamps itself, then takes its share of the arrivals)
constexpr int ARR_CTAS = 32, ARR_T = 256;
struct ArrivalBufs {
  const unsigned char *left0, *left1, *right0, *right1;  // my receive buffers, one per step parity
};
__global__ void __launch_bounds__(ARR_T) arrivals_kernel(Geom g, ArrivalBufs bufs, int cap, const int *__restrict__ xstep,
                                                         unsigned long long timeout_ns,
                                                         double *__restrict__ r, double *__restrict__ v,
                                                         int *__restrict__ id, const int *__restrict__ n_old_dev,
                                                         long long own_cap, int *__restrict__ n_after,
                                                         int *__restrict__ cellid, int *__restrict__ rank,
                                                         int *__restrict__ cell_count, int *__restrict__ err) {
  __shared__ int s_ok;
  // the integration kernel of this step has moved the exchange counter on: its value is this step's stamp
  const int stamp = *xstep;
  const bool odd = (stamp - 1) & 1;
  const unsigned char *__restrict__ recv_left = odd ? bufs.left1 : bufs.left0;
  const unsigned char *__restrict__ recv_right = odd ? bufs.right1 : bufs.right0;
  if (threadIdx.x == 0) {
    s_ok = err[5] == 0 && acquire_stamp(reinterpret_cast<const int *>(recv_left), stamp, timeout_ns) &&
           acquire_stamp(reinterpret_cast<const int *>(recv_right), stamp, timeout_ns);
    if (!s_ok) atomicOr(&err[5], 1);
  }
  __syncthreads();
  if (!s_ok) return;
  const int n_old = *n_old_dev;
  const int cl = min(*reinterpret_cast<const int *>(recv_left), cap);
  const int cr = min(*reinterpret_cast<const int *>(recv_right), cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *n_after = n_old + cl + cr;
    if ((long long)n_old + cl + cr > own_cap) atomicOr(&err[1], 1);
  }
  const double scd = (double)g.sc;
  for (int k = blockIdx.x * ARR_T + threadIdx.x; k < cl + cr; k += ARR_CTAS * ARR_T) {
    const bool right = k >= cl;
    const long long dst = (long long)n_old + k;
    if (dst >= own_cap) break;
    const double *s = reinterpret_cast<const double *>((right ? recv_right : recv_left) + sizeof(int) * HDR_INTS) +
                      7 * (size_t)(right ? k - cl : k);
    const double x = s[0], y = s[1], z = s[2];  // wrapped by the sender
#pragma unroll
    for (int c = 0; c < 3; c++) r[3 * dst + c] = s[c], v[3 * dst + c] = s[3 + c];
    id[dst] = (int)s[6];
    // md_lj_ll_module.py:108, same separately rounded operations as cell_assign_kernel
    const long long c0 = (long long)floor(__dmul_rn(__dadd_rn(x, 0.5), scd));
    const long long c1 = (long long)floor(__dmul_rn(__dadd_rn(y, 0.5), scd));
    const long long c2 = (long long)floor(__dmul_rn(__dadd_rn(z, 0.5), scd));
    const int lx = (int)c0 - g.x_lo + g.h;
    const bool bad = c0 < 0 || c0 >= g.sc || c1 < 0 || c1 >= g.sc || c2 < 0 || c2 >= g.sc || lx < g.h ||
                     lx >= g.h + g.nx_own;  // an arrival always belongs to an owned layer
    if (bad) {
      atomicOr(&err[0], 1);
      cellid[dst] = -1, rank[dst] = 0;
    } else {
      const int cid = (lx * g.sc + (int)c1) * g.sc + (int)c2;
      cellid[dst] = cid;
      rank[dst] = atomicAdd(&cell_count[cid], 1);
    }
  }
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

// ---- receive arena: migration [from: left, right][parity], then the two halo headers [from: left, right] ----------------
static inline unsigned char *mig_recv(unsigned char *arena, const MgState &m, int from, int parity) {
  return arena + (size_t)(2 * from + parity) * m.mig_bytes;
}
static inline int *halo_hdr(unsigned char *arena, const MgState &m, int from) {
  return reinterpret_cast<int *>(arena + 4 * m.mig_bytes + (size_t)from * sizeof(int) * HDR_INTS);
}
// where my message towards `dir` (0 = left, 1 = right) lands: the neighbour's "from the other side" buffer

// device-side row count of the tiled LJ kernel's partials (null for the other pair kernels)
static const int *tile2_rows(atlj_sim *s) {
  return sim_tile2(s) ? s->cb.err + 6 : nullptr;
}

// ---- the phases -----------------------------------------------------------------------------------------------------------
// The number of owned atoms changes every step and lives on the device (cnt_dev[1]); kernels are launched over the
// capacity and read it, so a step needs no host synchronisation.  The host copy s->n is refreshed at the end of a run.

// phase 1: ONE kernel: [trailing half kick + record of the finished step] + leading half kick + drift + wrap + cell
// assignment of the atoms that stay + leavers packed into the neighbours' buffers and published
// st: the stream to launch on (the sim's, or the capture stream); replay: nothing may depend on the step number - the
// record pointer follows the device counter cnt_dev[11] (rec = base of the records), as the stamps and buffer parities
// always follow cnt_dev[10]
static int phase1(atlj_sim *s, cudaStream_t st, double dt, bool first_kick, double *rec, int nb_prev, bool replay,
                  bool counts_zeroed, bool assign_only = false) {
  MgState &m = s->mg;
  SlabArgs sl;
  sl.assign_only = assign_only ? 1 : 0;
  sl.n_dev = m.cnt_dev + 1, sl.id = s->id[s->cur];
  sl.to_left0 = mig_recv(m.peer_arena[0], m, 1, 0), sl.to_left1 = mig_recv(m.peer_arena[0], m, 1, 1);
  sl.to_right0 = mig_recv(m.peer_arena[1], m, 0, 0), sl.to_right1 = mig_recv(m.peer_arena[1], m, 0, 1);
  sl.mig_cap = (int)m.mig_cap, sl.xstep = m.cnt_dev + 10, sl.mig_cnt = m.cnt_dev + 4;
  const size_t sc2p = (size_t)s->g.sc * s->g.sc + 1;
  sl.first_end = m.sorted ? s->cb.cell_start + (size_t)s->g.h * sc2p + (sc2p - 1) : nullptr;
  sl.last_begin = m.sorted ? s->cb.cell_start + (size_t)(s->g.h + s->g.nx_own - 1) * sc2p : nullptr;
  KdkLoop loop;
  loop.step_ctr = replay ? m.cnt_dev + 11 : nullptr;
  loop.tab0 = sim_tile2(s) ? s->tile_tab : nullptr;
  loop.counts_zeroed = counts_zeroed;
  loop.tail_in_scan = first_kick;  // phase 2's scan launch finishes the record of the step that just ended
  trace_mark(s, st, 0);
  Timer::begin(s, 2);
  const int nb = launch_kdk_assign(st, s->g, (int)m.own_cap, s->r[s->cur], s->v[s->cur], s->f, 0.5 * dt, dt, s->p.box,
                                   first_kick, s->cb, s->vf_partials, rec, s->partials, nb_prev, s->p.model, &sl,
                                   tile2_rows(s), &loop);
  m.tail = RecTail();
  if (nb >= 0) {
    mig_publish_kernel<<<1, 32, 0, st>>>(sl, s->cb.err);
    ATLJ_LAUNCHED();
    if (first_kick) {
      m.tail.rows = s->vf_partials, m.tail.nrows = nb, m.tail.rec = rec, m.tail.step_ctr = loop.step_ctr;
      m.tail.model = s->p.model, m.tail.has_pair = 1, m.tail.err = s->cb.err;
    }
  }
  Timer::end(s);
  trace_mark(s, st, 1);
  return nb < 0 ? nb : ATLJ_OK;
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

// phase 2: arrivals appended and assigned, scan, counting sort (its gather kernel sends the boundary layers to the
// neighbours' halo regions); the item table of the pair kernel runs beside the sort when a side stream is given
static int phase2(atlj_sim *s, cudaStream_t st, cudaStream_t side, double *rec, bool replay) {
  int rc;
  MgState &m = s->mg;
  const int cap = (int)m.mig_cap, nup = (int)m.own_cap;
  const int cur = s->cur, alt = 1 - cur;
  ArrivalBufs ab;
  ab.left0 = mig_recv(m.arena, m, 0, 0), ab.left1 = mig_recv(m.arena, m, 0, 1);
  ab.right0 = mig_recv(m.arena, m, 1, 0), ab.right1 = mig_recv(m.arena, m, 1, 1);
  Timer::begin(s, 4);
  arrivals_kernel<<<ARR_CTAS, ARR_T, 0, st>>>(s->g, ab, cap, m.cnt_dev + 10, m.timeout_ns, s->r[cur], s->v[cur],
                                               s->id[cur], m.cnt_dev + 1, (long long)m.own_cap, m.cnt_dev, s->cb.cellid,
                                               s->cb.rank, s->cb.cell_count, s->cb.err);
  ATLJ_LAUNCHED();
  Timer::end(s);
  trace_mark(s, st, 2);
  // my first layer is the left neighbour's RIGHT halo (its local layer nxl - 1, behind its left halo region); my last
  // layer is the right neighbour's LEFT halo (its local layer 0).  All ranks share own_cap / halo_cap and flip their
  // position buffers in step, so the neighbour's sorted buffer of this step is its r[alt] as well.  The sort + gather
  // kernel stores those two layers a second time, straight into the neighbours' memory, and publishes the stamps.
  const size_t sc2p = (size_t)s->g.sc * s->g.sc + 1;
  HaloOut ho;
  memset(&ho, 0, sizeof(ho));
  ho.base_left = (int)(m.own_cap + m.halo_cap), ho.base_right = (int)m.own_cap;
  ho.dst_r_left = m.peer_r[0][alt] + 3 * (size_t)ho.base_left, ho.dst_r_right = m.peer_r[1][alt] + 3 * (size_t)ho.base_right;
  ho.dst_cs_left = m.peer_cs[0] + (size_t)(m.peer_nxl[0] - 1) * sc2p, ho.dst_cs_right = m.peer_cs[1];
  ho.hdr_left = halo_hdr(m.peer_arena[0], m, 1), ho.hdr_right = halo_hdr(m.peer_arena[1], m, 0);
  ho.cap = (int)m.halo_cap, ho.xstep = m.cnt_dev + 10, ho.rec_ctr = replay ? m.cnt_dev + 11 : nullptr;
  ho.done = m.cnt_dev + 8, ho.n_own = m.cnt_dev + 1;
  ho.own_cap = (long long)m.own_cap, ho.rec = rec, ho.err = s->cb.err;
  Timer::begin(s, 1);
  if ((rc = launch_cell_scan(st, s->g, s->cb, 0, m.tail.nrows > 0 ? &m.tail : nullptr))) return rc;
  m.tail = RecTail();
  trace_mark(s, st, 3);
  const bool fork = side != nullptr && sim_tile2(s);
  if (fork) {  // the item table needs cell_start only: beside the sort + gather
    PairArgs a = pair_args(s);
    ATLJ_CUDA(cudaEventRecord(s->ev_fork, st));
    ATLJ_CUDA(cudaStreamWaitEvent(side, s->ev_fork, 0));
    if ((rc = launch_pair_tile2(side, a, s->cb.err + 2, 0, 1, true)) < 0) return rc;
    ATLJ_CUDA(cudaEventRecord(s->ev_join, side));
  }
  // (s->n: the owned count the host saw last, at the end of the previous run; migration changes it by ~1e-4 per step)
  if ((rc = launch_cell_sort_gather(st, s->g, s->cb, nup, 0, s->id[cur], s->r[cur], s->v[cur], s->r[alt], s->v[alt],
                                    s->id[alt], m.cnt_dev, &ho, true, (int)s->n)))
    return rc;
  Timer::end(s);
  trace_mark(s, st, 4);
  m.sorted = true;
  s->cur = alt;
  if (fork) ATLJ_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
  return ATLJ_OK;
}

// phase 3: pair kernel over the owned rows; the halo layers 0 and nxl-1 were written by the neighbours, and the kernel
// itself acquires their stamps when it reaches a boundary layer.  -> rows of partials written (the record is finished by
// the next phase 1 or by tail())
static int phase3(atlj_sim *s, cudaStream_t st, bool table_done, bool save_masks = false) {
  MgState &m = s->mg;
  PairArgs a = pair_args(s);
  if (save_masks) a.masks = s->masks;
  const int *hl = halo_hdr(m.arena, m, 0), *hr = halo_hdr(m.arena, m, 1);
  int nb;
  if (sim_tile2(s)) {
    a.halo.stamp_left = hl, a.halo.stamp_right = hr, a.halo.xstep = m.cnt_dev + 10, a.halo.timeout_ns = m.timeout_ns;
    a.halo.err = s->cb.err;
    Timer::begin(s, 0);
    nb = launch_pair_tile2(st, a, s->cb.err + 2, save_masks ? 1 : 0, table_done ? 2 : 0, true);
    Timer::end(s);
  } else {
    Timer::begin(s, 6);
    wait_stamp_kernel<<<1, 32, 0, st>>>(hl, hr, m.cnt_dev + 10, m.timeout_ns, s->cb.err);
    ATLJ_LAUNCHED();
    Timer::end(s);
    Timer::begin(s, 0);
    nb = launch_pair_tile(st, a, false);
    Timer::end(s);
  }
  trace_mark(s, st, 5);
  m.step++;
  return nb;
}

// the record of the last step of a run: pair totals, trailing half kick, sum v^2, sum f^2
static int tail(atlj_sim *s, double dt, double *rec, int nb) {
  int rc;
  if ((rc = launch_finalize(s->st, s->partials, nb, s->p.model, false, rec, s->cb.err + 2, nullptr, 0, tile2_rows(s))))
    return rc;
  Timer::begin(s, 2);
  rc = launch_kick_sums(s->st, 3ll * s->mg.own_cap, s->v[s->cur], s->f, 0.5 * dt, s->vf_partials, rec + 4, s->cb.err + 4,
                        s->mg.cnt_dev + 1);
  Timer::end(s);
  return rc;
}

// hessian of the step (md_lj_ll_module.py:227-357) after a phase 3 that kept its mask words and whose totals have been
// finalized (that launch also reset the work counter): boundary forces to the neighbours, then the hessian kernel walks
// the mask words; its boundary items acquire the neighbours' force stamps.  rec[6] receives this rank's share.
static int force_send(atlj_sim *s, cudaStream_t st) {
  MgState &m = s->mg;
  Timer::begin(s, 5);
  halo_force_send_kernel<<<148, 256, 0, st>>>(s->g, s->cb.cell_start, s->f,
                                             m.peer_f[0] + 3 * (size_t)(m.own_cap + m.halo_cap),
                                             m.peer_f[1] + 3 * (size_t)m.own_cap, (int)m.halo_cap,
                                             halo_hdr(m.peer_arena[0], m, 1), halo_hdr(m.peer_arena[1], m, 0),
                                             m.cnt_dev + 10, m.cnt_dev + 9, s->cb.err);
  ATLJ_LAUNCHED();
  Timer::end(s);
  return ATLJ_OK;
}
static int hess_launch(atlj_sim *s, cudaStream_t st, double *rec) {
  MgState &m = s->mg;
  int rc;
  PairArgs a = pair_args(s);
  a.masks = s->masks, a.fin = s->f, a.f = nullptr;
  // (the kernel reads stamp + 1: header word [2] is the force stamp)
  a.halo.stamp_left = halo_hdr(m.arena, m, 0) + 1, a.halo.stamp_right = halo_hdr(m.arena, m, 1) + 1;
  a.halo.xstep = m.cnt_dev + 10, a.halo.timeout_ns = m.timeout_ns, a.halo.err = s->cb.err;
  Timer::begin(s, 3);
  const int nb = launch_pair_tile2(st, a, s->cb.err + 2, 2, 0, true);
  Timer::end(s);
  if (nb < 0) return nb;
  if ((rc = launch_finalize(st, s->partials, nb, s->p.model, true, rec, s->cb.err + 2, nullptr, 0, tile2_rows(s)))) return rc;
  return ATLJ_OK;
}

// host copy of the owned count <-> device
static int push_n(atlj_sim *s);
static int pull_n(atlj_sim *s);

// ---- building blocks for the other integrators on slabs (dynamics.cu): the integrator's own kernels move the owned
// atoms (count on the device: mg_n_dev), then mg_assign_exchange bins them, migrates the leavers, sorts, sends the halo
// layers and runs the pair kernel; the totals of the step go to rec (LOCAL sums) ------------------------------------
const int *mg_n_dev(const atlj_sim *s) { return s->mg.cnt_dev + 1; }
long long mg_own_cap(const atlj_sim *s) { return s->mg.own_cap; }
int mg_begin(atlj_sim *s) { return push_n(s); }
int mg_end(atlj_sim *s) { return pull_n(s); }
int mg_assign_exchange_force(atlj_sim *s, double *rec, bool first) {
  int rc, nb;
  if ((rc = phase1(s, s->st, 0.0, false, nullptr, 0, false, !first, true)) || (rc = phase2(s, s->st, nullptr, rec, false)))
    return rc;
  if ((nb = phase3(s, s->st, false)) < 0) return nb;
  s->fin = FinTail();
  if (s->defer_fin) {  // the caller's half kick finishes the totals in an extra CTA (FinTail, pair_math.cuh)
    s->defer_fin = false;
    s->fin.partials = s->partials, s->fin.nb = nb, s->fin.model = s->p.model, s->fin.rec = rec;
    s->fin.reset_counter = s->cb.err + 2, s->fin.rows_dev = tile2_rows(s);
    return ATLJ_OK;
  }
  return launch_finalize(s->st, s->partials, nb, s->p.model, false, rec, s->cb.err + 2, nullptr, 0, tile2_rows(s));
}
// emulation (ranks of one process in lockstep): phase 1 of a step whose atoms were moved by another integrator
int mg_phase1_assign(atlj_sim *s, bool first) {
  if (!s->mg.peer_arena[0]) return ATLJ_ERR_ARG;
  int rc = phase1(s, s->st, 0.0, false, nullptr, 0, false, !first, true);
  if (rc) return rc;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  return ATLJ_OK;
}
// sum over the ranks, in place, on the sim's stream (NCCL; every rank ends up with the same bits)
int mg_allreduce(atlj_sim *s, double *dev, size_t count) {
  if (!s->mg.comm) {
    set_error("slab run: connect the ranks first (atlj_sim_mg_connect)");
    return ATLJ_ERR_ARG;
  }
  ATLJ_NCCL(g_nccl.AllReduce(dev, dev, count, ncclDouble, ncclSum, (ncclComm_t)s->mg.comm, s->st));
  return ATLJ_OK;
}
bool mg_connected(const atlj_sim *s) { return s->mg.comm != nullptr; }

static int push_n(atlj_sim *s) {
  s->mg.cnt_host[0] = (int)s->n;
  s->mg.cnt_host[1] = (int)s->mg.step;  // exchange step counter (stamps, buffer parity)
  s->mg.cnt_host[2] = 0;                // record row of the replayed launches
  ATLJ_CUDA(cudaMemcpyAsync(s->mg.cnt_dev + 1, s->mg.cnt_host, sizeof(int), cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(s->mg.cnt_dev + 10, s->mg.cnt_host + 1, 2 * sizeof(int), cudaMemcpyHostToDevice, s->st));
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
  for (int k = 0; k < m.n_ipc_mapped; k++) cudaIpcCloseMemHandle(m.ipc_mapped[k]);
  m.n_ipc_mapped = 0;
  for (int d = 0; d < 2; d++)
    m.peer_arena[d] = nullptr, m.peer_r[d][0] = m.peer_r[d][1] = nullptr, m.peer_cs[d] = nullptr, m.peer_f[d] = nullptr;
  cudaFree(m.arena);
  m.arena = nullptr;
  cudaFree(m.cnt_dev);
  m.cnt_dev = nullptr;
  cudaFree(m.trace);
  m.trace = nullptr;
  if (m.cnt_host) cudaFreeHost(m.cnt_host);
  m.cnt_host = nullptr;
  if (m.comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)m.comm);
  m.comm = nullptr;
  m.on = false;
}
void sim_mg_destroy(atlj_sim *s) { mg_free(s); }

static const unsigned long long MG_MAGIC = 0x61746c6a6d670002ull;  // the signature of api.cu's dev_alloc_ipc

// everything the captured launches of the slab step bake in
struct MgGraphKey {
  int64_t n_total;
  int cur, sc, x_lo, nx_own, nb, pad;
  double dt, box, r_cut_box_sq, pot_cut;
  void *r0, *rec, *partials, *peer, *tab;
};

// what a neighbour needs to write into this rank's memory (exchanged once, by any means)
struct MgBlob {
  cudaIpcMemHandle_t arena, r0, r1, cs, f;
  unsigned long long arena_bytes, r_bytes, cs_bytes, f_bytes;
  long long own_cap, halo_cap, mig_cap;
  int nxl, sc;
};

}  // namespace atlj

using namespace atlj;

extern "C" {

int atlj_sim_mg_enable(atlj_sim *s, int64_t mig_cap, int64_t halo_cap) {
  if (!s || !s->configured || mig_cap < 1 || halo_cap < 1 || s->g.h != 1) {
    set_error("mg_enable: configure the sim with a proper sub-range of x layers first");
    return ATLJ_ERR_ARG;
  }
  halo_cap += halo_cap & 1;  // even: the neighbours' 16-byte stores into the halo regions stay aligned
  int64_t own_cap = s->cap - 2 * halo_cap;
  own_cap -= own_cap & 1;
  if (own_cap < s->n || own_cap < 2) {
    set_error("mg_enable: capacity %lld cannot hold %lld owned atoms and two halo layers of %lld", (long long)s->cap,
              (long long)s->n, (long long)halo_cap);
    return ATLJ_ERR_ARG;
  }
  mg_free(s);
  MgState &m = s->mg;
  m.mig_cap = mig_cap, m.halo_cap = halo_cap, m.own_cap = own_cap;
  const char *to = getenv("ATLJ_PEER_TIMEOUT_MS");
  if (to && atof(to) > 0) m.timeout_ns = (unsigned long long)(atof(to) * 1e6);
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  m.mig_bytes = up(sizeof(int) * HDR_INTS + sizeof(double) * 7 * (size_t)mig_cap);
  // 256 B tail holds a signature the neighbours verify after mapping.  At least 8 MB and a multiple of 2 MB, so that
  // the arena is an allocation of its own: cudaIpcGetMemHandle exports the whole underlying allocation, and small
  // cudaMalloc blocks are carved out of shared 2 MB pages.
  m.arena_bytes = 4 * m.mig_bytes + 2 * sizeof(int) * HDR_INTS + 256;
  m.arena_bytes = ((m.arena_bytes > (8u << 20) ? m.arena_bytes : (8u << 20)) + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  ATLJ_CUDA(cudaMalloc((void **)&m.arena, m.arena_bytes));
  ATLJ_CUDA(cudaMemsetAsync(m.arena, 0, m.arena_bytes, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(m.arena + m.arena_bytes - 256, &MG_MAGIC, sizeof(MG_MAGIC), cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  ATLJ_CUDA(cudaMalloc((void **)&m.cnt_dev, 16 * sizeof(int)));
  ATLJ_CUDA(cudaMemsetAsync(m.cnt_dev, 0, 16 * sizeof(int), s->st));
  ATLJ_CUDA(cudaHostAlloc((void **)&m.cnt_host, 16 * sizeof(int), cudaHostAllocDefault));
  const char *tr = getenv("ATLJ_MG_TRACE");
  if (tr && tr[0] == '1') {
    ATLJ_CUDA(cudaMalloc((void **)&m.trace, sizeof(unsigned long long) * TRACE_STEPS * TRACE_SLOTS));
    ATLJ_CUDA(cudaMemsetAsync(m.trace, 0, sizeof(unsigned long long) * TRACE_STEPS * TRACE_SLOTS, s->st));
  }
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

/* peers living in the same process (several ranks emulated on one GPU): plain device pointers */
int atlj_sim_mg_set_peers(atlj_sim *s, atlj_sim *left, atlj_sim *right) {
  if (!s || !s->mg.on || !left || !right || !left->mg.on || !right->mg.on) return ATLJ_ERR_ARG;
  MgState &m = s->mg;
  atlj_sim *peer[2] = {left, right};
  for (int d = 0; d < 2; d++) {
    const MgState &pm = peer[d]->mg;
    if (pm.own_cap != m.own_cap || pm.halo_cap != m.halo_cap || pm.mig_cap != m.mig_cap || peer[d]->g.sc != s->g.sc) {
      set_error("mg_set_peers: all ranks must use the same capacities and cell grid");
      return ATLJ_ERR_ARG;
    }
    m.peer_arena[d] = pm.arena, m.peer_r[d][0] = peer[d]->r[0], m.peer_r[d][1] = peer[d]->r[1];
    m.peer_cs[d] = peer[d]->cb.cell_start, m.peer_nxl[d] = peer[d]->g.nxl, m.peer_f[d] = peer[d]->f;
  }
  return ATLJ_OK;
}

/* peers in other processes (one process per GPU): the blob names this rank's receive arena, position buffers and
 * cell table (CUDA IPC handles) and its layout */
int64_t atlj_sim_mg_blob_bytes(void) { return (int64_t)sizeof(MgBlob); }

int atlj_sim_mg_ipc_blob(atlj_sim *s, char *out) {
  if (!s || !s->mg.on || !out) return ATLJ_ERR_ARG;
  if (!s->ipc_safe) {
    set_error("mg_ipc_blob: create the sim with ATLJ_SIM_IPC so that its buffers can be exported");
    return ATLJ_ERR_ARG;
  }
  MgBlob b;
  memset(&b, 0, sizeof(b));
  ATLJ_CUDA(cudaIpcGetMemHandle(&b.arena, s->mg.arena));
  ATLJ_CUDA(cudaIpcGetMemHandle(&b.r0, s->r[0]));
  ATLJ_CUDA(cudaIpcGetMemHandle(&b.r1, s->r[1]));
  ATLJ_CUDA(cudaIpcGetMemHandle(&b.cs, s->cb.cell_start));
  ATLJ_CUDA(cudaIpcGetMemHandle(&b.f, s->f));
  b.arena_bytes = s->mg.arena_bytes, b.r_bytes = s->r_alloc_bytes, b.cs_bytes = s->cs_alloc_bytes;
  b.f_bytes = s->f_alloc_bytes;
  b.own_cap = s->mg.own_cap, b.halo_cap = s->mg.halo_cap, b.mig_cap = s->mg.mig_cap, b.nxl = s->g.nxl, b.sc = s->g.sc;
  memcpy(out, &b, sizeof(b));
  return ATLJ_OK;
}

int atlj_sim_mg_ipc_open(atlj_sim *s, const char *left_blob, const char *right_blob) {
  if (!s || !s->mg.on || !left_blob || !right_blob) return ATLJ_ERR_ARG;
  MgState &m = s->mg;
  const bool same = memcmp(left_blob, right_blob, sizeof(MgBlob)) == 0;  // two ranks: both neighbours are the one peer
  for (int d = 0; d < 2; d++) {
    if (d == 1 && same) {
      m.peer_arena[1] = m.peer_arena[0], m.peer_r[1][0] = m.peer_r[0][0], m.peer_r[1][1] = m.peer_r[0][1];
      m.peer_cs[1] = m.peer_cs[0], m.peer_nxl[1] = m.peer_nxl[0], m.peer_f[1] = m.peer_f[0];
      break;
    }
    MgBlob b;
    memcpy(&b, d == 0 ? left_blob : right_blob, sizeof(b));
    if (b.own_cap != m.own_cap || b.halo_cap != m.halo_cap || b.mig_cap != m.mig_cap || b.sc != s->g.sc) {
      set_error("mg_ipc_open: all ranks must use the same capacities and cell grid");
      return ATLJ_ERR_ARG;
    }
    struct {
      cudaIpcMemHandle_t *h;
      unsigned long long bytes;
      void **out;
    } maps[5] = {{&b.arena, b.arena_bytes, (void **)&m.peer_arena[d]},
                 {&b.r0, b.r_bytes, (void **)&m.peer_r[d][0]},
                 {&b.r1, b.r_bytes, (void **)&m.peer_r[d][1]},
                 {&b.cs, b.cs_bytes, (void **)&m.peer_cs[d]},
                 {&b.f, b.f_bytes, (void **)&m.peer_f[d]}};
    for (auto &mp : maps) {
      void *p = nullptr;
      ATLJ_CUDA(cudaIpcOpenMemHandle(&p, *mp.h, cudaIpcMemLazyEnablePeerAccess));
      m.ipc_mapped[m.n_ipc_mapped++] = p;
      unsigned long long magic = 0;  // the mapping must show the neighbour's buffer, not some other part of its memory
      ATLJ_CUDA(cudaMemcpy(&magic, (unsigned char *)p + mp.bytes - 256, sizeof(magic), cudaMemcpyDeviceToHost));
      if (magic != MG_MAGIC) {
        set_error("mg_ipc_open: a mapped peer buffer does not carry its signature (IPC mapping offset)");
        return ATLJ_ERR_CUDA;
      }
      *mp.out = p;
    }
    m.peer_nxl[d] = b.nxl;
  }
  return ATLJ_OK;
}

static bool mg_ready(const atlj_sim *s) {
  return s && s->mg.on && s->mg.peer_arena[0] && s->mg.peer_arena[1] && s->mg.peer_r[0][0] && s->mg.peer_r[1][0] &&
         s->mg.peer_cs[0] && s->mg.peer_cs[1] && s->mg.peer_f[0] && s->mg.peer_f[1];
}

int atlj_sim_mg_phase1(atlj_sim *s, double dt) {
  if (!mg_ready(s)) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, 1);
  if (rc || (rc = push_n(s))) return rc;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC, s->st));
  if ((rc = phase1(s, s->st, dt, false, nullptr, 0, false, false))) return rc;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  return ATLJ_OK;
}
int atlj_sim_mg_phase2(atlj_sim *s) {
  if (!mg_ready(s)) return ATLJ_ERR_ARG;
  int rc = phase2(s, s->st, nullptr, s->rec_dev, false);
  if (rc || (rc = pull_n(s))) return rc;
  return sim_check_flags(s);
}
int atlj_sim_mg_phase3(atlj_sim *s, double dt, double *rec_host) {
  if (!mg_ready(s)) return ATLJ_ERR_ARG;
  const int nb = phase3(s, s->st, false);
  if (nb < 0) return nb;
  int rc = tail(s, dt, s->rec_dev, nb);
  if (rc) return rc;
  return sim_copy_rec(s, rec_host, 1);
}

/* hessian on emulated ranks: phase 3 split in two so that every rank has sent its boundary forces before any rank's
 * hessian kernel waits for them */
int atlj_sim_mg_phase3_force(atlj_sim *s) {
  if (!mg_ready(s) || !sim_tile2(s)) return ATLJ_ERR_ARG;
  int rc = sim_ensure_masks(s);
  if (rc) return rc;
  const int nb = phase3(s, s->st, false, true);
  if (nb < 0) return nb;
  if ((rc = launch_finalize(s->st, s->partials, nb, s->p.model, false, s->rec_dev, s->cb.err + 2, nullptr, 0, tile2_rows(s))) ||
      (rc = force_send(s, s->st)))
    return rc;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  return ATLJ_OK;
}
int atlj_sim_mg_phase4_hess(atlj_sim *s, double dt, double *rec_host) {
  if (!mg_ready(s) || !sim_tile2(s)) return ATLJ_ERR_ARG;
  int rc = hess_launch(s, s->st, s->rec_dev);
  if (rc) return rc;
  Timer::begin(s, 2);
  rc = launch_kick_sums(s->st, 3ll * s->mg.own_cap, s->v[s->cur], s->f, 0.5 * dt, s->vf_partials, s->rec_dev + 4,
                        s->cb.err + 4, s->mg.cnt_dev + 1);
  Timer::end(s);
  if (rc) return rc;
  return sim_copy_rec(s, rec_host, 1);
}

/* ATLJ_MG_TRACE=1: the %globaltimer marks of the last `nsteps` exchange steps, out[nsteps][8] (ns); see trace_kernel */
int atlj_sim_mg_trace(atlj_sim *s, int nsteps, unsigned long long *out) {
  if (!s || !s->mg.on || !s->mg.trace || !out || nsteps < 1 || nsteps > TRACE_STEPS) return ATLJ_ERR_ARG;
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  std::vector<unsigned long long> all((size_t)TRACE_STEPS * TRACE_SLOTS);
  ATLJ_CUDA(cudaMemcpy(all.data(), s->mg.trace, sizeof(unsigned long long) * all.size(), cudaMemcpyDeviceToHost));
  for (int k = 0; k < nsteps; k++) {
    const long long step = s->mg.step - nsteps + k;
    if (step < 0) return ATLJ_ERR_ARG;
    memcpy(out + (size_t)k * TRACE_SLOTS, all.data() + (size_t)(step & (TRACE_STEPS - 1)) * TRACE_SLOTS,
           sizeof(unsigned long long) * TRACE_SLOTS);
  }
  return ATLJ_OK;
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
  if (!mg_ready(s) || !s->mg.comm || nsteps < 0) {
    set_error("run_nve_mg: enable the slab, map the neighbours' buffers and connect first");
    return ATLJ_ERR_ARG;
  }
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc) return rc;
  MgState &m = s->mg;
  if ((rc = push_n(s))) return rc;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC * (size_t)nsteps, s->st));
  int nb = 0, k = 0;
  double *rec_prev = nullptr;
  // step 0 launched kernel by kernel (nothing to close, cell_count zeroed by a memset, kernel attributes warmed up)
  if (nsteps > 0) {
    if ((rc = phase1(s, s->st, dt, false, nullptr, 0, false, false)) || (rc = phase2(s, s->st, nullptr, s->rec_dev, false)))
      return rc;
    if ((nb = phase3(s, s->st, false)) < 0) return nb;
    rec_prev = s->rec_dev;
    k = 1;
  }
  // pairs of steps as ONE replayed CUDA graph: stamps, buffer parities and record rows follow device counters
  if (!s->timing && !s->no_graph && nsteps - k >= 2) {
    MgGraphKey key;
    memset(&key, 0, sizeof(key));
    key.n_total = s->n_total, key.cur = s->cur, key.sc = s->g.sc, key.x_lo = s->g.x_lo, key.nx_own = s->g.nx_own, key.nb = nb;
    key.dt = dt, key.box = s->p.box, key.r_cut_box_sq = s->p.r_cut_box_sq, key.pot_cut = s->p.pot_cut;
    key.r0 = s->r[0], key.rec = s->rec_dev, key.partials = s->partials, key.peer = m.peer_arena[0], key.tab = s->tile_tab;
    static_assert(sizeof(MgGraphKey) <= sizeof(s->nve_graph_key), "key storage");
    bool have = s->nve_graph && memcmp(&key, s->nve_graph_key, sizeof(key)) == 0;
    if (!have) {
      if (s->nve_graph) cudaGraphExecDestroy(s->nve_graph), s->nve_graph = nullptr;
      if (!s->st_cap) {
        ATLJ_CUDA(cudaStreamCreateWithFlags(&s->st_cap, cudaStreamNonBlocking));
        ATLJ_CUDA(cudaStreamCreateWithFlags(&s->st_side, cudaStreamNonBlocking));
        ATLJ_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        ATLJ_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
      }
      const int64_t l0 = g_launches;
      const int cur0 = s->cur;
      const long long step0 = m.step;
      ATLJ_CUDA(cudaStreamBeginCapture(s->st_cap, cudaStreamCaptureModeThreadLocal));
      g_pdl_off = s->n_total > 44000;  // (common.cuh: programmatic launches inside a graph)
      int nbc = nb;
      for (int j = 0; j < 2 && rc == ATLJ_OK; j++) {
        if ((rc = phase1(s, s->st_cap, dt, true, s->rec_dev, nbc, true, true)) ||
            (rc = phase2(s, s->st_cap, s->st_side, s->rec_dev, true)))
          break;
        if ((nbc = phase3(s, s->st_cap, sim_tile2(s))) < 0) rc = nbc;
      }
      g_pdl_off = 0;
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(s->st_cap, &graph);
      s->nve_graph_launches = (int)(g_launches - l0);
      g_launches = l0, m.step = step0;
      if (rc == ATLJ_OK && (e != cudaSuccess || s->cur != cur0 || nbc != nb)) {
        set_error("CUDA graph capture of the slab step failed: %s", cudaGetErrorString(e));
        rc = ATLJ_ERR_CUDA;
      }
      s->cur = cur0;
      if (rc == ATLJ_OK && cudaGraphInstantiate(&s->nve_graph, graph, 0) != cudaSuccess) {
        s->nve_graph = nullptr;
        rc = ATLJ_ERR_CUDA;
      }
      if (graph) cudaGraphDestroy(graph);
      if (rc == ATLJ_OK) memcpy(s->nve_graph_key, &key, sizeof(key)), have = true;
      else cudaGetLastError(), rc = ATLJ_OK;  // fall back to plain launches
    }
    if (have) {
      for (; nsteps - k >= 2; k += 2) {
        ATLJ_CUDA(cudaGraphLaunch(s->nve_graph, s->st));
        g_launches += s->nve_graph_launches;
        m.step += 2;
      }
      rec_prev = s->rec_dev + (size_t)(k - 1) * ATLJ_NREC;
    }
  }
  for (; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    if ((rc = phase1(s, s->st, dt, true, rec_prev, nb, false, true)) || (rc = phase2(s, s->st, nullptr, rec, false)))
      return rc;
    if ((nb = phase3(s, s->st, false)) < 0) return nb;
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
  if (dt != 0.0) s->step_count += nsteps;  // a dt = 0 "step" is a force evaluation: the RNG counter does not move
  if ((rc = pull_n(s))) return rc;
  rc = sim_copy_rec(s, rec_host, nsteps);
  if (rec_host)
    for (int k = 0; k < nsteps; k++)
      if (rec_host[(size_t)k * ATLJ_NREC + 7] > 0.0) rec_host[(size_t)k * ATLJ_NREC + 7] = 1.0;
  return rc;
}

// md_nve_lj.py:178-192 with the hessian of md_nve_lj.py:46 in every step's record (rec[6]): the unfused sequence of the
// single-GPU hessian loop on a slab - integrate, exchange + sort, force (mask words kept), boundary forces to the
// neighbours, hessian, trailing half kick.  One all-reduce of the records per run.
int atlj_sim_run_nve_mg_hess(atlj_sim *s, double dt, int nsteps, double *rec_host) {
  if (!mg_ready(s) || !s->mg.comm || nsteps < 0) {
    set_error("run_nve_mg_hess: enable the slab, map the neighbours' buffers and connect first");
    return ATLJ_ERR_ARG;
  }
  if (!sim_tile2(s)) {
    set_error("run_nve_mg_hess: the hessian on slabs runs on the tiled LJ kernels only");
    return ATLJ_ERR_ARG;
  }
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc || (rc = sim_ensure_masks(s)) || (rc = push_n(s))) return rc;
  MgState &m = s->mg;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC * (size_t)(nsteps > 0 ? nsteps : 1), s->st));
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    if ((rc = phase1(s, s->st, dt, false, nullptr, 0, false, k > 0)) || (rc = phase2(s, s->st, nullptr, rec, false))) return rc;
    const int nb = phase3(s, s->st, false, true);
    if (nb < 0) return nb;
    if ((rc = launch_finalize(s->st, s->partials, nb, s->p.model, false, rec, s->cb.err + 2, nullptr, 0, tile2_rows(s))) ||
        (rc = force_send(s, s->st)) || (rc = hess_launch(s, s->st, rec)))
      return rc;
    Timer::begin(s, 2);
    rc = launch_kick_sums(s->st, 3ll * m.own_cap, s->v[s->cur], s->f, 0.5 * dt, s->vf_partials, rec + 4, s->cb.err + 4,
                          m.cnt_dev + 1);
    Timer::end(s);
    if (rc) return rc;
  }
  if (nsteps > 0) {
    Timer::begin(s, 3);
    ATLJ_NCCL(g_nccl.AllReduce(s->rec_dev, s->rec_dev, (size_t)nsteps * ATLJ_NREC, ncclDouble, ncclSum,
                               (ncclComm_t)m.comm, s->st));
    Timer::end(s);
  }
  if (dt != 0.0) s->step_count += nsteps;
  if ((rc = pull_n(s))) return rc;
  rc = sim_copy_rec(s, rec_host, nsteps);
  if (rec_host)
    for (int k = 0; k < nsteps; k++)
      if (rec_host[(size_t)k * ATLJ_NREC + 7] > 0.0) rec_host[(size_t)k * ATLJ_NREC + 7] = 1.0;
  return rc;
}

// md_nve_lj.py:178-192 with the rank's state in host memory (bench.py's end-to-end arm at N > 1): what comes in is what
// the previous call handed out, so the device-side tables of the last sort (boundary tiles, halo regions) still apply
int64_t atlj_sim_mg_own_cap(const atlj_sim *s) { return s && s->mg.on ? (int64_t)s->mg.own_cap : -1; }

int atlj_sim_nve_step_host_mg(atlj_sim *s, double dt, double *r_host, double *v_host, double *f_host, int32_t *id_host,
                              int64_t *n_io, double *rec_host) {
  if (!mg_ready(s) || !s->mg.comm || !r_host || !v_host || !f_host || !id_host || !n_io || *n_io < 0 ||
      *n_io > (int64_t)s->mg.own_cap) {
    set_error("nve_step_host_mg: bad arguments (slab enabled, neighbours mapped, ranks connected, n <= owned capacity)");
    return ATLJ_ERR_ARG;
  }
  int64_t n = *n_io;
  s->n = n;
  size_t b = sizeof(double) * 3 * (size_t)n;
  ATLJ_CUDA(cudaMemcpyAsync(s->r[s->cur], r_host, b, cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(s->v[s->cur], v_host, b, cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(s->f, f_host, b, cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(s->id[s->cur], id_host, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s->st));
  int rc = atlj_sim_run_nve_mg(s, dt, 1, rec_host);  // ends with the owned count back on the host (synchronised)
  if (rc) return rc;
  n = s->n;
  b = sizeof(double) * 3 * (size_t)n;
  ATLJ_CUDA(cudaMemcpyAsync(r_host, s->r[s->cur], b, cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(v_host, s->v[s->cur], b, cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(f_host, s->f, b, cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaMemcpyAsync(id_host, s->id[s->cur], sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  *n_io = n;
  return ATLJ_OK;
}

}  // extern "C"
