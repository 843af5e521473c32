// Link-cell construction as a counting sort.
//
// Reference: python_examples/md_lj_ll_module.py:88,106-120 (wrap, cell triplets, per-cell atom lists in C order of
// cells and ascending atom index) and link_list_module.f90:103-161 (make_list / c_index / create_in_list).
// The integer results (triplets, per-cell membership and order) are bit-identical to the reference's; the
// algorithm is not the reference's O(N * cells) mask sweep nor the Fortran head/list chain:
//   1. cell_assign : wrap, triplet, local cell id; one atomicAdd per atom gives the per-cell histogram and an
//                    arrival rank inside the cell                                  (reads 24 B, writes 8 B / atom)
//   2. cell_scan   : exclusive scan of the histogram -> cell_start       (4 B / cell, one decoupled-look-back kernel)
//   3. cell_scatter: atom id -> slot cell_start + arrival rank (every cell's id list complete, in arrival order)
//   4. cell_rankgather: one thread per source atom: slot = cell_start + (atoms of the cell with a smaller id) ->
//                    "ascending atom index" of the reference whatever the arrival order was; it moves its r, v, id there
// All HBM-bound; see DESIGN.md for bytes per atom.
#include <cstring>

#include "common.cuh"
#include "integrate.cuh"
#include "pair_math.cuh"

namespace atlj {

// ---- 1. wrap + assign ---------------------------------------------------------------------------------------
// The triplet must be floor((w+0.5)*sc) with the add and the multiply rounded separately (NumPy evaluates
// (r+0.5)*sc as two ufunc passes), so the intrinsics below forbid FMA contraction.
__global__ void __launch_bounds__(256) cell_assign_kernel(Geom g, const double *__restrict__ r_in,
                                                          double *__restrict__ r_wr, int n, int le, double strain,
                                                          int *__restrict__ cellid, int *__restrict__ rank,
                                                          int *__restrict__ cell_count, int *__restrict__ err,
                                                          long long *__restrict__ c_out,
                                                          const int *__restrict__ n_dev, int drop_adjacent,
                                                          const int *__restrict__ i0_dev) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i0_dev) {  // a sub-range [*i0_dev, *n_dev) of the atoms (arrivals appended behind the old ones); n = its bound
    if (i >= n) return;
    i += *i0_dev;
    n = *n_dev;
  } else if (n_dev) {
    n = min(n, *n_dev);
  }
  if (i >= n) return;
  double x = r_in[3 * i], y = r_in[3 * i + 1], z = r_in[3 * i + 2];
  if (le) x = __dsub_rn(x, __dmul_rn(rint(y), strain));  // md_lj_llle_module.py:94
  x = __dsub_rn(x, rint(x));                              // md_lj_ll_module.py:88
  y = __dsub_rn(y, rint(y));
  z = __dsub_rn(z, rint(z));
  if (r_wr) {
    r_wr[3 * i] = x;
    r_wr[3 * i + 1] = y;
    r_wr[3 * i + 2] = z;
  }
  double scd = (double)g.sc;
  long long c0 = (long long)floor(__dmul_rn(__dadd_rn(x, 0.5), scd));  // md_lj_ll_module.py:108
  long long c1 = (long long)floor(__dmul_rn(__dadd_rn(y, 0.5), scd));
  long long c2 = (long long)floor(__dmul_rn(__dadd_rn(z, 0.5), scd));
  if (c_out) {
    c_out[3 * i] = c0;
    c_out[3 * i + 1] = c1;
    c_out[3 * i + 2] = c2;
  }
  bool bad = c0 < 0 || c0 >= g.sc || c1 < 0 || c1 >= g.sc || c2 < 0 || c2 >= g.sc;  // md_lj_ll_module.py:109
  int lx = (int)c0 - g.x_lo + g.h;
  if (!bad && (lx < g.h || lx >= g.h + g.nx_own)) {  // not an owned layer
    const int left = g.x_lo == 0 ? g.sc - 1 : g.x_lo - 1, right = g.x_lo + g.nx_own == g.sc ? 0 : g.x_lo + g.nx_own;
    if (drop_adjacent && ((int)c0 == left || (int)c0 == right)) {  // packed for migration by migrate_pack_kernel
      cellid[i] = -1;
      rank[i] = 0;
      return;
    }
    bad = true;
  }
  if (bad) {
    atomicOr(&err[0], 1);
    cellid[i] = -1;
    rank[i] = 0;
    return;
  }
  int cid = (lx * g.sc + (int)c1) * g.sc + (int)c2;
  cellid[i] = cid;
  rank[i] = atomicAdd(&cell_count[cid], 1);
}

int launch_cell_assign(cudaStream_t st, const Geom &g, const double *r_in, double *r_wrapped, int n, bool le,
                       double strain, const CellBufs &b, long long *c_out, const int *n_dev, bool drop_adjacent,
                       const int *i0_dev) {
  if (!i0_dev) ATLJ_CUDA(cudaMemsetAsync(b.cell_count, 0, sizeof(int) * (size_t)g.ncell(), st));
  if (n > 0) {
    cell_assign_kernel<<<(n + 255) / 256, 256, 0, st>>>(g, r_in, r_wrapped, n, le ? 1 : 0, strain, b.cellid, b.rank,
                                                         b.cell_count, b.err, c_out, n_dev, drop_adjacent ? 1 : 0,
                                                         i0_dev);
    ATLJ_LAUNCHED();
  }
  return ATLJ_OK;
}

// ---- fused integration + assignment (single-GPU velocity-Verlet loop) ------------------------------------------------
// One HBM pass per step for everything between two force evaluations (md_nve_lj.py:190 of step k-1, then :182-185 of
// step k, then md_lj_ll_module.py:88,106-109):  v += hdt*f [sum v^2, sum f^2 of the finished step]  ->  v += hdt*f
// -> r += dt*v/box -> wrap -> cell triplet, histogram, arrival rank.  Same separately-rounded operations as
// kick_sums_kernel + kick_drift_kernel + cell_assign_kernel, so results are bit-identical to the unfused sequence.
#ifndef ATLJ_KDK_CTAS
#define ATLJ_KDK_CTAS 6
#endif
template <bool RCP, bool SLAB>
__global__ void __launch_bounds__(IT, ATLJ_KDK_CTAS) kdk_assign_kernel(Geom g, int n, double *__restrict__ r, double *__restrict__ v,
                                                        const double *__restrict__ f, double half_dt, double dt,
                                                        double box, double inv_box, int first_kick,
                                                        int *__restrict__ cellid,
                                                        int *__restrict__ rank, int *__restrict__ cell_count,
                                                        int *__restrict__ err, double *__restrict__ partials,
                                                        double *__restrict__ rec_vf,
                                                        const double *__restrict__ pair_partials, int pair_rows,
                                                        int model, SlabArgs sl, const int *__restrict__ pair_rows_dev,
                                                        int *__restrict__ step_ctr, int *__restrict__ tab0,
                                                        int no_assign, int tail_in_scan) {
  pdl_wait();  // (a no-op by default: this kernel is not launched programmatically, see common.cuh)
  pdl_launch_dependents();
  if (pair_rows_dev) pair_rows = min(pair_rows, *pair_rows_dev);
  // replayed launches (CUDA graph): the record of the finished step is rec_vf + NREC * *step_ctr; the CTA that finishes
  // last moves the counter on.  tab0: word 0 of the pair kernel's item table, zeroed here for this step's table kernel
  if (step_ctr) rec_vf += (size_t)ATLJ_NREC * (size_t)*step_ctr;
  if (tab0 && blockIdx.x == 0 && threadIdx.x == 0) *tab0 = 0;
  // tiles of IT atoms: the 3*IT components are streamed with perfectly coalesced accesses (one component per thread
  // and pass), the wrapped coordinates (and, on a slab, the velocities) are parked in shared memory, then one thread
  // per atom does the integer part
  __shared__ double xs[3 * IT];
  __shared__ double vs[3 * IT];
  constexpr bool slab = SLAB;
  if (slab && err[5]) return;  // a neighbour's message timed out earlier in this run: the state stays as it is
  const int xstep = slab ? *sl.xstep : 0;  // exchange step: buffer parity and stamp (read before the last CTA moves it on)
  unsigned char *const to_left = (xstep & 1) ? sl.to_left1 : sl.to_left0;
  unsigned char *const to_right = (xstep & 1) ? sl.to_right1 : sl.to_right0;
  if (sl.n_dev) n = min(n, *sl.n_dev);
  double s[2] = {0.0, 0.0};
  const double scd = (double)g.sc;
  const int ntiles = (n + IT - 1) / IT;
  // slab: tiles [0, nF) hold the first owned layer, [ntiles - nL, ntiles) the last one; they are taken first
  int nF = 0, nL = 0, nbt = 0;
  if (slab) {
    nF = ntiles, nL = 0;
    if (sl.first_end) {
      const int f1 = min(max(*sl.first_end, 0), n), l0 = min(max(*sl.last_begin, 0), n);
      nF = (f1 + IT - 1) / IT, nL = ntiles - l0 / IT;
      if (nF + nL >= ntiles) nF = ntiles, nL = 0;
    }
    nbt = nF + nL;
  }
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tile = !slab || t < nF ? t : (t < nbt ? ntiles - nL + (t - nF) : t - nL);
    const bool btile = slab && t < nbt;
    const size_t e0 = 3 * (size_t)tile * IT;
    const int ne = min(3 * IT, 3 * n - (int)e0 > 0 ? (int)(3 * (size_t)n - e0) : 0);
    if (slab && sl.assign_only) {
      // another integrator has moved (and wrapped) the atoms: cell assignment + migration only
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int l = threadIdx.x + k * IT;
        if (l < ne) {
          xs[l] = r[e0 + l];
          if (btile) vs[l] = v[e0 + l];
        }
      }
    } else {
      // all nine loads of a thread are issued before the first of them is used (the kernel is bound by the number of
      // bytes in flight: with one component at a time ncu showed 66 % of the stall cycles on the loads' scoreboard)
      double ff[3], vv[3], rr[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int l = threadIdx.x + k * IT;
        const size_t e = e0 + min(l, ne > 0 ? ne - 1 : 0);  // clamped: a valid address, the value is not used
        ff[k] = f[e], vv[k] = v[e], rr[k] = r[e];
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int l = threadIdx.x + k * IT;
        if (l < ne) {
          const size_t e = e0 + l;
          double w = vv[k];
          if (first_kick) {
            w = __dadd_rn(w, __dmul_rn(half_dt, ff[k]));
            s[0] += w * w;
            s[1] += ff[k] * ff[k];
          }
          w = __dadd_rn(w, __dmul_rn(half_dt, ff[k]));
          // (dt*v)/box, correctly rounded without the division subroutine (and its 64 registers): with y = RN(1/box)
          // from the host, q = RN(a*y), r = a - q*box (exact in an fma), RN(q + r*y) is RN(a/box) (Markstein); the
          // host passes inv_box = 0 for the one family of divisors the theorem excludes (significand all ones)
          const double a = __dmul_rn(dt, w);
          double q;
          if (RCP) {
            q = __dmul_rn(a, inv_box);
            q = __fma_rn(__fma_rn(-q, box, a), inv_box, q);
          } else {
            q = __ddiv_rn(a, box);
          }
          double x = wrap1(__dadd_rn(rr[k], q));
          x = __dsub_rn(x, rint(x));  // the force routine wraps again (md_lj_ll_module.py:88); idempotent
          v[e] = w;
          r[e] = x;
          xs[l] = x;
          if (slab) vs[l] = w;
        }
      }
    }
    __syncthreads();
    const int i = tile * IT + threadIdx.x;
    if (i < n && !no_assign) {
      const double w0 = xs[3 * threadIdx.x], w1 = xs[3 * threadIdx.x + 1], w2 = xs[3 * threadIdx.x + 2];
      const long long c0 = (long long)floor(__dmul_rn(__dadd_rn(w0, 0.5), scd));  // md_lj_ll_module.py:108
      const long long c1 = (long long)floor(__dmul_rn(__dadd_rn(w1, 0.5), scd));
      const long long c2 = (long long)floor(__dmul_rn(__dadd_rn(w2, 0.5), scd));
      bool bad = c0 < 0 || c0 >= g.sc || c1 < 0 || c1 >= g.sc || c2 < 0 || c2 >= g.sc;  // :109
      const bool foreign = !bad && (c0 < g.x_lo || c0 >= g.x_lo + g.nx_own);
      if (foreign) {
        // the atom left the slab: hand it (r, v, id) to the neighbour, straight into its receive buffer
        const int left = g.x_lo == 0 ? g.sc - 1 : g.x_lo - 1;
        const int right = g.x_lo + g.nx_own == g.sc ? 0 : g.x_lo + g.nx_own;
        // (an atom of an interior layer cannot reach a neighbour in one step; its messages are out already)
        const int dir = !btile ? -1 : (c0 == left ? 0 : (c0 == right ? 1 : -1));
        if (dir < 0) {
          bad = true;  // single domain: cannot happen; slab: moved more than one layer in a step
        } else {
          const int slot = atomicAdd(&sl.mig_cnt[dir], 1);
          if (slot >= sl.mig_cap) {
            atomicOr(&err[1], 1);
          } else {
            double *d = reinterpret_cast<double *>((dir ? to_right : to_left) + 64) + 7 * (size_t)slot;
            d[0] = w0, d[1] = w1, d[2] = w2;
            d[3] = vs[3 * threadIdx.x], d[4] = vs[3 * threadIdx.x + 1], d[5] = vs[3 * threadIdx.x + 2];
            d[6] = (double)sl.id[i];  // (no fence here: mig_publish_kernel, mg.cu, publishes after this kernel)
          }
          cellid[i] = -1;
          rank[i] = 0;
        }
      }
      if (bad) {
        atomicOr(&err[0], 1);
        cellid[i] = -1;
        rank[i] = 0;
      } else if (!foreign) {
        const int cid = (((int)c0 + g.h - g.x_lo) * g.sc + (int)c1) * g.sc + (int)c2;
        cellid[i] = cid;
        rank[i] = atomicAdd(&cell_count[cid], 1);
      }
    }
    __syncthreads();
  }
  if (!first_kick && !slab) return;
  // ---- finish the record of the step that just ended; publish the migration messages ------------------------------
  // level 1: this CTA's {sum v^2, sum f^2} and its fixed slice of the pair kernel's per-item rows -> one row of 12
  // level 2: the CTA that finishes last sums the rows in fixed order and writes the record (deterministic)
  constexpr int ROW = 12;
  if (first_kick) {
    block_reduce_store<2>(s, partials + ROW * (size_t)blockIdx.x);
    if (threadIdx.x >= 32 && threadIdx.x < 39) {
      const int k = threadIdx.x - 32;
      const int r0 = (int)((long long)pair_rows * blockIdx.x / gridDim.x);
      const int r1 = (int)((long long)pair_rows * (blockIdx.x + 1) / gridDim.x);
      // (eight rows requested before the first is added - same order of summation, adding +0.0 for the rows past the end
      // is exact: the plain loop waited for one L2 round trip per row, ~5 us at the end of every CTA's life)
      double t = 0.0;
      for (int b = r0; b < r1; b += 8) {
        double a[8];
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] = b + u < r1 ? __ldcg(pair_partials + 8 * (size_t)(b + u) + k) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++) t += a[u];
      }
      partials[ROW * (size_t)blockIdx.x + 2 + k] = t;
    }
  }
  __shared__ int s_last;
  __shared__ double fin[IT / 32][9];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&err[4], 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // (the migration messages - header counts and stamps - are published by mig_publish_kernel right after this kernel:
  // system-scope fences inside it, one per leaving atom's thread plus two in the tile that completed the boundary
  // layers, kept those CTAs waiting for ~15 us each: 70 us on a slab against 63.5 us for the single domain)
  if (slab && threadIdx.x == 0) *sl.xstep = xstep + 1;
  if (tail_in_scan) {  // level 2, the pair kernel's work counter and the step counter: the scan launch (RecTail)
    if (threadIdx.x == 0) err[4] = 0;
    return;
  }
  if (first_kick) {
    double t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += IT) {
#pragma unroll
      for (int k = 0; k < 9; k++) t[k] += __ldcg(partials + ROW * (size_t)b + k);
    }
#pragma unroll
    for (int k = 0; k < 9; k++) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < 9; k++) fin[threadIdx.x >> 5][k] = t[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 0; k < 9; k++) {
        double a = 0.0;
        for (int w = 0; w < IT / 32; w++) a += fin[w][k];
        t[k] = a;
      }
      rec_vf[4] = t[0], rec_vf[5] = t[1];
      if (pair_partials) finalize_totals(t + 2, model, false, rec_vf);
      err[2] = 0;  // work counter of the tiled pair kernel
    }
  }
  if (threadIdx.x == 0) {
    err[4] = 0;
    if (step_ctr) *step_ctr += 1;  // every CTA read it at its start, i.e. before the last one got here
  }
}

// -> number of CTAs (= rows of vf partials written when first_kick)
int launch_kdk_assign(cudaStream_t st, const Geom &g, int n, double *r, double *v, const double *f, double half_dt,
                      double dt, double box, bool first_kick, const CellBufs &b, double *vf_partials, double *rec_vf,
                      const double *pair_partials, int pair_rows, int model, const SlabArgs *slab,
                      const int *pair_rows_dev, const KdkLoop *loop) {
  // (in the replayed loop the counts were zeroed by the previous step's sort + gather, word 0 of the item table is
  // zeroed by this kernel and the record pointer follows a device counter: no memset nodes, no per-step arguments)
  const int no_assign = loop && loop->no_assign ? 1 : 0;
  if (!no_assign && (!loop || !loop->counts_zeroed))
    ATLJ_CUDA(cudaMemsetAsync(b.cell_count, 0, sizeof(int) * (size_t)g.ncell(), st));
  int *step_ctr = loop ? loop->step_ctr : nullptr, *tab0 = loop ? loop->tab0 : nullptr;
  int nb = integ_blocks(n);
  if (nb > 148 * ATLJ_KDK_CTAS) nb = 148 * ATLJ_KDK_CTAS;  // one resident wave of the grid-stride loop
  SlabArgs sl;
  memset(&sl, 0, sizeof(sl));
  if (slab) sl = *slab;
  // y = RN(1/box) for the division-free exact quotient; not for a divisor whose significand is all ones, nor when
  // the quotient could leave the normal range (never for a box length, but cheap to check)
  unsigned long long bits;
  memcpy(&bits, &box, sizeof(bits));
  const bool ones = (bits & 0xfffffffffffffull) == 0xfffffffffffffull;
  const double inv_box = (!ones && box > 1e-100 && box < 1e100) ? 1.0 / box : 0.0;
#define ATLJ_KDK(RCP, SLAB)                                                                                          \
  launch_pdl(PDL_KDK, kdk_assign_kernel<RCP, SLAB>, dim3(nb), dim3(IT), 0, st, g, n, r, v, f, half_dt, dt, box, inv_box,        \
             first_kick ? 1 : 0, b.cellid, b.rank, b.cell_count, b.err, vf_partials, rec_vf, pair_partials, pair_rows,  \
             model, sl, pair_rows_dev, step_ctr, tab0, no_assign, (loop && loop->tail_in_scan && first_kick) ? 1 : 0)
  if (inv_box != 0.0) {
    if (slab)
      ATLJ_KDK(true, true);
    else
      ATLJ_KDK(true, false);
  } else {
    if (slab)
      ATLJ_KDK(false, true);
    else
      ATLJ_KDK(false, false);
  }
#undef ATLJ_KDK
  ATLJ_LAUNCHED();
  return nb;
}

// ---- 2. exclusive scan over the owned cells ------------------------------------------------------------------
constexpr int SCAN_T = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_T * SCAN_ITEMS;  // 4096 cells per block

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem /*[SCAN_T/32]*/, int *total) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    int s = lane < SCAN_T / 32 ? smem[lane] : 0;
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    if (lane < SCAN_T / 32) smem[lane] = si - s;  // exclusive warp offsets
    if (lane == 31) *total = si;
  }
  __syncthreads();
  return incl - v + smem[w];
}

// Single-pass scan (decoupled look-back): every CTA scans its tile of 4096 cells, publishes the tile's aggregate, then
// adds up the aggregates / inclusive prefixes of the tiles before it, 32 at a time (CTAs are scheduled in index
// order, so a CTA only ever waits for CTAs that are already running).  One launch instead of three; integer sums, so
// the result does not depend on the order in which the look-back finds its terms.
// state[t] = value << 2 | flag (0 = nothing yet, 1 = aggregate of tile t, 2 = inclusive prefix up to tile t);
// state[SCAN_STATE_CTR] counts finished CTAs: the last one clears the states for the next launch.
// cell_start[first + c] = base_off + prefix, written into the padded table of the whole local grid (Geom::cs_idx: the
// k-th owned cell sits in layer h + k / sc2; one end sentinel per layer).
constexpr int SCAN_STATE_CTR = 4095;  // CellBufs::block_sums holds 4096 64-bit words
__global__ void __launch_bounds__(SCAN_T) scan_lookback_kernel(const int *__restrict__ cnt, int n, int base_off,
                                                                int *__restrict__ start, int sc2, int h,
                                                                unsigned long long *__restrict__ state, RecTail tail) {
  __shared__ int sm[SCAN_T / 32];
  __shared__ int total;
  __shared__ int s_before;
  pdl_wait();
  pdl_launch_dependents();
  const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if ((int)blockIdx.x == ntiles) {
    // the extra CTA: second level of the step record's reduction (RecTail), same order of summation as the last CTA
    // of the integration kernel used (IT threads stride over the rows, xor-shuffles, warps in index order)
    constexpr int ROW = 12;
    __shared__ double fin[IT / 32][9];
    double *rec = tail.rec + (tail.step_ctr ? (size_t)ATLJ_NREC * (size_t)*tail.step_ctr : 0);
    double t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (threadIdx.x < IT) {
      for (int b = threadIdx.x; b < tail.nrows; b += IT) {
#pragma unroll
        for (int k = 0; k < 9; k++) t[k] += __ldcg(tail.rows + ROW * (size_t)b + k);
      }
#pragma unroll
      for (int k = 0; k < 9; k++) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
      }
      if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) fin[threadIdx.x >> 5][k] = t[k];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 0; k < 9; k++) {
        double a = 0.0;
        for (int w = 0; w < IT / 32; w++) a += fin[w][k];
        t[k] = a;
      }
      rec[4] = t[0], rec[5] = t[1];
      if (tail.has_pair) finalize_totals(t + 2, tail.model, false, rec);
      tail.err[2] = 0;  // work counter of the tiled pair kernel
      if (tail.step_ctr) *tail.step_ctr += 1;
    }
    return;
  }
  const int tile = blockIdx.x;
  const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = base + k < n ? cnt[base + k] : 0;
    s += v[k];
  }
  const int local = block_exclusive_scan(s, sm, &total);
  if (threadIdx.x == 0) {
    const unsigned long long word = ((unsigned long long)(unsigned)total << 2) | (tile == 0 ? 2ull : 1ull);
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(state + tile), "l"(word) : "memory");
  }
  if (threadIdx.x < 32) {
    int before = 0;
    int j0 = tile - 1;  // look-back window [j0 - 31, j0]
    while (j0 >= 0) {
      const int j = j0 - (int)threadIdx.x;
      unsigned long long w = 2ull;  // lanes past tile 0: an empty inclusive prefix
      if (j >= 0) {
        do {
          asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(state + j) : "memory");
        } while ((w & 3ull) == 0ull);
      }
      // the nearest tile that already knows its inclusive prefix ends the look-back
      const unsigned has_p = __ballot_sync(0xffffffffu, (w & 3ull) == 2ull);
      const int stop = has_p ? __ffs(has_p) - 1 : 32;
      int val = ((int)threadIdx.x <= stop) ? (int)(w >> 2) : 0;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
      before += val;
      if (has_p) break;
      j0 -= 32;
    }
    if (threadIdx.x == 0) {
      s_before = before;
      if (tile > 0) {
        const unsigned long long word = ((unsigned long long)(unsigned)(before + total) << 2) | 2ull;
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(state + tile), "l"(word) : "memory");
      }
    }
  }
  __syncthreads();
  int off = local + s_before + base_off;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int c = base + k;
    if (c < n) {
      const int layer = c / sc2, w = c - layer * sc2;
      const int idx = (h + layer) * (sc2 + 1) + w;
      start[idx] = off;
      if (w == sc2 - 1) start[idx + 1] = off + v[k];  // end sentinel of this layer
    }
    off += v[k];
  }
  // the CTA that finishes last clears the states: every other CTA is past its look-back by then
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + SCAN_STATE_CTR, 1ull) == (unsigned long long)ntiles - 1ull) s_before = -1;
  }
  __syncthreads();
  if (s_before == -1) {
    for (int t = threadIdx.x; t < ntiles; t += SCAN_T) state[t] = 0ull;
    if (threadIdx.x == 0) state[SCAN_STATE_CTR] = 0ull;
  }
}

int launch_cell_scan(cudaStream_t st, const Geom &g, const CellBufs &b, int base, const RecTail *tail) {
  int n = g.n_own_cells();
  int first = g.first_own_cell();
  int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (ntiles > SCAN_STATE_CTR) {
    set_error("cell grid too large for the scan (%d cells)", n);
    return ATLJ_ERR_ARG;
  }
  RecTail rt;
  if (tail) rt = *tail;
  launch_pdl(PDL_SCAN, scan_lookback_kernel, dim3(ntiles + (rt.nrows > 0 ? 1 : 0)), dim3(SCAN_T), 0, st,
             b.cell_count + first, n, base, b.cell_start, g.sc * g.sc, g.h,
             reinterpret_cast<unsigned long long *>(b.block_sums), rt);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

// ---- 3. scatter the atom ids into their cells --------------------------------------------------------------------
// ids[cell_start + arrival rank] = id: after this kernel every cell's id list is complete (in arrival order, which is
// scheduling dependent - the next kernel turns it into the reference's ascending order).
__global__ void __launch_bounds__(256) cell_scatter_kernel(Geom g, int n, const int *__restrict__ n_dev, int base,
                                                           const int *__restrict__ cellid,
                                                           const int *__restrict__ rank,
                                                           const int *__restrict__ cell_start,
                                                           const int *__restrict__ id_in,
                                                           unsigned *__restrict__ ids) {
  pdl_wait();
  pdl_launch_dependents();
  if (n_dev) n = min(n, *n_dev);
  // (grid-stride: on a slab the grid is sized from the host's last known atom count, not from the capacity)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int cid = cellid[i];
    if (cid >= 0) ids[cell_start[g.cs_idx(cid)] - base + rank[i]] = id_in ? (unsigned)id_in[i] : (unsigned)i;
  }
}

// ---- 4. rank inside the cell + gather --------------------------------------------------------------------------------
// One thread per SOURCE atom: its slot in the cell-sorted arrays is cell_start + (number of atoms of its cell with a
// smaller id) - "ascending atom index" of md_lj_ll_module.py:118-119, independent of the arrival order - and it moves
// its own r, v, id there.  The cell's ids (a dozen, one L1 line shared by the neighbouring threads, which mostly sit
// in the same cell because the old order was the previous step's cell order) are counted, not sorted: no shared
// memory, no barrier, no phases - the kernel streams (reads coalesced by source, writes nearly sequential) and its
// run time does not depend on how the grid divides into waves of CTAs (the phased sort + gather kernel it replaces
// took 56 us with 1,300 CTAs and 105 us with 1,340 on 148 SMs).
// Slab ranks (ho.dst_r_left != null): slots [F0, F1) are the first owned layer = the left neighbour's right halo,
// [L0, L1) the last owned layer = the right neighbour's left halo; those positions are stored a second time, straight
// into the neighbours' position arrays (peer memory over NVLink), the layers' cell tables - rebased to where the
// positions land - into the halo rows of the neighbours' cell_start, and the CTA that finishes last publishes the step
// stamp in both neighbours' headers.
constexpr int RG_T = 256;
template <bool HALO>
__global__ void __launch_bounds__(RG_T, HALO ? 6 : 8) cell_rankgather_kernel(Geom g, int n, const int *__restrict__ n_dev, int base,
                                                                const int *__restrict__ cellid,
                                                                const int *__restrict__ id_in,
                                                                const int *__restrict__ cell_start,
                                                                int *__restrict__ cell_count,
                                                                const unsigned *__restrict__ ids,
                                                                const double *__restrict__ r_in,
                                                                const double *__restrict__ v_in,
                                                                double *__restrict__ r_out,
                                                                double *__restrict__ v_out,
                                                                int *__restrict__ id_out, int zero_counts,
                                                                HaloOut ho) {
  pdl_wait();
  pdl_launch_dependents();
  if (n_dev) n = min(n, *n_dev);
  const int nown = g.n_own_cells(), first = g.first_own_cell(), sc2 = g.sc * g.sc;
  constexpr bool halo = HALO;
  // threads needed: every source atom, every owned cell (zero_counts), every entry of a halo cell table.  On a slab
  // the atom count lives on the device and the grid is sized from the host's last known count: a grid-stride loop
  // takes whatever lies beyond (a grid over the CAPACITY was 2,300 idle CTAs)
  int cover = n;
  if (zero_counts && nown > cover) cover = nown;
  if (halo && sc2 + 1 > cover) cover = sc2 + 1;
  int F0 = 0, F1 = 0, L0 = 0, L1 = 0;
  for (int i = blockIdx.x * RG_T + threadIdx.x; i < cover; i += gridDim.x * RG_T) {
    if (zero_counts && i < nown) cell_count[first + i] = 0;  // ready for the next step's histogram
    int cid = -1;
    // the atom's own data is requested first: those loads are in flight while the rank is being counted
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0;
    if (i < n) {
      cid = cellid[i];
      x0 = r_in[3 * (size_t)i], x1 = r_in[3 * (size_t)i + 1], x2 = r_in[3 * (size_t)i + 2];
      if (v_in) w0 = v_in[3 * (size_t)i], w1 = v_in[3 * (size_t)i + 1], w2 = v_in[3 * (size_t)i + 2];
    }
    const int *rowF = cell_start + (size_t)g.h * (sc2 + 1), *rowL = cell_start + (size_t)(g.h + g.nx_own - 1) * (sc2 + 1);
    // slab: only atoms of the first / last owned layer (6 % of them: one comparison on the cell id tells) and the
    // threads that copy the cell tables need the extent of those layers in the sorted arrays
    const bool inF = halo && cid >= 0 && cid < (g.h + 1) * sc2, inL = halo && cid >= (g.h + g.nx_own - 1) * sc2;
    if (halo && (inF || inL || i <= sc2)) {
      F0 = rowF[0], F1 = min(rowF[sc2], F0 + ho.cap), L0 = rowL[0], L1 = min(rowL[sc2], L0 + ho.cap);
      if (i <= sc2) {  // the two layers' cell tables, rebased (entry sc2 = end sentinel)
        ho.dst_cs_left[i] = ho.base_left + min(rowF[i] - F0, F1 - F0);
        ho.dst_cs_right[i] = ho.base_right + min(rowL[i] - L0, L1 - L0);
      }
      if (i == 0) {
        const int n_own = rowL[sc2];  // end of the last owned layer = owned atoms after the sort
        *ho.n_own = n_own;
        if (ho.rec) ho.rec[(ho.rec_ctr ? (size_t)ATLJ_NREC * (size_t)*ho.rec_ctr : 0) + 11] = (double)n_own;
        if (n_own > ho.own_cap || rowL[sc2] - L0 > ho.cap || rowF[sc2] - F0 > ho.cap)
          atomicOr(ho.err + 1, 1);  // more atoms than the owned region / the neighbours' halo regions hold
      }
    }
    if (cid >= 0) {
      const unsigned myid = id_in ? (unsigned)id_in[i] : (unsigned)i;
      const int ci = g.cs_idx(cid);
      const int s0 = cell_start[ci] - base, s1 = cell_start[ci + 1] - base;
      int slot = s0;
      for (int s = s0; s < s1; s++) slot += ids[s] < myid;
      const size_t d = 3 * (size_t)slot;
      r_out[d] = x0, r_out[d + 1] = x1, r_out[d + 2] = x2;
      if (v_in) v_out[d] = w0, v_out[d + 1] = w1, v_out[d + 2] = w2;
      if (id_out) id_out[slot] = (int)myid;
      if (inF || inL) {
        const int gs = slot + base;
        if (gs >= F0 && gs < F1) {
          double *p = ho.dst_r_left + 3 * (size_t)(gs - F0);
          p[0] = x0, p[1] = x1, p[2] = x2;
        }
        if (gs >= L0 && gs < L1) {
          double *p = ho.dst_r_right + 3 * (size_t)(gs - L0);
          p[0] = x0, p[1] = x1, p[2] = x2;
        }
      }
    }
  }
}

// Slab ranks: the stamps of the two halo messages.  The rank + gather kernel stores the boundary layers into the
// neighbours' arrays with plain stores and NO fence: a system-scope fence inside that kernel (one per CTA that had stored
// to a neighbour in round 1, one per warp in a later attempt) kept every fencing warp resident for ~15 us - ncu showed
// 27 % of the kernel's stall samples on MEMBAR.SYS/ERRBAR and the slab form of the kernel at 57 us against 38 us for
// the single domain.  The kernel boundary orders those stores before this one-warp kernel, which fences once and
// publishes (release, system scope) the exchange step counter - already moved on by the integration kernel - as the
// stamp in both neighbours' headers.
__global__ void halo_publish_kernel(Geom g, const int *__restrict__ cell_start, HaloOut ho) {
  if (threadIdx.x != 0) return;
  const int sc2 = g.sc * g.sc;
  const int *rowF = cell_start + (size_t)g.h * (sc2 + 1), *rowL = cell_start + (size_t)(g.h + g.nx_own - 1) * (sc2 + 1);
  const int F0 = rowF[0], F1 = min(rowF[sc2], F0 + ho.cap), L0 = rowL[0], L1 = min(rowL[sc2], L0 + ho.cap);
  const int stamp = *ho.xstep;
  ho.hdr_left[0] = F1 - F0, ho.hdr_right[0] = L1 - L0;
  __threadfence_system();  // ONE release fence for both stamps
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(ho.hdr_left + 1), "r"(stamp) : "memory");
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(ho.hdr_right + 1), "r"(stamp) : "memory");
}

// n: atoms before the sort, or an upper bound when the count lives on the device (n_dev)
// halo (may be null): slab ranks - the first / last owned layer go to the neighbours from inside the gather kernel
// zero_counts: cell_count is left zeroed for the next step's histogram
int launch_cell_sort_gather(cudaStream_t st, const Geom &g, const CellBufs &b, int n, int base, const int *id_in,
                            const double *r_in, const double *v_in, double *r_out, double *v_out, int *id_out,
                            const int *n_dev, const HaloOut *halo, bool zero_counts, int n_hint) {
  if (n <= 0) return ATLJ_OK;
  unsigned *ids = reinterpret_cast<unsigned *>(b.keys);
  // threads to launch: n when it is exact; when the count lives on the device, the host's last known count plus a
  // margin (the kernels loop grid-stride over anything beyond), never more than the upper bound n
  int nt = n;
  if (n_dev && n_hint > 0) {
    const long long want = (long long)n_hint + n_hint / 64 + 4096;
    if (want < nt) nt = (int)want;
  }
  launch_pdl(PDL_SCATTER, cell_scatter_kernel, dim3((nt + 255) / 256), dim3(256), 0, st, g, n, n_dev, base, b.cellid,
             b.rank, b.cell_start, id_in, ids);
  ATLJ_LAUNCHED();
  HaloOut ho;
  memset(&ho, 0, sizeof(ho));
  if (halo) ho = *halo;
  int cover = nt;  // threads: every source atom, every owned cell (zero_counts), every entry of a halo cell table
  if (zero_counts && g.n_own_cells() > cover) cover = g.n_own_cells();
  if (halo && g.sc * g.sc + 1 > cover) cover = g.sc * g.sc + 1;
  if (halo) {
    launch_pdl(PDL_GATHER, cell_rankgather_kernel<true>, dim3((cover + RG_T - 1) / RG_T), dim3(RG_T), 0, st, g, n, n_dev,
               base, b.cellid, id_in, b.cell_start, b.cell_count, ids, r_in, v_in, r_out, v_out, id_out,
               zero_counts ? 1 : 0, ho);
    ATLJ_LAUNCHED();
    halo_publish_kernel<<<1, 32, 0, st>>>(g, b.cell_start, ho);
  } else
    launch_pdl(PDL_GATHER, cell_rankgather_kernel<false>, dim3((cover + RG_T - 1) / RG_T), dim3(RG_T), 0, st, g, n, n_dev,
               base, b.cellid, id_in, b.cell_start, b.cell_count, ids, r_in, v_in, r_out, v_out, id_out,
               zero_counts ? 1 : 0, ho);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

}  // namespace atlj
