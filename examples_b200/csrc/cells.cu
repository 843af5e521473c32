// Link-cell construction as a counting sort.
//
// Reference: python_examples/md_lj_ll_module.py:88,106-120 (wrap, cell triplets, per-cell atom lists in C order of
// cells and ascending atom index) and link_list_module.f90:103-161 (make_list / c_index / create_in_list).
// The integer results (triplets, per-cell membership and order) are bit-identical to the reference's; the
// algorithm is not the reference's O(N * cells) mask sweep nor the Fortran head/list chain:
//   1. cell_assign : wrap, triplet, local cell id; one atomicAdd per atom gives the per-cell histogram and an
//                    arrival rank inside the cell                                  (reads 24 B, writes 8 B / atom)
//   2. cell_scan   : exclusive scan of the histogram -> cell_start       (4 B / cell, one decoupled-look-back kernel)
//   3. cell_scatter: key = (atom id << 32 | source index) -> slot cell_start + rank
//   4. cell_sortgather: per block of 32..128 cells, keys to shared memory, one thread per cell orders its <= ~30 keys
//                    by atom id (the arrival rank of step 1 is scheduling dependent, the sorted result is not) ->
//                    "ascending atom index" of the reference; then r (and v, id) are gathered into cell order
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
template <bool RCP>
__global__ void __launch_bounds__(IT, 8) kdk_assign_kernel(Geom g, int n, double *__restrict__ r, double *__restrict__ v,
                                                        const double *__restrict__ f, double half_dt, double dt,
                                                        double box, double inv_box, int first_kick,
                                                        int *__restrict__ cellid,
                                                        int *__restrict__ rank, int *__restrict__ cell_count,
                                                        int *__restrict__ err, double *__restrict__ partials,
                                                        double *__restrict__ rec_vf,
                                                        const double *__restrict__ pair_partials, int pair_rows,
                                                        int model, SlabArgs sl, const int *__restrict__ pair_rows_dev,
                                                        int *__restrict__ step_ctr, int *__restrict__ tab0) {
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
  const bool slab = sl.to_left[0] != nullptr;
  if (slab && err[5]) return;  // a neighbour's message timed out earlier in this run: the state stays as it is
  const int xstep = slab ? *sl.xstep : 0;  // exchange step: buffer parity and stamp (read before the last CTA moves it on)
  unsigned char *const to_left = sl.to_left[xstep & 1], *const to_right = sl.to_right[xstep & 1];
  if (sl.n_dev) n = min(n, *sl.n_dev);
  double s[2] = {0.0, 0.0};
  const double scd = (double)g.sc;
  const int ntiles = (n + IT - 1) / IT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t e0 = 3 * (size_t)tile * IT;
    const int ne = min(3 * IT, 3 * n - (int)e0 > 0 ? (int)(3 * (size_t)n - e0) : 0);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int l = threadIdx.x + k * IT;
      if (l < ne) {
        const size_t e = e0 + l;
        const double ff = f[e];
        double vv = v[e];
        if (first_kick) {
          vv = __dadd_rn(vv, __dmul_rn(half_dt, ff));
          s[0] += vv * vv;
          s[1] += ff * ff;
        }
        vv = __dadd_rn(vv, __dmul_rn(half_dt, ff));
        // (dt*v)/box, correctly rounded without the division subroutine (and its 64 registers): with y = RN(1/box)
        // from the host, q = RN(a*y), r = a - q*box (exact in an fma), RN(q + r*y) is RN(a/box) (Markstein); the host
        // passes inv_box = 0 for the one family of divisors the theorem excludes (significand all ones)
        const double a = __dmul_rn(dt, vv);
        double q;
        if (RCP) {
          q = __dmul_rn(a, inv_box);
          q = __fma_rn(__fma_rn(-q, box, a), inv_box, q);
        } else {
          q = __ddiv_rn(a, box);
        }
        double x = wrap1(__dadd_rn(r[e], q));
        x = __dsub_rn(x, rint(x));  // the force routine wraps again (md_lj_ll_module.py:88); idempotent
        v[e] = vv;
        r[e] = x;
        xs[l] = x;
        if (slab) vs[l] = vv;
      }
    }
    __syncthreads();
    const int i = tile * IT + threadIdx.x;
    if (i < n) {
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
        const int dir = !slab ? -1 : (c0 == left ? 0 : (c0 == right ? 1 : -1));
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
            d[6] = (double)sl.id[i];
            __threadfence_system();
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
      double t = 0.0;
      for (int b = r0; b < r1; b++) t += pair_partials[8 * (size_t)b + k];
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
  if (slab && threadIdx.x == 0) {
    // header: [0] = number of atoms, [1] = step stamp, released at system scope after the count
    for (int dir = 0; dir < 2; dir++) {
      int *hdr = reinterpret_cast<int *>(dir ? to_right : to_left);
      hdr[0] = min(atomicAdd(&sl.mig_cnt[dir], 0), sl.mig_cap);
      __threadfence_system();
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(hdr + 1), "r"(xstep + 1) : "memory");
      sl.mig_cnt[dir] = 0;
    }
    *sl.xstep = xstep + 1;
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
  if (!loop || !loop->counts_zeroed) ATLJ_CUDA(cudaMemsetAsync(b.cell_count, 0, sizeof(int) * (size_t)g.ncell(), st));
  int *step_ctr = loop ? loop->step_ctr : nullptr, *tab0 = loop ? loop->tab0 : nullptr;
  const int nb = integ_blocks(n);
  SlabArgs sl;
  memset(&sl, 0, sizeof(sl));
  if (slab) sl = *slab;
  // y = RN(1/box) for the division-free exact quotient; not for a divisor whose significand is all ones, nor when
  // the quotient could leave the normal range (never for a box length, but cheap to check)
  unsigned long long bits;
  memcpy(&bits, &box, sizeof(bits));
  const bool ones = (bits & 0xfffffffffffffull) == 0xfffffffffffffull;
  const double inv_box = (!ones && box > 1e-100 && box < 1e100) ? 1.0 / box : 0.0;
  if (inv_box != 0.0)
    kdk_assign_kernel<true><<<nb, IT, 0, st>>>(g, n, r, v, f, half_dt, dt, box, inv_box, first_kick ? 1 : 0, b.cellid,
                                               b.rank, b.cell_count, b.err, vf_partials, rec_vf, pair_partials,
                                               pair_rows, model, sl, pair_rows_dev, step_ctr, tab0);
  else
    kdk_assign_kernel<false><<<nb, IT, 0, st>>>(g, n, r, v, f, half_dt, dt, box, inv_box, first_kick ? 1 : 0, b.cellid,
                                                b.rank, b.cell_count, b.err, vf_partials, rec_vf, pair_partials,
                                                pair_rows, model, sl, pair_rows_dev, step_ctr, tab0);
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
                                                                unsigned long long *__restrict__ state) {
  __shared__ int sm[SCAN_T / 32];
  __shared__ int total;
  __shared__ int s_before;
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
    if (atomicAdd(state + SCAN_STATE_CTR, 1ull) == (unsigned long long)gridDim.x - 1ull) s_before = -1;
  }
  __syncthreads();
  if (s_before == -1) {
    for (int t = threadIdx.x; t < (int)gridDim.x; t += SCAN_T) state[t] = 0ull;
    if (threadIdx.x == 0) state[SCAN_STATE_CTR] = 0ull;
  }
}

int launch_cell_scan(cudaStream_t st, const Geom &g, const CellBufs &b, int base) {
  int n = g.n_own_cells();
  int first = g.first_own_cell();
  int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (ntiles > SCAN_STATE_CTR) {
    set_error("cell grid too large for the scan (%d cells)", n);
    return ATLJ_ERR_ARG;
  }
  scan_lookback_kernel<<<ntiles, SCAN_T, 0, st>>>(b.cell_count + first, n, base, b.cell_start, g.sc * g.sc, g.h,
                                                  reinterpret_cast<unsigned long long *>(b.block_sums));
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

// ---- 3. scatter keys into their cells ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cell_scatter_kernel(Geom g, int n, const int *__restrict__ n_dev, int base,
                                                           const int *__restrict__ cellid,
                                                           const int *__restrict__ rank,
                                                           const int *__restrict__ cell_start,
                                                           const int *__restrict__ id_in,
                                                           unsigned long long *__restrict__ keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  int cid = cellid[i];
  if (cid < 0) return;
  unsigned id = id_in ? (unsigned)id_in[i] : (unsigned)i;
  int slot = cell_start[g.cs_idx(cid)] - base + rank[i];
  keys[slot] = ((unsigned long long)id << 32) | (unsigned)i;
}

// ---- 4. order each cell by atom id -----------------------------------------------------------------------------
// insertion sort for ordinary cells (keys arrive nearly ordered when the atoms were in cell order already), heap
// sort for crowded ones (dense blobs); in place, k may point to shared or global memory
__device__ __forceinline__ void sort_cell_keys(unsigned long long *k, int n) {
  if (n < 2) return;
  if (n <= 32) {
    for (int a = 1; a < n; a++) {
      unsigned long long x = k[a];
      int b = a - 1;
      while (b >= 0 && k[b] > x) {
        k[b + 1] = k[b];
        b--;
      }
      k[b + 1] = x;
    }
  } else {
    for (int start = n / 2 - 1; start >= 0; start--) {
      int root = start;
      for (;;) {
        int child = 2 * root + 1;
        if (child >= n) break;
        if (child + 1 < n && k[child] < k[child + 1]) child++;
        if (k[root] >= k[child]) break;
        unsigned long long t = k[root];
        k[root] = k[child];
        k[child] = t;
        root = child;
      }
    }
    for (int end = n - 1; end > 0; end--) {
      unsigned long long t = k[0];
      k[0] = k[end];
      k[end] = t;
      int root = 0;
      for (;;) {
        int child = 2 * root + 1;
        if (child >= end) break;
        if (child + 1 < end && k[child] < k[child + 1]) child++;
        if (k[root] >= k[child]) break;
        unsigned long long u = k[root];
        k[root] = k[child];
        k[child] = u;
        root = child;
      }
    }
  }
}

// ---- 4+5. order each cell by atom id and gather r, v, id into cell order ----------------------------------------
// One CTA takes cpb (32..128) consecutive owned cells: their keys are one contiguous run of slots, which is loaded into
// shared memory with coalesced reads, sorted there (one thread per cell; the arrival rank of step 1 is scheduling
// dependent, the sorted result is not -> "ascending atom index" of md_lj_ll_module.py:118-119), and then drives the
// gather: one thread per (atom, component), so the writes of r, v in cell order are perfectly coalesced.  A run that
// does not fit (crowded cells) is sorted in global memory instead.
constexpr int SG_T = 256;      // most threads per CTA: the first `cpb` of them sort one cell each, all of them gather
constexpr int SG_KEYS = 3072;  // 24 KB of keys per CTA

// cells per CTA: 128 on large grids; fewer on small ones so that the gather still spreads over every SM
static inline int sg_cells_per_block(int ncells) { return ncells >= 148 * 2 * 128 ? 128 : (ncells >= 148 * 2 * 64 ? 64 : 32); }

__global__ void __launch_bounds__(SG_T) cell_sortgather_kernel(Geom g, int base, int block0, int cpb,
                                                                    const int *__restrict__ cell_start,
                                                                    int *cell_count,
                                                                    unsigned long long *__restrict__ keys,
                                                                    const double *__restrict__ r_in,
                                                                    const double *__restrict__ v_in,
                                                                    double *__restrict__ r_out,
                                                                    double *__restrict__ v_out,
                                                                    int *__restrict__ id_out, int zero_counts,
                                                                    HaloOut ho) {
  __shared__ unsigned long long sk[SG_KEYS];
  __shared__ int s_lo, s_hi;
  const int nown = g.n_own_cells(), first = g.first_own_cell();
  // slab ranks: the blocks of the first and the last owned layer run first, so that the halo data is on its way to
  // the neighbours while the interior is still being sorted
  int blk = block0 + blockIdx.x;
  if (ho.dst_r_left && ho.n_first + ho.n_last <= (int)gridDim.x) {
    const int b = blockIdx.x, nblk = gridDim.x;
    blk = b < ho.n_first ? b : (b < ho.n_first + ho.n_last ? nblk - ho.n_last + (b - ho.n_first) : b - ho.n_last);
  }
  const int c0 = blk * cpb;
  const int c = c0 + threadIdx.x;
  int mystart = 0, mycnt = 0;
  if (threadIdx.x < cpb && c < nown) {
    mystart = cell_start[g.cs_idx(first + c)] - base;
    mycnt = cell_count[first + c];
    if (zero_counts) cell_count[first + c] = 0;  // ready for the next step's histogram
  }
  if (threadIdx.x == 0) s_lo = mystart;
  const int clast = min(c0 + cpb, nown) - 1;
  if (threadIdx.x < cpb && c == clast) s_hi = mystart + mycnt;
  __syncthreads();
  const int lo = s_lo, tot = s_hi - s_lo;
  const bool fits = tot <= SG_KEYS;
  unsigned long long *kbase = fits ? sk : keys + lo;
  if (fits) {
    for (int s = threadIdx.x; s < tot; s += blockDim.x) sk[s] = keys[lo + s];
    __syncthreads();
  }
  sort_cell_keys(kbase + (mystart - lo), mycnt);
  __syncthreads();
  // slab ranks: slots [F0, F1) are the first owned layer = the left neighbour's right halo, [L0, L1) the last owned
  // layer = the right neighbour's left halo; their positions are stored a second time, straight into the neighbours'
  // position arrays (peer memory over NVLink), and the layers' cell tables, rebased to where the positions land, into
  // the halo rows of the neighbours' cell_start.  The CTA that completes a layer publishes the step stamp.
  int F0 = 0, F1 = 0, L0 = 0, L1 = 0;
  const int sc2 = g.sc * g.sc;
  if (ho.dst_r_left) {
    const int *rowF = cell_start + (size_t)g.h * (sc2 + 1), *rowL = cell_start + (size_t)(g.h + g.nx_own - 1) * (sc2 + 1);
    F0 = rowF[0], F1 = min(rowF[sc2], F0 + ho.cap), L0 = rowL[0], L1 = min(rowL[sc2], L0 + ho.cap);
  }
  for (int e = threadIdx.x; e < 3 * tot; e += blockDim.x) {
    const int s = e / 3, k = e - 3 * s;
    const unsigned long long key = kbase[s];
    const size_t src = (size_t)(key & 0xffffffffull);
    const int slot = lo + s;
    const size_t dst = 3 * (size_t)slot + k;
    const double x = r_in[3 * src + k];
    r_out[dst] = x;
    if (slot >= F0 && slot < F1) ho.dst_r_left[3 * (size_t)(slot - F0) + k] = x;
    if (slot >= L0 && slot < L1) ho.dst_r_right[3 * (size_t)(slot - L0) + k] = x;
    if (v_in) v_out[dst] = v_in[3 * src + k];
    if (k == 0 && id_out) id_out[slot] = (int)(key >> 32);
  }
  if (!ho.dst_r_left) return;
  if (threadIdx.x < cpb && c < nown) {
    const int cl = c - (g.nx_own - 1) * sc2;  // cell index inside the last layer
    if (c < sc2) {
      ho.dst_cs_left[c] = ho.base_left + min(mystart + base - F0, F1 - F0);
      if (c == sc2 - 1) ho.dst_cs_left[sc2] = ho.base_left + (F1 - F0);
    }
    if (cl >= 0) {
      ho.dst_cs_right[cl] = ho.base_right + min(mystart + base - L0, L1 - L0);
      if (cl == sc2 - 1) {
        ho.dst_cs_right[sc2] = ho.base_right + (L1 - L0);
        const int n_own = cell_start[(size_t)(g.h + g.nx_own - 1) * (sc2 + 1) + sc2];  // end of the last owned layer
        *ho.n_own = n_own;
        if (ho.rec) ho.rec[(ho.rec_ctr ? (size_t)ATLJ_NREC * (size_t)*ho.rec_ctr : 0) + 11] = (double)n_own;
        if (n_own > ho.own_cap || n_own - L0 > ho.cap || cell_start[(size_t)g.h * (sc2 + 1) + sc2] - F0 > ho.cap)
          atomicOr(ho.err + 1, 1);  // more atoms than the neighbours' halo regions / the owned region hold
      }
    }
  }
  const bool inF = c0 < sc2, inL = clast >= (g.nx_own - 1) * sc2;
  if (!inF && !inL) return;
  __shared__ int s_pub;
  __syncthreads();  // every thread's peer stores are ordered before thread 0's system-scope fence (cumulativity)
  if (threadIdx.x == 0) {
    __threadfence_system();
    int pub = 0;
    if (inF && atomicAdd(ho.done, 1) == ho.n_first - 1) pub |= 1;
    if (inL && atomicAdd(ho.done + 1, 1) == ho.n_last - 1) pub |= 2;
    s_pub = pub;
    const int stamp = *ho.xstep;  // the exchange step counter, already moved on by the integration kernel
    if (pub & 1) {
      ho.hdr_left[0] = F1 - F0;
      __threadfence_system();
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ho.hdr_left + 1), "r"(stamp) : "memory");
      ho.done[0] = 0;
    }
    if (pub & 2) {
      ho.hdr_right[0] = L1 - L0;
      __threadfence_system();
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ho.hdr_right + 1), "r"(stamp) : "memory");
      ho.done[1] = 0;
    }
  }
}

// n_dev (may be null): device-side count of atoms BEFORE the sort (n is then an upper bound); the sort + gather
// walks cells, not atoms, so it needs no count at all.
// halo (may be null): slab ranks - the first / last owned layer go to the neighbours from inside this kernel
int launch_cell_sort_gather(cudaStream_t st, const Geom &g, const CellBufs &b, int n, int base, const int *id_in,
                            const double *r_in, const double *v_in, double *r_out, double *v_out, int *id_out,
                            const int *n_dev, const HaloOut *halo, bool zero_counts) {
  if (n <= 0) return ATLJ_OK;
  cell_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(g, n, n_dev, base, b.cellid, b.rank, b.cell_start, id_in, b.keys);
  ATLJ_LAUNCHED();
  const int nc = g.n_own_cells(), sc2 = g.sc * g.sc;
  const int cpb = sg_cells_per_block(nc);
  const int nblk = (nc + cpb - 1) / cpb;
  HaloOut ho;
  memset(&ho, 0, sizeof(ho));
  if (halo) {
    ho = *halo;
    ho.n_first = (sc2 - 1) / cpb + 1;                 // blocks that hold cells of the first owned layer
    ho.n_last = nblk - ((g.nx_own - 1) * sc2) / cpb;  // ... of the last owned layer
  }
  // large grids: 128 cells and 128 threads per CTA (measured best at 2 M atoms); small grids: fewer cells per CTA
  // and twice the threads for the gather
  cell_sortgather_kernel<<<nblk, cpb == 128 ? 128 : SG_T, 0, st>>>(g, base, 0, cpb, b.cell_start, b.cell_count, b.keys,
                                                                   r_in, v_in, r_out, v_out, id_out, zero_counts ? 1 : 0,
                                                                   ho);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

}  // namespace atlj
