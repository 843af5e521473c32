// Pair-force / energy / hessian kernels.
//
// Reference arithmetic: python_examples/md_lj_module.py:96-115 (== md_lj_ll_module.py:134-165),
// md_lj_le_module.py:93-112 (WCA + Lees-Edwards image), md_lj_module.py:176-191 (hessian); final scaling
// md_lj_module.py:143-147.  Order of operations that decides a PREDICATE (in-range, overlap) is reproduced
// exactly: separate rounding of every add/mul via __dadd_rn/__dmul_rn (no FMA
// contraction), rint = round-half-even, IEEE reciprocal.  Everything downstream of the predicates is ordinary
// fp64 and agrees with NumPy to summation-order rounding.
//
// This file holds the all-pairs kernel (small systems, sc < 4: the reference's N = 256 case) and the fixed-order
// final reduction; the link-cell kernels are pair_tile.cu (WCA / Lees-Edwards, hessian) and pair_tile2.cu (LJ).
// FULL stencil everywhere: every atom gathers its own force, so no atomics and a deterministic result; the
// double-counted totals are halved exactly like the reference's own same-cell blocks (md_lj_ll_module.py:156-159).
#include "common.cuh"
#include "pair_math.cuh"

namespace atlj {


int pair_grid_blocks() { return 148 * 16; }

// ---- all-pairs kernel (small systems: sc < 4, e.g. the reference's N=256 case) ------------------------------------
// One thread per atom i, all j != i streamed through a shared-memory tile; exact same pair arithmetic.
constexpr int AP_T = 128;

template <int MODEL, bool HESS>
__global__ void __launch_bounds__(AP_T) pair_all_kernel(const double *__restrict__ r, const double *__restrict__ fin,
                                                        int n, Phys p, double *__restrict__ f,
                                                        double *__restrict__ partials) {
  constexpr bool LE = (MODEL == ATLJ_MODEL_WCA);
  __shared__ double xs[AP_T], ys[AP_T], zs[AP_T];
  __shared__ double gxs[HESS ? AP_T : 1], gys[HESS ? AP_T : 1], gzs[HESS ? AP_T : 1];
  const int i = blockIdx.x * AP_T + threadIdx.x;
  const bool vi = i < n;
  double rix = 0, riy = 0, riz = 0, fix = 0, fiy = 0, fiz = 0;
  if (vi) {
    rix = r[3 * i], riy = r[3 * i + 1], riz = r[3 * i + 2];
    if (HESS) fix = fin[3 * i], fiy = fin[3 * i + 1], fiz = fin[3 * i + 2];
  }
  Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0ull, 0};
  double fx = 0.0, fy = 0.0, fz = 0.0;
  for (int j0 = 0; j0 < n; j0 += AP_T) {
    __syncthreads();
    int j = j0 + threadIdx.x;
    if (j < n) {
      xs[threadIdx.x] = r[3 * j], ys[threadIdx.x] = r[3 * j + 1], zs[threadIdx.x] = r[3 * j + 2];
      if (HESS) gxs[threadIdx.x] = fin[3 * j], gys[threadIdx.x] = fin[3 * j + 1], gzs[threadIdx.x] = fin[3 * j + 2];
    }
    __syncthreads();
    int m = min(AP_T, n - j0);
    if (vi) {
      for (int t = 0; t < m; t++) {
        if (j0 + t == i) continue;
        double dx, dy, dz, s;
        if (pair_sep<LE>(rix, riy, riz, xs[t], ys[t], zs[t], p.strain, p.r_cut_box_sq, dx, dy, dz, s)) {
          if (HESS)
            pair_hess_term(p, dx, dy, dz, s, fix - gxs[t], fiy - gys[t], fiz - gzs[t], acc);
          else
            pair_force_term<MODEL>(p, dx, dy, dz, s, fx, fy, fz, acc);
        }
      }
    }
  }
  if (vi && !HESS) {
    f[3 * i] = 24.0 * fx;
    f[3 * i + 1] = 24.0 * fy;
    f[3 * i + 2] = 24.0 * fz;
  }
  block_store_partials<AP_T / 32>(acc, partials + 8 * (size_t)blockIdx.x);
}

int launch_pair_all(cudaStream_t st, const double *r, const double *fin, int n, const Phys &p, double *f,
                    double *partials, bool hessian, int *nblocks_out) {
  int blocks = (n + AP_T - 1) / AP_T;
  if (blocks < 1) blocks = 1;
  if (blocks > pair_grid_blocks()) {
    set_error("all-pairs route limited to %d atoms (got %d); use the cell route", pair_grid_blocks() * AP_T, n);
    return ATLJ_ERR_ARG;
  }
  if (p.model == ATLJ_MODEL_LJ) {
    if (hessian)
      pair_all_kernel<ATLJ_MODEL_LJ, true><<<blocks, AP_T, 0, st>>>(r, fin, n, p, f, partials);
    else
      pair_all_kernel<ATLJ_MODEL_LJ, false><<<blocks, AP_T, 0, st>>>(r, fin, n, p, f, partials);
  } else {
    if (hessian) {
      set_error("hessian is defined for the LJ model only");
      return ATLJ_ERR_ARG;
    }
    pair_all_kernel<ATLJ_MODEL_WCA, false><<<blocks, AP_T, 0, st>>>(r, fin, n, p, f, partials);
  }
  ATLJ_LAUNCHED();
  *nblocks_out = blocks;
  return ATLJ_OK;
}

// ---- direct link-cell kernel for cells of ~1 atom (WCA cut-off) --------------------------------------------------------
// With r_cut = 2^(1/6) the reference's cells hold about one atom and an atom has only ~29 candidates in its 27 (LE: up
// to 30) neighbour cells.  Staging row segments in shared memory (pair_tile*.cu) buys nothing there; one thread per
// atom walks the neighbour cells straight from the cell-sorted arrays (L1/L2 hits) with the reference's exact
// arithmetic for every candidate.  Selected when the grid holds fewer than 3 atoms per cell.
constexpr int DC_T = 128;

constexpr int DC_LIST = 48;  // candidate slots per atom in shared memory (the mean is ~25 at the WCA cut-off)

template <int MODEL, bool HESS>
__global__ void __launch_bounds__(DC_T) pair_direct_kernel(PairArgs a, int n) {
  constexpr bool LE = (MODEL == ATLJ_MODEL_WCA);
  __shared__ int cand[DC_LIST * DC_T];  // [slot][thread]: conflict-free
  const Geom g = a.g;
  const Phys p = a.p;
  const int i = blockIdx.x * DC_T + threadIdx.x;
  Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0ull, 0};
  if (i < n) {
    const double rix = a.r[3 * (size_t)i], riy = a.r[3 * (size_t)i + 1], riz = a.r[3 * (size_t)i + 2];
    const double scd = (double)g.sc;
    // the atom's own cell, with the arithmetic of the cell assignment (md_lj_ll_module.py:108)
    const int c0 = (int)floor(__dmul_rn(__dadd_rn(rix, 0.5), scd));
    const int cy = (int)floor(__dmul_rn(__dadd_rn(riy, 0.5), scd));
    const int cz = (int)floor(__dmul_rn(__dadd_rn(riz, 0.5), scd));
    const int lx = c0 - g.x_lo + g.h;
    double fix = 0.0, fiy = 0.0, fiz = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    if (HESS) fix = a.fin[3 * (size_t)i], fiy = a.fin[3 * (size_t)i + 1], fiz = a.fin[3 * (size_t)i + 2];
    auto one = [&](int j) {
      double dx, dy, dz, s;
      if (pair_sep<LE>(rix, riy, riz, a.r[3 * (size_t)j], a.r[3 * (size_t)j + 1], a.r[3 * (size_t)j + 2], p.strain,
                       p.r_cut_box_sq, dx, dy, dz, s)) {
        if (HESS)
          pair_hess_term(p, dx, dy, dz, s, fix - a.fin[3 * (size_t)j], fiy - a.fin[3 * (size_t)j + 1],
                         fiz - a.fin[3 * (size_t)j + 2], acc);
        else
          pair_force_term<MODEL>(p, dx, dy, dz, s, fx, fy, fz, acc);
      }
    };
    // Pass 1 (cheap, divergent): collect the candidate indices.  The three z neighbours of one (dx, dy) slot are
    // adjacent cells, i.e. ONE index range of the sorted arrays unless z wraps.
    int cnt = 0;
    auto range = [&](int j0, int j1) {
      for (int j = j0; j < j1; j++) {
        if (j == i) continue;
        if (cnt < DC_LIST)
          cand[(cnt++) * DC_T + threadIdx.x] = j;
        else
          one(j);  // list full (denser than expected): evaluate on the spot
      }
    };
    // a slot outside the grid only happens after an index error (reported by the cell build): contribute nothing
    const bool ok = lx >= g.h && lx < g.h + g.nx_own && cy >= 0 && cy < g.sc && cz >= 0 && cz < g.sc;
    for (int kk = 0; ok && kk < MAX_NB / 3; kk++) {
      const int nc = neighbour_cell(g, LE, p.shift, lx, cy, cz, 3 * kk + 1);  // the dz = 0 cell of the slot
      if (nc < 0) continue;
      const int idx = g.cs_idx(nc);
      const int lo = cz > 0 ? idx - 1 : idx, hi = cz < g.sc - 1 ? idx + 2 : idx + 1;
      range(a.cell_start[lo], a.cell_start[hi]);
      if (cz == 0 || cz == g.sc - 1) {
        const int w = cz == 0 ? idx + g.sc - 1 : idx - (g.sc - 1);
        range(a.cell_start[w], a.cell_start[w + 1]);
      }
    }
    // Pass 2 (expensive, nearly uniform trip count): the reference's arithmetic on every candidate
    for (int c = 0; c < cnt; c++) one(cand[c * DC_T + threadIdx.x]);
    if (!HESS) {
      a.f[3 * (size_t)i] = 24.0 * fx;  // md_lj_module.py:143
      a.f[3 * (size_t)i + 1] = 24.0 * fy;
      a.f[3 * (size_t)i + 2] = 24.0 * fz;
    }
  }
  block_store_partials<DC_T / 32>(acc, a.partials + 8 * (size_t)blockIdx.x);
}

int pair_direct_blocks(int n) { return (n + DC_T - 1) / DC_T > 0 ? (n + DC_T - 1) / DC_T : 1; }

// -> CTAs launched (= rows of partials written) or a negative status; n = owned atoms (they come first in a.r)
int launch_pair_direct(cudaStream_t st, const PairArgs &a, bool hessian, int n) {
  const int blocks = pair_direct_blocks(n);
  if (a.p.model == ATLJ_MODEL_LJ) {
    if (hessian)
      pair_direct_kernel<ATLJ_MODEL_LJ, true><<<blocks, DC_T, 0, st>>>(a, n);
    else
      pair_direct_kernel<ATLJ_MODEL_LJ, false><<<blocks, DC_T, 0, st>>>(a, n);
  } else {
    if (hessian) {
      set_error("hessian is defined for the LJ model only");
      return ATLJ_ERR_ARG;
    }
    pair_direct_kernel<ATLJ_MODEL_WCA, false><<<blocks, DC_T, 0, st>>>(a, n);
  }
  ATLJ_LAUNCHED();
  return blocks;
}

// ---- small systems: eight lanes per atom --------------------------------------------------------------------------------
// Small systems cannot fill a B200 with the tiled kernels (N = 32,000 is 250 chunks of 128 atoms for 592 warp
// schedulers) and a step is bound by the latency of ONE chunk.  Here an atom's candidates are spread over MK_L = 8
// lanes of a warp (4 atoms per warp, 16 per CTA; a whole warp per atom below 4,096 atoms).  Every lane runs the reference's exact
// arithmetic (pair_sep, md_lj_module.py:96-99) on its share of the candidates, parks the in-range ones in a small
// shared-memory list (pass 1: divergent but cheap), evaluates them (pass 2: nearly uniform trip count), and the eight
// partial forces are added with warp shuffles in a fixed order (md_lj_module.py:109,113).  CELLS: candidates are the
// atom's 27 (LE: up to 30) neighbour cells on the cell-sorted arrays (md_lj_ll_module.py:122-165); !CELLS: all atoms
// (md_lj_module.py:95-115, the reference's N = 256 case).
constexpr int MK_T = 128, MK_LIST = 20;  // lanes per atom (MK_L): 8, or 32 for systems of a few thousand atoms

template <int MODEL, bool HESS, bool CELLS, int MK_L>
__global__ void __launch_bounds__(MK_T) pair_multi_kernel(PairArgs a, int n) {
  constexpr bool LE = (MODEL == ATLJ_MODEL_WCA);
  __shared__ int cand[MK_LIST * MK_T];  // [slot][thread]: conflict-free
  const Geom g = a.g;
  const Phys p = a.p;
  const int sub = threadIdx.x & (MK_L - 1);
  const int i = (int)(((long long)blockIdx.x * MK_T + threadIdx.x) / MK_L);
  const bool act = i < n;
  Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0ull, 0};
  double fx = 0.0, fy = 0.0, fz = 0.0;
  if (act) {
    const double rix = a.r[3 * (size_t)i], riy = a.r[3 * (size_t)i + 1], riz = a.r[3 * (size_t)i + 2];
    double fix = 0.0, fiy = 0.0, fiz = 0.0;
    if (HESS) fix = a.fin[3 * (size_t)i], fiy = a.fin[3 * (size_t)i + 1], fiz = a.fin[3 * (size_t)i + 2];
    auto one = [&](int j) {
      double dx, dy, dz, s;
      if (pair_sep<LE>(rix, riy, riz, a.r[3 * (size_t)j], a.r[3 * (size_t)j + 1], a.r[3 * (size_t)j + 2], p.strain,
                       p.r_cut_box_sq, dx, dy, dz, s)) {
        if (HESS)
          pair_hess_term(p, dx, dy, dz, s, fix - a.fin[3 * (size_t)j], fiy - a.fin[3 * (size_t)j + 1],
                         fiz - a.fin[3 * (size_t)j + 2], acc);
        else
          pair_force_term<MODEL>(p, dx, dy, dz, s, fx, fy, fz, acc);
      }
    };
    int cnt = 0;
    auto range = [&](int j0, int j1) {  // this lane's share of the candidates [j0, j1)
      for (int j = j0 + sub; j < j1; j += MK_L) {
        if (j == i) continue;
        double dx, dy, dz, s;
        if (pair_sep<LE>(rix, riy, riz, a.r[3 * (size_t)j], a.r[3 * (size_t)j + 1], a.r[3 * (size_t)j + 2], p.strain,
                         p.r_cut_box_sq, dx, dy, dz, s)) {
          if (cnt < MK_LIST)
            cand[(cnt++) * MK_T + threadIdx.x] = j;
          else
            one(j);  // list full (a crowded neighbourhood): evaluate on the spot
        }
      }
    };
    if (CELLS) {
      const double scd = (double)g.sc;
      // the atom's own cell, with the arithmetic of the cell assignment (md_lj_ll_module.py:108)
      const int c0 = (int)floor(__dmul_rn(__dadd_rn(rix, 0.5), scd));
      const int cy = (int)floor(__dmul_rn(__dadd_rn(riy, 0.5), scd));
      const int cz = (int)floor(__dmul_rn(__dadd_rn(riz, 0.5), scd));
      const int lx = c0 - g.x_lo + g.h;
      // a slot outside the grid only happens after an index error (reported by the cell build): contribute nothing
      const bool ok = lx >= g.h && lx < g.h + g.nx_own && cy >= 0 && cy < g.sc && cz >= 0 && cz < g.sc;
      for (int kk = 0; ok && kk < MAX_NB / 3; kk++) {
        const int nc = neighbour_cell(g, LE, p.shift, lx, cy, cz, 3 * kk + 1);  // the dz = 0 cell of the slot
        if (nc < 0) continue;
        // the three z neighbours of one (dx, dy) slot are ONE index range of the sorted arrays unless z wraps
        const int idx = g.cs_idx(nc);
        const int lo = cz > 0 ? idx - 1 : idx, hi = cz < g.sc - 1 ? idx + 2 : idx + 1;
        range(a.cell_start[lo], a.cell_start[hi]);
        if (cz == 0 || cz == g.sc - 1) {
          const int w = cz == 0 ? idx + g.sc - 1 : idx - (g.sc - 1);
          range(a.cell_start[w], a.cell_start[w + 1]);
        }
      }
    } else {
      range(0, n);
    }
    for (int c = 0; c < cnt; c++) one(cand[c * MK_T + threadIdx.x]);
  }
  if (!HESS) {
    // the eight lanes of an atom are adjacent lanes of one warp: fixed-order butterfly over lane bits 2, 1, 0
#pragma unroll
    for (int o = MK_L / 2; o >= 1; o >>= 1) {
      fx += __shfl_xor_sync(FULL, fx, o);
      fy += __shfl_xor_sync(FULL, fy, o);
      fz += __shfl_xor_sync(FULL, fz, o);
    }
    if (act && sub == 0) {
      a.f[3 * (size_t)i] = 24.0 * fx;  // md_lj_module.py:143
      a.f[3 * (size_t)i + 1] = 24.0 * fy;
      a.f[3 * (size_t)i + 2] = 24.0 * fz;
    }
  }
  block_store_partials<MK_T / 32>(acc, a.partials + 8 * (size_t)blockIdx.x);
}

static inline int multi_lanes(int n) { return n <= 4096 ? 32 : 8; }
int pair_multi_blocks(int n) {
  const long long b = ((long long)n * multi_lanes(n) + MK_T - 1) / MK_T;
  return b > 0 ? (int)b : 1;
}

// -> CTAs launched (= rows of partials written) or a negative status; cells: candidates from the cell table of a.g,
// otherwise all n atoms of a.r
int launch_pair_multi(cudaStream_t st, const PairArgs &a, bool hessian, int n, bool cells) {
  const int blocks = pair_multi_blocks(n);
  const bool wide = multi_lanes(n) == 32;
#define ATLJ_MK(MODEL, HESS)                                                         \
  do {                                                                               \
    if (cells && wide)                                                               \
      pair_multi_kernel<MODEL, HESS, true, 32><<<blocks, MK_T, 0, st>>>(a, n);       \
    else if (cells)                                                                  \
      pair_multi_kernel<MODEL, HESS, true, 8><<<blocks, MK_T, 0, st>>>(a, n);        \
    else if (wide)                                                                   \
      pair_multi_kernel<MODEL, HESS, false, 32><<<blocks, MK_T, 0, st>>>(a, n);      \
    else                                                                             \
      pair_multi_kernel<MODEL, HESS, false, 8><<<blocks, MK_T, 0, st>>>(a, n);       \
  } while (0)
  if (a.p.model == ATLJ_MODEL_LJ) {
    if (hessian)
      ATLJ_MK(ATLJ_MODEL_LJ, true);
    else
      ATLJ_MK(ATLJ_MODEL_LJ, false);
  } else {
    if (hessian) {
      set_error("hessian is defined for the LJ model only");
      return ATLJ_ERR_ARG;
    }
    ATLJ_MK(ATLJ_MODEL_WCA, false);
  }
#undef ATLJ_MK
  ATLJ_LAUNCHED();
  return blocks;
}

// ---- force_1: one atom against a set of partners (smc_lj_module.py:100-191) ---------------------------------------
// One thread per partner j with the reference's arithmetic; f_partial[j] = force on i due to j (times 24, :190); the
// per-CTA rows go through the usual fixed-order finalize.  The caller zeroes everything when the overlap flag comes
// back set (:146-149).
__global__ void __launch_bounds__(128) force_1_kernel(double rix, double riy, double riz, const double *__restrict__ r,
                                                      int nj, Phys p, double *__restrict__ f_partial,
                                                      double *__restrict__ partials) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0ull, 0};
  if (j < nj) {
    double dx, dy, dz, s, fx = 0.0, fy = 0.0, fz = 0.0;
    if (pair_sep<false>(rix, riy, riz, r[3 * (size_t)j], r[3 * (size_t)j + 1], r[3 * (size_t)j + 2], 0.0,
                        p.r_cut_box_sq, dx, dy, dz, s))
      pair_force_term<ATLJ_MODEL_LJ>(p, dx, dy, dz, s, fx, fy, fz, acc);
    f_partial[3 * (size_t)j] = 24.0 * fx;
    f_partial[3 * (size_t)j + 1] = 24.0 * fy;
    f_partial[3 * (size_t)j + 2] = 24.0 * fz;
  }
  // finalize halves the sums (it serves kernels that visit every pair from both ends): each pair is visited once here
  acc.a *= 2.0, acc.pot *= 2.0, acc.vir *= 2.0, acc.lap *= 2.0, acc.np *= 2ull;
  block_store_partials<4>(acc, partials + 8 * (size_t)blockIdx.x);
}

int launch_force_1(cudaStream_t st, const double ri[3], const double *r, int nj, const Phys &p, double *f_partial,
                   double *partials) {
  const int blocks = (nj + 127) / 128 > 0 ? (nj + 127) / 128 : 1;
  force_1_kernel<<<blocks, 128, 0, st>>>(ri[0], ri[1], ri[2], r, nj, p, f_partial, partials);
  ATLJ_LAUNCHED();
  return blocks;
}

// ---- fixed-order reduction of the CTA partials + the reference's final scaling ------------------------------------
constexpr int FIN_T = 1024;
__global__ void __launch_bounds__(FIN_T) finalize_kernel(const double *__restrict__ partials, int nblocks, int model,
                                                         int hessian, double *__restrict__ rec,
                                                         int *__restrict__ reset_counter,
                                                         const double *__restrict__ vf_partials, int nvf,
                                                         const int *__restrict__ rows_dev,
                                                         const int *__restrict__ step_ctr) {
  if (rows_dev) nblocks = min(nblocks, *rows_dev);
  if (step_ctr) rec += (size_t)ATLJ_NREC * (size_t)*step_ctr;  // replayed launches: the row follows a device counter
  if (threadIdx.x == 0 && reset_counter) *reset_counter = 0;
  __shared__ double sm[FIN_T / 32][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += FIN_T) {
#pragma unroll
    for (int k = 0; k < 7; k++) v[k] += partials[8 * (size_t)b + k];
  }
  if (vf_partials)
    for (int b = threadIdx.x; b < nvf; b += FIN_T) v[7] += vf_partials[4 * (size_t)b], v[8] += vf_partials[4 * (size_t)b + 1];
#pragma unroll
  for (int k = 0; k < 9; k++) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 9; k++) sm[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[9];
    for (int k = 0; k < 9; k++) {
      double s = 0.0;
      for (int w = 0; w < FIN_T / 32; w++) s += sm[w][k];
      t[k] = s;
    }
    finalize_totals(t, model, hessian != 0, rec);
    if (vf_partials) rec[4] = t[7], rec[5] = t[8];
  }
}

int launch_finalize(cudaStream_t st, const double *partials, int nblocks, int model, bool hessian, double *rec,
                    int *reset_counter, const double *vf_partials, int nvf, const int *rows_dev, const int *step_ctr) {
  finalize_kernel<<<1, FIN_T, 0, st>>>(partials, nblocks, model, hessian ? 1 : 0, rec, reset_counter, vf_partials, nvf,
                                       rows_dev, step_ctr);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

}  // namespace atlj
