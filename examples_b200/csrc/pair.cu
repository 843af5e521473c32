// Pair-force / energy / hessian kernels.
//
// Reference arithmetic: python_examples/md_lj_module.py:96-115 (== md_lj_ll_module.py:134-165),
// md_lj_le_module.py:93-112 (WCA + Lees-Edwards image), md_lj_module.py:176-191 (hessian); final scaling
// md_lj_module.py:143-147.  Order of operations that decides a PREDICATE (in-range, overlap) is reproduced
// exactly: separate rounding of every add/mul via __dadd_rn/__dmul_rn (no FMA
// contraction), rint = round-half-even, IEEE reciprocal.  Everything downstream of the predicates is ordinary
// fp64 and agrees with NumPy to summation-order rounding.
//
// Design (not the reference's): FULL stencil, every atom gathers its own force, so no atomics and a
// deterministic result; the double-counted totals are halved exactly like the reference's own same-cell blocks
// (md_lj_ll_module.py:156-159).  One CTA per i-cell, PW warps.  The candidates of the 27 (LE: up to 30) neighbour
// cells are staged once per cell into shared memory: fp64 coordinates for the exact arithmetic plus an fp32
// copy relative to the cell centre (cell units).  For each i atom one warp sweeps the candidates 32 at a time
// with a cheap fp32 distance prefilter (7 FP32 ops / candidate instead of 15 FP64), compacts the survivors into
// a ring buffer with ballot/popc, and evaluates them 32 at a time with every lane busy in fp64.  The prefilter
// threshold carries a 1e-4 relative margin (fp32 error bound ~1.5e-6), so it never rejects a pair the
// exact fp64 test accepts; the exact test then decides.  Per-atom forces are finished with a warp-shuffle
// reduction, totals accumulate per lane in fp64 and are reduced per CTA, then by launch_finalize in fixed order.
#include "common.cuh"
#include "pair_math.cuh"

namespace atlj {

constexpr int PW = 4;      // warps per CTA
constexpr int CAP = 512;   // staged candidates per chunk (27 cells x ~12-15 atoms fits; larger sets loop)

int pair_grid_blocks() { return 148 * 16; }

// ---- link-cell kernel --------------------------------------------------------------------------------------------
template <int MODEL, bool HESS>
__global__ void __launch_bounds__(32 * PW) pair_cells_kernel(PairArgs a) {
  constexpr bool LE = (MODEL == ATLJ_MODEL_WCA);
  __shared__ double xs[CAP], ys[CAP], zs[CAP];
  __shared__ float4 qs[CAP];
  __shared__ double gxs[HESS ? CAP : 1], gys[HESS ? CAP : 1], gzs[HESS ? CAP : 1];
  __shared__ int nb_start[32];
  __shared__ int nb_off[33];
  __shared__ int queue[PW][64];

  const Geom g = a.g;
  const Phys p = a.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int sc = g.sc;
  const double scd = (double)sc;
  const double sci = 1.0 / scd;
  const int n_own = g.n_own_cells();
  const int first = g.first_own_cell();
  int *q = queue[warp];

  Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0ull, 0};

  for (int oc = blockIdx.x; oc < n_own; oc += gridDim.x) {
    const int cell = first + oc;
    const int ni = a.cell_count[cell];
    if (ni == 0) continue;  // uniform over the CTA
    const int istart = a.cell_start[cell];
    const int lx = cell / (sc * sc), cy = (cell / sc) % sc, cz = cell % sc;

    __syncthreads();  // the previous cell's consumers are done with shared memory
    if (warp == 0) {
      int nc = lane < MAX_NB ? neighbour_cell(g, LE, p.shift, lx, cy, cz, lane) : -1;
      int cnt = nc >= 0 ? a.cell_count[nc] : 0;
      int st = nc >= 0 ? a.cell_start[nc] : 0;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
      }
      nb_start[lane] = st;
      nb_off[lane] = incl - cnt;
      if (lane == 31) nb_off[32] = incl;
    }
    __syncthreads();
    const int M = nb_off[32];

    // centre of the i-cell in box units (approximate geometry only: used for the fp32 prefilter copy)
    const double ctrx = ((lx - g.h + g.x_lo) + 0.5) * sci - 0.5;
    const double ctry = (cy + 0.5) * sci - 0.5;
    const double ctrz = (cz + 0.5) * sci - 0.5;

    for (int c0 = 0; c0 < M; c0 += CAP) {
      const int mc = min(CAP, M - c0);
      if (c0 > 0) __syncthreads();
      // ---- stage candidates c0 .. c0+mc
      for (int s = threadIdx.x; s < mc; s += 32 * PW) {
        int gi = c0 + s;
        int k = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1)
          if (nb_off[k + step] <= gi) k += step;
        int j = nb_start[k] + (gi - nb_off[k]);
        double x = a.r[3 * (size_t)j], y = a.r[3 * (size_t)j + 1], z = a.r[3 * (size_t)j + 2];
        xs[s] = x;
        ys[s] = y;
        zs[s] = z;
        double ux = x - ctrx, uy = y - ctry, uz = z - ctrz;  // nearest image of j relative to the cell centre
        if (LE) ux -= rint(uy) * p.strain;
        ux -= rint(ux);
        uy -= rint(uy);
        uz -= rint(uz);
        qs[s] = make_float4((float)(ux * scd), (float)(uy * scd), (float)(uz * scd), __int_as_float(j));
        if (HESS) {
          gxs[s] = a.fin[3 * (size_t)j];
          gys[s] = a.fin[3 * (size_t)j + 1];
          gzs[s] = a.fin[3 * (size_t)j + 2];
        }
      }
      __syncthreads();

      // ---- each warp takes i atoms of the cell in turn
      for (int ii = warp; ii < ni; ii += PW) {
        const int i = istart + ii;
        const double rix = a.r[3 * (size_t)i], riy = a.r[3 * (size_t)i + 1], riz = a.r[3 * (size_t)i + 2];
        const float qix = (float)((rix - ctrx) * scd), qiy = (float)((riy - ctry) * scd),
                    qiz = (float)((riz - ctrz) * scd);
        double fix = 0.0, fiy = 0.0, fiz = 0.0;
        if (HESS) {
          fix = a.fin[3 * (size_t)i];
          fiy = a.fin[3 * (size_t)i + 1];
          fiz = a.fin[3 * (size_t)i + 2];
        }
        double fx = 0.0, fy = 0.0, fz = 0.0;
        int qhead = 0, qn = 0;

        auto dense = [&](int sj, bool act) {
          double dx, dy, dz, s;
          bool in = pair_sep<LE>(rix, riy, riz, xs[sj], ys[sj], zs[sj], p.strain, p.r_cut_box_sq, dx, dy, dz, s);
          if (act && in) {
            if (HESS)
              pair_hess_term(p, dx, dy, dz, s, fix - gxs[sj], fiy - gys[sj], fiz - gzs[sj], acc);
            else
              pair_force_term<MODEL>(p, dx, dy, dz, s, fx, fy, fz, acc);
          }
        };

        for (int b = 0; b < mc; b += 32) {
          const int s = b + lane;
          bool pass = false;
          if (s < mc) {
            float4 c = qs[s];
            float ddx = qix - c.x, ddy = qiy - c.y, ddz = qiz - c.z;
            float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
            pass = (d2 < p.rc2_cell) && (__float_as_int(c.w) != i);
          }
          const unsigned m = __ballot_sync(FULL, pass);
          if (pass) q[(qhead + qn + __popc(m & lt_mask)) & 63] = s;
          qn += __popc(m);
          if (qn >= 32) {
            __syncwarp();
            dense(q[(qhead + lane) & 63], true);
            qhead = (qhead + 32) & 63;
            qn -= 32;
          }
        }
        if (qn > 0) {
          __syncwarp();
          const bool act = lane < qn;
          dense(act ? q[(qhead + lane) & 63] : 0, act);
        }
        __syncwarp();
        if (!HESS) {
          fx = warp_sum(fx);
          fy = warp_sum(fy);
          fz = warp_sum(fz);
          if (lane < 3) {
            double v = 24.0 * (lane == 0 ? fx : (lane == 1 ? fy : fz));  // md_lj_module.py:143
            double *dst = a.f + 3 * (size_t)i + lane;
            *dst = (c0 == 0) ? v : *dst + v;
          }
        }
      }
    }
  }
  block_store_partials<PW>(acc, a.partials + 8 * (size_t)blockIdx.x);
}

int launch_pair_cells(cudaStream_t st, const PairArgs &a, bool hessian, int natoms_hint) {
  (void)natoms_hint;
  int blocks = a.g.n_own_cells() < pair_grid_blocks() ? a.g.n_own_cells() : pair_grid_blocks();
  if (blocks < 1) blocks = 1;
  if (a.p.model == ATLJ_MODEL_LJ) {
    if (hessian)
      pair_cells_kernel<ATLJ_MODEL_LJ, true><<<blocks, 32 * PW, 0, st>>>(a);
    else
      pair_cells_kernel<ATLJ_MODEL_LJ, false><<<blocks, 32 * PW, 0, st>>>(a);
  } else {
    if (hessian) {
      set_error("hessian is defined for the LJ model only (md_lj_le_module.py has none)");
      return ATLJ_ERR_ARG;
    }
    pair_cells_kernel<ATLJ_MODEL_WCA, false><<<blocks, 32 * PW, 0, st>>>(a);
  }
  ATLJ_LAUNCHED();
  return blocks;
}

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

// ---- fixed-order reduction of the CTA partials + the reference's final scaling ------------------------------------
__global__ void __launch_bounds__(256) finalize_kernel(const double *__restrict__ partials, int nblocks, int model,
                                                       int hessian, double *__restrict__ rec,
                                                       int *__restrict__ reset_counter) {
  if (threadIdx.x == 0 && reset_counter) *reset_counter = 0;
  __shared__ double sm[256][7];
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += 256) {
#pragma unroll
    for (int k = 0; k < 7; k++) v[k] += partials[8 * (size_t)b + k];
  }
#pragma unroll
  for (int k = 0; k < 7; k++) sm[threadIdx.x][k] = v[k];
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int k = 0; k < 7; k++) sm[threadIdx.x][k] += sm[threadIdx.x + o][k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // every pair was visited from both ends: halve (exact), then scale as md_lj_module.py:143-147
    double a = sm[0][0] * 0.5, pot = sm[0][1] * 0.5, vir = sm[0][2] * 0.5, lap = sm[0][3] * 0.5;
    if (hessian) {
      rec[6] = sm[0][4] * 0.5;
    } else {
      if (model == ATLJ_MODEL_LJ) {
        rec[0] = a * 4.0;            // cut
        rec[1] = pot * 4.0;          // pot
        rec[2] = vir * 24.0 / 3.0;   // vir
        rec[3] = lap * 24.0 * 2.0;   // lap
      } else {
        rec[0] = pot * 4.0;          // md_lj_le_module.py:144-148
        rec[1] = vir * 24.0 / 3.0;
        rec[2] = lap * 24.0 * 2.0;
        rec[3] = a * 24.0;           // pyx
      }
      rec[7] = sm[0][6] > 0.0 ? 1.0 : 0.0;  // ovr
      rec[8] = sm[0][5] * 0.5;              // in-range pairs, counted once
    }
  }
}

int launch_finalize(cudaStream_t st, const double *partials, int nblocks, int model, bool hessian, double *rec,
                    int *reset_counter) {
  finalize_kernel<<<1, 256, 0, st>>>(partials, nblocks, model, hessian ? 1 : 0, rec, reset_counter);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

}  // namespace atlj
