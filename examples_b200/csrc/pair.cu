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

// ---- fixed-order reduction of the CTA partials + the reference's final scaling ------------------------------------
constexpr int FIN_T = 1024;
__global__ void __launch_bounds__(FIN_T) finalize_kernel(const double *__restrict__ partials, int nblocks, int model,
                                                         int hessian, double *__restrict__ rec,
                                                         int *__restrict__ reset_counter,
                                                         const double *__restrict__ vf_partials, int nvf,
                                                         const int *__restrict__ rows_dev) {
  if (rows_dev) nblocks = min(nblocks, *rows_dev);
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
                    int *reset_counter, const double *vf_partials, int nvf, const int *rows_dev) {
  finalize_kernel<<<1, FIN_T, 0, st>>>(partials, nblocks, model, hessian ? 1 : 0, rec, reset_counter, vf_partials, nvf,
                                       rows_dev);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

}  // namespace atlj
