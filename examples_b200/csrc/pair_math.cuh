// Pair arithmetic shared by the pair kernels (internal).
//
// Reference: python_examples/md_lj_module.py:96-115 (== md_lj_ll_module.py:134-165), md_lj_le_module.py:93-112,
// md_lj_module.py:176-191 (hessian).  pair_sep reproduces the reference's operation order exactly (SURVEY.md
// Appendix A): every add / multiply rounded separately (no FMA contraction), rint = round-half-even.
#pragma once
#include "common.cuh"

namespace atlj {

constexpr unsigned FULL = 0xffffffffu;

struct Acc {
  double a, pot, vir, lap, hes;
  unsigned long long np;
  int ovr;
};

// separation with the reference's exact operation order; returns the in-range predicate
template <bool LE>
__device__ __forceinline__ bool pair_sep(double rix, double riy, double riz, double xj, double yj, double zj,
                                         double strain, double rcbsq, double &dx, double &dy, double &dz, double &s) {
  dx = __dsub_rn(rix, xj);
  dy = __dsub_rn(riy, yj);
  dz = __dsub_rn(riz, zj);
  if (LE) dx = __dsub_rn(dx, __dmul_rn(rint(dy), strain));  // md_lj_le_module.py:94
  dx = __dsub_rn(dx, rint(dx));
  dy = __dsub_rn(dy, rint(dy));
  dz = __dsub_rn(dz, rint(dz));
  s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));  // ((x^2+y^2)+z^2), no FMA
  return s < rcbsq;
}

template <int MODEL>
__device__ __forceinline__ void pair_force_term(const Phys &p, double dx, double dy, double dz, double s, double &fx,
                                                double &fy, double &fz, Acc &acc) {
  double s2 = __dmul_rn(s, p.box_sq);
  dx *= p.box;
  dy *= p.box;
  dz *= p.box;
  double sr2 = __drcp_rn(s2);  // == 1.0/s2, correctly rounded (feeds the overlap predicate)
  acc.ovr |= (sr2 > 1.77) ? 1 : 0;
  double sr6 = sr2 * sr2 * sr2;
  double sr12 = sr6 * sr6;
  double cut = sr12 - sr6;
  double vir = cut + sr12;
  acc.lap += (22.0 * sr12 - 5.0 * sr6) * sr2;
  double fs = vir * sr2;
  double fxi = dx * fs;
  fx += fxi;
  fy += dy * fs;
  fz += dz * fs;
  acc.vir += vir;
  if (MODEL == ATLJ_MODEL_LJ) {
    acc.pot += cut - p.pot_cut;
    acc.a += cut;
  } else {
    acc.pot += cut + 0.25;  // md_lj_le_module.py:106
    acc.a += dy * fxi;      // pyx, md_lj_le_module.py:110
  }
  acc.np++;
}

// md_lj_module.py:176-191
__device__ __forceinline__ void pair_hess_term(const Phys &p, double dx, double dy, double dz, double s, double gx,
                                               double gy, double gz, Acc &acc) {
  double s2 = __dmul_rn(s, p.box_sq);
  dx *= p.box;
  dy *= p.box;
  dz *= p.box;
  double ff = gx * gx + gy * gy + gz * gz;
  double rf = dx * gx + dy * gy + dz * gz;
  double sr2 = __drcp_rn(s2);
  double sr6 = sr2 * sr2 * sr2;
  double sr8 = sr6 * sr2;
  double sr10 = sr8 * sr2;
  double v1 = 24.0 * (1.0 - 2.0 * sr6) * sr8;
  double v2 = 96.0 * (7.0 * sr6 - 2.0) * sr10;
  acc.hes += v1 * ff + v2 * (rf * rf);
  acc.np++;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// CTA-level reduction of the accumulators into one row of partials (fixed order: lanes by butterfly, warps 0..NW-1)
template <int NW>
__device__ __forceinline__ void block_store_partials(Acc acc, double *__restrict__ row) {
  __shared__ double red[NW][8];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[7] = {acc.a, acc.pot, acc.vir, acc.lap, acc.hes, (double)acc.np, (double)acc.ovr};
#pragma unroll
  for (int k = 0; k < 7; k++) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 7; k++) red[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double s = 0.0;
    for (int w = 0; w < NW; w++) s += red[w][threadIdx.x];
    row[threadIdx.x] = s;
  }
  if (threadIdx.x == 7) row[7] = 0.0;
}


// Final halving + scaling of the summed partials t = {a, pot, vir, lap, hes, npairs, ovr} into the step record
// (every pair was visited from both ends: halve, exactly; then md_lj_module.py:143-147 / md_lj_le_module.py:144-148)
__device__ __forceinline__ void finalize_totals(const double *t, int model, bool hessian, double *__restrict__ rec) {
  const double a = t[0] * 0.5, pot = t[1] * 0.5, vir = t[2] * 0.5, lap = t[3] * 0.5;
  if (hessian) {
    rec[6] = t[4] * 0.5;
    return;
  }
  if (model == ATLJ_MODEL_LJ) {
    rec[0] = a * 4.0;            // cut
    rec[1] = pot * 4.0;          // pot
    rec[2] = vir * 24.0 / 3.0;   // vir
    rec[3] = lap * 24.0 * 2.0;   // lap
  } else {
    rec[0] = pot * 4.0;
    rec[1] = vir * 24.0 / 3.0;
    rec[2] = lap * 24.0 * 2.0;
    rec[3] = a * 24.0;           // pyx
  }
  rec[7] = t[6] > 0.0 ? 1.0 : 0.0;  // ovr
  rec[8] = t[5] * 0.5;              // in-range pairs, counted once
}

// Fixed-order reduction of the pair kernels' per-CTA rows [nblocks][8] into the step record by ONE CTA of BT threads
// (md_lj_module.py:143-147 scaling included).  finalize_kernel is this and nothing else; the loops that launch an
// elementwise kernel right after the force evaluation (BAOAB's half kick, SLLOD's sums) run it as an extra CTA of that
// kernel instead (FinTail): 9 us with one SM working otherwise.
struct FinTail {
  const double *partials = nullptr;
  int nb = 0;                   // 0: nothing pending
  int model = 0;
  double *rec = nullptr;
  int *reset_counter = nullptr;
  const int *rows_dev = nullptr;
};
template <int BT>
__device__ __forceinline__ void finalize_rows_block(const FinTail &t) {
  int nblocks = t.nb;
  if (t.rows_dev) nblocks = min(nblocks, *t.rows_dev);
  if (threadIdx.x == 0 && t.reset_counter) *t.reset_counter = 0;
  __shared__ double fsm[BT / 32][7];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += BT) {
#pragma unroll
    for (int k = 0; k < 7; k++) v[k] += t.partials[8 * (size_t)b + k];
  }
#pragma unroll
  for (int k = 0; k < 7; k++) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 7; k++) fsm[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tt[7];
    for (int k = 0; k < 7; k++) {
      double a = 0.0;
      for (int w = 0; w < BT / 32; w++) a += fsm[w][k];
      tt[k] = a;
    }
    finalize_totals(tt, t.model, false, t.rec);
  }
}

// 1/x to ~1 ulp: MUFU.RCP64H seed (>= 20 bits) + two Newton steps; used only where no predicate depends on it
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

}  // namespace atlj
