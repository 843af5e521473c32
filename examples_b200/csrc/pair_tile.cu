// Tiled pair-force / hessian kernel over the link cells: the hot kernel of the device-resident loop
// (SURVEY.md rows a7-a9, a11).
//
// Reference arithmetic: python_examples/md_lj_ll_module.py:134-165 (pair block), md_lj_llle_module.py:144-177
// (WCA + Lees-Edwards), md_lj_ll_module.py:333-355 (hessian); scaling md_lj_module.py:143-147.
//
// Design (B200-first, not the reference's cell-pair blocks):
//   * FULL stencil, one THREAD per i atom: no atomics, no cross-lane reduction of forces, deterministic sums.
//   * Cells of one (x,y) row are contiguous in the cell-sorted arrays, so a CTA walks one row in chunks of whole
//     cells holding <= NT atoms and stages the 3 x 3 (LE: up to 10) neighbouring row segments into shared memory
//     with coalesced loads: fp64 coordinates relative to the chunk centre in sigma units (the image shift of the
//     periodic / sheared boundary is applied once here, per candidate, not per pair) and an fp32 copy
//     (x, y, z, |q|^2).
//   * Phase 1 (FP32 pipe): every thread tests the ~330 candidates of its own 3-cell window in each row with the
//     expanded form |qj|^2 - 2 qi.qj < thr - |qi|^2 : one broadcast LDS.128 + 3 FFMA + 1 compare per candidate;
//     the outcome is shifted into per-window bit masks.  The threshold carries a rigorous rounding margin, so the
//     filter never drops a pair that is in range.
//   * Phase 2 (FP64 pipe): each thread walks the set bits of its masks (~49 per atom in the rho = 0.75 liquid) and
//     evaluates those pairs in fp64 -- every lane busy with a pair that is (almost surely) in range.
//   * Exactness: the in-range and overlap decisions are the reference's.  The fast fp64 separation differs from
//     the reference's operation order by <= 1e-13 relative, so a pair is decided by the fast value only when it is
//     further than 1e-8 (relative) from the threshold; the few pairs inside that band, and every pair close to
//     the overlap threshold, are re-evaluated with the reference's exact operation sequence (pair_sep: separate
//     roundings, rint = half-even, no FMA contraction, IEEE reciprocal).  In-range pair counts are therefore
//     bit-exact; energies / forces agree to summation order (<= 1e-10 relative, tests/test_force_parity.py).
#include "common.cuh"
#include "pair_math.cuh"

namespace atlj {

constexpr int NT = 128;          // threads per CTA = i atoms per batch
constexpr int TCAP = 1536;       // staged candidates per chunk
constexpr int MAXC = 64;         // logical cells per chunk, including one halo cell at each end
constexpr int NROWMAX = 10;      // neighbour rows: 9 (LJ) or up to 10 (Lees-Edwards wide row)
constexpr int SOFF_STRIDE = MAXC + 1;

static inline size_t tile_smem_bytes(bool hess) {
  size_t b = sizeof(double) * 3 * TCAP + sizeof(float4) * TCAP;
  if (hess) b += sizeof(double) * 3 * TCAP;
  b += sizeof(int) * NROWMAX * SOFF_STRIDE;
  return b;
}

// parts per row so that small systems still fill the machine
static inline int tile_parts(const Geom &g) {
  int rows = g.nx_own * g.sc;
  int want = (148 * 6 + rows - 1) / rows;
  int maxp = g.sc / 4;
  if (maxp < 1) maxp = 1;
  return want < 1 ? 1 : (want > maxp ? maxp : want);
}
int pair_tile_grid(const Geom &g) { return g.nx_own * g.sc * tile_parts(g); }

template <int MODEL, bool HESS>
__global__ void __launch_bounds__(NT) pair_tile_kernel(PairArgs a, int parts) {
  constexpr bool LE = (MODEL == ATLJ_MODEL_WCA);
  constexpr int NR = LE ? 10 : 9;
  constexpr int ROWN = LE ? 1 : 4;  // slot of the i row itself (dx = dy = 0) among the neighbour rows
  constexpr int NW = 2 * NR;        // mask words per thread per pass (64 candidates per window)

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *ux = reinterpret_cast<double *>(smem_raw);
  double *uy = ux + TCAP;
  double *uz = uy + TCAP;
  float4 *q = reinterpret_cast<float4 *>(uz + TCAP);
  double *gx = reinterpret_cast<double *>(q + TCAP);
  double *gy = gx + (HESS ? TCAP : 0);
  double *gz = gy + (HESS ? TCAP : 0);
  int *soff = reinterpret_cast<int *>(gz + (HESS ? TCAP : 0));
  __shared__ int rowcell0[NROWMAX];
  __shared__ int pst[3 * NROWMAX + 2];
  __shared__ int pgl[3 * NROWMAX];
  __shared__ unsigned s_q2max;
  __shared__ int s_maxwin;

  const Geom g = a.g;
  const Phys p = a.p;
  const int sc = g.sc, tid = threadIdx.x, lane = tid & 31;
  const int rowid = blockIdx.x / parts, part = blockIdx.x - rowid * parts;
  const int lx = g.h + rowid / sc, cy = rowid % sc;
  const int own0 = g.cs_idx((lx * sc + cy) * sc);  // index of the row's first cell in the padded cell_start
  const int zA = (int)((long long)sc * part / parts), zB = (int)((long long)sc * (part + 1) / parts);
  const int *__restrict__ cs = a.cell_start;
  const double box = p.box;
  const double rc2s = p.r_cut_box_sq * p.box_sq;        // r_cut^2 in sigma^2 (classification only)
  const double band = 1e-8 * rc2s;
  const double sovr = (1.0 / 1.77) * (1.0 + 1e-6);      // below this s the overlap decision goes the exact way
  const double sci = 1.0 / (double)sc;
  const double ctrx = ((lx - g.h + g.x_lo) + 0.5) * sci - 0.5;
  const double ctry = (cy + 0.5) * sci - 0.5;

  if (tid < NROWMAX) {
    const int c = tid < NR ? neighbour_cell(g, LE, p.shift, lx, cy, 0, 3 * tid + 1) : -1;
    rowcell0[tid] = c < 0 ? -1 : g.cs_idx(c);
  }
  if (tid == 0) s_q2max = 0u, s_maxwin = 0;

  Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0ull, 0};
  double cutacc = 0.0;
  unsigned long long npf = 0ull;  // pairs taken by the fast path (their pot shift is applied at the end)

  for (int z = zA; z < zB;) {
    // ---- chunk extent: as many whole cells as hold <= NT atoms (at least one) ----------------------------------
    const int s0 = cs[own0 + z];
    int ok = 0;
    if (tid < MAXC - 2 && z + tid < zB) ok = (cs[own0 + z + tid + 1] - s0) <= NT;
    int ncz = __syncthreads_count(ok);  // barrier: the previous chunk is done with shared memory
    if (ncz == 0) ncz = 1;
    const int ni = cs[own0 + z + ncz] - s0;
    if (ni == 0) {
      z += ncz;
      continue;
    }
    const double ctrz = (z + 0.5 * ncz) * sci - 0.5;

    // ---- piece table: per row the cell below, the chunk's cells, the cell above --------------------------------
    if (tid < 32) {
      int len = 0, gl = 0;
      if (tid < 3 * NR) {
        const int r = tid / 3, pc = tid - 3 * r, rc0 = rowcell0[r];
        if (rc0 >= 0) {
          int c0, c1;
          if (pc == 0)
            c0 = (z == 0 ? sc - 1 : z - 1), c1 = c0 + 1;
          else if (pc == 1)
            c0 = z, c1 = z + ncz;
          else
            c0 = (z + ncz == sc ? 0 : z + ncz), c1 = c0 + 1;
          gl = cs[rc0 + c0];
          len = cs[rc0 + c1] - gl;
        }
        pgl[tid] = gl;
      }
      int incl = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (tid <= 3 * NR) pst[tid] = incl - len;
    }
    __syncthreads();
    const int total = pst[3 * NR];

    if (total > TCAP) {
      // ---- crowded neighbourhood (does not occur in a fluid): exact arithmetic straight from global memory ------
      for (int il = tid; il < ni; il += NT) {
        const int i = s0 + il;
        int c = 0;
        for (int step = 32; step >= 1; step >>= 1)
          if (c + step < ncz && cs[own0 + z + c + step] <= i) c += step;
        const int zi = z + c;
        const double rix = a.r[3 * (size_t)i], riy = a.r[3 * (size_t)i + 1], riz = a.r[3 * (size_t)i + 2];
        double fix = 0, fiy = 0, fiz = 0, fx = 0, fy = 0, fz = 0;
        if (HESS) fix = a.fin[3 * (size_t)i], fiy = a.fin[3 * (size_t)i + 1], fiz = a.fin[3 * (size_t)i + 2];
        for (int r = 0; r < NR; r++) {
          const int rc0 = rowcell0[r];
          if (rc0 < 0) continue;
          for (int dz = -1; dz <= 1; dz++) {
            int zj = zi + dz;
            zj = zj < 0 ? zj + sc : (zj >= sc ? zj - sc : zj);
            for (int j = cs[rc0 + zj]; j < cs[rc0 + zj + 1]; j++) {
              if (j == i) continue;
              double dx, dy, dzz, s;
              if (pair_sep<LE>(rix, riy, riz, a.r[3 * (size_t)j], a.r[3 * (size_t)j + 1], a.r[3 * (size_t)j + 2],
                               p.strain, p.r_cut_box_sq, dx, dy, dzz, s)) {
                if (HESS)
                  pair_hess_term(p, dx, dy, dzz, s, fix - a.fin[3 * (size_t)j], fiy - a.fin[3 * (size_t)j + 1],
                                 fiz - a.fin[3 * (size_t)j + 2], acc);
                else
                  pair_force_term<MODEL>(p, dx, dy, dzz, s, fx, fy, fz, acc);
              }
            }
          }
        }
        if (!HESS) {
          a.f[3 * (size_t)i] = 24.0 * fx;
          a.f[3 * (size_t)i + 1] = 24.0 * fy;
          a.f[3 * (size_t)i + 2] = 24.0 * fz;
        }
      }
      z += ncz;
      continue;
    }

    // ---- window table: soff[r][c] = staged offset of logical cell c (c = 0 is the cell below the chunk) ----------
    {
      int mw = 0;
      const int per = ncz + 3;
      for (int e = tid; e < NR * per; e += NT) {
        const int r = e / per, c = e - r * per, rc0 = rowcell0[r];
        int v;
        if (c == 0 || rc0 < 0)
          v = pst[3 * r];
        else if (c <= ncz + 1)
          v = pst[3 * r + 1] + (cs[rc0 + z + c - 1] - cs[rc0 + z]);
        else
          v = pst[3 * r + 3];
        soff[r * SOFF_STRIDE + c] = v;
      }
      // ---- stage the candidates ---------------------------------------------------------------------------------
      float q2m = 0.f;
      for (int s = tid; s < total; s += NT) {
        int t = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1)
          if (t + step < 3 * NR && pst[t + step] <= s) t += step;
        const int j = pgl[t] + (s - pst[t]);
        const int pc = t % 3;
        const double wz = (pc == 0 && z == 0) ? -1.0 : ((pc == 2 && z + ncz == sc) ? 1.0 : 0.0);
        double x = a.r[3 * (size_t)j] - ctrx, y = a.r[3 * (size_t)j + 1] - ctry;
        double zz = (a.r[3 * (size_t)j + 2] + wz) - ctrz;
        if (LE) x -= rint(y) * p.strain;
        x -= rint(x);
        y -= rint(y);
        x *= box, y *= box, zz *= box;
        ux[s] = x, uy[s] = y, uz[s] = zz;
        const float fx = (float)x, fy = (float)y, fz = (float)zz;
        const float n2 = fx * fx + fy * fy + fz * fz;
        q[s] = make_float4(fx, fy, fz, n2);
        q2m = fmaxf(q2m, n2);
        if (HESS) gx[s] = a.fin[3 * (size_t)j], gy[s] = a.fin[3 * (size_t)j + 1], gz[s] = a.fin[3 * (size_t)j + 2];
      }
      unsigned qb = __reduce_max_sync(0xffffffffu, __float_as_uint(q2m));
      if (lane == 0) atomicMax(&s_q2max, qb);
      __syncthreads();
      // longest window of the chunk -> number of 64-candidate passes
      for (int e = tid; e < NR * ncz; e += NT) {
        const int r = e / ncz, c = e - r * ncz;
        mw = max(mw, soff[r * SOFF_STRIDE + c + 3] - soff[r * SOFF_STRIDE + c]);
      }
      mw = __reduce_max_sync(0xffffffffu, mw);
      if (lane == 0) atomicMax(&s_maxwin, mw);
      __syncthreads();
    }
    const int npass = (s_maxwin + 63) >> 6;
    // filter threshold: |t_fp32 - t_exact| <= ~1.5e-6 * Q^2 (three fmaf roundings of terms <= 3 Q^2, the fp32
    // rounding of the coordinates and of the two norms); 4e-6 Q^2 is used
    const float q2max = __uint_as_float(s_q2max);
    const float thr = (float)(rc2s * (1.0 + 1e-6) + 4e-6 * (double)q2max) * (1.0f + 2e-7f);
    const int *soffn = soff + ROWN * SOFF_STRIDE;

    for (int ib = 0; ib < ni; ib += NT) {
      const int il = ib + tid;
      const bool act = il < ni;
      const int i = s0 + il;
      const int self = act ? soffn[1] + il : 0;
      int myc = 0;
      if (act) {
#pragma unroll
        for (int step = 32; step >= 1; step >>= 1)
          if (myc + step < ncz && soffn[1 + myc + step] <= self) myc += step;
      }
      const double uix = ux[self], uiy = uy[self], uiz = uz[self];
      const float4 qi = q[self];
      const float m2x = -2.f * qi.x, m2y = -2.f * qi.y, m2z = -2.f * qi.z, ti = thr - qi.w;
      double fix = 0.0, fiy = 0.0, fiz = 0.0;
      if (HESS) fix = gx[self], fiy = gy[self], fiz = gz[self];
      double fx = 0.0, fy = 0.0, fz = 0.0;

      for (int pass = 0; pass < npass; pass++) {
        unsigned m[NW];
        int top[NW];
        // ---- phase 1: fp32 filter into bit masks ---------------------------------------------------------------
#pragma unroll
        for (int r = 0; r < NR; r++) {
          int b = 0, e = 0;
          if (act) {
            b = soff[r * SOFF_STRIDE + myc] + 64 * pass;
            e = min(soff[r * SOFF_STRIDE + myc + 3], b + 64);
          }
          const int e0 = min(e, b + 32);
          unsigned w0 = 0u, w1 = 0u;
          for (int j = b; j < e0; j++) {
            const float4 c = q[j];
            const float t = fmaf(m2x, c.x, fmaf(m2y, c.y, fmaf(m2z, c.z, c.w)));
            w0 = w0 + w0 + (t < ti ? 1u : 0u);
          }
          for (int j = e0; j < e; j++) {
            const float4 c = q[j];
            const float t = fmaf(m2x, c.x, fmaf(m2y, c.y, fmaf(m2z, c.z, c.w)));
            w1 = w1 + w1 + (t < ti ? 1u : 0u);
          }
          if (r == ROWN) {  // never pair an atom with itself
            if (self >= b && self < e0)
              w0 &= ~(1u << (e0 - 1 - self));
            else if (self >= e0 && self < e)
              w1 &= ~(1u << (e - 1 - self));
          }
          m[2 * r] = w0, top[2 * r] = e0 - 1;
          m[2 * r + 1] = w1, top[2 * r + 1] = e - 1;
        }
        // ---- phase 2: fp64 evaluation of the surviving pairs ---------------------------------------------------
        int ww = 0;
        unsigned cm = m[0];
        int jt = top[0];
        for (;;) {
          while (cm == 0u && ++ww < NW) cm = m[ww], jt = top[ww];
          if (cm == 0u) break;
          const int b = 31 - __clz(cm);
          cm ^= 1u << b;
          const int j = jt - b;
          double dx = uix - ux[j], dy = uiy - uy[j], dz = uiz - uz[j];
          double s = fma(dz, dz, fma(dy, dy, dx * dx));
          bool in = s < rc2s;
          if (fabs(s - rc2s) <= band || s < sovr) {
            // borderline: decide with the reference's exact operation sequence on the original coordinates
            int t = 0;
            while (t + 1 < 3 * NR && pst[t + 1] <= j) t++;
            const size_t jg = (size_t)(pgl[t] + (j - pst[t]));
            double ex, ey, ez, es;
            in = pair_sep<LE>(a.r[3 * (size_t)i], a.r[3 * (size_t)i + 1], a.r[3 * (size_t)i + 2], a.r[3 * jg],
                              a.r[3 * jg + 1], a.r[3 * jg + 2], p.strain, p.r_cut_box_sq, ex, ey, ez, es);
            const double s2 = __dmul_rn(es, p.box_sq);
            if (in && __drcp_rn(s2) > 1.77) acc.ovr = 1;  // md_lj_module.py:103
            dx = ex * box, dy = ey * box, dz = ez * box, s = s2;
          }
          if (in) {
            const double sr2 = rcp_fast(s);
            const double sr6 = sr2 * sr2 * sr2;
            if (HESS) {
              const double hx = fix - gx[j], hy = fiy - gy[j], hz = fiz - gz[j];
              const double ff = fma(hz, hz, fma(hy, hy, hx * hx));
              const double rf = fma(dz, hz, fma(dy, hy, dx * hx));
              const double sr8 = sr6 * sr2, sr10 = sr8 * sr2;
              const double v1 = 24.0 * (1.0 - 2.0 * sr6) * sr8;   // md_lj_module.py:184-185
              const double v2 = 96.0 * (7.0 * sr6 - 2.0) * sr10;
              acc.hes += v1 * ff + v2 * (rf * rf);
            } else {
              const double sr12 = sr6 * sr6;
              const double cut = sr12 - sr6;
              const double vir = cut + sr12;
              acc.lap = fma(fma(22.0, sr6, -5.0) * sr6, sr2, acc.lap);  // (22 sr12 - 5 sr6) sr2
              const double fs = vir * sr2;
              const double fxi = dx * fs;
              fx += fxi;
              fy = fma(dy, fs, fy);
              fz = fma(dz, fs, fz);
              acc.vir += vir;
              cutacc += cut;
              if (LE) acc.a = fma(dy, fxi, acc.a);  // pyx, md_lj_le_module.py:110
            }
            npf++;
          }
        }
      }
      if (!HESS && act) {
        a.f[3 * (size_t)i] = 24.0 * fx;  // md_lj_module.py:143
        a.f[3 * (size_t)i + 1] = 24.0 * fy;
        a.f[3 * (size_t)i + 2] = 24.0 * fz;
      }
    }
    __syncthreads();
    if (tid == 0) s_q2max = 0u, s_maxwin = 0;
    z += ncz;
  }
  if (!HESS) {
    // pot = sum(cut - pot_cut) (md_lj_module.py:107) / sum(cut + 0.25) (md_lj_le_module.py:106) over this thread's pairs
    if (LE) {
      acc.pot += cutacc + 0.25 * (double)npf;
    } else {
      acc.pot += cutacc - p.pot_cut * (double)npf;
      acc.a += cutacc;
    }
  }
  acc.np += npf;
  block_store_partials<NT / 32>(acc, a.partials + 8 * (size_t)blockIdx.x);
}

template <int MODEL, bool HESS>
static int launch_one(cudaStream_t st, const PairArgs &a, int parts, int grid) {
  static bool configured = false;
  size_t smem = tile_smem_bytes(HESS);
  if (!configured) {
    ATLJ_CUDA(cudaFuncSetAttribute(pair_tile_kernel<MODEL, HESS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    configured = true;
  }
  pair_tile_kernel<MODEL, HESS><<<grid, NT, smem, st>>>(a, parts);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

// returns the number of CTAs (= rows of partials written) or a negative status
int launch_pair_tile(cudaStream_t st, const PairArgs &a, bool hessian) {
  const int parts = tile_parts(a.g);
  const int grid = a.g.nx_own * a.g.sc * parts;
  int rc;
  if (a.p.model == ATLJ_MODEL_LJ) {
    rc = hessian ? launch_one<ATLJ_MODEL_LJ, true>(st, a, parts, grid)
                 : launch_one<ATLJ_MODEL_LJ, false>(st, a, parts, grid);
  } else {
    if (hessian) {
      set_error("hessian is defined for the LJ model only (md_lj_le_module.py has none)");
      return ATLJ_ERR_ARG;
    }
    rc = launch_one<ATLJ_MODEL_WCA, false>(st, a, parts, grid);
  }
  return rc ? rc : grid;
}

}  // namespace atlj
