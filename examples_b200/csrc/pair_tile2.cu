// Second-generation tiled LJ pair-force kernel: the hot kernel of the device-resident loop.
//
// Reference arithmetic: python_examples/md_lj_ll_module.py:134-165; scaling md_lj_module.py:143-147.
// Exactness scheme: a cheap filter that never drops an in-range pair -> fp64 evaluation of the survivors, which decides;
// the reference's exact operation sequence for every pair within 1e-8 of the cutoff or near the overlap threshold.
// Structure, each step chosen from an ncu profile (profiles/r1_pair_kernel_history.md):
//   * a CTA owns TWO adjacent y rows of cells, so the staged cross-section is 3 x 4 = 12 rows for two i rows
//     (6 per i row instead of 9): a third less shared memory per atom and a third less staging work;
//   * staged candidates are laid out CELL-major (for each z cell: 4 y-offsets x 3 x-offsets), so the candidates of
//     one thread are ONE contiguous run per z cell (3 runs, ~111 candidates each) instead of 9 short windows;
//   * the filter is d^2 = dx^2 + dy^2 + dz^2 in packed fp16 (HADD2/HFMA2) on fp16 copies of the chunk-centred
//     coordinates: one LDS.128 + 6 half2 operations + 3 shifts per TWO candidates, the sign bit of d^2 - thr shifted
//     straight into left-aligned 32-bit mask words; thr carries a proven bound of the fp16 error (the kernel is bound
//     by the shared-memory pipe: the fp32 expanded form it replaced needed two LDS.128 per two candidates);
//   * 43 KB of shared memory and 96 registers per thread: 5 CTAs (20 warps) per SM;
//   * work items (chunks of <= 128 atoms of a row pair, not cell-aligned) are handed out through an atomic counter so
//     the tail of the grid stays busy; every item writes its own row of partial sums, so totals stay bitwise
//     reproducible.
#include <cuda_fp16.h>

#include "common.cuh"
#include "pair_math.cuh"

namespace atlj {

// Tunables (A/B builds: make variant TAG=... XFLAGS=-DATLJ_U_...; measured values in profiles/r*_pair_kernel_history.md)
#ifndef ATLJ_U_NT
#define ATLJ_U_NT 128
#endif
constexpr int U_NT = ATLJ_U_NT;         // threads per CTA = i atoms per batch
#ifndef ATLJ_U_CTAS
#define ATLJ_U_CTAS 5
#endif
constexpr int U_CTAS = U_NT == 128 ? ATLJ_U_CTAS : 2;      // resident CTAs per SM (registers: 128 per thread)
#ifndef ATLJ_U_CAP
#define ATLJ_U_CAP 1280
#endif
constexpr int U_CAP = U_NT == 128 ? ATLJ_U_CAP : 2304;           // staged candidates per chunk (nine cell planes of a liquid)
constexpr int U_MAXC = U_NT == 128 ? 24 : 32;              // logical cells per chunk, incl. one halo cell at each end
constexpr int U_K = 12;                 // staged rows per z cell: 4 y-offsets (outer) x 3 x-offsets (inner)
constexpr int U_NE = U_MAXC * U_K;      // (cell,row) entries per chunk
constexpr int U_WW = 6;                 // mask words per window (<= 191 candidates per window)
constexpr int U_NWORDS = 3 * U_WW;
// stride of the staged coordinate arrays: slot U_CAP holds a dummy candidate far away (1e150 in every coordinate) that
// the unused slot of a phase-2 iteration reads: its separation (3e300) is classified out of range by the same tests as
// a real pair and contributes exactly 0 (sr^6 underflows), so a dead slot needs no instructions of its own
constexpr int U_STR = U_CAP + 2;
constexpr unsigned U_YOFF = 8u * 31u;  // mask words carry the address of the candidate of bit 0 (walked with bfind)

// Kernel modes.  T2_FORCE: forces + totals.  T2_FORCE_SAVE: the same launch with a mask buffer - every atom's mask
// words (the filter's survivors) are kept in global memory.  T2_HESS: the hessian sum (md_lj_module.py:176-191) from those mask words - no
// filter pass; the chunk boundaries and staged indices are the same because they depend on the cell table only, so a
// mask word means the same candidates in both kernels.  The forces of the candidates take the place of the fp16 copy in
// shared memory (48 B per candidate instead of 32: 3 CTAs per SM).
enum { T2_FORCE = 0, T2_FORCE_SAVE = 1, T2_HESS = 2 };

static inline size_t tile2_smem_bytes(int mode) {
  const size_t tab = sizeof(int) * (U_NE + 4) + sizeof(int) * U_NE;
  if (mode == T2_HESS) return sizeof(double) * 6 * U_STR + tab;
  return sizeof(double) * 3 * U_STR + sizeof(uint4) * (U_CAP / 2 + 1) + tab;
}
size_t pair_tile2_mask_bytes(long long natoms) { return sizeof(uint2) * (size_t)U_NWORDS * (size_t)natoms; }

#ifndef ATLJ_U_SE
#define ATLJ_U_SE 3
#endif
#ifndef ATLJ_U_IPC
#define ATLJ_U_IPC 8  // work items per resident CTA the item size aims at
#endif
#ifndef ATLJ_U_P1U
#define ATLJ_U_P1U 16
#endif
constexpr int U_P1U = ATLJ_U_P1U;  // unroll of the filter loop
constexpr int U_GK_MAX = 4;  // chunks per work item: 1 (small systems, keeps every SM busy) .. 4

// Work items.  The atoms of a row pair (two adjacent y rows of one x layer) are taken in "interleaved" order: for each
// z cell, its row-A atoms then its row-B atoms.  A chunk is the next <= U_NT atoms of that order, cut short only where
// it would span more than ncz_max cells (the staged candidate set must fit U_CAP) - chunks are NOT cell-aligned, so
// every chunk but the last of a row pair is (nearly) full: 99 % of the lanes carry an atom instead of 89.5 % with
// whole-cell chunks.  gk consecutive chunks are one work item; tile2_table_kernel replays the cutting rule and records
// where every item starts (tab layout: [0] = largest number of items of any row pair, [16 .. 16+nrp) = items per row
// pair, then [nrp][gstride] start positions in interleaved order, closed by the row pair's atom count).
// Item t = (group t / nrp, row pair t % nrp).
static inline int tile2_gstride(const Geom &g) { return g.sc + 2; }
int pair_tile2_tab_ints(const Geom &g) {
  const int nrp = g.nx_own * ((g.sc + 1) / 2);
  return 16 + nrp + 2 * nrp * tile2_gstride(g);  // start positions, then the z cell each item starts in
}
// upper bound of the partial rows a launch over this geometry writes
int pair_tile2_items(const Geom &g) { return g.nx_own * ((g.sc + 1) / 2) * (tile2_gstride(g) - 1); }

__global__ void __launch_bounds__(128) tile2_table_kernel(Geom g, const int *__restrict__ cs, int nyb, int nrp,
                                                          int gstride, int gk, int ncz_max, int catoms,
                                                          int *__restrict__ tab) {
  __shared__ int cnt[4][1024];
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int rp = blockIdx.x * 4 + w;
  if (rp >= nrp) return;
  const int sc = g.sc;
  const int xo = rp / nyb, yb = rp - xo * nyb;
  const int lx = g.h + xo, y0 = 2 * yb;
  const bool two = (y0 + 1 < sc);
  const int ownA0 = g.cs_idx((lx * sc + y0) * sc), ownB0 = ownA0 + sc;
  for (int c = lane; c < sc; c += 32)
    cnt[w][c] = (cs[ownA0 + c + 1] - cs[ownA0 + c]) + (two ? cs[ownB0 + c + 1] - cs[ownB0 + c] : 0);
  __syncwarp();
  if (lane != 0) return;
  int *zb = tab + 16 + nrp + (size_t)rp * gstride;
  int *zc = zb + (size_t)nrp * gstride;  // the z cell that holds the item's first atom (saves the kernel a search)
  const int T = (cs[ownA0 + sc] - cs[ownA0]) + (two ? cs[ownB0 + sc] - cs[ownB0] : 0);
  int z0 = 0, cum0 = 0, gpos = 0, nch = 0, ng = 0;
  while (gpos < T) {
    while (cum0 + cnt[w][z0] <= gpos) cum0 += cnt[w][z0], z0++;  // the cell that holds atom gpos
    // past gstride - 1 items (only a pathologically crowded row pair gets there) the last item takes all the rest
    if (nch % gk == 0 && ng < gstride - 1) zb[ng] = gpos, zc[ng] = z0, ng++;
    int lim = cum0;
    for (int c = 0; c < ncz_max && z0 + c < sc; c++) lim += cnt[w][z0 + c];
    gpos = min(gpos + catoms, lim);  // same rule as the kernel's chunk extent
    nch++;
  }
  zb[ng] = T;
  tab[16 + rp] = ng;
  atomicMax(&tab[0], ng);
}

#ifdef ATLJ_TILE2_STATS
// debug build only (make stats): where do the phase-2 iterations go?
__device__ unsigned long long g_t2stats[8];
#endif

__device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned *>(&h); }
__device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2 *>(&u); }

__device__ __forceinline__ unsigned long long halo_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// A pair the fast fp64 path could not classify safely (within 1e-8 of the cutoff, or close to the overlap
// threshold): the reference's exact operation sequence on the original coordinates decides, then it is accumulated
// like md_lj_module.py:100-115.  Rare (probability ~1e-8 per pair in a fluid), hence out of line.
struct ExactOut {
  double dx, dy, dz, s;  // separation and its square in sigma units (reference operation order)
  int in, ovr;           // in range (md_lj_module.py:99), overlap (md_lj_module.py:103)
};
__device__ __noinline__ ExactOut exact_pair(const double *__restrict__ r, Phys p, const int *S, const int *egl, int NE,
                                            int i, int j) {
  ExactOut o;
  int en = 0;
  for (int step = 256; step >= 1; step >>= 1)
    if (en + step < NE && S[en + step] <= j) en += step;
  const size_t jg = (size_t)(egl[en] + (j - S[en]));
  double dx, dy, dz, s;
  o.in = pair_sep<false>(r[3 * (size_t)i], r[3 * (size_t)i + 1], r[3 * (size_t)i + 2], r[3 * jg], r[3 * jg + 1],
                         r[3 * jg + 2], 0.0, p.r_cut_box_sq, dx, dy, dz, s);
  o.s = __dmul_rn(s, p.box_sq);
  o.ovr = o.in && (__drcp_rn(o.s) > 1.77);
  o.dx = dx * p.box, o.dy = dy * p.box, o.dz = dz * p.box;
  return o;
}

// sr^2 = 1/s: MUFU.RCP64H seed (relative error <= 2^-23) + one Newton step -> <= 2^-46
__device__ __forceinline__ double rcp46(double s) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  return fma(y, fma(-s, y, 1.0), y);
}

// one in-range pair into the running sums: forces (md_lj_module.py:109,113) and sr^6, sr^12, sr^8, sr^14
#define U_ACCUM(dx_, dy_, dz_, s_)                                                     \
  do {                                                                                 \
    const double sr2_ = rcp46(s_);                                                     \
    const double sr6_ = sr2_ * sr2_ * sr2_, sr12_ = sr6_ * sr6_;                       \
    const double sr8_ = sr6_ * sr2_, sr14_ = sr12_ * sr2_;                             \
    const double fs_ = fma(2.0, sr14_, -sr8_); /* (cut + sr12) * sr2 */                \
    fx = fma(dx_, fs_, fx), fy = fma(dy_, fs_, fy), fz = fma(dz_, fs_, fz);            \
    a6 += sr6_, a12 += sr12_, a8 += sr8_, a14 += sr14_;                                \
  } while (0)
#define U_EXACT(jj)                                                                    \
  do {                                                                                 \
    const ExactOut o_ = exact_pair(a.r, p, S, egl, NE, i, (jj));                       \
    ovr |= o_.ovr;                                                                     \
    if (o_.in) {                                                                       \
      U_ACCUM(o_.dx, o_.dy, o_.dz, o_.s);                                              \
      npf++;                                                                           \
    }                                                                                  \
  } while (0)

// hessian term of one in-range pair (md_lj_module.py:176-191): rij in sigma units, fij = f_i - f_j
#define U_HESS(dx_, dy_, dz_, s_, gx_, gy_, gz_)                                       \
  do {                                                                                 \
    const double sr2_ = rcp46(s_);                                                     \
    const double sr6_ = sr2_ * sr2_ * sr2_;                                            \
    const double sr8_ = sr6_ * sr2_, sr10_ = sr8_ * sr2_;                              \
    const double ff_ = fma(gz_, gz_, fma(gy_, gy_, gx_ * gx_));                        \
    const double rf_ = fma(dz_, gz_, fma(dy_, gy_, dx_ * gx_));                        \
    const double v1_ = 24.0 * fma(-2.0, sr6_, 1.0) * sr8_;                             \
    const double v2_ = 96.0 * fma(7.0, sr6_, -2.0) * sr10_;                            \
    hs = fma(v1_, ff_, fma(v2_, rf_ * rf_, hs));                                       \
  } while (0)
#define U_EXACT_H(jj)                                                                  \
  do {                                                                                 \
    const ExactOut o_ = exact_pair(a.r, p, S, egl, NE, i, (jj));                       \
    if (o_.in) {                                                                       \
      const double ex_ = fix - gx[(jj)], ey_ = fiy - gy[(jj)], ez_ = fiz - gz[(jj)];   \
      U_HESS(o_.dx, o_.dy, o_.dz, o_.s, ex_, ey_, ez_);                                \
      npf++;                                                                           \
    }                                                                                  \
  } while (0)

template <int MODE, bool SPLIT = false>
__global__ void __launch_bounds__(U_NT, MODE == T2_HESS ? 3 : U_CTAS)
    pair_tile2_kernel(PairArgs a, int nyb, int nrp, int gstride, int ncz_max, const int *__restrict__ tab,
                      int *__restrict__ work_counter, int *__restrict__ rows_out, double rc2s, int catoms) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *ux = reinterpret_cast<double *>(smem_raw);
  double *uy = ux + U_STR;
  double *uz = uy + U_STR;
  double *gx = uz + U_STR, *gy = gx + U_STR, *gz = gy + U_STR;  // T2_HESS only: forces of the staged candidates
  // fp16 copy for the filter: record p holds candidates 2p (low halves) and 2p+1 (high halves): (x, y, z, unused)
  uint4 *hp = reinterpret_cast<uint4 *>(uz + U_STR);
  // [U_NE + 4] staged offset of entry e = cell*12 + row
  int *S = MODE == T2_HESS ? reinterpret_cast<int *>(gz + U_STR) : reinterpret_cast<int *>(hp + U_CAP / 2 + 1);
  int *egl = S + U_NE + 4;                               // [U_NE] global index of the first atom of entry e
  __shared__ int rowcell0[U_K];
  __shared__ int cumloc[U_MAXC], sAloc[U_MAXC], sBloc[U_MAXC];  // cells z0 .. of the chunk: atoms before, row starts
  __shared__ int wsum[U_NT / 32];
  __shared__ double s_ctr[3];  // centre of the staged region (the compiler re-derived it for every staged entry)
  __shared__ int s_item, s_halo;

  pdl_wait();  // launched while the sort drains (common.cuh)
  pdl_launch_dependents();
  const Geom g = a.g;
  const Phys p = a.p;
  const bool save_masks = MODE == T2_FORCE && a.masks != nullptr;  // T2_FORCE_SAVE is T2_FORCE with a mask buffer
  // Small systems (fewer chunks than CTA slots: the kernel lasts as long as ONE chunk's dependent instruction stream):
  // chunks of 64 atoms and TWO threads per atom - warps 0, 1 take the lower half of every atom's candidate windows,
  // warps 2, 3 the upper half - and the two partial forces are added through shared memory in a fixed order.
  // (a template parameter: as a run-time flag the tests in the loops cost the large systems 0.8 %)
  constexpr bool split = SPLIT && MODE == T2_FORCE;
  const int part = split ? (int)(threadIdx.x >> 6) : 0;
  const int sc = g.sc, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int *__restrict__ cs = a.cell_start;
  const double box = p.box;
  // rc2s = r_cut_box_sq * box_sq comes as a kernel argument: a constant-bank operand of the one DADD that uses it in the
  // phase-2 loop (as a product of two members of `a` the compiler re-loaded both factors in every iteration)
  const double band = 1e-8 * rc2s;
  const double sovr = (1.0 / 1.77) * (1.0 + 1e-6);
  const unsigned band2 = 2u * (unsigned)__double2hiint(band);
  const int sovr_hi = __double2hiint(sovr);
  const double sci = 1.0 / (double)sc;
  const __half *hq = reinterpret_cast<const __half *>(hp);
  const unsigned ux_s = (unsigned)__cvta_generic_to_shared(ux);
  // Slab ranks: the x layers that read a halo layer (the first and the last owned one) come LAST in the item order,
  // so that the neighbours' halo writes are hidden behind the interior layers; a CTA acquires the neighbours' stamps
  // when it takes its first boundary item.
  const int nbl = g.h ? min(2, g.nx_own) : 0;  // boundary layers
  const int nrp_int = (g.nx_own - nbl) * nyb, nrp_bnd = nbl * nyb;
  if (a.halo.err && a.halo.err[5]) return;  // a neighbour's message timed out earlier in this run
  if (tid == 0) s_halo = a.halo.stamp_left == nullptr;
  if (tid == 0) {  // the dummy candidate (never overwritten: the staged set stays below U_CAP)
    ux[U_CAP] = uy[U_CAP] = uz[U_CAP] = 1e150;
    if (MODE == T2_HESS) gx[U_CAP] = gy[U_CAP] = gz[U_CAP] = 0.0;
  }

  for (;;) {
    // ---- next work item ----------------------------------------------------------------------------------------
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(work_counter, 1);
    const int halo_ok = s_halo;  // read between the two barriers: thread 0 only writes it after the second one
    __syncthreads();
    const int item = s_item;
    const int nitems = tab[0] * nrp;
    if (item == 0 && tid == 0) *rows_out = nitems;
    if (item >= nitems) break;
    int grp, xo, yb;  // item -> (group of chunks, owned x layer, row pair of that layer)
    const int n_int = tab[0] * nrp_int;
    if (item < n_int) {
      grp = item / nrp_int;
      const int q = item - grp * nrp_int;
      xo = q / nyb, yb = q - xo * nyb;
      xo += nbl ? 1 : 0;
    } else {
      const int t = item - n_int;
      grp = t / nrp_bnd;
      const int q = t - grp * nrp_bnd;
      xo = q / nyb, yb = q - xo * nyb;
      xo = xo ? g.nx_own - 1 : 0;
      if (!halo_ok) {  // CTA-uniform
        if (tid == 0) {
          const unsigned long long t0 = halo_now_ns();
          const int stamp = *a.halo.xstep;
          int ok = 0;
          for (;;) {
            int vl, vr;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(vl) : "l"(a.halo.stamp_left + 1) : "memory");
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(vr) : "l"(a.halo.stamp_right + 1) : "memory");
            if (vl == stamp && vr == stamp) {
              ok = 1;
              break;
            }
            if (halo_now_ns() - t0 > a.halo.timeout_ns) break;
          }
          if (!ok) atomicOr(&a.halo.err[5], 1);
          s_halo = ok ? 1 : -1;
        }
        __syncthreads();
        if (s_halo < 0) break;  // timed out: the run is void (err[5]); leave without touching the halo data
      }
    }
    const int rp = xo * nyb + yb;
    if (grp >= tab[16 + rp]) {  // this row pair has fewer items than the longest one: an all-zero row of partials
      if (tid < 8) a.partials[8 * (size_t)item + tid] = 0.0;
      continue;
    }
    const int lx = g.h + xo, y0 = 2 * yb;
    const bool two = (y0 + 1 < sc);
    // atoms [gA, gB) of the row pair's interleaved order
    const int gA = tab[16 + nrp + (size_t)rp * gstride + grp], gB = tab[16 + nrp + (size_t)rp * gstride + grp + 1];
    const int ownA0 = g.cs_idx((lx * sc + y0) * sc), ownB0 = ownA0 + sc;  // rows in the padded cell_start
    const int csA0 = cs[ownA0], csB0 = two ? cs[ownB0] : 0;
    // atoms of the row pair in cells [0, zz)
    auto cumf = [&](int zz) { return (cs[ownA0 + zz] - csA0) + (two ? cs[ownB0 + zz] - csB0 : 0); };
    if (tid < U_K) {
      // row k = yi*3 + xi holds cells (lx-1+xi, y0-1+yi); yi = 0..2 are the dy = -1,0,1 neighbours of row y0 and
      // yi = 1..3 those of row y0+1; the ids come from the stencil function the bit-exact tests cover
      const int yi = tid / 3, xi = tid - 3 * yi;
      int c;
      if (yi < 3)
        c = neighbour_cell(g, false, 0, lx, y0, 0, xi * 9 + yi * 3 + 1);
      else
        c = two ? neighbour_cell(g, false, 0, lx, y0 + 1, 0, xi * 9 + 2 * 3 + 1) : -1;
      rowcell0[tid] = c < 0 ? -1 : g.cs_idx(c);
    }
    const double ctrx = ((lx - g.h + g.x_lo) + 0.5) * sci - 0.5;
    const double ctry = (y0 + (two ? 1.0 : 0.5)) * sci - 0.5;
    // the staged x / y range crosses the periodic boundary only for the outermost layers / rows: elsewhere the minimum
    // image of a centred coordinate is the coordinate itself (|x - ctrx| <= 1.5 cells, |y - ctry| <= 2 cells) and
    // x -= rint(x) is skipped - bit-identical, rint is exactly 0 there
    const int glx = lx - g.h + g.x_lo;
    const bool wrapxy = glx == 0 || glx == sc - 1 || y0 == 0 || y0 + 2 >= sc;
    if (tid == 0) s_ctr[0] = ctrx, s_ctr[1] = ctry;  // (read after the barriers of the chunk's entry table)

    // sums of sr^6, sr^12, sr^8, sr^14 over this thread's pairs; cut, vir, lap follow at the end of the item
    double a6 = 0.0, a12 = 0.0, a8 = 0.0, a14 = 0.0, hs = 0.0;
    unsigned npf = 0u;
    int ovr = 0;

    // the cell that holds atom g0: cumf(z) <= g0 < cumf(z + 1) (recorded by the table kernel)
    int z = tab[16 + nrp + (size_t)nrp * gstride + (size_t)rp * gstride + grp];
    for (int g0 = gA; g0 < gB;) {
      // ---- chunk extent: the next <= U_NT atoms, over at most ncz_max cells ----------------------------------
      __syncthreads();  // the previous chunk is done with shared memory
      if (tid < U_MAXC) {
        const int zz = min(z + tid, sc);
        cumloc[tid] = cumf(zz);
        sAloc[tid] = cs[ownA0 + zz];
        sBloc[tid] = two ? cs[ownB0 + zz] : 0;
      }
      __syncthreads();
      int g1 = min(min(g0 + catoms, gB), cumloc[min(ncz_max, U_MAXC - 2)]);
      int ncz = 1;  // cells that hold atoms [g0, g1)
      while (cumloc[ncz] < g1) ncz++;

      int total, NE;
      for (;;) {
        // ---- entry table: counts -> exclusive scan (3 entries per thread) -------------------------------------
        NE = (ncz + 2) * U_K;
        int cnt[3], sum = 0;
#pragma unroll
        for (int u = 0; u < 3; u++) {
          const int e = 3 * tid + u;
          cnt[u] = 0;
          if (e < NE) {
            const int c = e / U_K, k = e - c * U_K, rc0 = rowcell0[k];
            int zc = z - 1 + c;
            zc = zc < 0 ? zc + sc : (zc >= sc ? zc - sc : zc);
            if (rc0 >= 0) {
              const int gs = cs[rc0 + zc];
              cnt[u] = cs[rc0 + zc + 1] - gs;
              egl[e] = gs;
            } else {
              egl[e] = 0;
            }
          }
          sum += cnt[u];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int off = incl - sum;
#pragma unroll
        for (int w = 0; w < U_NT / 32; w++)
          if (w < warp) off += wsum[w];
        total = 0;
#pragma unroll
        for (int w = 0; w < U_NT / 32; w++) total += wsum[w];
#pragma unroll
        for (int u = 0; u < 3; u++) {
          const int e = 3 * tid + u;
          if (e <= NE) S[e] = off;
          off += cnt[u];
        }
        __syncthreads();
        if (total <= U_CAP || ncz == 1) break;
        ncz = max(1, min(ncz - 1, (int)((long long)ncz * U_CAP / total)));
      }
      g1 = min(g1, cumloc[ncz]);  // ncz may have been cut to fit U_CAP
      if (tid == 0) s_ctr[2] = (z + 0.5 * ncz) * sci - 0.5;  // (the barrier of the window check follows)
      const int ni = g1 - g0;
      // atom il of the chunk -> (cell myc of the chunk, row, global index i, offset in its cell row)
      auto locate = [&](int il, int &myc, int &i, int &k0, int &off) {
        const int gp = g0 + il;
        myc = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1)
          if (myc + step < ncz && cumloc[myc + step] <= gp) myc += step;
        const int idx = gp - cumloc[myc], nAc = sAloc[myc + 1] - sAloc[myc];
        const bool inA = idx < nAc;
        off = inA ? idx : idx - nAc;
        i = (inA ? sAloc[myc] : sBloc[myc]) + off;
        k0 = inA ? 0 : 3;
      };
      // longest window (9 consecutive entries of 3 consecutive cells are three windows; check each)
      bool too_long;
      {
        int mw = 0;
        for (int e = tid; e < (ncz + 2) * 2; e += U_NT) {
          const int c = e >> 1, k0 = (e & 1) * 3;
          mw = max(mw, S[c * U_K + k0 + 9] - S[c * U_K + k0]);
        }
        // (the barrier itself carries the one bit the chunk needs: is any window too long for its mask words?)
        mw = __syncthreads_or(mw > 32 * U_WW - 2);
        too_long = mw != 0;
      }
      const bool slow = total > U_CAP || too_long;

      if (slow) {
        // ---- crowded neighbourhood (does not occur in a fluid): exact arithmetic straight from global memory ----
        for (int il = tid; il < ni; il += U_NT) {
          int c, i, k0, off;
          locate(il, c, i, k0, off);
          const int zi = z + c;
          const double rix = a.r[3 * (size_t)i], riy = a.r[3 * (size_t)i + 1], riz = a.r[3 * (size_t)i + 2];
          double fx = 0, fy = 0, fz = 0, fix = 0, fiy = 0, fiz = 0;
          if (MODE == T2_HESS) fix = a.fin[3 * (size_t)i], fiy = a.fin[3 * (size_t)i + 1], fiz = a.fin[3 * (size_t)i + 2];
          for (int k = k0; k < k0 + 9; k++) {
            const int rc0 = rowcell0[k];
            if (rc0 < 0) continue;
            for (int dz = -1; dz <= 1; dz++) {
              int zj = zi + dz;
              zj = zj < 0 ? zj + sc : (zj >= sc ? zj - sc : zj);
              for (int j = cs[rc0 + zj]; j < cs[rc0 + zj + 1]; j++) {
                if (j == i) continue;
                double dx, dy, dzz, s;
                if (pair_sep<false>(rix, riy, riz, a.r[3 * (size_t)j], a.r[3 * (size_t)j + 1],
                                    a.r[3 * (size_t)j + 2], 0.0, p.r_cut_box_sq, dx, dy, dzz, s)) {
                  const double s2 = __dmul_rn(s, p.box_sq);
                  if (MODE == T2_HESS) {
                    const double ex = fix - a.fin[3 * (size_t)j], ey = fiy - a.fin[3 * (size_t)j + 1],
                                 ez = fiz - a.fin[3 * (size_t)j + 2];
                    U_HESS(dx * box, dy * box, dzz * box, s2, ex, ey, ez);
                  } else {
                    if (__drcp_rn(s2) > 1.77) ovr = 1;  // md_lj_module.py:103
                    U_ACCUM(dx * box, dy * box, dzz * box, s2);
                  }
                  npf++;
                }
              }
            }
          }
          if (MODE != T2_HESS) {
            a.f[3 * (size_t)i] = 24.0 * fx;
            a.f[3 * (size_t)i + 1] = 24.0 * fy;
            a.f[3 * (size_t)i + 2] = 24.0 * fz;
          }
        }
        __syncthreads();
        g0 = g1;
        z += ncz - 1;
        while (z < sc - 1 && cumf(z + 1) <= g0) z++;
        continue;
      }

      // ---- stage the candidates: half a warp per (cell,row) entry, two entries in flight ------------------------------
      {
        const int hw = tid >> 4, hl = tid & 15;
        const double ctrz = s_ctr[2];
        auto put = [&](int s, double x, double y, double zz, double wz) {
          x -= s_ctr[0], y -= s_ctr[1], zz = (zz + wz) - ctrz;
          if (wrapxy) {  // CTA-uniform
            asm volatile("");  // (keeps this a branch: if-converted, both rints would run for every candidate)
            x -= rint(x);
            y -= rint(y);
          }
          x *= box, y *= box, zz *= box;
          ux[s] = x, uy[s] = y, uz[s] = zz;
          if (MODE != T2_HESS) {
            __half *d = reinterpret_cast<__half *>(hp + (s >> 1)) + (s & 1);
            d[0] = __float2half_rn((float)x), d[2] = __float2half_rn((float)y), d[4] = __float2half_rn((float)zz);
          }
        };
        auto putf = [&](int s, size_t j) {  // T2_HESS: the candidate's force next to its position
          gx[s] = a.fin[3 * j], gy[s] = a.fin[3 * j + 1], gz[s] = a.fin[3 * j + 2];
        };
        constexpr int HWS = U_NT / 16;  // half warps per CTA
        constexpr int SE = ATLJ_U_SE;   // entries in flight per half warp (3: 0.819 ms, 2: 0.836, 4 no better than 2)
        for (int e0 = hw; e0 < NE; e0 += SE * HWS) {
          // entries e0, e0 + HWS, ...: issue the global loads of all of them before any is converted, so that a half
          // warp pays one load latency per SE entries (ncu: the staging loop held 13 % of the stall samples)
          int sE[SE], cE[SE], gE[SE];
          double xE[SE], yE[SE], zE[SE], wE[SE], fxE[SE], fyE[SE], fzE[SE];
#pragma unroll
          for (int u = 0; u < SE; u++) {
            const int e = e0 + u * HWS;
            const bool ok = e < NE;
            sE[u] = ok ? S[e] : 0;
            cE[u] = ok ? S[e + 1] - sE[u] : 0;
            gE[u] = ok ? egl[e] : 0;
            const int lc = z - 1 + e / U_K;
            wE[u] = lc < 0 ? -1.0 : (lc >= sc ? 1.0 : 0.0);
          }
#pragma unroll
          for (int u = 0; u < SE; u++) {
            xE[u] = yE[u] = zE[u] = 0.0;
            if (hl < cE[u]) {
              const size_t j = (size_t)(gE[u] + hl);
              xE[u] = a.r[3 * j], yE[u] = a.r[3 * j + 1], zE[u] = a.r[3 * j + 2];
              if (MODE == T2_HESS) fxE[u] = a.fin[3 * j], fyE[u] = a.fin[3 * j + 1], fzE[u] = a.fin[3 * j + 2];
            }
          }
#pragma unroll
          for (int u = 0; u < SE; u++) {
            if (hl < cE[u]) {
              put(sE[u] + hl, xE[u], yE[u], zE[u], wE[u]);
              if (MODE == T2_HESS) gx[sE[u] + hl] = fxE[u], gy[sE[u] + hl] = fyE[u], gz[sE[u] + hl] = fzE[u];
            }
            for (int k = hl + 16; k < cE[u]; k += 16) {  // cells with more than 16 atoms
              const size_t j = (size_t)(gE[u] + k);
              put(sE[u] + k, a.r[3 * j], a.r[3 * j + 1], a.r[3 * j + 2], wE[u]);
              if (MODE == T2_HESS) putf(sE[u] + k, j);
            }
          }
        }
      }
      __syncthreads();
      // Filter threshold.  The filter evaluates d^2 = dx^2 + dy^2 + dz^2 in fp16 (unit roundoff u = 2^-11) on fp16 copies
      // of the chunk-centred coordinates, |coordinate| <= Q.  For a pair within the cutoff R: each difference is off by
      // at most ec = 2uQ (two quantised coordinates) + uR (the subtraction), the sum of squares by at most
      // 2 sqrt(3) R ec + 3 ec^2, and the three fp16 operations that form it by at most 6 u R^2 more.  With 10 % on
      // top and rounded up to fp16, no pair with d^2 < R^2 can fail the test; what passes in excess (about 3 % at
      // Q = 11 sigma) is rejected by the fp64 evaluation.
      // Q from the geometry of the staged region instead of a running maximum over the staged atoms: x within 1.5 cells of
      // the centre, y within 2, z within ncz / 2 + 1 (one halo plane at each end)
      const float Qm = (float)(box * sci) * fmaxf(2.0f, 0.5f * (float)ncz + 1.0f) * 1.001f;
      float thr;
      {
        const float u = 4.8828125e-4f, R = sqrtf((float)rc2s) * 1.0001f;
        const float ec = 2.f * u * Qm + u * R;
        const float E = 3.4642f * R * ec + 3.f * ec * ec + 6.f * u * R * R;
        thr = (float)rc2s * (1.0f + 1e-5f) + 1.1f * E;
      }
      const __half thr_h = __float2half_ru(thr);

      for (int ib = 0; ib < ni; ib += U_NT) {
        const int il = ib + (split ? (tid & 63) : tid);
        const bool act = il < ni;
        int myc = 0, self = 0, i = 0, k0 = 0;
        if (act) {
          int off;
          locate(il, myc, i, k0, off);
          self = S[(myc + 1) * U_K + k0 + 4] + off;
        }
        const double uix = ux[self], uiy = uy[self], uiz = uz[self];
        // ---- phase 1: packed fp16 filter -> left-aligned mask words --------------------------------------------
        // (mask word, shared-memory address of ux[c] for the candidate c of bit 31); bit p <-> candidate c + 31 - p
        uint2 mw[U_NWORDS + 2];  // closed by two zero words: the walk needs no end test of its own
        int nw = 0;
        double fix = 0.0, fiy = 0.0, fiz = 0.0;  // T2_HESS: this atom's force
        if constexpr (MODE == T2_HESS) {
          fix = gx[self], fiy = gy[self], fiz = gz[self];
          if (act) {  // the survivors of the force kernel's filter, same staged indices
            const uint2 *mp = a.masks + (size_t)i * U_NWORDS;
#pragma unroll 1
            for (; nw < U_NWORDS; nw++) {
              const uint2 m = mp[nw];
              if (m.x == 0u) break;
              mw[nw] = make_uint2(m.x, ux_s + 8u * m.y + U_YOFF);  // kept as staged index of bit 31's candidate
            }
          }
        } else {
          const __half *qs = hq + 8 * (self >> 1) + (self & 1);
          const __half2 xi2 = __half2half2(qs[0]), yi2 = __half2half2(qs[2]), zi2 = __half2half2(qs[4]);
          const __half2 nthr2 = __half2half2(__hneg(thr_h));
          if (act) {
#pragma unroll 1
            for (int w = 0; w < 3; w++) {
              const int eb = (myc + w) * U_K + k0;
              const int b = S[eb], e = S[eb + 9];
              if (e <= b) continue;
              const int b0 = b & ~1;
              int npairs = ((e + 1) >> 1) - (b0 >> 1);
              const uint4 *pp = hp + (b0 >> 1);
              int base = 0;
              if (split) {  // this thread's share of the window's 16-record groups
                const int half = 16 * ((((npairs + 15) >> 4) + 1) >> 1);
                if (part == 0)
                  npairs = w == 0 ? npairs : (w == 1 ? min(npairs, half) : 0);
                else
                  base = w == 2 ? 0 : (w == 1 ? half : npairs);
              }
              for (; base < npairs; base += 16) {
                const int n = min(16, npairs - base);
                unsigned wd = 0u;
#pragma unroll U_P1U
                for (int k = 0; k < n; k++) {
                  const uint4 A = pp[base + k];
                  const __half2 dx = __hsub2(xi2, u2h(A.x)), dy = __hsub2(yi2, u2h(A.y)), dz = __hsub2(zi2, u2h(A.z));
                  __half2 t = __hfma2(dx, dx, nthr2);
                  t = __hfma2(dy, dy, t);
                  t = __hfma2(dz, dz, t);
                  const unsigned tu = h2u(t);  // sign bits: candidate 2k at bit 15, candidate 2k+1 at bit 31
                  wd = __funnelshift_l(tu << 16, wd, 1);
                  wd = __funnelshift_l(tu, wd, 1);
                }
                wd <<= 2 * (16 - n);
                // bit p <-> candidate c31 - p.  Candidates outside [b, e) that rode along in the first / last pair, and
                // the atom itself, are cleared here; a word without survivors is not stored at all (ncu/stat build: 4.1
                // of the 10.7 words per atom were empty and cost a full phase-2 iteration each)
                const int c31 = b0 + 2 * base + 31;
                if (base == 0 && (b & 1)) wd &= 0x7fffffffu;
                const unsigned pe = (unsigned)(c31 - e), ps = (unsigned)(c31 - self);
                if (pe < 32u) wd &= ~(1u << pe);
                if (w == 1 && ps < 32u) wd &= ~(1u << ps);
                if (wd != 0u) {
                  // second member: shared-memory address of the candidate a leading-zero count of 0 stands for
                  mw[nw] = make_uint2(wd, ux_s + 8u * (unsigned)(c31 - 31) + U_YOFF);
                  nw++;
                }
              }
            }
          }

          if (save_masks && act) {  // kept for the hessian kernel (a zero word ends the list)
            uint2 *mp = a.masks + (size_t)i * U_NWORDS;
            for (int k = 0; k < nw; k++) mp[k] = make_uint2(mw[k].x, (mw[k].y - U_YOFF - ux_s) >> 3);
            if (nw < U_NWORDS) mp[nw] = make_uint2(0u, 0u);
          }
        }

        // ---- phase 2: fp64 evaluation of the surviving pairs ----------------------------------------------------
        // Every lane walks its own mask words, so in almost every iteration SOME lane moves on to its next word:
        // that step is branch-free (selects + one always-issued local load whose result is needed an iteration
        // later at the earliest).  Each iteration takes up to TWO pairs off the current word and evaluates them as
        // two independent fp64 chains (the kernel is latency-, not fp64-throughput-bound); a dead slot gets
        // s = 1e300, whose sr^6 underflows to exactly 0, so no branch is needed around the arithmetic.
        double fx = 0.0, fy = 0.0, fz = 0.0;
        // (opaque to the compiler: it re-derived this address - five instructions - in every iteration of the loop)
        unsigned dummy_addr;
        asm volatile("add.u32 %0, %1, %2;" : "=r"(dummy_addr) : "r"(ux_s), "n"(8 * U_CAP));
        int nd = 0;
        int defer[4];
        int ww = 0;
        uint2 cur = make_uint2(0u, 0u), nxt = make_uint2(0u, 0u);  // (mask, address of the candidate a leading-zero count of 0 stands for)
        mw[nw] = make_uint2(0u, 0u), mw[nw + 1] = make_uint2(0u, 0u);
        cur = mw[0], nxt = mw[1];
#ifdef ATLJ_TILE2_STATS
        {
          int pc = 0, emp = 0, half = 0;
          for (int q = 0; q < nw; q++) {
            const int c = __popc(mw[q].x);
            pc += c, emp += c == 0, half += (c + 1) >> 1;
          }
          const int mx = __reduce_max_sync(FULL, half + emp), mxh = __reduce_max_sync(FULL, (pc + 1) >> 1);
          atomicAdd(&g_t2stats[1], (unsigned long long)pc);
          atomicAdd(&g_t2stats[2], (unsigned long long)nw);
          atomicAdd(&g_t2stats[3], (unsigned long long)emp);
          atomicAdd(&g_t2stats[4], (unsigned long long)half);
          if (lane == 0) atomicAdd(&g_t2stats[5], 1ull), atomicAdd(&g_t2stats[6], (unsigned long long)mx),
              atomicAdd(&g_t2stats[7], (unsigned long long)mxh);
        }
        int iters = 0;
#endif
        // (the loop ends with the last bit of the last word: testing ww < nw alone costs every warp one dead iteration)
        while ((cur.x | nxt.x) != 0u) {  // only words with survivors are stored: nxt.x == 0 <=> no further word
#ifdef ATLJ_TILE2_STATS
          iters++;
#endif
          const bool adv = cur.x == 0u;
          if (adv) ww++;
          if (adv) cur = nxt;
          nxt = mw[ww + 1];  // re-read every iteration (the same word unless the walk has just moved on): no selects
          unsigned cm = cur.x;
          bool live0 = cm != 0u;
          // cur.y is the address of the candidate of bit 0 + 8 * 31: the candidate of bit p sits at cur.y - 8 p, so the
          // position of the highest set bit is used as it comes (bfind; -1 for an empty word, whose shift clamps to 0)
          unsigned p0, p1, b0m, b1m;
          asm("bfind.u32 %0, %1;" : "=r"(p0) : "r"(cm));
          asm("shl.b32 %0, 1, %1;" : "=r"(b0m) : "r"(p0));
          cm &= ~b0m;
          bool live1 = cm != 0u;
          asm("bfind.u32 %0, %1;" : "=r"(p1) : "r"(cm));
          asm("shl.b32 %0, 1, %1;" : "=r"(b1m) : "r"(p1));
          cm &= ~b1m;
          cur.x = cm;
          const unsigned a0 = live0 ? cur.y - 8u * p0 : dummy_addr;
          const unsigned a1 = live1 ? cur.y - 8u * p1 : dummy_addr;
          const double dx0 = uix - lds_f64(a0), dy0 = uiy - lds_f64(a0 + 8u * U_STR),
                       dz0 = uiz - lds_f64(a0 + 16u * U_STR);
          const double dx1 = uix - lds_f64(a1), dy1 = uiy - lds_f64(a1 + 8u * U_STR),
                       dz1 = uiz - lds_f64(a1 + 16u * U_STR);
          double s0 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0)), s1 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
          const int e0 = __double2hiint(s0 - rc2s), e1 = __double2hiint(s1 - rc2s);
          // In-range pairs are counted here.  A false positive of the filter (about 3 % of the survivors: the fp16
          // margin) and the dummy of a dead slot are simply not in range - no branch, most warp iterations have one.
          // |s - rc2s| <= band or s < sovr (tested conservatively on the high words: integer pipe, not fp64) is rare
          // and leaves the hot path: such pairs are set aside and decided with the reference's exact operation sequence.
          const bool b0 = (((unsigned)e0 << 1) <= band2) | (__double2hiint(s0) <= sovr_hi);
          const bool b1 = (((unsigned)e1 << 1) <= band2) | (__double2hiint(s1) <= sovr_hi);
          // (a deferred pair is switched off in the rare branch below: the hot path selects on the sign alone)
          // counted by the sign of s - rc2s alone (one add per slot); a deferred pair is taken back below
          npf += ((unsigned)e0 >> 31) + ((unsigned)e1 >> 31);
          if (b0 | b1) {
            if (b0) npf -= (unsigned)e0 >> 31, s0 = 3e300;
            if (b1) npf -= (unsigned)e1 >> 31, s1 = 3e300;
            if (b0) defer[nd++] = (int)((a0 - ux_s) >> 3);
            if (b1) defer[nd++] = (int)((a1 - ux_s) >> 3);
            if (nd >= 3) {
              for (int k = 0; k < nd; k++) {
                if constexpr (MODE == T2_HESS)
                  U_EXACT_H(defer[k]);
                else
                  U_EXACT(defer[k]);
              }
              nd = 0;
            }
          }
          // out of range: s becomes ~1e300 (only the high word needs replacing), whose sr^6 underflows to exactly 0
          s0 = __hiloint2double(e0 < 0 ? __double2hiint(s0) : 0x7e37e43c, __double2loint(s0));
          s1 = __hiloint2double(e1 < 0 ? __double2hiint(s1) : 0x7e37e43c, __double2loint(s1));
          if constexpr (MODE == T2_HESS) {
            const double ex0 = fix - lds_f64(a0 + 24u * U_STR), ey0 = fiy - lds_f64(a0 + 32u * U_STR),
                         ez0 = fiz - lds_f64(a0 + 40u * U_STR);
            const double ex1 = fix - lds_f64(a1 + 24u * U_STR), ey1 = fiy - lds_f64(a1 + 32u * U_STR),
                         ez1 = fiz - lds_f64(a1 + 40u * U_STR);
            U_HESS(dx0, dy0, dz0, s0, ex0, ey0, ez0);
            U_HESS(dx1, dy1, dz1, s1, ex1, ey1, ez1);
          } else {
            U_ACCUM(dx0, dy0, dz0, s0);
            U_ACCUM(dx1, dy1, dz1, s1);
          }
        }
#ifdef ATLJ_TILE2_STATS
        {
          const int mi = __reduce_max_sync(FULL, iters);
          if (lane == 0) atomicAdd(&g_t2stats[0], (unsigned long long)mi);
        }
#endif
        for (int k = 0; k < nd; k++) {
          if constexpr (MODE == T2_HESS)
            U_EXACT_H(defer[k]);
          else
            U_EXACT(defer[k]);
        }
        if (split) {
          // upper-half threads hand their partial force over through the (now idle) fp16 region; fixed order of the
          // two-term sum, so the result is reproducible
          double *fsc = reinterpret_cast<double *>(hp);
          __syncthreads();  // every warp is past phase 1 (the last reader of that region)
          if (part == 1 && act) fsc[3 * il] = fx, fsc[3 * il + 1] = fy, fsc[3 * il + 2] = fz;
          __syncthreads();
          if (part == 0 && act) fx += fsc[3 * il], fy += fsc[3 * il + 1], fz += fsc[3 * il + 2];
        }
        if (MODE != T2_HESS && act && part == 0) {
          a.f[3 * (size_t)i] = 24.0 * fx;  // md_lj_module.py:143
          a.f[3 * (size_t)i + 1] = 24.0 * fy;
          a.f[3 * (size_t)i + 2] = 24.0 * fz;
        }
      }
      __syncthreads();
      g0 = g1;
      // the cell that holds the next chunk's first atom (skipping empty cells): cumloc[k] = cumf(z + k), k < U_MAXC,
      // is still this chunk's table (it is rewritten after the barrier at the top of the chunk loop)
      int zn = z + ncz - 1;
      while (zn < sc - 1 && (zn + 1 - z < U_MAXC ? cumloc[zn + 1 - z] : cumf(zn + 1)) <= g0) zn++;
      z = zn;
    }
    // cut = sr12 - sr6, pot = cut - pot_cut, vir = cut + sr12, lap = (22 sr12 - 5 sr6) sr2 (md_lj_module.py:104-108)
    Acc acc;
    {
      const double cutacc = a12 - a6;
      acc.pot = cutacc - p.pot_cut * (double)npf;
      acc.a = cutacc;
      acc.vir = cutacc + a12;
      acc.lap = 22.0 * a14 - 5.0 * a8;
      acc.hes = hs;
      acc.np = npf;
      acc.ovr = ovr;
    }
    block_store_partials<U_NT / 32>(acc, a.partials + 8 * (size_t)item);
  }
}

template <int MODE>
static int tile2_grid() {
  static int grid = 0;
  if (grid == 0) {
    const size_t smem = tile2_smem_bytes(MODE);
    if (cudaFuncSetAttribute(pair_tile2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return -1;
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pair_tile2_kernel<MODE>, U_NT, smem) != cudaSuccess ||
        cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return -1;
    grid = (per_sm < 1 ? 1 : per_sm) * sms;
  }
  return grid;
}

// -> upper bound of the partial rows written, or a negative status.  The exact number of rows (items) is known on
// the device only: the kernel stores it in *rows_out (work_counter + 4), which the reducers read.
// mode: T2_FORCE, T2_FORCE_SAVE (a.masks receives the mask words) or T2_HESS (a.fin = forces, a.masks from the
// T2_FORCE_SAVE launch on the same cell table; the item table of that launch is reused).
int launch_pair_tile2(cudaStream_t st, const PairArgs &a, int *work_counter, int mode, int part, bool tab_zeroed) {
  const int grid = mode == T2_HESS ? tile2_grid<T2_HESS>()
                                   : tile2_grid<T2_FORCE>();
  if (grid < 0) {
    set_error("pair_tile2: kernel attributes -> %s", cudaGetErrorString(cudaGetLastError()));
    return ATLJ_ERR_CUDA;
  }
  const size_t smem = tile2_smem_bytes(mode);
  if (a.g.sc > 1024 || !a.tile_tab || (mode != T2_FORCE && !a.masks) || (mode == T2_HESS && !a.fin)) {
    set_error("pair_tile2: sc = %d exceeds 1024 cells per edge, or a table / mask / force buffer is missing", a.g.sc);
    return ATLJ_ERR_ARG;
  }
  const int nyb = (a.g.sc + 1) / 2, nrp = a.g.nx_own * nyb, gstride = tile2_gstride(a.g);
  // chunks per item: about 8 items per resident CTA (of the force kernel, so that both modes cut the same items), at
  // least 1 (small systems), at most U_GK_MAX
  const int grid_f = tile2_grid<T2_FORCE>();
  const long long natoms = a.natoms > 0 ? (long long)a.natoms : 12ll * a.g.n_own_cells();
  const double per_cell =
      a.per_cell > 0.0 ? a.per_cell : (double)natoms / (double)(a.g.n_own_cells() > 0 ? a.g.n_own_cells() : 1);
  // two threads per atom and 64-atom chunks (split mode of the force kernel) while the WHOLE system - all ranks: the
  // rule must not depend on the decomposition, or slab and single-domain forces would differ in the last bit - has
  // fewer 64-atom chunks than the GPU has CTA slots; never together with the mask words the hessian kernel walks
  static long long split_max = -1;
  if (split_max < 0) {
    const char *sm = getenv("ATLJ_SPLIT_MAX");
    split_max = sm ? atoll(sm) : 44000;
  }
  const double n_system = per_cell * (double)a.g.sc * (double)a.g.sc * (double)a.g.sc;
  const int catoms = (mode == T2_FORCE && U_NT == 128 && n_system <= (double)split_max) ? 64 : U_NT;
  const long long chunks = natoms / (catoms - catoms / 16) + nrp;
  int gk = (int)(chunks / ((long long)ATLJ_U_IPC * (grid_f > 0 ? grid_f : grid)));
  gk = gk < 1 ? 1 : (gk > U_GK_MAX ? U_GK_MAX : gk);
  // cells a chunk may span: the staged cross-section (12 rows per cell, one halo cell at each end) has to fit U_CAP
  int ncz_max = (int)(0.96 * U_CAP / (U_K * (per_cell > 0.05 ? per_cell : 0.05))) - 2;
  // fp16 range: the chunk-centred z coordinates stay below 100 sigma
  const double cell = a.p.box / a.g.sc;
  if (ncz_max > (int)(200.0 / cell) - 2) ncz_max = (int)(200.0 / cell) - 2;
  ncz_max = ncz_max < 1 ? 1 : (ncz_max > U_MAXC - 2 ? U_MAXC - 2 : ncz_max);
  if (mode != T2_HESS && part != 2) {
    if (!tab_zeroed) ATLJ_CUDA(cudaMemsetAsync(a.tile_tab, 0, sizeof(int), st));
    launch_pdl(PDL_TABLE, tile2_table_kernel, dim3((nrp + 3) / 4), dim3(128), 0, st, a.g, a.cell_start, nyb, nrp, gstride, gk,
               ncz_max, catoms, a.tile_tab);
    ATLJ_LAUNCHED();
  }
  const int bound = nrp * (gstride - 1);
  if (part == 1) return bound;
  const int nblk = grid < bound ? grid : bound;
  if (mode == T2_HESS)
    launch_pdl(PDL_PAIR, pair_tile2_kernel<T2_HESS>, dim3(nblk), dim3(U_NT), smem, st, a, nyb, nrp, gstride, ncz_max, a.tile_tab,
               work_counter, work_counter + 4, a.p.r_cut_box_sq * a.p.box_sq, U_NT);
  else
  {
    PairArgs af = a;
    if (mode == T2_FORCE) af.masks = nullptr;
    if (catoms < U_NT) {
      static bool attr = false;
      if (!attr) {
        ATLJ_CUDA(cudaFuncSetAttribute(pair_tile2_kernel<T2_FORCE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        attr = true;
      }
      launch_pdl(PDL_PAIR, pair_tile2_kernel<T2_FORCE, true>, dim3(nblk), dim3(U_NT), smem, st, af, nyb, nrp, gstride, ncz_max,
                 a.tile_tab, work_counter, work_counter + 4, a.p.r_cut_box_sq * a.p.box_sq, catoms);
    } else {
      launch_pdl(PDL_PAIR, pair_tile2_kernel<T2_FORCE>, dim3(nblk), dim3(U_NT), smem, st, af, nyb, nrp, gstride, ncz_max,
                 a.tile_tab, work_counter, work_counter + 4, a.p.r_cut_box_sq * a.p.box_sq, catoms);
    }
  }
  ATLJ_LAUNCHED();
  return bound;
}

}  // namespace atlj

#ifdef ATLJ_TILE2_STATS
extern "C" int atlj_debug_tile2_stats(unsigned long long out[8]) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, atlj::g_t2stats, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(atlj::g_t2stats, z, sizeof(z));
  return 0;
}
#endif
