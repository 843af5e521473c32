// Shared declarations of libatlj's translation units (internal; the public ABI is include/atlj.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/atlj.h"

namespace atlj {

// ---- error plumbing -----------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int device_ok();  // ATLJ_OK, or ATLJ_ERR_CUDA with the message set (the library has no CPU path)
extern int64_t g_launches;  // kernels launched by this library (atlj_kernel_launches)

#define ATLJ_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      atlj::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
      return ATLJ_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

#define ATLJ_LAUNCHED()                                                                          \
  do {                                                                                           \
    atlj::g_launches++;                                                                          \
    cudaError_t e_ = cudaPeekAtLastError();                                                      \
    if (e_ != cudaSuccess) {                                                                     \
      atlj::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));  \
      return ATLJ_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

// ---- programmatic dependent launch ---------------------------------------------------------------------------
// The kernels of the step loop are tiny next to the pair kernel and strictly dependent; with the attribute below the
// next kernel of the stream is launched while its predecessor drains (its CTAs become resident as slots free up) and
// blocks in pdl_wait() - the first statement of every such kernel - until the predecessor has completed and its
// memory is visible.  Launched without the attribute (ATLJ_PDL=0, or a kernel that follows a copy), pdl_wait() and
// pdl_launch_dependents() are no-ops.
extern int g_pdl;  // bit mask of the kernels that use the attribute (ATLJ_PDL; 0: plain stream order everywhere)
enum { PDL_KDK = 1, PDL_SCAN = 2, PDL_SCATTER = 4, PDL_GATHER = 8, PDL_TABLE = 16, PDL_PAIR = 32 };
// Measured (profiles/r2_hbm_kernels.md): stream-launched loops gain 1.5-3.6 %; inside a replayed graph the attribute
// only pays for systems below ~44,000 atoms and costs ~1.5 us per step above, and on the integration kernel (which
// would queue up behind the persistent pair kernel) it costs 17 us at 2 M atoms.  Default mask: scan, scatter, gather,
// pair; switched off while a graph of a larger system is being captured (g_pdl_off).
extern int g_pdl_off;
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.launch_dependents;");
#endif
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(int which, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = ((g_pdl & which) && !g_pdl_off) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- local cell grid of one rank ------------------------------------------------------------------------------
// Global grid: sc x sc x sc cells, id = (c0*sc + c1)*sc + c2 (C order, md_lj_ll_module.py:116).
// A rank owns global x-layers [x_lo, x_lo+nx_own); its local grid has nxl = nx_own + 2h layers where h = 1
// when the slab has halo layers (multi-GPU) and 0 when x wraps inside the array (single GPU).
struct Geom {
  int sc;      // cells per box edge (global)
  int x_lo;    // first owned global x layer
  int nx_own;  // owned layers
  int h;       // halo layers on each side (0 or 1)
  int nxl;     // nx_own + 2h
  __host__ __device__ int ncell() const { return nxl * sc * sc; }
  // cell_start carries one end sentinel PER x LAYER: entry of cell c is cs_idx(c) = c + c / sc^2.  Layers can then
  // sit anywhere in the sorted arrays ([owned | left halo | right halo] on a slab) and a z row never reads past its
  // own layer.
  __host__ __device__ int cs_size() const { return nxl * (sc * sc + 1); }
  __host__ __device__ int cs_idx(int cell) const { return cell + cell / (sc * sc); }
  __host__ __device__ int n_own_cells() const { return nx_own * sc * sc; }
  __host__ __device__ int first_own_cell() const { return h * sc * sc; }
};

constexpr int MAX_NB = 30;  // 27 (LJ) or up to 9+12+9 (Lees-Edwards top/bottom layer)

// Full-stencil neighbour cell k of local cell (lx,cy,cz); returns local cell id or -1 if slot k is unused.
// LJ: the 14 offsets of md_lj_ll_module.py:84-86 and their mirror images = all 27 offsets.
// LE: md_lj_llle_module.py:88-92,131-135: i-cells in the top y layer look up (dy=+1) at dx in {1,0,-1,-2}-shift;
//     the mirror relation makes bottom-layer cells look down (dy=-1) at dx in {-1,0,1,2}+shift.
__host__ __device__ inline int neighbour_cell(const Geom &g, bool le, int shift, int lx, int cy, int cz, int k) {
  int dx, dy, dz;
  if (!le) {
    if (k >= 27) return -1;
    dx = k / 9 - 1;
    dy = (k / 3) % 3 - 1;
    dz = k % 3 - 1;
  } else {
    // slots 0..8: dy = 0; slots 9..20: the "wide" row (4 x-offsets: top layer looking up / bottom layer looking
    // down through the sheared boundary; interior cells use 3 of the 4); slots 21..29: the opposite plain row.
    if (k >= 30) return -1;
    const bool bottom = (cy == 0), top = (cy == g.sc - 1);
    if (k < 9) {
      dy = 0;
      dx = k / 3 - 1;
      dz = k % 3 - 1;
    } else if (k < 21) {
      int m = k - 9;
      dz = m % 3 - 1;
      int q = m / 3;  // 0..3
      if (top) {
        dy = 1;
        dx = (1 - q) - shift;  // 1,0,-1,-2 minus shift   (md_lj_llle_module.py:88-92,133)
      } else if (bottom) {
        dy = -1;
        dx = (q - 1) + shift;  // mirror image: -1,0,1,2 plus shift
      } else {
        if (q == 3) return -1;
        dy = 1;
        dx = 1 - q;
      }
    } else {
      int m = k - 21;  // 0..8
      dz = m % 3 - 1;
      dx = m / 3 - 1;
      dy = (bottom && !top) ? 1 : -1;
    }
  }
  int sc = g.sc;
  int ny = cy + dy;
  ny = ny < 0 ? ny + sc : (ny >= sc ? ny - sc : ny);
  int nz = cz + dz;
  nz = nz < 0 ? nz + sc : (nz >= sc ? nz - sc : nz);
  int nx = lx + dx;
  if (g.h == 0) {
    nx %= sc;
    if (nx < 0) nx += sc;
  } else if (nx < 0 || nx >= g.nxl) {
    return -1;  // cannot happen for LJ with one halo layer
  }
  return (nx * sc + ny) * sc + nz;
}

// ---- physics constants of one force evaluation -------------------------------------------------------------------
struct Phys {
  double box, box_sq, r_cut_box_sq, pot_cut, strain;
  float rc2_cell;  // prefilter threshold: (r_cut/box*sc)^2 * (1+margin), cell units
  int shift;       // floor(strain*sc), md_lj_llle_module.py:112
  int model;       // ATLJ_MODEL_*
};

// ---- launchers implemented in cells.cu / pair.cu / integrate.cu -------------------------------------------------
struct CellBufs {
  int *cellid;          // [cap] local cell id of each atom (current order)
  int *rank;            // [cap] arrival rank inside its cell
  unsigned long long *keys;  // [cap] scratch of the counting sort: the cells' atom ids (32-bit) in arrival order
  int *cell_count;      // [ncell]
  int *cell_start;      // [ncell+1]
  int *block_sums;      // scan scratch
  int *err;             // [4] device flags: [0] index error, [1] capacity overflow
};

// wrap + cell assignment + per-cell counts (+ optional int64 triplets out)
// n_dev (may be null): device-side atom count, n is then only an upper bound; drop_adjacent: atoms that sit one layer
// outside the slab are skipped silently (they were packed for migration) instead of raising the index error
int launch_cell_assign(cudaStream_t st, const Geom &g, const double *r_in, double *r_wrapped, int n, bool le,
                       double strain, const CellBufs &b, long long *c_out, const int *n_dev = nullptr,
                       bool drop_adjacent = false, const int *i0_dev = nullptr);
// slab extras of the fused integration kernel (all null / 0 on a single GPU)
struct SlabArgs {
  const int *n_dev;                  // owned atoms, kept on the device (n is then an upper bound)
  const int *id;                     // global ids of the atoms held
  unsigned char *to_left0, *to_left1, *to_right0, *to_right1;  // the neighbours' migration buffers (peer memory), one
                                     // per step parity: leavers are packed into them
  int mig_cap;                       // atoms per message
  int *xstep;                        // device counter of exchange steps: parity = *xstep & 1 selects the buffer, *xstep + 1
                                     // is the stamp published in the header when the message is complete; the kernel's
                                     // last CTA moves the counter on (so the launch arguments do not depend on the step)
  int assign_only;                   // the atoms were moved (and wrapped) by another integrator kernel (BAOAB, Nose-Hoover):
                                     // no kick, no drift - cell assignment and migration only
  int *mig_cnt;                      // [0], [1] slot counters, [2] finished boundary tiles
  // Only atoms of the first and the last owned layer can leave the slab in one step.  The sorted arrays hold the
  // layers one after the other, so the tiles of those two layers are integrated FIRST and the migration messages are
  // published as soon as they are done - the rest of the kernel hides the neighbours' latency and skew.
  const int *first_end;              // &cell_start[end sentinel of the first owned layer] of the previous sort
  const int *last_begin;             // &cell_start[first cell of the last owned layer]      (both null: no sort yet,
                                     // every tile counts as a boundary tile)
};
// fused kick [+ sums] + kick + drift + wrap + cell assignment of the NVE loop; returns the number of CTAs
// (= rows of vf_partials written when first_kick) or a negative status
// rec_vf (used when first_kick): the finished step's record: its [4], [5] = sum v^2, sum f^2 and, when pair_partials
// is given, the pair kernel's totals (per-item rows reduced in fixed order) are written by the kernel's last CTA
// loop (may be null): the device-resident NVE loop, whose launches must not depend on the step number so that two
// steps can be captured once and replayed as a CUDA graph
struct KdkLoop {
  int *step_ctr;       // device counter: rec_vf is the BASE of the records and the finished step's record is row *step_ctr
                       // (null: rec_vf is that record itself)
  int *tab0;           // word 0 of the tiled pair kernel's item table, zeroed by the kernel (null: the launcher of the
                       // pair kernel does it)
  bool counts_zeroed;  // cell_count was left zeroed by the previous step's launch_cell_sort_gather(zero_counts)
  bool no_assign = false;  // all-pairs route (sc < 4): kick + sums + kick + drift + wrap only, no cell assignment
  bool tail_in_scan = false;  // the record's final fixed-order sum is left to the scan launch that follows (RecTail)
};
// The second level of the step record's reduction (the per-CTA rows of the integration kernel -> the record), done by
// one extra CTA of the scan launch that follows it: in the integration kernel's last CTA it was 8 us with one SM
// working; beside the scan it costs nothing.  Same fixed order of summation, so the records keep their bits.
struct RecTail {
  const double *rows = nullptr;  // [nrows][12] rows the integration kernel's CTAs left behind
  int nrows = 0;                 // 0: no tail
  double *rec = nullptr;         // the record, or the base of the records when step_ctr is given
  int *step_ctr = nullptr;       // device counter of replayed steps: row of the record; moved on by the tail
  int model = 0;
  int has_pair = 0;              // the rows carry the pair kernel's totals as well
  int *err = nullptr;            // device flags: [2] = work counter of the tiled pair kernel, reset here
};
int launch_kdk_assign(cudaStream_t st, const Geom &g, int n, double *r, double *v, const double *f, double half_dt,
                      double dt, double box, bool first_kick, const CellBufs &b, double *vf_partials, double *rec_vf,
                      const double *pair_partials, int pair_rows, int model, const SlabArgs *slab = nullptr,
                      const int *pair_rows_dev = nullptr, const KdkLoop *loop = nullptr);
// exclusive scan of cell_count over the OWNED cells -> cell_start (offset by base)
int launch_cell_scan(cudaStream_t st, const Geom &g, const CellBufs &b, int base, const RecTail *tail = nullptr);
// slab ranks: where the sort + gather kernel sends the first / last owned layer (the neighbours' halo regions, mg.cu)
struct HaloOut {
  double *dst_r_left, *dst_r_right;  // the neighbours' position arrays at their halo regions (peer memory)
  int *dst_cs_left, *dst_cs_right;   // the halo rows of their cell tables
  int base_left, base_right;         // atom index at which the positions land over there
  int *hdr_left, *hdr_right;         // message headers: [0] = atoms, [1] = stamp
  int cap;                           // atoms a halo region holds
  const int *xstep;                  // device counter of exchange steps (already moved on by the integration kernel):
                                     // its value is the stamp published when a layer is complete
  const int *rec_ctr;                // when not null, rec is the base of the records and the row is *rec_ctr
  int *done;                         // units (atoms + table entries) stored so far (left 0 by the kernel)
  int *n_own;                        // device copy of the owned-atom count after the sort
  long long own_cap;
  double *rec;                       // step record: [11] = owned atoms (may be null)
  int *err;
};
// scatter ids into cells, sort each cell by id, gather r (and v) into the sorted arrays
int launch_cell_sort_gather(cudaStream_t st, const Geom &g, const CellBufs &b, int n, int base, const int *id_in,
                            const double *r_in, const double *v_in, double *r_out, double *v_out, int *id_out,
                            const int *n_dev = nullptr, const HaloOut *halo = nullptr, bool zero_counts = false,
                            int n_hint = 0);

// slab ranks: the halo layers are written by the neighbours; a pair kernel that reaches a boundary x layer acquires
// their stamps first (null pointers: nothing to wait for)
struct HaloWait {
  const int *stamp_left = nullptr, *stamp_right = nullptr;  // message headers: [1] = stamp
  const int *xstep = nullptr;  // device counter of exchange steps: its value is the stamp to wait for
  unsigned long long timeout_ns = 0;
  int *err = nullptr;  // device flags; [5] is raised on a timeout and turns the rest of the run into no-ops
};

struct PairArgs {
  const double *r;       // sorted positions (owned + halo), AoS
  const double *fin;     // sorted forces (hessian) or nullptr
  const int *cell_start;
  const int *cell_count;
  double *f;             // sorted forces out (force mode)
  double *partials;      // [rows][8]
  int *tile_tab;         // work-item table of pair_tile2 (pair_tile2_tab_ints)
  uint2 *masks;          // pair_tile2: per-atom mask words kept between the force and the hessian kernel, or nullptr
  int natoms;            // owned atoms (a hint for the work-item size; 0 = unknown)
  double per_cell;       // mean atoms per cell of the WHOLE system (0 = unknown): fixes the chunk rule of pair_tile2, so
                         // that a slab cuts the same chunks - and sums in the same order - as the single domain
  Geom g;
  Phys p;
  HaloWait halo;
};
int pair_grid_blocks();  // upper bound of the all-pairs kernel's CTAs (partials has at least this many rows)
int pair_tile_grid(const Geom &g);  // CTAs of the tiled kernel (rows of partials it writes)
int launch_pair_tile(cudaStream_t st, const PairArgs &a, bool hessian);  // -> CTAs launched or a negative status
int pair_tile2_items(const Geom &g);  // work items (= rows of partials) of the second-generation LJ kernel
// -> items (= rows of partials written) or a negative status; *work_counter must be 0 on entry (launch_finalize and
// the fused integration kernel reset it)
// The number of rows actually written is stored by the kernel in work_counter[4] (device); the reducers take that
// pointer (rows_dev) next to the upper bound returned here.
// mode 0: forces; 1: forces, and a.masks receives every atom's mask words; 2: hessian sum from a.fin and the masks of a
// mode-1 launch on the same cell table
// part 0: item table + pair kernel; 1: the item table only (it depends on cell_start alone, so the replayed loop runs it
// beside the sort + gather); 2: the pair kernel only.  tab_zeroed: word 0 of the table is already 0 (KdkLoop::tab0)
int launch_pair_tile2(cudaStream_t st, const PairArgs &a, int *work_counter, int mode = 0, int part = 0,
                      bool tab_zeroed = false);
size_t pair_tile2_mask_bytes(long long natoms);
int pair_tile2_tab_ints(const Geom &g);  // ints of PairArgs::tile_tab for this geometry
// direct link-cell kernel (cells of ~1 atom): one thread per owned atom, no staging
int pair_direct_blocks(int n);
int launch_pair_direct(cudaStream_t st, const PairArgs &a, bool hessian, int n);
// small systems: eight lanes per atom, exact arithmetic, candidates from the cell table (cells) or all n atoms
int pair_multi_blocks(int n);
int launch_pair_multi(cudaStream_t st, const PairArgs &a, bool hessian, int n, bool cells);
// one atom against nj partners (smc_lj_module.force_1): -> rows of partials written
int launch_force_1(cudaStream_t st, const double ri[3], const double *r, int nj, const Phys &p, double *f_partial,
                   double *partials);
int launch_pair_all(cudaStream_t st, const double *r, const double *fin, int n, const Phys &p, double *f,
                    double *partials, bool hessian, int *nblocks_out);
// fixed-order reduction of partials -> rec[0..3], rec[7], rec[8] (force) or rec[6] (hessian)
// (also zeroes *reset_counter, the work counter of the tiled kernel, when it is not null; rows_dev, when not null,
// holds the number of rows on the device and nblocks is only an upper bound)
// vf_partials (may be null): [nvf][4] partial {sum v^2, sum f^2} of the integration kernel, reduced into rec[4], rec[5]
int launch_finalize(cudaStream_t st, const double *partials, int nblocks, int model, bool hessian, double *rec,
                    int *reset_counter = nullptr, const double *vf_partials = nullptr, int nvf = 0,
                    const int *rows_dev = nullptr, const int *step_ctr = nullptr);

}  // namespace atlj
