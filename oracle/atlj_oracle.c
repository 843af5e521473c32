/*
 * oracle/atlj_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU restatement of the Lennard-Jones pair-force hot path of
 * Allen-Tildesley/examples.  It exists to CHECK the CUDA path (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).
 * Nothing under examples_b200/ may import, link or call it.
 *
 * Parity pin: the reference holds no golden vectors of its own,
 * so this restatement is pinned against outputs of the reference Python
 * modules imported in the build container (oracle/make_golden.py ->
 * tests/golden/*.npz) and against the known-answer values of SURVEY.md
 * Appendix B.  The Fortran twins cannot be compiled here (no gfortran), so
 * there is no oracle/_ref binary; see DESIGN.md.
 *
 * Semantics follow the PYTHON reference where Python and Fortran differ:
 *   - rint() = round-half-even         (md_lj_ll_module.py:88 vs ANINT in .f90:160)
 *   - out-of-range cell index = error  (md_lj_ll_module.py:109 vs clamp in
 *                                       link_list_module.f90:146-148)
 *   - atoms inside a cell in ASCENDING atom index (md_lj_ll_module.py:116-120;
 *     the Fortran head/list walk of link_list_module.f90:152-161 is descending)
 * The traversal (cells c0 outermost .. c2 innermost, 14/17-entry half stencil in
 * the reference's order, i<j inside the home cell) is the scalar form
 * md_lj_ll_module.py:173-215 == md_lj_ll_module.f90:168-231.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, optional -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ATO_OK 0
#define ATO_ERR_SMALL -1 /* 'System is too small for cells' (md_lj_ll_module.py:107) */
#define ATO_ERR_INDEX -2 /* 'Index error'                   (md_lj_ll_module.py:109) */
#define ATO_ERR_ALLOC -3

static const double SR2_OVR = 1.77; /* md_lj_module.py:79 */

/* md_lj_ll_module.py:84-86 == md_lj_ll_module.f90:143-148 */
static const int D_LJ[14][3] = {{0, 0, 0},  {1, 0, 0},  {1, 1, 0},   {-1, 1, 0}, {0, 1, 0},
                                {0, 0, 1},  {-1, 0, 1}, {1, 0, 1},   {-1, -1, 1}, {0, -1, 1},
                                {1, -1, 1}, {-1, 1, 1}, {0, 1, 1},   {1, 1, 1}};
/* md_lj_llle_module.py:88-92 */
static const int D_LE[17][3] = {{0, 0, 0},   {1, 0, 0},  {1, 0, 1},  {-1, 0, 1}, {0, 0, 1},  {1, 1, -1},
                                {1, 1, 0},   {1, 1, 1},  {0, 1, -1}, {0, 1, 0},  {0, 1, 1},  {-1, 1, -1},
                                {-1, 1, 0},  {-1, 1, 1}, {-2, 1, -1}, {-2, 1, 0}, {-2, 1, 1}};

int ato_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py sets the team size explicitly: a launcher (torch.distributed.run) exports OMP_NUM_THREADS=1 */
void ato_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static inline int imod(int a, int m) {
  int r = a % m;
  return r < 0 ? r + m : r;
}

/* ---- host scalars, same expressions as md_lj_module.py:80-88 ------------------------- */
void ato_lj_scalars(double box, double r_cut, double *r_cut_box_sq, double *box_sq, double *pot_cut) {
  double r_cut_box = r_cut / box;
  *r_cut_box_sq = r_cut_box * r_cut_box;
  *box_sq = box * box;
  double sr2 = 1.0 / (r_cut * r_cut);
  double sr6 = sr2 * sr2 * sr2;
  double sr12 = sr6 * sr6;
  *pot_cut = sr12 - sr6;
}

/* ---- a4/a5: wrap, cell index triplets, CSR cell lists ------------------------------- */
/* rw[n*3] receives r - rint(r) (md_lj_ll_module.py:88); c[n*3] the int64 cell triplets
 * (md_lj_ll_module.py:108); perm/cell_start the per-cell atom lists in C order of
 * (c0,c1,c2), ascending atom index inside a cell (md_lj_ll_module.py:116-120).
 * le != 0 applies the Lees-Edwards pre-wrap of md_lj_llle_module.py:94 first. */
int ato_cells(int64_t n, const double *r, int sc, int le, double strain, double *rw, int64_t *c,
              int32_t *perm, int32_t *cell_start) {
  if (sc < 3) return ATO_ERR_SMALL;
  int64_t ncell = (int64_t)sc * sc * sc;
  int bad = 0;
  for (int64_t i = 0; i < n; i++) {
    double x = r[3 * i], y = r[3 * i + 1], z = r[3 * i + 2];
    if (le) x = x - nearbyint(y) * strain;
    double w[3] = {x - nearbyint(x), y - nearbyint(y), z - nearbyint(z)};
    for (int k = 0; k < 3; k++) {
      rw[3 * i + k] = w[k];
      double t = (w[k] + 0.5) * (double)sc;
      int64_t ck = (int64_t)floor(t);
      c[3 * i + k] = ck;
      if (ck < 0 || ck >= sc) bad = 1;
    }
  }
  if (bad) return ATO_ERR_INDEX;
  memset(cell_start, 0, (size_t)(ncell + 1) * sizeof(int32_t));
  for (int64_t i = 0; i < n; i++) {
    int64_t id = (c[3 * i] * sc + c[3 * i + 1]) * sc + c[3 * i + 2];
    cell_start[id + 1]++;
  }
  for (int64_t k = 0; k < ncell; k++) cell_start[k + 1] += cell_start[k];
  int32_t *fill = (int32_t *)malloc((size_t)ncell * sizeof(int32_t));
  if (!fill) return ATO_ERR_ALLOC;
  memcpy(fill, cell_start, (size_t)ncell * sizeof(int32_t));
  for (int64_t i = 0; i < n; i++) { /* ascending i => stable */
    int64_t id = (c[3 * i] * sc + c[3 * i + 1]) * sc + c[3 * i + 2];
    perm[fill[id]++] = (int32_t)i;
  }
  free(fill);
  return ATO_OK;
}

/* a6: neighbour-cell ids of one cell, reference order (md_lj_ll_module.py:127-129,
 * md_lj_llle_module.py:131-141).  Returns the number written (14, or 17 for an LE
 * top-layer cell). */
int ato_neighbours(int sc, int ci1, int le, int shift, int32_t *out) {
  int c0 = ci1 / (sc * sc), c1 = (ci1 / sc) % sc, c2 = ci1 % sc;
  if (!le) {
    for (int k = 0; k < 14; k++)
      out[k] = (imod(c0 + D_LJ[k][0], sc) * sc + imod(c1 + D_LJ[k][1], sc)) * sc + imod(c2 + D_LJ[k][2], sc);
    return 14;
  }
  int top = (c1 == sc - 1);
  int nk = top ? 17 : 14;
  for (int k = 0; k < nk; k++) {
    int d0 = D_LE[k][0];
    if (top && k >= 5) d0 -= shift;
    out[k] = (imod(c0 + d0, sc) * sc + imod(c1 + D_LE[k][1], sc)) * sc + imod(c2 + D_LE[k][2], sc);
  }
  return nk;
}

/* ---- a7: the pair arithmetic of SURVEY.md Appendix A ---------------------------------- */
typedef struct {
  double a, pot, vir, lap; /* a = cut (LJ) or pyx (LE) */
  int ovr;
  int64_t npair;
} acc_t;

/* One pair i,j; adds to acc and to fi / fj.  le: WCA + strain image. */
static inline void pair_term(const double *ri, const double *rj, int le, double strain, double r_cut_box_sq,
                             double box, double box_sq, double pot_cut, acc_t *acc, double *fi, double *fj) {
  double dx = ri[0] - rj[0], dy = ri[1] - rj[1], dz = ri[2] - rj[2];
  if (le) dx = dx - nearbyint(dy) * strain; /* md_lj_le_module.py:94 */
  dx = dx - nearbyint(dx);
  dy = dy - nearbyint(dy);
  dz = dz - nearbyint(dz);
  double s = (dx * dx + dy * dy) + dz * dz;
  if (!(s < r_cut_box_sq)) return;
  s = s * box_sq;
  dx = dx * box;
  dy = dy * box;
  dz = dz * box;
  double sr2 = 1.0 / s;
  if (sr2 > SR2_OVR) acc->ovr = 1;
  double sr6 = sr2 * sr2 * sr2;
  double sr12 = sr6 * sr6;
  double cut = sr12 - sr6;
  double vir = cut + sr12;
  double lap = (22.0 * sr12 - 5.0 * sr6) * sr2;
  double fs = vir * sr2;
  double fx = dx * fs, fy = dy * fs, fz = dz * fs;
  if (le) {
    acc->pot += cut + 0.25; /* md_lj_le_module.py:106 */
    acc->a += dy * fx;      /* pyx, md_lj_le_module.py:110 */
  } else {
    acc->pot += cut - pot_cut;
    acc->a += cut;
  }
  acc->vir += vir;
  acc->lap += lap;
  acc->npair++;
  fi[0] += fx;
  fi[1] += fy;
  fi[2] += fz;
  fj[0] -= fx;
  fj[1] -= fy;
  fj[2] -= fz;
}

/* a8 final scaling, md_lj_module.py:143-147 / md_lj_le_module.py:143-148 */
static void finish(int64_t n, double *f, const acc_t *acc, int le, double *totals, int *ovr, int64_t *npair) {
  for (int64_t k = 0; k < 3 * n; k++) f[k] = f[k] * 24.0;
  if (le) {
    totals[0] = acc->pot * 4.0;
    totals[1] = acc->vir * 24.0 / 3.0;
    totals[2] = acc->lap * 24.0 * 2.0;
    totals[3] = acc->a * 24.0; /* pyx */
  } else {
    totals[0] = acc->a * 4.0; /* cut */
    totals[1] = acc->pot * 4.0;
    totals[2] = acc->vir * 24.0 / 3.0;
    totals[3] = acc->lap * 24.0 * 2.0;
  }
  *ovr = acc->ovr;
  if (npair) *npair = acc->npair;
}

/* a3 / a10: all pairs, md_lj_module.py:117-140 / md_lj_le_module.py:117-141.
 * totals = {cut,pot,vir,lap} (LJ) or {pot,vir,lap,pyx} (LE). */
int ato_force_allpairs(int64_t n, const double *r, double box, double r_cut, int le, double strain, double *f,
                       double *totals, int *ovr, int64_t *npair) {
  double rcbsq, bsq, pc;
  ato_lj_scalars(box, r_cut, &rcbsq, &bsq, &pc);
  acc_t acc = {0, 0, 0, 0, 0, 0};
  memset(f, 0, (size_t)(3 * n) * sizeof(double));
  for (int64_t i = 0; i + 1 < n; i++)
    for (int64_t j = i + 1; j < n; j++)
      pair_term(r + 3 * i, r + 3 * j, le, strain, rcbsq, box, bsq, pc, &acc, f + 3 * i, f + 3 * j);
  finish(n, f, &acc, le, totals, ovr, npair);
  return ATO_OK;
}

/* a7 / a11 over cell lists.  Serial build follows the reference order exactly;
 * with OpenMP (baseline timing only) i-cells are distributed over threads with a
 * private force array per thread, as md_lj_omp_module.f90:174-178 does. */
int ato_force_cells(int64_t n, const double *r, double box, double r_cut, int le, double strain, double *f,
                    double *totals, int *ovr, int64_t *npair) {
  double rcbsq, bsq, pc;
  ato_lj_scalars(box, r_cut, &rcbsq, &bsq, &pc);
  int sc = (int)floor(box / r_cut);
  if (sc < 3) return ATO_ERR_SMALL;
  int64_t ncell = (int64_t)sc * sc * sc;
  double *rw = (double *)malloc((size_t)(3 * n) * sizeof(double));
  int64_t *c = (int64_t *)malloc((size_t)(3 * n) * sizeof(int64_t));
  int32_t *perm = (int32_t *)malloc((size_t)n * sizeof(int32_t));
  int32_t *cs = (int32_t *)malloc((size_t)(ncell + 1) * sizeof(int32_t));
  if (!rw || !c || !perm || !cs) return ATO_ERR_ALLOC;
  int rc = ato_cells(n, r, sc, le, strain, rw, c, perm, cs);
  if (rc != ATO_OK) {
    free(rw), free(c), free(perm), free(cs);
    return rc;
  }
  int shift = le ? (int)floor(strain * sc) : 0; /* md_lj_llle_module.py:112 */
  acc_t acc = {0, 0, 0, 0, 0, 0};
  int nthr = ato_num_threads();
  double *fpriv = (double *)calloc((size_t)(3 * n) * (size_t)nthr, sizeof(double));
  if (!fpriv) return ATO_ERR_ALLOC;
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double *ft = fpriv + (size_t)tid * (size_t)(3 * n);
    acc_t a = {0, 0, 0, 0, 0, 0};
    int32_t nb[17];
#pragma omp for schedule(static, 1)
    for (int64_t ci1 = 0; ci1 < ncell; ci1++) {
      int nk = ato_neighbours(sc, (int)ci1, le, shift, nb);
      for (int32_t ii = cs[ci1]; ii < cs[ci1 + 1]; ii++) {
        int32_t ki = perm[ii];
        for (int k = 0; k < nk; k++) {
          int32_t cj1 = nb[k];
          int32_t j0 = (cj1 == ci1) ? ii + 1 : cs[cj1]; /* look up-list only in the home cell */
          int32_t j1 = (cj1 == ci1) ? cs[ci1 + 1] : cs[cj1 + 1];
          if (cj1 == ci1 && k != 0) continue; /* cannot happen for sc>=3 (LJ) / sc>=4 (LE) */
          for (int32_t jj = j0; jj < j1; jj++) {
            int32_t kj = perm[jj];
            pair_term(rw + 3 * ki, rw + 3 * kj, le, strain, rcbsq, box, bsq, pc, &a, ft + 3 * ki, ft + 3 * kj);
          }
        }
      }
    }
#pragma omp critical
    {
      acc.a += a.a, acc.pot += a.pot, acc.vir += a.vir, acc.lap += a.lap;
      acc.ovr |= a.ovr, acc.npair += a.npair;
    }
  }
  for (int64_t k = 0; k < 3 * n; k++) {
    double s = 0.0;
    for (int t = 0; t < nthr; t++) s += fpriv[(size_t)t * (size_t)(3 * n) + k];
    f[k] = s;
  }
  finish(n, f, &acc, le, totals, ovr, npair);
  free(fpriv), free(rw), free(c), free(perm), free(cs);
  return ATO_OK;
}

/* ---- a9: hessian, md_lj_module.py:197-211 (all pairs) / md_lj_ll_module.py:333-355 -------- */
static inline double hess_term(const double *ri, const double *rj, const double *fi, const double *fj,
                               double r_cut_box_sq, double box, double box_sq) {
  double dx = ri[0] - rj[0], dy = ri[1] - rj[1], dz = ri[2] - rj[2];
  dx = dx - nearbyint(dx);
  dy = dy - nearbyint(dy);
  dz = dz - nearbyint(dz);
  double s = (dx * dx + dy * dy) + dz * dz;
  if (!(s < r_cut_box_sq)) return 0.0;
  s = s * box_sq;
  dx *= box, dy *= box, dz *= box;
  double gx = fi[0] - fj[0], gy = fi[1] - fj[1], gz = fi[2] - fj[2];
  double ff = (gx * gx + gy * gy) + gz * gz;
  double rf = (dx * gx + dy * gy) + dz * gz;
  double sr2 = 1.0 / s;
  double sr6 = sr2 * sr2 * sr2;
  double sr8 = sr6 * sr2;
  double sr10 = sr8 * sr2;
  double v1 = 24.0 * (1.0 - 2.0 * sr6) * sr8;
  double v2 = 96.0 * (7.0 * sr6 - 2.0) * sr10;
  return v1 * ff + v2 * (rf * rf);
}

int ato_hessian_allpairs(int64_t n, const double *r, const double *f, double box, double r_cut, double *hes) {
  double rcbsq, bsq, pc;
  ato_lj_scalars(box, r_cut, &rcbsq, &bsq, &pc);
  double h = 0.0;
  for (int64_t i = 0; i + 1 < n; i++)
    for (int64_t j = i + 1; j < n; j++) h += hess_term(r + 3 * i, r + 3 * j, f + 3 * i, f + 3 * j, rcbsq, box, bsq);
  *hes = h;
  return ATO_OK;
}

int ato_hessian_cells(int64_t n, const double *r, const double *f, double box, double r_cut, double *hes) {
  double rcbsq, bsq, pc;
  ato_lj_scalars(box, r_cut, &rcbsq, &bsq, &pc);
  int sc = (int)floor(box / r_cut);
  if (sc < 3) return ATO_ERR_SMALL;
  int64_t ncell = (int64_t)sc * sc * sc;
  double *rw = (double *)malloc((size_t)(3 * n) * sizeof(double));
  int64_t *c = (int64_t *)malloc((size_t)(3 * n) * sizeof(int64_t));
  int32_t *perm = (int32_t *)malloc((size_t)n * sizeof(int32_t));
  int32_t *cs = (int32_t *)malloc((size_t)(ncell + 1) * sizeof(int32_t));
  if (!rw || !c || !perm || !cs) return ATO_ERR_ALLOC;
  int rc = ato_cells(n, r, sc, 0, 0.0, rw, c, perm, cs);
  if (rc != ATO_OK) {
    free(rw), free(c), free(perm), free(cs);
    return rc;
  }
  double h = 0.0;
#pragma omp parallel for schedule(static, 1) reduction(+ : h)
  for (int64_t ci1 = 0; ci1 < ncell; ci1++) {
    int32_t nb[17];
    int nk = ato_neighbours(sc, (int)ci1, 0, 0, nb);
    for (int32_t ii = cs[ci1]; ii < cs[ci1 + 1]; ii++) {
      int32_t ki = perm[ii];
      for (int k = 0; k < nk; k++) {
        int32_t cj1 = nb[k];
        if (cj1 == ci1 && k != 0) continue;
        int32_t j0 = (cj1 == ci1) ? ii + 1 : cs[cj1];
        int32_t j1 = (cj1 == ci1) ? cs[ci1 + 1] : cs[cj1 + 1];
        for (int32_t jj = j0; jj < j1; jj++) {
          int32_t kj = perm[jj];
          h += hess_term(rw + 3 * ki, rw + 3 * kj, f + 3 * ki, f + 3 * kj, rcbsq, box, bsq);
        }
      }
    }
  }
  *hes = h;
  free(rw), free(c), free(perm), free(cs);
  return ATO_OK;
}

/* ---- a12/a16: velocity-Verlet NVE loop around the cell force, md_nve_lj.py:178-192 -------
 * rec[step*6 + {0..5}] = cut, pot, vir, lap, sum v^2, sum f^2 after each step.
 * r (box units), v, f are updated in place; f must hold the forces of r on entry. */
int ato_nve_steps(int64_t n, double *r, double *v, double *f, double box, double r_cut, double dt, int nsteps,
                  int use_cells, double *rec, int *ovr_any) {
  double totals[4];
  int ovr = 0;
  *ovr_any = 0;
  for (int s = 0; s < nsteps; s++) {
    for (int64_t k = 0; k < 3 * n; k++) {
      v[k] = v[k] + 0.5 * dt * f[k];
      double x = r[k] + dt * v[k] / box;
      r[k] = x - nearbyint(x);
    }
    int rc = use_cells ? ato_force_cells(n, r, box, r_cut, 0, 0.0, f, totals, &ovr, NULL)
                       : ato_force_allpairs(n, r, box, r_cut, 0, 0.0, f, totals, &ovr, NULL);
    if (rc != ATO_OK) return rc;
    *ovr_any |= ovr;
    double vsq = 0.0, fsq = 0.0;
    for (int64_t k = 0; k < 3 * n; k++) {
      v[k] = v[k] + 0.5 * dt * f[k];
      vsq += v[k] * v[k];
      fsq += f[k] * f[k];
    }
    if (rec) {
      rec[6 * s + 0] = totals[0], rec[6 * s + 1] = totals[1], rec[6 * s + 2] = totals[2];
      rec[6 * s + 3] = totals[3], rec[6 * s + 4] = vsq, rec[6 * s + 5] = fsq;
    }
  }
  return ATO_OK;
}
