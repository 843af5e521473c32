// Thermostatted / sheared integrators kept resident on the device.
//
// Reference loop bodies: bd_nvt_lj.py:91-127,221-233 (BAOAB Langevin), md_nvt_lj.py:101-159,270-288 (Nose-Hoover
// chain), md_nvt_lj_le.py:84-129,226-240 (isokinetic SLLOD with Lees-Edwards boundaries).
// Every elementwise statement between two force calls / global reductions is fused into one HBM pass; the global
// sums are two-stage and fixed-order (deterministic); the handful of scalar coefficients that depend on those sums
// (NHC chain, SLLOD c1,c2,g / alpha,beta,h,e) are computed by one-thread kernels on the device so a step never
// waits for the host.  The elementwise arithmetic keeps NumPy's operation order (separately rounded _rn ops).
#include <math.h>

#include "common.cuh"
#include "integrate.cuh"
#include "sim.cuh"

namespace atlj {

// ---- counter-based Gaussian noise: Philox4x32-10 + Box-Muller ---------------------------------------------------
// (replaces np.random.randn(n,3) of bd_nvt_lj.py:127; keyed by (seed), counter (atom id, step, draw) so the noise
//  of an atom does not depend on which GPU or in which order it is processed)
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                              unsigned k1, unsigned out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

__device__ __forceinline__ double u53(unsigned hi, unsigned lo) {  // uniform in (0,1), 53 bits
  const unsigned long long x = ((unsigned long long)hi << 21) ^ (unsigned long long)(lo >> 11);
  return ((double)(x & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void gauss3(unsigned long long seed, unsigned id, unsigned step, double g[3]) {
  unsigned a[4], b[4];
  philox4x32_10(id, step, 0u, 0u, (unsigned)seed, (unsigned)(seed >> 32), a);
  philox4x32_10(id, step, 1u, 0u, (unsigned)seed, (unsigned)(seed >> 32), b);
  const double r0 = sqrt(-2.0 * log(u53(a[0], a[1]))), r1 = sqrt(-2.0 * log(u53(b[0], b[1])));
  double s0, c0, s1, c1;
  sincospi(2.0 * u53(a[2], a[3]), &s0, &c0);
  sincospi(2.0 * u53(b[2], b[3]), &s1, &c1);
  g[0] = r0 * c0, g[1] = r0 * s0, g[2] = r1 * c1;
  (void)s1;
}

// ---- BAOAB: B(t) A(t) O(2t) A(t) fused, one thread per atom (bd_nvt_lj.py:223-226) ---------------------------------
__global__ void __launch_bounds__(IT) baoab_kernel(int n, double *__restrict__ r, double *__restrict__ v,
                                                   const double *__restrict__ f, const int *__restrict__ id, double t,
                                                   BoxDiv box, double decay, double amp,
                                                   const double *__restrict__ noise, unsigned long long seed,
                                                   unsigned step, const int *__restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);  // slab ranks keep the owned-atom count on the device
  for (int i = blockIdx.x * IT + threadIdx.x; i < n; i += gridDim.x * IT) {
    double g[3];
    const int gid = id[i];
    if (noise) {
      g[0] = noise[3 * (size_t)gid], g[1] = noise[3 * (size_t)gid + 1], g[2] = noise[3 * (size_t)gid + 2];
    } else {
      gauss3(seed, (unsigned)gid, step, g);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const size_t e = 3 * (size_t)i + k;
      double vv = __dadd_rn(v[e], __dmul_rn(t, f[e]));                          // b_propagator, :109
      double x = wrap1(__dadd_rn(r[e], div_box(__dmul_rn(t, vv), box)));      // a_propagator, :100-101
      vv = __dadd_rn(__dmul_rn(vv, decay), __dmul_rn(amp, g[k]));               // o_propagator, :127
      x = wrap1(__dadd_rn(x, div_box(__dmul_rn(t, vv), box)));                // a_propagator
      v[e] = vv;
      r[e] = x;
    }
  }
}

// ---- Nose-Hoover chain scalars (md_nvt_lj.py:119-159), one thread -------------------------------------------------
// st[0] = sum v^2 of the logical velocities, st[1] = pending velocity scale (applied by the next kick kernel),
// st[2..2+m) = eta, st[10..10+m) = p_eta, st[18..18+m) = q.
__device__ double exprel_neg(double x) {  // scipy.special.exprel(-x) = (1 - exp(-x)) / x
  return fabs(x) < 1e-16 ? 1.0 : -expm1(-x) / x;
}
__device__ void nhc_u4(double *st, int m, double t, double temperature, double g, bool inwards) {
  double *p_eta = st + 10;
  const double *q = st + 18;
  for (int jj = 0; jj < m; jj++) {
    const int j = inwards ? m - 1 - jj : jj;
    const double gj = j == 0 ? st[0] - g * temperature : (p_eta[j - 1] * p_eta[j - 1] / q[j - 1]) - temperature;
    if (j == m - 1) {
      p_eta[j] = p_eta[j] + t * gj;
    } else {
      const double x = t * p_eta[j + 1] / q[j + 1];
      const double c = exprel_neg(x);
      p_eta[j] = p_eta[j] * exp(-x) + t * gj * c;
    }
  }
}
__global__ void nhc_chain_kernel(double *st, int m, double dt, double temperature, double g, const double *ksq_in,
                                 double *rec) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (ksq_in) st[0] = *ksq_in;  // fresh sum v^2 from the kick kernel
  nhc_u4(st, m, 0.25 * dt, temperature, g, true);
  {  // u3 (md_nvt_lj.py:129-130): the elementwise v *= s is applied lazily by the next kernel that touches v
    const double t = 0.5 * dt;
    const double s = exp(-t * st[10] / st[18]);
    for (int j = 0; j < m; j++) st[2 + j] += t * st[10 + j] / st[18 + j];
    st[0] *= s * s;
    st[1] *= s;
  }
  nhc_u4(st, m, 0.25 * dt, temperature, g, false);
  if (rec) {
    double ext = 0.0, es = 0.0;
    for (int j = 0; j < m; j++) ext += 0.5 * st[10 + j] * st[10 + j] / st[18 + j];
    for (int j = 1; j < m; j++) es += st[2 + j];
    rec[4] = st[0];                                        // sum v^2 after the last thermostat scaling
    rec[10] = ext + temperature * (g * st[2] + es);        // md_nvt_lj.py:44
  }
}

// v *= *scale ; *scale = 1 afterwards (flush of the pending thermostat factor)
__global__ void __launch_bounds__(IT) scale_v_kernel(long long n3, double *__restrict__ v, const double *scale) {
  const double s = *scale;
  for (long long t = (long long)blockIdx.x * IT + threadIdx.x; t < n3; t += (long long)gridDim.x * IT)
    v[t] = __dmul_rn(v[t], s);
}
__global__ void set_scalar_kernel(double *p, double val) { *p = val; }

// ---- SLLOD (md_nvt_lj_le.py:84-129) ------------------------------------------------------------------------------
// The coefficients of b1 (g) and b2 (prefactor, dt_factor) depend on three global sums each.  The kernel that needs a
// coefficient forms it itself from the per-CTA rows its predecessor left behind (every CTA adds up the same <= 296 rows
// in the same fixed order, so all CTAs get the same bits): no reduction kernel and no one-thread kernel in between.
constexpr int SL_CTAS = 296;  // CTAs of the SLLOD kernels = rows of partial sums (2 per SM)

// out[k] = sum over rows of partials[row][k], k < 3; every thread of the CTA returns the same values
__device__ __forceinline__ void sllod_reduce_rows(const double *__restrict__ partials, int rows, double out[3]) {
  __shared__ double red[IT / 32][3];
  __shared__ double tot[3];
  double a[3] = {0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < rows; b += IT) {
#pragma unroll
    for (int k = 0; k < 3; k++) a[k] += __ldcg(partials + 4 * (size_t)b + k);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) red[threadIdx.x >> 5][k] = a[k];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < IT / 32; w++) t += red[w][threadIdx.x];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  out[0] = tot[0], out[1] = tot[1], out[2] = tot[2];
  __syncthreads();  // red / tot are reused by block_reduce_store below
}
// g of b1_propagator from (Sxy, Syy, Svv), :108-111
__device__ __forceinline__ double sllod_b1_g(const double sums[3], double x) {
  const double c1 = x * sums[0] / sums[2];
  const double c2 = (x * x) * sums[1] / sums[2];
  return 1.0 / sqrt(1.0 - 2.0 * c1 + c2);
}
// (prefactor, dt_factor) of b2_propagator from (Sfv, Sff, Svv), :120-127
__device__ __forceinline__ void sllod_b2_pd(const double sums[3], double t, double &pre, double &dtf) {
  const double alpha = sums[0] / sums[2];
  const double beta = sqrt(sums[1] / sums[2]);
  const double h = (alpha + beta) / (alpha - beta);
  const double e = exp(-beta * t);
  dtf = (1.0 + h - e - h / e) / ((1.0 - h) * beta);
  pre = (1.0 - h) / (e - h / e);
}

// a(t) and b1(t) on one atom; b1_first selects the order (b1 then a at the end of a step, a then b1 at its start)
__global__ void __launch_bounds__(IT) sllod_ab1_kernel(int n, double *__restrict__ r, double *__restrict__ v, double t,
                                                       double x, double strain_new, BoxDiv box,
                                                       const double *__restrict__ rows_in, int nrows, int b1_first,
                                                       double *__restrict__ partials) {
  double sums[3];
  sllod_reduce_rows(rows_in, nrows, sums);  // Sxy, Syy, Svv of the current velocities
  const double g = sllod_b1_g(sums, x);
  double s[3] = {0.0, 0.0, 0.0};  // sum vx*vy, sum vy^2, sum v^2 of the velocities left behind
  for (int i = blockIdx.x * IT + threadIdx.x; i < n; i += gridDim.x * IT) {
    const size_t e = 3 * (size_t)i;
    double vx = v[e], vy = v[e + 1], vz = v[e + 2];
    double rx = r[e], ry = r[e + 1], rz = r[e + 2];
    if (b1_first) {  // b1_propagator, :112-113
      vx = __dsub_rn(vx, __dmul_rn(x, vy));
      vx = __dmul_rn(g, vx), vy = __dmul_rn(g, vy), vz = __dmul_rn(g, vz);
    }
    // a_propagator, :92-98
    rx = __dadd_rn(rx, __dmul_rn(x, ry));
    rx = __dadd_rn(rx, div_box(__dmul_rn(t, vx), box));
    ry = __dadd_rn(ry, div_box(__dmul_rn(t, vy), box));
    rz = __dadd_rn(rz, div_box(__dmul_rn(t, vz), box));
    rx = __dsub_rn(rx, __dmul_rn(rint(ry), strain_new));
    rx = wrap1(rx), ry = wrap1(ry), rz = wrap1(rz);
    if (!b1_first) {
      vx = __dsub_rn(vx, __dmul_rn(x, vy));
      vx = __dmul_rn(g, vx), vy = __dmul_rn(g, vy), vz = __dmul_rn(g, vz);
    }
    r[e] = rx, r[e + 1] = ry, r[e + 2] = rz;
    v[e] = vx, v[e + 1] = vy, v[e + 2] = vz;
    s[0] += vx * vy;
    s[1] += vy * vy;
    s[2] += vx * vx + vy * vy + vz * vz;
  }
  block_reduce_store<3>(s, partials + 4 * (size_t)blockIdx.x);
}

// sums for b2: sum f.v, sum f^2, sum v^2
__global__ void __launch_bounds__(IT) sllod_fv_sums_kernel(long long n3, const double *__restrict__ v,
                                                           const double *__restrict__ f,
                                                           double *__restrict__ partials, FinTail fin) {
  // the extra CTA (fin.nb > 0): the pair totals of the force evaluation that has just ended
  const int nblk = (int)gridDim.x - (fin.nb > 0 ? 1 : 0);
  if ((int)blockIdx.x == nblk) {
    finalize_rows_block<IT>(fin);
    return;
  }
  double s[3] = {0.0, 0.0, 0.0};
  for (long long t = (long long)blockIdx.x * IT + threadIdx.x; t < n3; t += (long long)nblk * IT) {
    const double vv = v[t], ff = f[t];
    s[0] += ff * vv;
    s[1] += ff * ff;
    s[2] += vv * vv;
  }
  block_reduce_store<3>(s, partials + 4 * (size_t)blockIdx.x);
}

// b2_propagator, :129: v = prefactor * (v + dt_factor * f); sums sum vx*vy, sum vy^2, sum v^2 of the new v
// rows_in = the rows of sllod_fv_sums_kernel; nullptr: identity (plain reduction of the current velocities);
// sums_f_out (CTA 0): the (Sfv, Sff, Svv) the coefficients came from, for the step record
__global__ void __launch_bounds__(IT) sllod_b2_kernel(int n, double *__restrict__ v, const double *__restrict__ f,
                                                      const double *__restrict__ rows_in, int nrows, double t,
                                                      double *__restrict__ partials, double *__restrict__ sums_f_out) {
  double pre = 1.0, dtf = 0.0;
  if (rows_in) {
    double sums[3];
    sllod_reduce_rows(rows_in, nrows, sums);
    sllod_b2_pd(sums, t, pre, dtf);
    if (blockIdx.x == 0 && threadIdx.x < 3 && sums_f_out) sums_f_out[threadIdx.x] = sums[threadIdx.x];
  }
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * IT + threadIdx.x; i < n; i += gridDim.x * IT) {
    const size_t e = 3 * (size_t)i;
    const double vx = __dmul_rn(pre, __dadd_rn(v[e], __dmul_rn(dtf, f[e])));
    const double vy = __dmul_rn(pre, __dadd_rn(v[e + 1], __dmul_rn(dtf, f[e + 1])));
    const double vz = __dmul_rn(pre, __dadd_rn(v[e + 2], __dmul_rn(dtf, f[e + 2])));
    v[e] = vx, v[e + 1] = vy, v[e + 2] = vz;
    s[0] += vx * vy;
    s[1] += vy * vy;
    s[2] += vx * vx + vy * vy + vz * vz;
  }
  block_reduce_store<3>(s, partials + 4 * (size_t)blockIdx.x);
}

__global__ void __launch_bounds__(IT) sllod_rec_kernel(const double *__restrict__ rows_v, int nrows,
                                                       const double *sums_f /*Sfv,Sff,Svv*/, double strain,
                                                       double *rec) {
  double sv[3];
  sllod_reduce_rows(rows_v, nrows, sv);  // Sxy, Syy, Svv after the closing a(dt/2)
  if (threadIdx.x != 0) return;
  rec[4] = sv[2];
  rec[5] = sums_f[1];
  rec[9] = sv[0];
  rec[10] = strain;
}

static int atom_blocks(int64_t n) { return integ_blocks(n); }

}  // namespace atlj

using namespace atlj;

// scalar slots in s->scal (64 doubles)
enum { SC_NHC = 0, SC_SUMV = 32, SC_SUMF = 36, SC_COEF = 40 };

static int baoab_run(atlj_sim *s, double dt, double decay, double amp, const double *noise_dev, uint64_t seed,
                     int nsteps, double *rec_host) {
  int rc;
  if ((rc = sim_ensure_forces(s, 0.0))) return rc;
  const int n = (int)s->n;
  const long long n3 = 3 * s->n;
  const double t = dt / 2;  // bd_nvt_lj.py:223
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    Timer::begin(s, 2);
    baoab_kernel<<<atom_blocks(n), IT, 0, s->st>>>(n, s->r[s->cur], s->v[s->cur], s->f, s->id[s->cur], t, make_box_div(s->p.box),
                                                   decay, amp,
                                                   noise_dev ? noise_dev + (size_t)k * 3 * (size_t)s->n_total : nullptr,
                                                   seed, (unsigned)(s->step_count + k), nullptr);
    ATLJ_LAUNCHED();
    Timer::end(s);
    s->defer_fin = true;  // the half kick below finishes the force totals in an extra CTA (FinTail)
    if ((rc = sim_force(s, 0.0, false, rec, false))) return rc;
    Timer::begin(s, 2);
    if ((rc = launch_kick_sums(s->st, n3, s->v[s->cur], s->f, t, s->vf_partials, rec + 4, s->cb.err + 4, nullptr, nullptr,
                               s->fin.nb > 0 ? &s->fin : nullptr)))
      return rc;
    s->fin = FinTail();
    Timer::end(s);
  }
  s->step_count += nsteps;
  return sim_copy_rec(s, rec_host, nsteps);
}

// ---- the thermostats on slab ranks (widened) -------------------------------------------------------
// The integrator's own elementwise kernels run on the owned atoms (count on the device); mg.cu then bins them,
// migrates the leavers, sorts, sends the halo layers and runs the pair kernel.  BAOAB draws its noise from Philox keyed
// by (seed; GLOBAL atom id, step), so a trajectory does not depend on the decomposition; the Nose-Hoover chain needs
// the GLOBAL sum v^2 once per step: the step record is all-reduced (NCCL, 96 bytes) before the chain kernel reads it.
static int baoab_move(atlj_sim *s, double dt, double decay, double amp, const double *noise_dev, uint64_t seed,
                      unsigned step) {
  const int cap = (int)mg_own_cap(s);
  baoab_kernel<<<atom_blocks(cap), IT, 0, s->st>>>(cap, s->r[s->cur], s->v[s->cur], s->f, s->id[s->cur], dt / 2,
                                                   make_box_div(s->p.box), decay, amp, noise_dev, seed, step, mg_n_dev(s));
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

static int nhc_open(atlj_sim *s, int m, const double *eta, const double *p_eta, const double *q) {
  double h[32];
  memset(h, 0, sizeof(h));
  h[1] = 1.0;
  for (int j = 0; j < m; j++) h[2 + j] = eta[j], h[10 + j] = p_eta[j], h[18 + j] = q[j];
  ATLJ_CUDA(cudaMemcpyAsync(s->scal + SC_NHC, h, sizeof(h), cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));  // h is on the stack
  return ATLJ_OK;
}
// chain (first call of a run: with the global sum v^2 in *ksq) + u2(dt/2) u1(dt) with the pending factor
static int nhc_move(atlj_sim *s, double dt, double temperature, double g, int m, const double *ksq) {
  int rc;
  double *st = s->scal + SC_NHC;
  nhc_chain_kernel<<<1, 32, 0, s->st>>>(st, m, dt, temperature, g, ksq, nullptr);
  ATLJ_LAUNCHED();
  if ((rc = launch_kick_drift(s->st, 3 * mg_own_cap(s), s->r[s->cur], s->v[s->cur], s->f, 0.5 * dt, dt, s->p.box, st + 1,
                              mg_n_dev(s))))
    return rc;
  set_scalar_kernel<<<1, 1, 0, s->st>>>(st + 1, 1.0);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}
// the closing chain update from the GLOBAL record of the step (rec[4] = sum v^2 over all ranks)
static int nhc_close(atlj_sim *s, double dt, double temperature, double g, int m, double *rec) {
  nhc_chain_kernel<<<1, 32, 0, s->st>>>(s->scal + SC_NHC, m, dt, temperature, g, rec + 4, rec);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}
static int nhc_finish(atlj_sim *s, int m, double *eta, double *p_eta) {
  double *st = s->scal + SC_NHC;
  double h[32];
  scale_v_kernel<<<integ_blocks(3 * mg_own_cap(s)), IT, 0, s->st>>>(3 * (long long)s->n, s->v[s->cur], st + 1);
  ATLJ_LAUNCHED();
  set_scalar_kernel<<<1, 1, 0, s->st>>>(st + 1, 1.0);
  ATLJ_LAUNCHED();
  ATLJ_CUDA(cudaMemcpyAsync(h, st, sizeof(h), cudaMemcpyDeviceToHost, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  for (int j = 0; j < m; j++) eta[j] = h[2 + j], p_eta[j] = h[10 + j];
  return ATLJ_OK;
}

extern "C" {

int atlj_sim_run_baoab_mg(atlj_sim *s, double dt, double o_decay, double o_noise, uint64_t seed, int nsteps,
                          double *rec_host) {
  if (!s || !s->configured || !s->mg.on || nsteps < 0) return ATLJ_ERR_ARG;
  if (!mg_connected(s)) {
    set_error("run_baoab_mg: connect the ranks first");
    return ATLJ_ERR_ARG;
  }
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc || (rc = mg_begin(s))) return rc;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC * (size_t)(nsteps > 0 ? nsteps : 1), s->st));
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    if ((rc = baoab_move(s, dt, o_decay, o_noise, nullptr, seed, (unsigned)(s->step_count + k)))) return rc;
    s->defer_fin = true;  // the half kick below finishes the force totals in an extra CTA (FinTail)
    if ((rc = mg_assign_exchange_force(s, rec, k == 0))) return rc;
    if ((rc = launch_kick_sums(s->st, 3 * mg_own_cap(s), s->v[s->cur], s->f, dt / 2, s->vf_partials, rec + 4,
                               s->cb.err + 4, mg_n_dev(s), nullptr, s->fin.nb > 0 ? &s->fin : nullptr)))
      return rc;
    s->fin = FinTail();
  }
  if (nsteps > 0 && (rc = mg_allreduce(s, s->rec_dev, (size_t)nsteps * ATLJ_NREC))) return rc;
  s->step_count += nsteps;
  if ((rc = mg_end(s))) return rc;
  rc = sim_copy_rec(s, rec_host, nsteps);
  if (rec_host)
    for (int k = 0; k < nsteps; k++)
      if (rec_host[(size_t)k * ATLJ_NREC + 7] > 0.0) rec_host[(size_t)k * ATLJ_NREC + 7] = 1.0;
  return rc;
}

int atlj_sim_run_nhc_mg(atlj_sim *s, double dt, double temperature, double g, int m, double *eta, double *p_eta,
                        const double *q, int nsteps, double *rec_host) {
  if (!s || !s->configured || !s->mg.on || nsteps < 0 || m < 1 || m > 8 || !eta || !p_eta || !q) return ATLJ_ERR_ARG;
  if (!mg_connected(s)) {
    set_error("run_nhc_mg: connect the ranks first");
    return ATLJ_ERR_ARG;
  }
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc || (rc = mg_begin(s)) || (rc = nhc_open(s, m, eta, p_eta, q))) return rc;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC * (size_t)(nsteps > 0 ? nsteps : 1), s->st));
  // global sum v^2 of the starting velocities
  double *ksq = s->scal + SC_SUMV;
  if ((rc = launch_kick_sums(s->st, 3 * mg_own_cap(s), s->v[s->cur], s->f, 0.0, s->vf_partials, ksq, s->cb.err + 4,
                             mg_n_dev(s))) ||
      (rc = mg_allreduce(s, ksq, 2)))
    return rc;
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    if ((rc = nhc_move(s, dt, temperature, g, m, k == 0 ? ksq : nullptr))) return rc;
    s->defer_fin = true;
    if ((rc = mg_assign_exchange_force(s, rec, k == 0))) return rc;
    if ((rc = launch_kick_sums(s->st, 3 * mg_own_cap(s), s->v[s->cur], s->f, 0.5 * dt, s->vf_partials, rec + 4,
                               s->cb.err + 4, mg_n_dev(s), nullptr, s->fin.nb > 0 ? &s->fin : nullptr)))
      return rc;
    s->fin = FinTail();
    // the chain needs the GLOBAL sum v^2: the whole record of the step is summed over the ranks here
    if ((rc = mg_allreduce(s, rec, ATLJ_NREC)) || (rc = nhc_close(s, dt, temperature, g, m, rec))) return rc;
  }
  s->step_count += nsteps;
  if ((rc = mg_end(s)) || (rc = nhc_finish(s, m, eta, p_eta))) return rc;
  rc = sim_copy_rec(s, rec_host, nsteps);
  if (rec_host)
    for (int k = 0; k < nsteps; k++)
      if (rec_host[(size_t)k * ATLJ_NREC + 7] > 0.0) rec_host[(size_t)k * ATLJ_NREC + 7] = 1.0;
  return rc;
}

/* the same steps cut into phases for ranks emulated in one process (tests): <integrator>_phase1 on every rank, then
 * atlj_sim_mg_phase2 on every rank, then atlj_sim_mg_phase3 (pair kernel, closing half kick, LOCAL sums) on every rank;
 * Nose-Hoover: the caller sums the records over the ranks and hands the global record to atlj_sim_mg_nhc_close */
int atlj_sim_mg_phase1_baoab(atlj_sim *s, double dt, double o_decay, double o_noise, uint64_t seed, int first) {
  if (!s || !s->mg.on) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, 1);
  if (rc || (rc = mg_begin(s))) return rc;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC, s->st));
  if ((rc = baoab_move(s, dt, o_decay, o_noise, nullptr, seed, (unsigned)s->step_count))) return rc;
  s->step_count += 1;
  return mg_phase1_assign(s, first != 0);
}
int atlj_sim_mg_nhc_open(atlj_sim *s, int m, const double *eta, const double *p_eta, const double *q, double ksq_global) {
  if (!s || !s->mg.on || m < 1 || m > 8) return ATLJ_ERR_ARG;
  int rc = nhc_open(s, m, eta, p_eta, q);
  if (rc) return rc;
  ATLJ_CUDA(cudaMemcpyAsync(s->scal + SC_SUMV, &ksq_global, sizeof(double), cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  return ATLJ_OK;
}
int atlj_sim_mg_phase1_nhc(atlj_sim *s, double dt, double temperature, double g, int m, int first) {
  if (!s || !s->mg.on) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, 1);
  if (rc || (rc = mg_begin(s))) return rc;
  ATLJ_CUDA(cudaMemsetAsync(s->rec_dev, 0, sizeof(double) * ATLJ_NREC, s->st));
  if ((rc = nhc_move(s, dt, temperature, g, m, first ? s->scal + SC_SUMV : nullptr))) return rc;
  s->step_count += 1;
  return mg_phase1_assign(s, first != 0);
}
int atlj_sim_mg_nhc_close(atlj_sim *s, double dt, double temperature, double g, int m, double *rec_global_host) {
  if (!s || !s->mg.on || !rec_global_host) return ATLJ_ERR_ARG;
  ATLJ_CUDA(cudaMemcpyAsync(s->rec_dev, rec_global_host, sizeof(double) * ATLJ_NREC, cudaMemcpyHostToDevice, s->st));
  int rc = nhc_close(s, dt, temperature, g, m, s->rec_dev);
  if (rc) return rc;
  return sim_copy_rec(s, rec_global_host, 1);
}
int atlj_sim_mg_nhc_finish(atlj_sim *s, int m, double *eta, double *p_eta) {
  if (!s || !s->mg.on || m < 1 || m > 8) return ATLJ_ERR_ARG;
  return nhc_finish(s, m, eta, p_eta);
}

int atlj_sim_run_baoab(atlj_sim *s, double dt, double o_decay, double o_noise, uint64_t seed, int nsteps,
                       double *rec_host) {
  if (!s || !s->configured || nsteps < 0) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc) return rc;
  return baoab_run(s, dt, o_decay, o_noise, nullptr, seed, nsteps, rec_host);
}

int atlj_sim_run_baoab_noise(atlj_sim *s, double dt, double o_decay, double o_noise, const double *noise_host,
                             int nsteps, double *rec_host) {
  if (!s || !s->configured || nsteps < 0 || !noise_host) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc) return rc;
  const size_t need = (size_t)nsteps * 3 * (size_t)s->n_total;
  if ((int64_t)need > s->noise_cap) {
    cudaFree(s->noise_dev);
    s->noise_dev = nullptr;
    ATLJ_CUDA(cudaMalloc((void **)&s->noise_dev, sizeof(double) * (need ? need : 1)));
    s->noise_cap = (int64_t)need;
  }
  ATLJ_CUDA(cudaMemcpyAsync(s->noise_dev, noise_host, sizeof(double) * need, cudaMemcpyHostToDevice, s->st));
  return baoab_run(s, dt, o_decay, o_noise, s->noise_dev, 0, nsteps, rec_host);
}

int atlj_sim_run_nhc(atlj_sim *s, double dt, double temperature, double g, int m, double *eta, double *p_eta,
                     const double *q, int nsteps, double *rec_host) {
  if (!s || !s->configured || nsteps < 0 || m < 1 || m > 8 || !eta || !p_eta || !q) return ATLJ_ERR_ARG;
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc || (rc = sim_ensure_forces(s, 0.0))) return rc;
  const long long n3 = 3 * s->n;
  double h[32];
  memset(h, 0, sizeof(h));
  h[1] = 1.0;
  for (int j = 0; j < m; j++) h[2 + j] = eta[j], h[10 + j] = p_eta[j], h[18 + j] = q[j];
  double *st = s->scal + SC_NHC;
  ATLJ_CUDA(cudaMemcpyAsync(st, h, sizeof(h), cudaMemcpyHostToDevice, s->st));
  ATLJ_CUDA(cudaStreamSynchronize(s->st));  // h is on the stack
  // sum v^2 of the starting velocities -> st[0]
  if ((rc = launch_vf_sums(s->st, n3, s->v[s->cur], s->f, s->partials, s->scal + SC_SUMV))) return rc;
  const double *ksq = s->scal + SC_SUMV;
  const double half_dt = 0.5 * dt;
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    Timer::begin(s, 2);
    nhc_chain_kernel<<<1, 32, 0, s->st>>>(st, m, dt, temperature, g, k == 0 ? ksq : nullptr, nullptr);
    ATLJ_LAUNCHED();
    // u2(dt/2) u1(dt) with the pending thermostat factor folded in (md_nvt_lj.py:276-277)
    if ((rc = launch_kick_drift(s->st, n3, s->r[s->cur], s->v[s->cur], s->f, half_dt, dt, s->p.box, st + 1))) return rc;
    set_scalar_kernel<<<1, 1, 0, s->st>>>(st + 1, 1.0);
    ATLJ_LAUNCHED();
    Timer::end(s);
    s->defer_fin = true;  // the half kick below finishes the force totals in an extra CTA (FinTail)
    if ((rc = sim_force(s, 0.0, false, rec, false))) return rc;
    Timer::begin(s, 2);
    if ((rc = launch_kick_sums(s->st, n3, s->v[s->cur], s->f, half_dt, s->vf_partials, rec + 4, s->cb.err + 4, nullptr,
                               nullptr, s->fin.nb > 0 ? &s->fin : nullptr)))
      return rc;
    s->fin = FinTail();
    nhc_chain_kernel<<<1, 32, 0, s->st>>>(st, m, dt, temperature, g, rec + 4, rec);
    ATLJ_LAUNCHED();
    Timer::end(s);
  }
  // apply the pending factor so that the stored velocities are the logical ones
  scale_v_kernel<<<integ_blocks(n3), IT, 0, s->st>>>(n3, s->v[s->cur], st + 1);
  ATLJ_LAUNCHED();
  set_scalar_kernel<<<1, 1, 0, s->st>>>(st + 1, 1.0);
  ATLJ_LAUNCHED();
  ATLJ_CUDA(cudaMemcpyAsync(h, st, sizeof(h), cudaMemcpyDeviceToHost, s->st));
  rc = sim_copy_rec(s, rec_host, nsteps);  // synchronises
  for (int j = 0; j < m; j++) eta[j] = h[2 + j], p_eta[j] = h[10 + j];
  s->step_count += nsteps;
  return rc;
}

int atlj_sim_run_sllod(atlj_sim *s, double dt, double strain_rate, double *strain, int nsteps, double *rec_host) {
  if (!s || !s->configured || nsteps < 0 || !strain) return ATLJ_ERR_ARG;
  if (s->p.model != ATLJ_MODEL_WCA) {
    set_error("SLLOD needs the WCA / Lees-Edwards model (md_nvt_lj_le.py imports md_lj_le_module)");
    return ATLJ_ERR_ARG;
  }
  int rc = sim_ensure_rec(s, nsteps > 0 ? nsteps : 1);
  if (rc || (rc = sim_ensure_forces(s, *strain))) return rc;
  const int n = (int)s->n;
  const long long n3 = 3 * s->n;
  const int nb = atom_blocks(n) < SL_CTAS ? atom_blocks(n) : SL_CTAS;
  const int nb3 = integ_blocks(n3) < SL_CTAS ? integ_blocks(n3) : SL_CTAS;
  double *sumf = s->scal + SC_SUMF;
  // two sets of per-CTA rows, used alternately: a kernel reads the rows of its predecessor and writes its own
  double *rowsA = s->vf_partials, *rowsB = s->vf_partials + 4 * 2048;
  const double t = dt / 2;              // md_nvt_lj_le.py:230
  const double x = t * strain_rate;     // :89,105
  double st = *strain;
  auto advance = [&]() {                // :94-95
    st = st + x;
    st = st - rint(st);
  };
  const BoxDiv bd = make_box_div(s->p.box);
  // sums of the starting velocities for the first b1: the b2 kernel as an identity -> rowsA
  sllod_b2_kernel<<<nb, IT, 0, s->st>>>(n, s->v[s->cur], s->f, nullptr, 0, dt, rowsA, nullptr);
  ATLJ_LAUNCHED();
  for (int k = 0; k < nsteps; k++) {
    double *rec = s->rec_dev + (size_t)k * ATLJ_NREC;
    Timer::begin(s, 2);
    // a(dt/2) then b1(dt/2); g from rowsA
    advance();
    sllod_ab1_kernel<<<nb, IT, 0, s->st>>>(n, s->r[s->cur], s->v[s->cur], t, x, st, bd, rowsA, nb, 0, rowsB);
    ATLJ_LAUNCHED();
    Timer::end(s);
    s->defer_fin = true;  // the sums kernel below finishes the force totals in an extra CTA (FinTail)
    if ((rc = sim_force(s, st, false, rec, false))) return rc;
    Timer::begin(s, 2);
    // b2(dt): sums of f.v, f^2, v^2 -> rowsA; the kernel that applies b2 forms its coefficients from them
    sllod_fv_sums_kernel<<<nb3 + (s->fin.nb > 0 ? 1 : 0), IT, 0, s->st>>>(n3, s->v[s->cur], s->f, rowsA, s->fin);
    ATLJ_LAUNCHED();
    s->fin = FinTail();
    sllod_b2_kernel<<<nb, IT, 0, s->st>>>(n, s->v[s->cur], s->f, rowsA, nb3, dt, rowsB, sumf);
    ATLJ_LAUNCHED();
    // b1(dt/2) then a(dt/2); g from rowsB, the sums of the velocities it leaves behind -> rowsA
    advance();
    sllod_ab1_kernel<<<nb, IT, 0, s->st>>>(n, s->r[s->cur], s->v[s->cur], t, x, st, bd, rowsB, nb, 1, rowsA);
    ATLJ_LAUNCHED();
    sllod_rec_kernel<<<1, IT, 0, s->st>>>(rowsA, nb, sumf, st, rec);
    ATLJ_LAUNCHED();
    Timer::end(s);
  }
  *strain = st;
  s->step_count += nsteps;
  return sim_copy_rec(s, rec_host, nsteps);
}

}  // extern "C"
