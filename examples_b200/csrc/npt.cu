// Constant-NPT molecular dynamics around the device-resident force: the step loop of
// python_examples/md_npt_lj.py:344-366 with its propagators (:101-214).
//
// The box length changes every step (u1_propagator, :125-126), and with it the reference's host scalars
// (md_lj_module.py:80-88) and, now and then, the cell grid sc = floor(box/r_cut) (md_lj_ll_module.py:106).  These
// scalars and the thermostat / barostat chains are a few dozen flops on 2m + 3 numbers: they stay on the HOST, with
// the reference's own expressions, and the library re-configures the force path each step (atlj_sim_configure only
// grows the cell buffers when sc grows).  The per-atom work is three fused passes on the device,
//     v <- (v*s3)*e2 + (c2*t)*f ; r <- wrap(r + ((c1*dt)*v)/box)     u3 (pending) . u2 . u1      (:166,:149,:120-121)
//     force(box_new)
//     v <- v*e2' + (c2'*t)*f ; sum v^2, sum f^2                      u2                          (:149)
// and the host reads ONE record per step (sum v^2 and the virial feed u2', u4: :162-163,:199), i.e. one stream
// synchronisation per step - the price of a box that is an input of the next launch's geometry.
#include <cmath>
#include <cstring>

#include "integrate.cuh"
#include "sim.cuh"

namespace atlj {

__global__ void __launch_bounds__(IT) npt_kick_drift_kernel(long long n3, double *__restrict__ r,
                                                            double *__restrict__ v, const double *__restrict__ f,
                                                            double s3, double e2, double c2t, double c1t, BoxDiv box) {
  long long t = (long long)blockIdx.x * IT + threadIdx.x;
  const long long stride = (long long)gridDim.x * IT;
  for (; t < n3; t += stride) {
    double vv = __dmul_rn(v[t], s3);                                      // u3, md_npt_lj.py:166 (applied lazily)
    vv = __dadd_rn(__dmul_rn(vv, e2), __dmul_rn(c2t, f[t]));             // u2, :149
    const double x = __dadd_rn(r[t], div_box(__dmul_rn(c1t, vv), box));  // u1, :120
    v[t] = vv;
    r[t] = wrap1(x);                                                      // :121
  }
}

__global__ void __launch_bounds__(IT) npt_kick_sums_kernel(long long n3, double *__restrict__ v,
                                                           const double *__restrict__ f, double e2, double c2t,
                                                           double *__restrict__ partials) {
  long long t = (long long)blockIdx.x * IT + threadIdx.x;
  const long long stride = (long long)gridDim.x * IT;
  double s[2] = {0.0, 0.0};
  for (; t < n3; t += stride) {
    const double ff = f[t];
    const double vv = __dadd_rn(__dmul_rn(v[t], e2), __dmul_rn(c2t, ff));  // u2, :149
    v[t] = vv;
    s[0] += vv * vv;
    s[1] += ff * ff;
  }
  block_reduce_store<2>(s, partials + 4 * (size_t)blockIdx.x);
}

__global__ void __launch_bounds__(IT) npt_scale_kernel(long long n3, double *__restrict__ v, double s3) {
  for (long long t = (long long)blockIdx.x * IT + threadIdx.x; t < n3; t += (long long)gridDim.x * IT)
    v[t] = __dmul_rn(v[t], s3);
}

// scipy.special.exprel(-x) = (1 - exp(-x)) / x, "preserving accuracy for small x" (md_npt_lj.py:118)
static double exprel_neg(double x) { return std::fabs(x) < 2.220446049250313e-16 ? 1.0 : std::expm1(-x) / (-x); }

// the host-side state of the chains and the barostat
struct NptState {
  int m;
  double temperature, pressure, g, w_eps, box0, box, eps, p_eps;
  double eta[8], p_eta[8], q[8], eta_baro[8], p_eta_baro[8], q_baro[8];
  double ksq;     // sum v^2 of the logical velocities (stored velocities times the pending factor)
  double pend;    // pending u3 factor, folded into the next pass over v
};

// u4 and u4' (md_npt_lj.py:182-214)
static void npt_u4(NptState &z, double t, bool inwards) {
  const int m = z.m;
  for (int jj = 0; jj < m; jj++) {
    const int j = inwards ? m - 1 - jj : jj;
    const double gj = j == 0 ? z.ksq - z.g * z.temperature : (z.p_eta[j - 1] * z.p_eta[j - 1] / z.q[j - 1]) - z.temperature;
    if (j == m - 1) {
      z.p_eta[j] = z.p_eta[j] + t * gj;
    } else {
      const double x = t * z.p_eta[j + 1] / z.q[j + 1];
      z.p_eta[j] = z.p_eta[j] * std::exp(-x) + t * gj * exprel_neg(x);
    }
  }
  for (int jj = 0; jj < m; jj++) {
    const int j = inwards ? m - 1 - jj : jj;
    const double gj = j == 0 ? z.p_eps * z.p_eps / z.w_eps - z.temperature
                             : (z.p_eta_baro[j - 1] * z.p_eta_baro[j - 1] / z.q_baro[j - 1]) - z.temperature;
    if (j == m - 1) {
      z.p_eta_baro[j] = z.p_eta_baro[j] + t * gj;
    } else {
      const double x = t * z.p_eta_baro[j + 1] / z.q_baro[j + 1];
      z.p_eta_baro[j] = z.p_eta_baro[j] * std::exp(-x) + t * gj * exprel_neg(x);
    }
  }
}

// u3 and u3' (md_npt_lj.py:166-172); the elementwise v *= s is deferred (z.pend)
static void npt_u3(NptState &z, double t) {
  const double s = std::exp(-t * z.p_eta[0] / z.q[0]);
  z.pend *= s;
  z.ksq *= s * s;
  for (int j = 0; j < z.m; j++) z.eta[j] += t * z.p_eta[j] / z.q[j];
  for (int j = 0; j < z.m; j++) z.eta_baro[j] += t * z.p_eta_baro[j] / z.q_baro[j];
  z.p_eps = z.p_eps * std::exp(-t * z.p_eta_baro[0] / z.q_baro[0]);
}

// u2' (md_npt_lj.py:160-163)
static void npt_u2p(NptState &z, double t, double vir) {
  const double alpha = 1.0 + 3.0 / z.g;
  const double pv = alpha * z.ksq / 3.0 + vir;
  z.p_eps = z.p_eps + 3.0 * (pv - z.pressure * (z.box * z.box * z.box)) * t;
}

static void npt_thermo(NptState &z, double dt) {  // :346-348 and :361-363
  npt_u4(z, 0.25 * dt, true);
  npt_u3(z, 0.5 * dt);
  npt_u4(z, 0.25 * dt, false);
}

}  // namespace atlj

using namespace atlj;

extern "C" int atlj_sim_run_npt(atlj_sim *s, double dt, double temperature, double pressure, double r_cut, double g,
                                int m, double *eta, double *p_eta, const double *q, double *eta_baro,
                                double *p_eta_baro, const double *q_baro, double w_eps, double box0, double *eps,
                                double *p_eps, int nsteps, double *rec_host) {
  if (!s || !s->configured || nsteps < 0 || m < 1 || m > 8 || !eta || !p_eta || !q || !eta_baro || !p_eta_baro ||
      !q_baro || !eps || !p_eps || !(r_cut > 0.0) || !(box0 > 0.0))
    return ATLJ_ERR_ARG;
  if (s->mg.on || s->p.model != ATLJ_MODEL_LJ) {
    set_error("run_npt: single-domain LJ only");
    return ATLJ_ERR_ARG;
  }
  int rc = sim_ensure_rec(s, 1);
  if (rc) return rc;
  const long long n3 = 3 * s->n;
  NptState z;
  memset(&z, 0, sizeof(z));
  z.m = m, z.temperature = temperature, z.pressure = pressure, z.g = g, z.w_eps = w_eps, z.box0 = box0;
  z.eps = *eps, z.p_eps = *p_eps, z.box = s->p.box, z.pend = 1.0;
  for (int j = 0; j < m; j++) {
    z.eta[j] = eta[j], z.p_eta[j] = p_eta[j], z.q[j] = q[j];
    z.eta_baro[j] = eta_baro[j], z.p_eta_baro[j] = p_eta_baro[j], z.q_baro[j] = q_baro[j];
  }
  // sum v^2 and the virial of the starting state: the forces in s->f belong to the current positions (the caller
  // ran atlj_sim_force, as md_npt_lj.py:335 does); its record is still in rec_dev only if nothing ran in between, so
  // evaluate once more (positions unchanged: same forces)
  double h[ATLJ_NREC];
  if ((rc = sim_force(s, 0.0, false, s->rec_dev, false))) return rc;
  if ((rc = launch_vf_sums(s->st, n3, s->v[s->cur], s->f, s->partials, s->rec_dev + 4))) return rc;
  ATLJ_CUDA(cudaMemcpyAsync(h, s->rec_dev, sizeof(h), cudaMemcpyDeviceToHost, s->st));
  if ((rc = sim_check_flags(s))) return rc;  // synchronises
  z.ksq = h[4];
  double vir = h[2];
  const double alpha = 1.0 + 3.0 / g;
  const int nb = integ_blocks(n3);
  for (int k = 0; k < nsteps; k++) {
    const double t = 0.5 * dt;
    npt_thermo(z, dt);
    npt_u2p(z, t, vir);
    // u2(dt/2) (:145-149) and u1(dt) (:115-121) in one pass, with the box of BEFORE the strain update
    double x2 = t * alpha * z.p_eps / z.w_eps;
    double e2 = std::exp(-x2), c2t = exprel_neg(x2) * t;
    const double x1 = dt * z.p_eps / z.w_eps;
    const double c1t = exprel_neg(x1) * dt;
    Timer::begin(s, 2);
    npt_kick_drift_kernel<<<nb, IT, 0, s->st>>>(n3, s->r[s->cur], s->v[s->cur], s->f, z.pend, e2, c2t, c1t, make_box_div(z.box));
    ATLJ_LAUNCHED();
    Timer::end(s);
    z.pend = 1.0;
    z.eps = z.eps + x1;                  // :125
    z.box = z.box0 * std::exp(z.eps);    // :126
    {  // the force path's host scalars for the new box, with the reference's expressions (md_lj_module.py:80-88)
      const double r_cut_box = r_cut / z.box;
      // pot_cut depends on r_cut only (md_lj_module.py:85-88): unchanged
      const int sc = (int)std::floor(z.box / r_cut);
      if ((rc = atlj_sim_configure(s, s->p.model, z.box, r_cut_box * r_cut_box, z.box * z.box, s->p.pot_cut,
                                   sc < 1 ? 1 : sc, 0, sc < 1 ? 1 : sc)))
        return rc;
    }
    if ((rc = sim_force(s, 0.0, false, s->rec_dev, false))) return rc;
    x2 = t * alpha * z.p_eps / z.w_eps;  // p_eps unchanged since the first u2: same factors, kept for clarity
    e2 = std::exp(-x2), c2t = exprel_neg(x2) * t;
    Timer::begin(s, 2);
    npt_kick_sums_kernel<<<nb, IT, 0, s->st>>>(n3, s->v[s->cur], s->f, e2, c2t, s->vf_partials);
    ATLJ_LAUNCHED();
    if ((rc = launch_sums_final(s->st, s->vf_partials, nb, 2, s->rec_dev + 4))) return rc;
    Timer::end(s);
    ATLJ_CUDA(cudaMemcpyAsync(h, s->rec_dev, sizeof(h), cudaMemcpyDeviceToHost, s->st));
    if ((rc = sim_check_flags(s))) return rc;  // synchronises
    z.ksq = h[4];
    vir = h[2];
    npt_u2p(z, t, vir);
    npt_thermo(z, dt);
    if (rec_host) {
      double *o = rec_host + (size_t)k * ATLJ_NREC;
      memcpy(o, h, sizeof(h));
      double ext = 0.0, es = 0.0;
      for (int j = 0; j < m; j++) ext += 0.5 * z.p_eta[j] * z.p_eta[j] / z.q[j];
      for (int j = 0; j < m; j++) ext += 0.5 * z.p_eta_baro[j] * z.p_eta_baro[j] / z.q_baro[j];
      ext += 0.5 * z.p_eps * z.p_eps / z.w_eps;
      for (int j = 1; j < m; j++) es += z.eta[j];
      for (int j = 0; j < m; j++) es += z.eta_baro[j];
      o[4] = z.ksq;                                          // after the closing thermostat half step
      o[9] = z.box;                                          // box length the step ended with
      o[10] = ext + temperature * (g * z.eta[0] + es);       // md_npt_lj.py:44-45
    }
  }
  // flush the pending thermostat factor so that the stored velocities are the logical ones
  if (z.pend != 1.0) {
    npt_scale_kernel<<<nb, IT, 0, s->st>>>(n3, s->v[s->cur], z.pend);
    ATLJ_LAUNCHED();
  }
  ATLJ_CUDA(cudaStreamSynchronize(s->st));
  for (int j = 0; j < m; j++) eta[j] = z.eta[j], p_eta[j] = z.p_eta[j], eta_baro[j] = z.eta_baro[j],
                              p_eta_baro[j] = z.p_eta_baro[j];
  *eps = z.eps, *p_eps = z.p_eps;
  s->step_count += nsteps;
  return ATLJ_OK;
}
