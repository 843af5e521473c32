// Integrator kernels kept resident on the device between steps.
//
// Reference loop bodies: md_nve_lj.py:178-192 (velocity Verlet), bd_nvt_lj.py:91-127,221-233 (BAOAB),
// md_nvt_lj.py:101-159,270-288 (Nose-Hoover chain), md_nvt_lj_le.py:84-129,226-240 (isokinetic SLLOD).
// All are elementwise over the flat (3N) arrays -> pure HBM streaming; each kernel fuses every elementwise
// statement between two global reductions / force calls of the reference loop.  The elementwise arithmetic keeps
// NumPy's operation order and rounding (e.g. v + (0.5*dt)*f as multiply-then-add, (dt*v)/box as multiply-then-
// divide) via the _rn intrinsics so short trajectories track the reference to rounding of the force sums only.
// Global sums are two-stage (per-CTA partials, then one CTA in fixed order) => run-to-run deterministic.
#include "common.cuh"
#include "integrate.cuh"
#include "pair_math.cuh"

namespace atlj {

// ---- velocity Verlet ---------------------------------------------------------------------------------------------
// md_nve_lj.py:182-185: v = v + 0.5*dt*f ; r = r + dt*v/box ; r = r - rint(r)     (reads r,v,f, writes r,v)
// vscale (device scalar, may be null) multiplies v first: pending Nose-Hoover U3 factor (md_nvt_lj.py:129).
__global__ void __launch_bounds__(IT) kick_drift_kernel(long long n3, double *__restrict__ r, double *__restrict__ v,
                                                        const double *__restrict__ f, double half_dt, double dt,
                                                        BoxDiv box, const double *__restrict__ vscale,
                                                        const int *__restrict__ n_dev) {
  if (n_dev) n3 = min(n3, 3ll * *n_dev);  // atom count kept on the device (slab path): n3 is an upper bound
  long long t = (long long)blockIdx.x * IT + threadIdx.x;
  long long stride = (long long)gridDim.x * IT;
  double sc = vscale ? *vscale : 1.0;
  for (; t < n3; t += stride) {
    double vv = v[t];
    if (vscale) vv = __dmul_rn(vv, sc);
    vv = __dadd_rn(vv, __dmul_rn(half_dt, f[t]));
    double x = __dadd_rn(r[t], div_box(__dmul_rn(dt, vv), box));
    v[t] = vv;
    r[t] = wrap1(x);
  }
}

// md_nve_lj.py:190 + the sums of :43-45: v = v + 0.5*dt*f ; sum v^2, sum f^2.  Two-level fixed-order reduction in ONE
// launch: every CTA writes its row, the CTA that finishes last adds the rows up (deterministic) and writes out2[0..1].
__global__ void __launch_bounds__(IT) kick_sums_kernel(long long n3, double *__restrict__ v,
                                                       const double *__restrict__ f, double half_dt,
                                                       double *__restrict__ partials, const int *__restrict__ n_dev,
                                                       int *__restrict__ done, double *__restrict__ out2,
                                                       int *__restrict__ step_ctr, FinTail fin) {
  // the extra CTA (fin.nb > 0): the pair totals of the force evaluation that has just ended
  const int nblk = (int)gridDim.x - (fin.nb > 0 ? 1 : 0);
  if ((int)blockIdx.x == nblk) {
    finalize_rows_block<IT>(fin);
    return;
  }
  if (n_dev) n3 = min(n3, 3ll * *n_dev);
  // replayed launches: out2 is relative to the record of step *step_ctr; this kernel ends the step and moves the
  // counter on (its last CTA, after every CTA has read it)
  if (step_ctr) out2 += (size_t)ATLJ_NREC * (size_t)*step_ctr;
  long long t = (long long)blockIdx.x * IT + threadIdx.x;
  long long stride = (long long)nblk * IT;
  double s[2] = {0.0, 0.0};
  for (; t < n3; t += stride) {
    double ff = f[t];
    double vv = __dadd_rn(v[t], __dmul_rn(half_dt, ff));
    v[t] = vv;
    s[0] += vv * vv;
    s[1] += ff * ff;
  }
  block_reduce_store<2>(s, partials + 4 * (size_t)blockIdx.x);
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done, 1) == nblk - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double a[2] = {0.0, 0.0};
  for (int b = threadIdx.x; b < nblk; b += IT) {
    a[0] += __ldcg(partials + 4 * (size_t)b);
    a[1] += __ldcg(partials + 4 * (size_t)b + 1);
  }
  block_reduce_store<2>(a, out2);
  if (threadIdx.x == 0) {
    *done = 0;
    if (step_ctr) *step_ctr += 1;
  }
}

// generic second stage: out[k] = sum_b partials[b][k], k < K <= 4, fixed order
__global__ void __launch_bounds__(256) sums_final_kernel(const double *__restrict__ partials, int nblocks, int K,
                                                         double *__restrict__ out) {
  __shared__ double sm[256][4];
  double v[4] = {0, 0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += 256)
    for (int k = 0; k < K; k++) v[k] += partials[4 * (size_t)b + k];
  for (int k = 0; k < 4; k++) sm[threadIdx.x][k] = v[k];
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 4; k++) sm[threadIdx.x][k] += sm[threadIdx.x + o][k];
    __syncthreads();
  }
  if (threadIdx.x < K) out[threadIdx.x] = sm[0][threadIdx.x];
}

int launch_sums_final(cudaStream_t st, const double *partials, int nblocks, int K, double *out) {
  sums_final_kernel<<<1, 256, 0, st>>>(partials, nblocks, K, out);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

int integ_blocks(long long n3) {
  long long b = (n3 + IT - 1) / IT;
  long long cap = 148 * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int launch_kick_drift(cudaStream_t st, long long n3, double *r, double *v, const double *f, double half_dt, double dt,
                      double box, const double *vscale, const int *n_dev) {
  if (n3 <= 0) return ATLJ_OK;
  kick_drift_kernel<<<integ_blocks(n3), IT, 0, st>>>(n3, r, v, f, half_dt, dt, make_box_div(box), vscale, n_dev);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

int launch_kick_sums(cudaStream_t st, long long n3, double *v, const double *f, double half_dt, double *partials,
                     double *out2, int *done, const int *n_dev, int *step_ctr, const FinTail *fin) {
  int nb = integ_blocks(n3);
  FinTail ft;
  if (fin) ft = *fin;  // (its rows must not be `partials`: this kernel's CTAs write theirs there)
  kick_sums_kernel<<<nb + (ft.nb > 0 ? 1 : 0), IT, 0, st>>>(n3, v, f, half_dt, partials, n_dev, done, out2, step_ctr, ft);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

// sums of v only (initial record / after upload): out = {sum v^2, sum f^2}
__global__ void __launch_bounds__(IT) vf_sums_kernel(long long n3, const double *__restrict__ v,
                                                     const double *__restrict__ f, double *__restrict__ partials) {
  long long t = (long long)blockIdx.x * IT + threadIdx.x;
  long long stride = (long long)gridDim.x * IT;
  double s[2] = {0.0, 0.0};
  for (; t < n3; t += stride) {
    double vv = v[t], ff = f[t];
    s[0] += vv * vv;
    s[1] += ff * ff;
  }
  block_reduce_store<2>(s, partials + 4 * (size_t)blockIdx.x);
}

int launch_vf_sums(cudaStream_t st, long long n3, const double *v, const double *f, double *partials, double *out2) {
  int nb = integ_blocks(n3);
  vf_sums_kernel<<<nb, IT, 0, st>>>(n3, v, f, partials);
  ATLJ_LAUNCHED();
  sums_final_kernel<<<1, 256, 0, st>>>(partials, nb, 2, out2);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

// ---- back to the caller's atom order: out[id[s]] = in[s] ------------------------------------------------------------
__global__ void __launch_bounds__(256) unsort_kernel(int n, const int *__restrict__ id, const double *__restrict__ in,
                                                     double *__restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  int s = t / 3, k = t - 3 * s;
  out[3 * (size_t)id[s] + k] = in[t];
}

int launch_unsort(cudaStream_t st, int n, const int *id, const double *in, double *out) {
  if (n <= 0) return ATLJ_OK;
  unsort_kernel<<<(3 * n + 255) / 256, 256, 0, st>>>(n, id, in, out);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

__global__ void iota_kernel(int n, int *id) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) id[i] = i;
}
int launch_iota(cudaStream_t st, int n, int *id) {
  if (n <= 0) return ATLJ_OK;
  iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, id);
  ATLJ_LAUNCHED();
  return ATLJ_OK;
}

// ---- fp64 peak probe: register-resident DFMA chains on every SM ------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
      x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;  // keep the chain alive
}

int measure_fp64_peak(cudaStream_t st, double *scratch, double *flops_per_s) {
  const int blocks = 148 * 8, iters = 4096;
  cudaEvent_t e0, e1;
  ATLJ_CUDA(cudaEventCreate(&e0));
  ATLJ_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    ATLJ_CUDA(cudaEventRecord(e0, st));
    dfma_peak_kernel<<<blocks, 256, 0, st>>>(scratch, iters, 0.999999, 1e-9);
    ATLJ_LAUNCHED();
    ATLJ_CUDA(cudaEventRecord(e1, st));
    ATLJ_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    ATLJ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
    double rate = flops / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *flops_per_s = best;
  return ATLJ_OK;
}

}  // namespace atlj
