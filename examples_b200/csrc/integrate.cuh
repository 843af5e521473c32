// Launchers of integrate.cu (internal).
#pragma once
#include <cstring>

#include "common.cuh"

namespace atlj {

constexpr int IT = 256;

__device__ __forceinline__ double wrap1(double x) { return __dsub_rn(x, rint(x)); }

// a / box, correctly rounded, without the division subroutine: with inv = RN(1/box) from the host, q = RN(a*inv),
// r = a - q*box (exact in an fma), RN(q + r*inv) = RN(a/box) (Markstein).  inv = 0 selects the division: the host does
// that for the one family of divisors the theorem excludes (significand all ones) and outside the normal range.
struct BoxDiv {
  double box, inv;
};
inline BoxDiv make_box_div(double box) {
  unsigned long long bits;
  memcpy(&bits, &box, sizeof(bits));
  const bool ones = (bits & 0xfffffffffffffull) == 0xfffffffffffffull;
  BoxDiv b;
  b.box = box;
  b.inv = (!ones && box > 1e-100 && box < 1e100) ? 1.0 / box : 0.0;
  return b;
}
__device__ __forceinline__ double div_box(double a, const BoxDiv &b) {
  if (b.inv != 0.0) {
    const double q = __dmul_rn(a, b.inv);
    return __fma_rn(__fma_rn(-q, b.box, a), b.inv, q);
  }
  return __ddiv_rn(a, b.box);
}

// CTA reduction of K values, result row written by thread 0
template <int K>
__device__ __forceinline__ void block_reduce_store(double (&v)[K], double *__restrict__ row) {
  __shared__ double red[IT / 32][K];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) red[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
    for (int w = 0; w < IT / 32; w++) s += red[w][threadIdx.x];
    row[threadIdx.x] = s;
  }
}


int launch_sums_final(cudaStream_t st, const double *partials, int nblocks, int K, double *out);  // K <= 4

int integ_blocks(long long n3);
// n_dev (may be null): the atom count lives on the device and n3 is only an upper bound (slab path)
int launch_kick_drift(cudaStream_t st, long long n3, double *r, double *v, const double *f, double half_dt, double dt,
                      double box, const double *vscale, const int *n_dev = nullptr);  // box: plain length
// done: a device int that is 0 on entry (the kernel leaves it 0): counts finished CTAs for the in-kernel final sum
int launch_kick_sums(cudaStream_t st, long long n3, double *v, const double *f, double half_dt, double *partials,
                     double *out2, int *done, const int *n_dev = nullptr, int *step_ctr = nullptr,
                     const struct FinTail *fin = nullptr);
int launch_vf_sums(cudaStream_t st, long long n3, const double *v, const double *f, double *partials, double *out2);
int launch_unsort(cudaStream_t st, int n, const int *id, const double *in, double *out);
int launch_iota(cudaStream_t st, int n, int *id);
int measure_fp64_peak(cudaStream_t st, double *scratch, double *flops_per_s);

}  // namespace atlj
