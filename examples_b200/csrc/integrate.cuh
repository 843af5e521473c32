// Launchers of integrate.cu (internal).
#pragma once
#include "common.cuh"

namespace atlj {

int integ_blocks(long long n3);
int launch_kick_drift(cudaStream_t st, long long n3, double *r, double *v, const double *f, double half_dt, double dt,
                      double box, const double *vscale);
int launch_kick_sums(cudaStream_t st, long long n3, double *v, const double *f, double half_dt, double *partials,
                     double *out2);
int launch_vf_sums(cudaStream_t st, long long n3, const double *v, const double *f, double *partials, double *out2);
int launch_unsort(cudaStream_t st, int n, const int *id, const double *in, double *out);
int launch_iota(cudaStream_t st, int n, int *id);
int measure_fp64_peak(cudaStream_t st, double *scratch, double *flops_per_s);

}  // namespace atlj
