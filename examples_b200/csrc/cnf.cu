// Configuration files of multi-million-atom systems: the text format of the reference's
// config_io_module.py (write_cnf_atoms :74-95, read_cnf_atoms :30-48), produced and parsed on the GPU.
//
// write_cnf_atoms is np.savetxt(fmt='%15.10f'): every value becomes a 15-character field followed by one separator
// (' ' between columns, '\n' after the last), i.e. a file body of n*cols fixed 16-byte fields - byte work that
// parallelises perfectly: one thread per value, one 16-byte store (format) or one 16-byte load (parse).
//   * format: '%.10f' must be the correctly rounded (ties-to-even on the EXACT binary value, as glibc printf does)
//     10-decimal expansion.  |x| = m * 2^e exactly, so N = round_half_even(m * 10^10 * 2^e) is computed in 128-bit
//     integer arithmetic; no floating-point rounding is involved anywhere.
//   * parse: a field holds at most 15 significant digits D and f <= 22 decimals, so D and 10^f are both exact doubles
//     and the single IEEE division D / 10^f is the correctly rounded value (what strtod / np.loadtxt return).
// Values that do not fit a 15-character field (|x| >= 10^4, or <= -10^3), NaN/Inf, and fields in any other notation
// are reported as ATLJ_ERR_ARG with a message - never silently approximated.
#include "common.cuh"
#include "sim.cuh"

namespace atlj {

constexpr int CNF_T = 256;
constexpr unsigned long long P10_10 = 10000000000ull;

// -> the 16 bytes of one field (15 characters + sep), or ok = false
__device__ __forceinline__ uint4 cnf_field(double x, char sep, bool &ok) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
  const bool neg = (bits >> 63) != 0;
  const int ex = (int)((bits >> 52) & 0x7ff);
  unsigned long long m = bits & 0xfffffffffffffull;
  int e;
  if (ex == 0)
    e = -1074;
  else
    m |= 1ull << 52, e = ex - 1075;
  ok = ex != 0x7ff;
  unsigned long long N = 0;
  if (ok && m != 0) {
    // |x| < 10^4 needs e <= -39 for a normal number (m >= 2^52)
    if (e > -39) {
      ok = false;
    } else {
      const int k = -e;
      if (k < 128) {
        const unsigned __int128 P = (unsigned __int128)m * P10_10;  // < 2^87
        unsigned __int128 q = P >> k;
        const unsigned __int128 rem = P & ((((unsigned __int128)1) << k) - 1), half = ((unsigned __int128)1) << (k - 1);
        if (rem > half || (rem == half && ((unsigned long long)q & 1ull))) q += 1;
        N = (unsigned long long)q;
      }
    }
  }
  unsigned long long ip = N / P10_10, fp = N - ip * P10_10;
  if (ip > (neg ? 999ull : 9999ull)) ok = false, ip = 0;
  char c[16];
#pragma unroll
  for (int k = 14; k >= 5; k--) {
    const unsigned long long d = fp / 10ull;
    c[k] = (char)('0' + (int)(fp - d * 10ull));
    fp = d;
  }
  c[4] = '.';
  // integer part, at least one digit; then the sign; then blanks
  unsigned v = (unsigned)ip;
  bool more = true, sign = neg;
#pragma unroll
  for (int k = 3; k >= 0; k--) {
    char ch = ' ';
    if (more) {
      ch = (char)('0' + (int)(v % 10u));
      v /= 10u;
      more = v != 0u;
    } else if (sign) {
      ch = '-';
      sign = false;
    }
    c[k] = ch;
  }
  c[15] = sep;
  uint4 o;
  o.x = (unsigned)(unsigned char)c[0] | ((unsigned)(unsigned char)c[1] << 8) | ((unsigned)(unsigned char)c[2] << 16) |
        ((unsigned)(unsigned char)c[3] << 24);
  o.y = (unsigned)(unsigned char)c[4] | ((unsigned)(unsigned char)c[5] << 8) | ((unsigned)(unsigned char)c[6] << 16) |
        ((unsigned)(unsigned char)c[7] << 24);
  o.z = (unsigned)(unsigned char)c[8] | ((unsigned)(unsigned char)c[9] << 8) | ((unsigned)(unsigned char)c[10] << 16) |
        ((unsigned)(unsigned char)c[11] << 24);
  o.w = (unsigned)(unsigned char)c[12] | ((unsigned)(unsigned char)c[13] << 8) |
        ((unsigned)(unsigned char)c[14] << 16) | ((unsigned)(unsigned char)c[15] << 24);
  return o;
}

// One thread per value.  a: (n,3) scaled by scale_a (the drivers pass r*box, md_nve_lj.py:196); b: (n,3) or null;
// id (or null): row k of the arrays is written as row id[k] of the file (undoes the cell sort).
__global__ void __launch_bounds__(CNF_T) cnf_format_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                           const int *__restrict__ id, double scale_a, long long n,
                                                           int cols, uint4 *__restrict__ out, int *__restrict__ err) {
  const long long t = (long long)blockIdx.x * CNF_T + threadIdx.x;
  if (t >= n * cols) return;
  const long long k = t / cols;
  const int c = (int)(t - k * cols);
  double x;
  if (c < 3)
    x = a[3 * k + c], x = scale_a == 1.0 ? x : __dmul_rn(x, scale_a);
  else
    x = b[3 * k + (c - 3)];
  bool ok;
  const uint4 f = cnf_field(x, c == cols - 1 ? '\n' : ' ', ok);
  if (!ok) atomicOr(err, 1);
  const long long row = id ? (long long)id[k] : k;
  out[row * cols + c] = f;
}

// One thread per field: [blanks][sign]digits[.digits] right up to the separator
__global__ void __launch_bounds__(CNF_T) cnf_parse_kernel(const uint4 *__restrict__ text, long long n, int cols,
                                                          double *__restrict__ a, double *__restrict__ b,
                                                          int *__restrict__ err) {
  const long long t = (long long)blockIdx.x * CNF_T + threadIdx.x;
  if (t >= n * cols) return;
  const long long k = t / cols;
  const int c = (int)(t - k * cols);
  const uint4 w = text[t];
  const unsigned u[4] = {w.x, w.y, w.z, w.w};
  unsigned long long D = 0;
  int nd = 0, nf = 0, state = 0;  // 0 blanks, 1 after sign, 2 integer digits, 3 fraction digits
  bool neg = false, bad = false;
#pragma unroll
  for (int i = 0; i < 15; i++) {
    const int ch = (int)((u[i >> 2] >> (8 * (i & 3))) & 0xffu);
    if (ch == ' ') {
      bad |= state != 0;
    } else if (ch == '-' || ch == '+') {
      bad |= state != 0;
      neg = ch == '-';
      state = 1;
    } else if (ch >= '0' && ch <= '9') {
      if (state < 2) state = 2;
      D = D * 10ull + (unsigned long long)(ch - '0');
      nd += (D != 0ull) ? 1 : 0;
      nf += state == 3 ? 1 : 0;
    } else if (ch == '.') {
      bad |= state == 3;
      state = 3;
    } else {
      bad = true;
    }
  }
  const int sep = (int)((u[3] >> 24) & 0xffu);
  bad |= (c == cols - 1) ? (sep != '\n') : (sep != ' ');
  bad |= state < 2 || nd > 15;
  if (bad) {
    atomicOr(err, 1);
    return;
  }
  double p = 1.0;  // 10^nf, exact (nf <= 14)
  for (int i = 0; i < nf; i++) p *= 10.0;
  double x = __ddiv_rn((double)D, p);
  if (neg) x = -x;
  if (c < 3)
    a[3 * k + c] = x;
  else if (c < 6 && b)
    b[3 * k + (c - 3)] = x;
}

static int cnf_err_check(cudaStream_t st, int *err_dev, const char *what) {
  int h = 0;
  ATLJ_CUDA(cudaMemcpyAsync(&h, err_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  ATLJ_CUDA(cudaStreamSynchronize(st));
  if (h) {
    set_error("%s", what);
    return ATLJ_ERR_ARG;
  }
  return ATLJ_OK;
}

struct DevBuf {  // scratch that lives for one call
  void *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
      p = nullptr;
      set_error("cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
      return ATLJ_ERR_CUDA;
    }
    return ATLJ_OK;
  }
};

static const char *FMT_MSG =
    "cnf: a value does not fit the reference's '%15.10f' field (|x| >= 1e4, x <= -1e3, NaN or Inf)";
static const char *PARSE_MSG =
    "cnf: the file body is not n*cols fixed 16-byte '%15.10f' fields as written by write_cnf_atoms";

}  // namespace atlj

using namespace atlj;

extern "C" {

int atlj_cnf_format(const double *r_host, const double *v_host, int64_t n, double scale_r, char *out_host) {
  int rc = device_ok();
  if (rc) return rc;
  if (!r_host || !out_host || n < 0) return ATLJ_ERR_ARG;
  if (n == 0) return ATLJ_OK;
  const int cols = v_host ? 6 : 3;
  const size_t b = sizeof(double) * 3 * (size_t)n, tb = 16 * (size_t)n * cols;
  DevBuf r, v, t, e;
  if ((rc = r.alloc(b)) || (v_host && (rc = v.alloc(b))) || (rc = t.alloc(tb)) || (rc = e.alloc(sizeof(int)))) return rc;
  cudaStream_t st = nullptr;
  ATLJ_CUDA(cudaMemsetAsync(e.p, 0, sizeof(int), st));
  ATLJ_CUDA(cudaMemcpyAsync(r.p, r_host, b, cudaMemcpyHostToDevice, st));
  if (v_host) ATLJ_CUDA(cudaMemcpyAsync(v.p, v_host, b, cudaMemcpyHostToDevice, st));
  const long long nt = (long long)n * cols;
  cnf_format_kernel<<<(unsigned)((nt + CNF_T - 1) / CNF_T), CNF_T, 0, st>>>((const double *)r.p, (const double *)v.p,
                                                                            nullptr, scale_r, n, cols, (uint4 *)t.p,
                                                                            (int *)e.p);
  ATLJ_LAUNCHED();
  ATLJ_CUDA(cudaMemcpyAsync(out_host, t.p, tb, cudaMemcpyDeviceToHost, st));
  return cnf_err_check(st, (int *)e.p, FMT_MSG);
}

int atlj_cnf_parse(const char *text_host, int64_t n, int cols, double *r_host, double *v_host) {
  int rc = device_ok();
  if (rc) return rc;
  if (!text_host || !r_host || n < 0 || cols < 3 || (v_host && cols < 6)) return ATLJ_ERR_ARG;
  if (n == 0) return ATLJ_OK;
  const size_t b = sizeof(double) * 3 * (size_t)n, tb = 16 * (size_t)n * cols;
  DevBuf r, v, t, e;
  if ((rc = r.alloc(b)) || (v_host && (rc = v.alloc(b))) || (rc = t.alloc(tb)) || (rc = e.alloc(sizeof(int)))) return rc;
  cudaStream_t st = nullptr;
  ATLJ_CUDA(cudaMemsetAsync(e.p, 0, sizeof(int), st));
  ATLJ_CUDA(cudaMemcpyAsync(t.p, text_host, tb, cudaMemcpyHostToDevice, st));
  const long long nt = (long long)n * cols;
  cnf_parse_kernel<<<(unsigned)((nt + CNF_T - 1) / CNF_T), CNF_T, 0, st>>>((const uint4 *)t.p, n, cols, (double *)r.p,
                                                                           (double *)v.p, (int *)e.p);
  ATLJ_LAUNCHED();
  ATLJ_CUDA(cudaMemcpyAsync(r_host, r.p, b, cudaMemcpyDeviceToHost, st));
  if (v_host) ATLJ_CUDA(cudaMemcpyAsync(v_host, v.p, b, cudaMemcpyDeviceToHost, st));
  return cnf_err_check(st, (int *)e.p, PARSE_MSG);
}

int atlj_sim_cnf_format(atlj_sim *s, int with_v, char *out_host) {
  if (!s || !s->configured || !out_host) return ATLJ_ERR_ARG;
  if (s->mg.on) {
    set_error("cnf: gather the slabs first (single-domain call)");
    return ATLJ_ERR_ARG;
  }
  const long long n = s->n;
  if (n == 0) return ATLJ_OK;
  const int cols = with_v ? 6 : 3;
  const size_t tb = 16 * (size_t)n * cols;
  int rc;
  DevBuf t;
  if ((rc = t.alloc(tb))) return rc;
  int *err = s->cb.err + 3;  // spare flag word of the sim
  ATLJ_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s->st));
  const long long nt = n * cols;
  // positions are held in box = 1 units and in cell order: scale by box (md_nve_lj.py:196 writes r*box) and put every
  // atom back on the row of its original index
  cnf_format_kernel<<<(unsigned)((nt + CNF_T - 1) / CNF_T), CNF_T, 0, s->st>>>(
      s->r[s->cur], with_v ? s->v[s->cur] : nullptr, s->id[s->cur], s->p.box, n, cols, (uint4 *)t.p, err);
  ATLJ_LAUNCHED();
  ATLJ_CUDA(cudaMemcpyAsync(out_host, t.p, tb, cudaMemcpyDeviceToHost, s->st));
  return cnf_err_check(s->st, err, FMT_MSG);
}

}  // extern "C"
