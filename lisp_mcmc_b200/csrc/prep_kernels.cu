// prep_kernels.cu -- the input side of walker-create on the device.
//
// walker-create (mcmc-fitting.lisp:1132-1163) normalises :data / :data-error on the host (clean-data, clean-data-error,
// to-double-floats :774-835); what the likelihood kernels stream is derived from that once per upload:
//     w_i = 1/sigma_i,  yw_i = y_i w_i,  c_i = -ln(2 pi)/2 - ln sigma_i  (log-normal :372-377 minus its theta-dependent part),
//     sum_i c_i,  the range of x (Model::safe),  whether y is a valid count vector (log-poisson :382-383), max y,
//     -sum_i log y_i!  (log-factorial :379-380, single-float accumulation, looked up in a table the host builds to max y).
// The first version did this with one host thread (a log per point: 3 ms for C2's 1e5 points, 0.2 s for C5's 1e7, most of
// the end-to-end fixed cost); here it is one pass over the uploaded arrays.
//
// Sums are compensated (two-sum error terms carried along) in a fixed shape -- thread-strided, then a block tree, then
// block order on the host -- so the constants are deterministic and within an ulp or two of exact: with a scalar
// :data-error every c_i is the same number and a plain running sum drifts like N eps (1.3e-10 on C5, found by
// tests/test_gpu_fullsize.py).  Compiled with -fmad=false like the chain kernels.
#include <cuda_runtime.h>
#include <math.h>
#include "mfit_internal.h"

namespace {

struct Comp {  // running sum s with the rounding errors of its additions collected in c
  double s, c;
  __device__ __forceinline__ void add(double v) {
    const double t = s + v;
    c += fabs(s) >= fabs(v) ? (s - t) + v : (v - t) + s;
    s = t;
  }
  __device__ __forceinline__ void merge(const Comp& o) {
    add(o.s);
    c += o.c;
  }
};

// one block's share of the per-dataset (or per-row) results
__device__ void block_reduce(Comp sum, double xmin, double xmax, double ymax, int bad_x, int bad_pois, MfitPrepBlock* out) {
  __shared__ Comp s_sum[MFIT_PREP_THREADS];
  __shared__ double s_min[MFIT_PREP_THREADS], s_max[MFIT_PREP_THREADS], s_ymax[MFIT_PREP_THREADS];
  __shared__ int s_bx[MFIT_PREP_THREADS], s_bp[MFIT_PREP_THREADS];
  const int t = threadIdx.x;
  s_sum[t] = sum; s_min[t] = xmin; s_max[t] = xmax; s_ymax[t] = ymax; s_bx[t] = bad_x; s_bp[t] = bad_pois;
  __syncthreads();
  for (int h = MFIT_PREP_THREADS / 2; h > 0; h >>= 1) {  // fixed-shape tree
    if (t < h) {
      s_sum[t].merge(s_sum[t + h]);
      s_min[t] = fmin(s_min[t], s_min[t + h]);
      s_max[t] = fmax(s_max[t], s_max[t + h]);
      s_ymax[t] = fmax(s_ymax[t], s_ymax[t + h]);
      s_bx[t] |= s_bx[t + h];
      s_bp[t] |= s_bp[t + h];
    }
    __syncthreads();
  }
  if (t == 0) {
    out->sum = s_sum[0].s; out->err = s_sum[0].c; out->xmin = s_min[0]; out->xmax = s_max[0]; out->ymax = s_ymax[0];
    out->bad_x = s_bx[0]; out->bad_pois = s_bp[0];
  }
}

// rows = 1: a shared data set, the blocks of the grid stride over its points, one MfitPrepBlock per block.
// rows > 1: a walker set with per-walker data, block b handles rows b, b + gridDim.x, ...: one MfitPrepBlock per ROW.
// sigma arrives in c (sig_rows rows: 1 = one vector for all rows) and is replaced by c_i; x (x_rows rows) is only read.
__global__ void __launch_bounds__(MFIT_PREP_THREADS)
k_prepare(const double* x, const double* y, double* yw, double* w, double* c, long long n, long long n_pad, long long rows,
          long long x_rows, long long sig_rows, int have_sigma, double half_log_2pi_neg, MfitPrepBlock* out) {
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  if (rows == 1) {
    Comp sum = {0.0, 0.0};
    double xmin = inf, xmax = -inf, ymax = 0.0;
    int bad_x = 0, bad_pois = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (long long)gridDim.x * blockDim.x) {
      if (i >= n) { yw[i] = 0.0; w[i] = 0.0; c[i] = 0.0; continue; }   // padding of the last tile (the host side zeroes that of x and y)
      const double sg = have_sigma ? c[i] : 1.0;   // (if data-error data-error 1) :1144
      const double wi = 1.0 / sg, yi = y[i], xi = x[i];
      w[i] = wi;
      yw[i] = yi * wi;
      const double ci = half_log_2pi_neg + (-1.0 * log(sg));
      c[i] = ci;
      sum.add(ci);
      if (!isfinite(xi)) bad_x = 1;
      xmin = fmin(xmin, xi); xmax = fmax(xmax, xi);
      if (!(yi >= 0.0) || yi != floor(yi) || yi > 5e6) bad_pois = 1;
      else ymax = fmax(ymax, yi);
    }
    block_reduce(sum, xmin, xmax, ymax, bad_x, bad_pois, out + blockIdx.x);
    return;
  }
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    Comp sum = {0.0, 0.0};
    double xmin = inf, xmax = -inf, ymax = 0.0;
    int bad_x = 0, bad_pois = 0;
    const double* xr = x + (x_rows == 1 ? 0 : r) * n_pad;
    const double* yr = y + r * n_pad;
    double* cr = c + (sig_rows == 1 ? 0 : r) * n_pad;
    double* wr = w + (sig_rows == 1 ? 0 : r) * n_pad;
    for (long long i = threadIdx.x; i < n_pad; i += blockDim.x) {
      if (i >= n) { yw[r * n_pad + i] = 0.0; continue; }
      // a shared sigma row is converted by the block of row 0 only; the others read it after the kernel (yw below uses sigma itself)
      const double sg = have_sigma ? cr[i] : 1.0;
      const double wi = 1.0 / sg, yi = yr[i];
      yw[r * n_pad + i] = yi * wi;
      const double ci = half_log_2pi_neg + (-1.0 * log(sg));
      sum.add(ci);
      if (sig_rows != 1) { wr[i] = wi; cr[i] = ci; }
      if (x_rows != 1 || r == 0) {
        const double xi = xr[i];
        if (!isfinite(xi)) bad_x = 1;
        xmin = fmin(xmin, xi); xmax = fmax(xmax, xi);
      }
      if (!(yi >= 0.0) || yi != floor(yi) || yi > 5e6) bad_pois = 1;
      else ymax = fmax(ymax, yi);
    }
    __syncthreads();
    block_reduce(sum, xmin, xmax, ymax, bad_x, bad_pois, out + r);
    __syncthreads();
  }
}

// a shared sigma vector of a walker set: sigma (in c) -> w, c, once every row has been served by k_prepare
__global__ void k_prepare_shared_sigma(double* w, double* c, long long n, long long n_pad, int have_sigma, double half_log_2pi_neg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i >= n) { w[i] = 0.0; c[i] = 0.0; return; }
  const double sg = have_sigma ? c[i] : 1.0;
  w[i] = 1.0 / sg;
  c[i] = half_log_2pi_neg + (-1.0 * log(sg));
}

// -sum_i table[y_i] per data set (rows = 1: one MfitPrepBlock per block) or per row (one per row), compensated
__global__ void __launch_bounds__(MFIT_PREP_THREADS)
k_poisson_const(const double* y, const double* table, long long n, long long n_pad, long long rows, MfitPrepBlock* out) {
  if (rows == 1) {
    Comp sum = {0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      sum.add(-table[(long long)y[i]]);
    block_reduce(sum, 0.0, 0.0, 0.0, 0, 0, out + blockIdx.x);
    return;
  }
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    Comp sum = {0.0, 0.0};
    for (long long i = threadIdx.x; i < n; i += blockDim.x) sum.add(-table[(long long)y[r * n_pad + i]]);
    __syncthreads();
    block_reduce(sum, 0.0, 0.0, 0.0, 0, 0, out + r);
    __syncthreads();
  }
}

// log y_i! per point from the host-built table (a likelihood lambda that calls log-factorial / log-poisson)
__global__ void k_logfact_fill(const double* y, const double* table, double* lf, long long n, long long n_pad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) lf[i] = i < n ? table[(long long)y[i]] : 0.0;
}

}  // namespace

void mfit_launch_logfact_fill(const double* y, const double* table, double* lf, long long n, long long n_pad, cudaStream_t st) {
  k_logfact_fill<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(y, table, lf, n, n_pad);
}

void mfit_launch_prepare(const double* x, const double* y, double* yw, double* w, double* c, long long n, long long n_pad,
                         long long rows, long long x_rows, long long sig_rows, int have_sigma, MfitPrepBlock* out, int blocks,
                         cudaStream_t st) {
  const double a = -0.5 * log(2.0 * 3.141592653589793);  // (* -1/2 (log (* 2 pi))) :377
  k_prepare<<<blocks, MFIT_PREP_THREADS, 0, st>>>(x, y, yw, w, c, n, n_pad, rows, x_rows, sig_rows, have_sigma, a, out);
  if (rows > 1 && sig_rows == 1)
    k_prepare_shared_sigma<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(w, c, n, n_pad, have_sigma, a);
}

void mfit_launch_poisson_const(const double* y, const double* table, long long n, long long n_pad, long long rows,
                               MfitPrepBlock* out, int blocks, cudaStream_t st) {
  k_poisson_const<<<blocks, MFIT_PREP_THREADS, 0, st>>>(y, table, n, n_pad, rows, out);
}
