// loop_probe.cu -- the inner loop of the likelihood kernel in isolation (no TMA ring, no work items): every CTA
// evaluates the same shared-memory tile `iters` times.  Separates "what the loop can do" from "what the persistent
// kernel around it costs".  Not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../lisp_mcmc_b200/csrc [-DMFIT_UNROLL_PAIRS=n] [-DCTAS=n] -o loop_probe loop_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "mfit_eval.cuh"
#include "mfit_models.cuh"

// the same pass with (x, y, w) synthesised on the integer pipe instead of read from shared memory
template <class Model, int LL, int NCH, class G>
__device__ __forceinline__ void tile_pass_noload(const Model* m, int cnt, double* acc, G* gd, int seed) {
  constexpr int UNR = MFIT_UNROLL_PAIRS;
#pragma unroll UNR
  for (int i = 0; i < (cnt >> 1); i++) {
    const int j = (i + seed) & 1023;
    double2 xv, yv, wv, cv;
    xv.x = __hiloint2double(0x3ff00000 + j, 0); xv.y = __hiloint2double(0x3ff10000 + j, 0);
    yv.x = __hiloint2double(0x40300000 + j, 0); yv.y = __hiloint2double(0x40310000 + j, 0);
    wv.x = __hiloint2double(0x40240000 + j, 0); wv.y = __hiloint2double(0x40250000 + j, 0);
    cv.x = cv.y = -1.0;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      const double f0 = m[c].template eval<false>(xv.x, gd[c]);
      const double f1 = m[c].template eval<false>(xv.y, gd[c]);
      acc[2 * c] = MfitTerm<LL>::template add<false>(acc[2 * c], f0, yv.x, wv.x, cv.x, gd[c]);
      acc[2 * c + 1] = MfitTerm<LL>::template add<false>(acc[2 * c + 1], f1, yv.y, wv.y, cv.y, gd[c]);
    }
  }
}

#ifndef CTAS
#define CTAS 4
#endif
#ifndef THREADS
#define THREADS 128
#endif

#ifndef PNC
#define PNC 1
#endif
template <class Model, int LL, class G>
__global__ void __launch_bounds__(THREADS, CTAS) k(MfitEvalArgs a, int iters, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* buf = reinterpret_cast<double*>(smem);
  mfit_math_stage(Model::USES | MfitTerm<LL>::USES, threadIdx.x, THREADS);
  for (int i = threadIdx.x; i < MFIT_TILE; i += THREADS) {
    buf[i] = a.x[i]; buf[MFIT_TILE + i] = a.yw[i]; buf[2 * MFIT_TILE + i] = a.w[i]; buf[3 * MFIT_TILE + i] = a.c[i];
  }
  __syncthreads();
  Model m[2];
  m[0].load(a, (long long)blockIdx.x * THREADS + threadIdx.x);
  m[1].load(a, (long long)(gridDim.x - 1 - blockIdx.x) * THREADS + threadIdx.x);
  m[0].setup();
  m[1].setup();
  double acc[4] = {0, 0, 0, 0};
  G gd[2];
  gd[0].reset();
  gd[1].reset();
  for (int it = 0; it < iters; it++) {
#ifdef NOLOAD
    tile_pass_noload<Model, LL, PNC>(m, MFIT_TILE, acc, gd, it);
#else
    mfit_tile_pass<Model, LL, false, PNC>(m, buf, buf + MFIT_TILE, buf + 2 * MFIT_TILE, buf + 3 * MFIT_TILE, MFIT_TILE, acc, gd);
#endif
  }
  out[(size_t)blockIdx.x * THREADS + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3] + ((gd[0].tripped() || gd[1].tripped()) ? 1.0 : 0.0);
}

template <class Model, int LL, class G>
void run(const char* name, int nparam, const double* th, int fp64_per_point) {
  int sms = 148, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int grid = sms * CTAS;
  const long long stride = (long long)grid * THREADS;
  double *x, *y, *w, *c, *theta, *out;
  double hx[MFIT_TILE], hy[MFIT_TILE], hw[MFIT_TILE], hc[MFIT_TILE];
  for (int i = 0; i < MFIT_TILE; i++) { hx[i] = -10.0 + 20.0 * i / MFIT_TILE; hy[i] = 30.0 + 0.01 * (i % 7); hw[i] = 10.0 + (i % 5); hc[i] = -1.0; }
  cudaMalloc(&x, sizeof hx); cudaMalloc(&y, sizeof hx); cudaMalloc(&w, sizeof hx); cudaMalloc(&c, sizeof hx);
  cudaMemcpy(x, hx, sizeof hx, cudaMemcpyHostToDevice); cudaMemcpy(y, hy, sizeof hx, cudaMemcpyHostToDevice);
  cudaMemcpy(w, hw, sizeof hx, cudaMemcpyHostToDevice); cudaMemcpy(c, hc, sizeof hx, cudaMemcpyHostToDevice);
  cudaMalloc(&theta, sizeof(double) * stride * nparam); cudaMalloc(&out, sizeof(double) * stride);
  double* ht = new double[stride * nparam];
  for (int p = 0; p < nparam; p++) for (long long i = 0; i < stride; i++) ht[p * stride + i] = th[p] * (1.0 + 1e-4 * (i % 97));
  cudaMemcpy(theta, ht, sizeof(double) * stride * nparam, cudaMemcpyHostToDevice);
  MfitEvalArgs a{};
  a.x = x; a.y = y; a.yw = y; a.w = w; a.c = c; a.n = MFIT_TILE; a.xmin = -10; a.xmax = 10; a.stride = stride; a.theta = theta; a.W = (int)stride;
  for (int k2 = 0; k2 < MFIT_MAXK; k2++) a.idx[k2] = k2;
  const size_t smem = 4 * MFIT_TILE * 8;
  cudaFuncSetAttribute(k<Model, LL, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<Model, LL, G>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<Model, LL, G>, THREADS, smem);
  const int iters = 800;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<Model, LL, G><<<grid, THREADS, smem>>>(a, 2, out);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<Model, LL, G><<<grid, THREADS, smem>>>(a, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  const double pts = (double)stride * iters * MFIT_TILE * PNC;
  const double rate = pts / (best * 1e-3);
  printf("%-22s NC %d ctas/sm %d (occ %d) threads %d unroll %d: %.4g point-evals/s, %.1f cyc/warp-point/SMSP, FP64 pipe %.1f%% of 64/clk/SM  %s\n", name,
         PNC, CTAS, occ, THREADS, MFIT_UNROLL_PAIRS, rate, (clk * 1e3) * sms * 4 * 32 / rate, 100.0 * rate * fp64_per_point / (clk * 1e3 * sms * 64.0),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(x); cudaFree(y); cudaFree(w); cudaFree(c); cudaFree(theta); cudaFree(out); delete[] ht;
}

int main() {
  const double g[] = {10.0, 0.5, 1.5, 2.0, 0.1};
  run<MfitGaussLin, MFITK_NORMAL, MfitNoGuard>("gauss_lin unguarded", 5, g, 14);
  run<MfitGaussLin, MFITK_NORMAL, MfitGuard>("gauss_lin guarded", 5, g, 14);
  const double p5[] = {1, -0.5, 0.8, 0.3, -0.2, 0.05};
  run<MfitPoly<5>, MFITK_NORMAL, MfitNoGuard>("poly5", 6, p5, 7);
  const double ex[] = {2.0, -0.7, 0.5};
  run<MfitExpo, MFITK_NORMAL, MfitNoGuard>("expo unguarded", 3, ex, 14);
  const double lo[] = {5.0, 0.5, 3.0, -1.0};
  run<MfitLorentz, MFITK_NORMAL, MfitNoGuard>("lorentz", 4, lo, 13);
  const double si[] = {20.0, 0.7, 0.3, 50.0};
  run<MfitSinus, MFITK_POISSON, MfitGuard>("sinus+poisson guarded", 4, si, 31);
  return 0;
}
