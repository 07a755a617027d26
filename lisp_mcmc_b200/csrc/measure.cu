// measure.cu -- live FP64 roofline denominators.
//
// MEASURED_PEAKS.json carries HBM and bf16-GEMM peaks but no FP64 figure, and the likelihood kernel
// is bound by the FP64 pipe.  These micro-kernels measure, on the device the
// context is bound to, the DFMA issue peak (8 independent chains per thread, every SM full) and
// the cost of the FP64 exp/log/sin/divide sequences in DFMA-issue equivalents, which is the
// conversion factor DESIGN.md uses to put transcendental-bearing models on the same roofline.
#include <cuda_runtime.h>
#include "mfit_chain.h"

namespace {

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double r0 = threadIdx.x * 1e-9, r1 = r0 + 1e-3, r2 = r0 + 2e-3, r3 = r0 + 3e-3;
  double r4 = r0 + 4e-3, r5 = r0 + 5e-3, r6 = r0 + 6e-3, r7 = r0 + 7e-3;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
      r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}

template <int OP>
__global__ void __launch_bounds__(256) k_fn(double* out, int iters, double a, double b) {
  double r0 = 0.1 + threadIdx.x * 1e-6, r1 = r0 + 0.01, r2 = r0 + 0.02, r3 = r0 + 0.03;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (OP == 0) { r0 = exp(r0 * a - b); r1 = exp(r1 * a - b); r2 = exp(r2 * a - b); r3 = exp(r3 * a - b); }
      if (OP == 1) { r0 = log(r0 * a + b); r1 = log(r1 * a + b); r2 = log(r2 * a + b); r3 = log(r3 * a + b); }
      if (OP == 2) { r0 = sin(r0 * a + b); r1 = sin(r1 * a + b); r2 = sin(r2 * a + b); r3 = sin(r3 * a + b); }
      if (OP == 3) { r0 = b / (r0 * a + b); r1 = b / (r1 * a + b); r2 = b / (r2 * a + b); r3 = b / (r3 * a + b); }
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = (r0 + r1) + (r2 + r3);
}

float time_launch(cudaStream_t st, void (*launch)(cudaStream_t, double*, int), double* buf, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch(st, buf, iters);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0, st);
    launch(st, buf, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

int g_blocks = 148 * 8;
void l_dfma(cudaStream_t st, double* b, int it) { k_dfma<<<g_blocks, 256, 0, st>>>(b, it, 0.999999, 1e-7); }
void l_exp(cudaStream_t st, double* b, int it) { k_fn<0><<<g_blocks, 256, 0, st>>>(b, it, 0.5, 0.3); }
void l_log(cudaStream_t st, double* b, int it) { k_fn<1><<<g_blocks, 256, 0, st>>>(b, it, 0.5, 1.3); }
void l_sin(cudaStream_t st, double* b, int it) { k_fn<2><<<g_blocks, 256, 0, st>>>(b, it, 3.5, 0.3); }
void l_div(cudaStream_t st, double* b, int it) { k_fn<3><<<g_blocks, 256, 0, st>>>(b, it, 0.5, 1.3); }

}  // namespace

int mfit_measure_fp64_dev(cudaStream_t st, double* dfma_flops, double* cost_exp, double* cost_log, double* cost_sin,
                          double* cost_div) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g_blocks = sms * 8;
  double* buf = nullptr;
  if (cudaMalloc(&buf, (size_t)g_blocks * 256 * sizeof(double)) != cudaSuccess) return -1;
  const double threads = (double)g_blocks * 256.0;
  const int it_dfma = 4096, it_fn = 256;
  float ms = time_launch(st, l_dfma, buf, it_dfma);
  const double dfma_per_s = threads * it_dfma * 64.0 / (ms * 1e-3);
  if (dfma_flops) *dfma_flops = 2.0 * dfma_per_s;
  struct { void (*l)(cudaStream_t, double*, int); double* out; } fns[] = {
      {l_exp, cost_exp}, {l_log, cost_log}, {l_sin, cost_sin}, {l_div, cost_div}};
  for (auto& f : fns) {
    float m = time_launch(st, f.l, buf, it_fn);
    const double calls_per_s = threads * it_fn * 16.0 / (m * 1e-3);
    // each call also carries one FMA for the argument; cost is reported in DFMA issue slots, argument FMA excluded
    if (f.out) *f.out = dfma_per_s / calls_per_s - 1.0;
  }
  cudaFree(buf);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
