// fp64_probe.cu -- how the B200 FP64 pipe behaves under the conditions of the likelihood kernel (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu && ./fp64_probe
// Prints DFMA/clk/SM for: warps per scheduler x independent chains per thread; constant-bank operands; integer
// instructions interleaved; shared-memory broadcast loads interleaved.
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double kc[8] = {0.999999, 1e-7, 0.5, 0.25, 0.125, 1.5, 2.5, 3.5};

template <int ILP, int MODE>
__global__ void k(double* out, int iters, double a, double b, int ia) {
  extern __shared__ double sm[];
  double r[ILP];
  int q[ILP];
  for (int j = 0; j < ILP; j++) { r[j] = threadIdx.x * 1e-9 + j * 1e-3; q[j] = threadIdx.x + j; }
  if (MODE == 3) { for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-6; __syncthreads(); }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int j = 0; j < ILP; j++) {
        if (MODE == 0) r[j] = fma(r[j], a, b);
        if (MODE == 1) r[j] = fma(r[j], kc[0], kc[1]);                       // constant-bank operands
        if (MODE == 2) { r[j] = fma(r[j], a, b); q[j] = (q[j] ^ ia) + (q[j] >> 3); }  // + 2-3 integer ops per DFMA
        if (MODE == 3) { r[j] = fma(r[j], a, sm[(i * 8 + u) & 1023]); }       // + broadcast LDS.64 per DFMA
        if (MODE == 4) { r[j] = fma(r[j], a, b); q[j] = q[j] * ia + u; }      // + 1 IMAD per DFMA
      }
    }
  }
  double s = 0; int t = 0;
  for (int j = 0; j < ILP; j++) { s += r[j]; t += q[j]; }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int ILP, int MODE>
double run(int warps_per_sm, double* buf, int sms) {
  // one CTA of (32 * warps_per_sm) threads per SM, at most 1024
  int threads = 32 * warps_per_sm, blocks_per_sm = 1;
  while (threads > 1024) { threads /= 2; blocks_per_sm *= 2; }
  const int iters = 2000;
  size_t smem = MODE == 3 ? 8192 : 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, MODE><<<sms * blocks_per_sm, threads, smem>>>(buf, 10, 0.999999, 1e-7, 3);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<ILP, MODE><<<sms * blocks_per_sm, threads, smem>>>(buf, iters, 0.999999, 1e-7, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double dfma = (double)sms * blocks_per_sm * threads * iters * 8.0 * ILP;
  return dfma / (best * 1e-3);   // DFMA per second
}

int main() {
  int sms = 148, clk = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double hz = clk * 1e3;
  double* buf; cudaMalloc(&buf, sizeof(double) * sms * 4096);
  printf("SMs %d, clock %.0f MHz; numbers are DFMA per clock per SM (nominal 64)\n", sms, hz / 1e6);
  int ws[] = {4, 8, 12, 16, 24, 32, 64};
  printf("%-28s", "mode / warps per SM:");
  for (int w : ws) printf("%8d", w);
  printf("\n");
#define ROW(ILP, MODE, name) { printf("%-28s", name); for (int w : ws) printf("%8.2f", run<ILP, MODE>(w, buf, sms) / hz / sms); printf("\n"); }
  ROW(1, 0, "reg ILP1") ROW(2, 0, "reg ILP2") ROW(4, 0, "reg ILP4") ROW(8, 0, "reg ILP8")
  ROW(4, 1, "const operands ILP4") ROW(8, 1, "const operands ILP8")
  ROW(4, 2, "+3 int ops/DFMA ILP4") ROW(8, 2, "+3 int ops/DFMA ILP8")
  ROW(4, 4, "+1 IMAD/DFMA ILP4") ROW(8, 4, "+1 IMAD/DFMA ILP8")
  ROW(4, 3, "+1 LDS.64 bcast/DFMA ILP4") ROW(8, 3, "+1 LDS.64 bcast/DFMA ILP8")
  return 0;
}
