// horner_probe.cu -- does a DFMA with three distinct register operands issue as fast as one with shared operands?
// Horner steps r_j = fma(x_j, r_j, c_u) over J independent chains, order pinned with volatile asm:
//   DEPTH = J : all J chains interleaved (distance between dependent DFMAs = J)
//   DEPTH = 2 : chains evaluated two at a time (distance 2), like ptxas schedules the real loop
#include <cstdio>
#include <cuda_runtime.h>

#define FMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))

template <int J, int DEPTH, int SHARED_C>
__global__ void k(double* out, int iters, const double* in) {
  double x[J], r[J], c[6];
  for (int j = 0; j < J; j++) { x[j] = in[j] + threadIdx.x * 1e-9; r[j] = in[8 + j]; }
  for (int u = 0; u < 6; u++) c[u] = in[16 + u] + (SHARED_C ? 0.0 : threadIdx.x * 1e-12);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int g = 0; g < J; g += DEPTH) {
#pragma unroll
      for (int u = 0; u < 6; u++) {
#pragma unroll
        for (int j = g; j < g + DEPTH; j++) FMA(r[j], x[j], r[j], c[u]);
      }
    }
  }
  double s = 0;
  for (int j = 0; j < J; j++) s += r[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int J, int DEPTH, int SHARED_C>
void run(const char* name, double* buf, double* in, int sms, double hz) {
  printf("%-34s", name);
  int ws[] = {4, 8, 12, 16, 32};
  for (int w : ws) {
    int threads = 32 * w, bps = 1;
    while (threads > 512) { threads /= 2; bps *= 2; }
    const int iters = 3000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<J, DEPTH, SHARED_C><<<sms * bps, threads>>>(buf, 10, in);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      k<J, DEPTH, SHARED_C><<<sms * bps, threads>>>(buf, iters, in);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double dfma = (double)sms * bps * threads * iters * 6.0 * J;
    printf("%8.2f", dfma / (best * 1e-3) / hz / sms);
  }
  printf("\n");
}

int main() {
  int sms = 148, clk = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double *buf, *in, h[32];
  for (int i = 0; i < 32; i++) h[i] = 0.3 + 0.01 * i;
  cudaMalloc(&buf, sizeof(double) * sms * 4096); cudaMalloc(&in, sizeof h); cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  printf("DFMA per clock per SM (nominal 64); columns = warps per SM 4 8 12 16 32\n");
  run<8, 8, 0>("J8 depth8 per-thread c", buf, in, sms, clk * 1e3);
  run<8, 4, 0>("J8 depth4 per-thread c", buf, in, sms, clk * 1e3);
  run<8, 2, 0>("J8 depth2 per-thread c", buf, in, sms, clk * 1e3);
  run<8, 1, 0>("J8 depth1 per-thread c", buf, in, sms, clk * 1e3);
  run<8, 8, 1>("J8 depth8 uniform c", buf, in, sms, clk * 1e3);
  run<8, 2, 1>("J8 depth2 uniform c", buf, in, sms, clk * 1e3);
  run<4, 4, 0>("J4 depth4 per-thread c", buf, in, sms, clk * 1e3);
  return 0;
}
