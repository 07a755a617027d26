// smid_probe.cu -- where do the CTAs of a one-wave persistent grid land, and do they run concurrently?
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 4) k(unsigned* smid, unsigned long long* t0, unsigned long long* t1, int iters, double* out) {
  extern __shared__ double sm[];
  unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  unsigned long long a; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a));
  double r0 = threadIdx.x, r1 = 1, r2 = 2, r3 = 3;
  for (int i = 0; i < iters; i++) { r0 = fma(r0, 0.999, 1e-3); r1 = fma(r1, 0.999, 1e-3); r2 = fma(r2, 0.999, 1e-3); r3 = fma(r3, 0.999, 1e-3); }
  unsigned long long b; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(b));
  if (threadIdx.x == 0) { smid[blockIdx.x] = s; t0[blockIdx.x] = a; t1[blockIdx.x] = b; }
  out[blockIdx.x * 128 + threadIdx.x] = r0 + r1 + r2 + r3;
}
int main(int argc, char** argv) {
  int smem_kb = argc > 1 ? atoi(argv[1]) : 45;
  const int G = 592;
  unsigned* smid; unsigned long long *t0, *t1; double* out;
  cudaMalloc(&smid, G * 4); cudaMalloc(&t0, G * 8); cudaMalloc(&t1, G * 8); cudaMalloc(&out, G * 128 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  if (argc > 2) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(argv[2]));
  int occ; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 128, smem_kb * 1024);
  k<<<G, 128, smem_kb * 1024>>>(smid, t0, t1, 1000, out);
  k<<<G, 128, smem_kb * 1024>>>(smid, t0, t1, 200000, out);
  cudaDeviceSynchronize();
  std::vector<unsigned> s(G); std::vector<unsigned long long> a(G), b(G);
  cudaMemcpy(s.data(), smid, G * 4, cudaMemcpyDeviceToHost); cudaMemcpy(a.data(), t0, G * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), t1, G * 8, cudaMemcpyDeviceToHost);
  int cnt[256] = {0};
  for (int i = 0; i < G; i++) cnt[s[i]]++;
  int hist[16] = {0};
  for (int i = 0; i < 148; i++) hist[cnt[i] > 15 ? 15 : cnt[i]]++;
  unsigned long long amin = *std::min_element(a.begin(), a.end()), amax = *std::max_element(a.begin(), a.end());
  unsigned long long bmin = *std::min_element(b.begin(), b.end()), bmax = *std::max_element(b.begin(), b.end());
  printf("smem %d KB occ(api) %d: CTAs per SM histogram:", smem_kb, occ);
  for (int i = 0; i < 16; i++) if (hist[i]) printf(" %d CTAs x %d SMs;", i, hist[i]);
  printf("\n start spread %.1f us, first end after %.1f us, last end after %.1f us (late starters: %d)\n", (amax - amin) * 1e-3, (bmin - amin) * 1e-3,
         (bmax - amin) * 1e-3, (int)std::count_if(a.begin(), a.end(), [&](unsigned long long v) { return v > bmin; }));
  return 0;
}
