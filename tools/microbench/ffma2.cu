// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a, and SHFL throughput.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters, float a, float b) {
  unsigned long long x[8], aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
  for (int i = 0; i < 8; ++i) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_shfl(float* out, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 3));
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
  const int iters = 20000, blocks = 148 * 4, threads = 512;
  for (int rep = 0; rep < 3; ++rep) {
    float ms;
    cudaEventRecord(e0); k_ffma<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * threads * iters * 16;
    printf("FFMA : %.3f ms  %.1f lane-FMA/clk/SM @%d kHz  (%.2f TFMA/s)\n", ms, fma / (ms * 1e-3) / 148 / (dev_clk * 1e3), dev_clk, fma / ms * 1e-9);
    cudaEventRecord(e0); k_ffma2<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA2: %.3f ms  %.1f lane-FMA/clk/SM (%.2f TFMA/s)\n", ms, fma / (ms * 1e-3) / 148 / (dev_clk * 1e3), fma / ms * 1e-9);
    cudaEventRecord(e0); k_shfl<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double sh = (double)blocks * (threads / 32) * iters * 8;
    printf("SHFL : %.3f ms  %.3f warp-shfl/clk/SM\n", ms, sh / (ms * 1e-3) / 148 / (dev_clk * 1e3));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
