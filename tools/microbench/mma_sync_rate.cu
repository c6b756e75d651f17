// Microbenchmark: legacy warp-level tensor-core MMA (mma.sync.m16n8k8 tf32, SASS HMMA.1688.F32.TF32) on sm_100a.
// Question it answers for the M = 16 covariance: how many m16n8k8 MMAs per clock does one SM retire, and do they
// overlap with FFMA2 work issued by other warps (separate pipe) or steal the FP32 pipe?
//   mode 0: MMA only, `warps` warps per CTA, one CTA per SM, 8 independent accumulators per warp
//   mode 1: FFMA2 only (the Jacobi / scan stand-in)
//   mode 2: half of the warps MMA, the other half FFMA2 -- compare with mode 0 and mode 1 at half the warps each
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(1024, 1) k_mix(float* out, int iters, int mode, long long* clk) {
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const bool do_mma = mode == 0 || (mode == 2 && warp < nw / 2);
  const long long t0 = clock64();
  float s = 0.f;
  if (do_mma) {
    float d[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    unsigned a0 = __float_as_uint(1.0f + threadIdx.x * 1e-3f), a1 = a0 ^ 0x1000u, a2 = a0 ^ 0x2000u, a3 = a0 ^ 0x3000u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mma_tf32(d[i], a0, a1, a2, a3, a0, a2);
    }
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  } else {
    unsigned long long x[8], aa, bb;
    const float a = 1.0001f, b = 0.5f;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    }
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = clock64() - t0;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
  long long* clk; cudaMallocManaged(&clk, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int mode = 0; mode < 3; ++mode) {
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_mix<<<148, warps * 32>>>(out, iters, mode, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      const double cycles = (double)clk[0];
      const int wm = mode == 0 ? warps : mode == 2 ? warps / 2 : 0, wf = mode == 1 ? warps : mode == 2 ? warps - warps / 2 : 0;
      printf("warps %2d mode %d: %.3f ms, %.0f clk | MMA m16n8k8 %.3f /clk/SM (%.1f dense TF32 TFLOP/s chip) | FFMA2 %.1f lane-FMA/clk/SM\n",
             warps, mode, ms, cycles, wm * 8.0 * iters / cycles, wm * 8.0 * iters * 2048.0 * 148 / (ms * 1e-3) * 1e-12,
             wf * 8.0 * iters * 64.0 / cycles);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
