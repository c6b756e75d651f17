// Throughput of the fma.rn.f32x2 operand forms the packed covariance uses: plain pairs, .F32 broadcast, .LO_HI.NP swizzle.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_forms ffma2_forms.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void k(const float* in, float* out, int iters, long long* clk) {
  constexpr int NA = 24;
  f32x2 acc[NA];
  float x[8], y[8];
  for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + 32 * i]; y[i] = in[threadIdx.x + 32 * i + 256]; }
  for (int i = 0; i < NA; ++i) acc[i] = pk2(x[i % 8], y[i % 8]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int r = i % 8, c = (i * 3 + 1) % 8;
      if (MODE == 0) acc[i] = fma2(pk2(x[r], y[r]), pk2(x[c], y[c]), acc[i]);            // plain pairs
      else if (MODE == 1) acc[i] = fma2(pk2(x[r], y[r]), pk2(x[c], x[c]), acc[i]);       // broadcast b
      else if (MODE == 2) acc[i] = fma2(pk2(y[r], -x[r]), pk2(y[c], y[c]), acc[i]);      // swapped/negated a + broadcast b
      else { float lo, hi; upk2(acc[i], lo, hi); lo = fmaf(x[r], x[c], lo); hi = fmaf(y[r], x[c], hi); acc[i] = pk2(lo, hi); }   // scalar
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < NA; ++i) { float lo, hi; upk2(acc[i], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[MODE] = t1 - t0;
}

int main() {
  float *in, *out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 64);
  cudaMemset(in, 0, 4096 * 4);
  const int iters = 4000;
  for (int warps : {4, 8, 16}) {
    k<0><<<148, warps * 32>>>(in, out, iters, clk);
    k<1><<<148, warps * 32>>>(in, out, iters, clk);
    k<2><<<148, warps * 32>>>(in, out, iters, clk);
    k<3><<<148, warps * 32>>>(in, out, iters, clk);
    cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, clk, 32, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 24 * warps;   // warp-level packed ops (or scalar pairs) per SM
    printf("%2d warps/SM: clk per warp-instr per SM: plain %.3f  broadcast %.3f  swap-neg+broadcast %.3f  | scalar pair (2 FFMA) %.3f   [%s]\n", warps,
           h[0] / n, h[1] / n, h[2] / n, h[3] / n, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
