// Rate and semantics of tcgen05 kind::tf32 M=128 N=64 K=8 with the A operand in shared memory (SS) vs in TMEM (TS).
//   1. correctness of TS: A written with tcgen05.st (lane = row, column = k), B from a swizzled smem tile; exact integer data
//   2. clocks per MMA for long back-to-back chains in both modes (no other smem traffic)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>

constexpr int MROWS = 128, NCOLS = 64, KTOT = 32;       // one 128-byte atom
constexpr int ATOM_BYTES = MROWS * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __host__ __forceinline__ uint32_t sw128_off(int r, int k) {
  const uint32_t kb = (uint32_t)k * 4u;
  const uint32_t chunk = (kb >> 4) ^ ((uint32_t)r & 7u);
  return ((uint32_t)r >> 3) * 1024u + ((uint32_t)r & 7u) * 128u + (chunk << 4) + (kb & 15u);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// out: D (128 x 64) of the TS product, clocks[0] = SS clocks for `reps` x 4 MMAs, clocks[1] = TS
__global__ void __launch_bounds__(128) umma_rate(const float* __restrict__ A, float* __restrict__ D, long long* clocks, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t mbar, mbar2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < MROWS * KTOT; i += blockDim.x) {
    const int r = i / KTOT, k = i % KTOT;
    *reinterpret_cast<float*>(smem + sw128_off(r, k)) = A[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t d_tmem = tmem_base, a_tmem = tmem_base + 192;       // A: 32 columns at column 192

  // A -> TMEM: thread (row) stores its 32 K-values, 8 columns at a time
  {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < KTOT; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[row * KTOT + c0 + j]);
      const uint32_t taddr = a_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");

  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOLS >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);
  const uint32_t sbase = smem_u32(smem);
  uint32_t parity = 0;
  if (tid == 0) {
    // timing: SS
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 4; ++k) mma_ss(d_tmem, make_desc(sbase + k * 32), make_desc(sbase + k * 32), idesc, 1);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    wait_bar(&mbar, parity); parity ^= 1;
    long long t1 = clock64();
    clocks[0] = t1 - t0;
    // timing: TS
    t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 4; ++k) mma_ts(d_tmem, a_tmem + k * 8, make_desc(sbase + k * 32), idesc, 1);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    wait_bar(&mbar, parity); parity ^= 1;
    t1 = clock64();
    clocks[1] = t1 - t0;
    {
      const uint32_t idesc128 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);
      t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int k = 0; k < 4; ++k) mma_ts(d_tmem, a_tmem + k * 8, make_desc(sbase + k * 32), idesc128, 1);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
      wait_bar(&mbar, parity); parity ^= 1;
      t1 = clock64();
      clocks[2] = t1 - t0;
      // mixed: N=128 then N=64 into columns 64.. of the same accumulator
      t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int k = 0; k < 4; ++k) {
          mma_ts(d_tmem, a_tmem + k * 8, make_desc(sbase + k * 32), idesc128, 1);
          mma_ts(d_tmem + 64, a_tmem + k * 8, make_desc(sbase + k * 32), idesc, 1);
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
      wait_bar(&mbar, parity); parity ^= 1;
      t1 = clock64();
      clocks[3] = t1 - t0;
    }
    // correctness: TS from a clean accumulator
    for (int k = 0; k < 4; ++k) mma_ts(d_tmem, a_tmem + k * 8, make_desc(sbase + k * 32), idesc, k != 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    wait_bar(&mbar, parity); parity ^= 1;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // two issuing threads (warps 0 and 1) at once, different accumulators
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(&mbar2))); }
  __syncthreads();
  if (tid == 0 || tid == 32) {
    const uint32_t dd = d_tmem + (tid == 0 ? 64u : 128u);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 4; ++k) mma_ts(dd, a_tmem + k * 8, make_desc(sbase + k * 32), idesc, 1);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar2)) : "memory");
    wait_bar(&mbar2, 0);
    long long t1 = clock64();
    clocks[4 + (tid >> 5)] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < NCOLS; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 8; ++j) D[row * NCOLS + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem_base));
}

int main() {
  std::vector<float> A(MROWS * KTOT), D(MROWS * NCOLS, -1.f), ref(MROWS * NCOLS);
  for (int r = 0; r < MROWS; ++r)
    for (int k = 0; k < KTOT; ++k) A[r * KTOT + k] = (float)(((r * 7 + k * 13) % 17) - 8);
  for (int r = 0; r < MROWS; ++r)
    for (int c = 0; c < NCOLS; ++c) {
      double s = 0;
      for (int k = 0; k < KTOT; ++k) s += (double)A[r * KTOT + k] * A[c * KTOT + k];
      ref[r * NCOLS + c] = (float)s;
    }
  float *dA, *dD; long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 64);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  const int smem = ATOM_BYTES + 1024, reps = 2000;
  cudaFuncSetAttribute(umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_rate<<<1, 128, smem>>>(dA, dD, dC, reps);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  long long c[8];
  cudaMemcpy(c, dC, 64, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
  printf("TS product: max |D - ref| = %g, mismatches = %d of %zu\n", maxerr, bad, D.size());
  printf("SS: %.1f clk per MMA   TS: %.1f clk per MMA  (M=128 N=64 K=8 tf32, %d MMAs each)\n", (double)c[0] / (4.0 * reps),
         (double)c[1] / (4.0 * reps), 4 * reps);
  printf("TS N=128: %.1f clk per MMA; mixed N=128 + N=64: %.1f clk per pair; two warps at once (N=64 each): %.1f / %.1f clk per MMA per warp\n",
         (double)c[2] / (4.0 * reps), (double)c[3] / (4.0 * reps), (double)c[4] / (4.0 * reps), (double)c[5] / (4.0 * reps));
  return bad != 0;
}
