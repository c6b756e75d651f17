// Step-A validation of the tcgen05 tf32 building blocks used by the large-array HERK (herk_tc.cu):
//   D (128 x 64, fp32, TMEM) = A (128 x K, K-major, 128B-swizzled smem) * B^T (B = rows 0..63 of A)
// Operands are small integers (exact in tf32), so the result must match the host product exactly.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_tf32_test umma_tf32_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int MROWS = 128, NCOLS = 64, KTOT = 64;       // K = 64 floats = 2 swizzle atoms of 32 floats (128 B)
constexpr int ATOM_BYTES = MROWS * 128;                 // one 128-row x 128-byte atom column

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of (row r, float k) inside a [128 rows][32 floats] K-major SWIZZLE_128B atom
__device__ __host__ __forceinline__ uint32_t sw128_off(int r, int k) {
  const uint32_t kb = (uint32_t)k * 4u;                          // byte within the 128-byte row
  const uint32_t chunk = (kb >> 4) ^ ((uint32_t)r & 7u);          // 16-byte chunk index XOR row-in-group
  return ((uint32_t)r >> 3) * 1024u + ((uint32_t)r & 7u) * 128u + (chunk << 4) + (kb & 15u);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO(16 B)>>4 << 16 | SBO(1024 B)>>4 << 32 | version 1 << 46 | layout 2 << 61
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) umma_test(const float* __restrict__ A, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // fill the two atoms: element (r, k) of A -> atom k/32, column k%32
  for (int i = tid; i < MROWS * KTOT; i += blockDim.x) {
    const int r = i / KTOT, k = i % KTOT;
    *reinterpret_cast<float*>(smem + (k / 32) * ATOM_BYTES + sw128_off(r, k % 32)) = A[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOLS >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);
    const uint32_t sbase = smem_u32(smem);
    uint32_t accum = 0;
    for (int atom = 0; atom < KTOT / 32; ++atom) {
      for (int k = 0; k < 4; ++k) {                                  // 4 MMAs of K = 8 per 128-byte atom
        const uint64_t da = make_desc(sbase + atom * ATOM_BYTES + k * 32);
        const uint64_t db = da;                                      // B = first 64 rows of the same tile
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(accum));
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: warp w reads TMEM lanes 32w..32w+31, 64 columns each (8 loads of 8 columns)
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < NCOLS; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 8; ++j) D[row * NCOLS + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem_base));
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
  float *dA, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, D.size() * 4);
  const int smem = 2 * ATOM_BYTES + 1024;
  cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_test<<<1, 128, smem>>>(dA, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
  printf("max |D - ref| = %g, mismatches = %d of %zu; D[0..3] = %g %g %g %g ref %g %g %g %g\n", maxerr, bad, D.size(), D[0], D[1], D[2], D[3],
         ref[0], ref[1], ref[2], ref[3]);
  printf("D[64*64+5]=%g ref=%g  D[127*64+63]=%g ref=%g\n", D[64 * 64 + 5], ref[64 * 64 + 5], D[127 * 64 + 63], ref[127 * 64 + 63]);
  return bad != 0;
}
