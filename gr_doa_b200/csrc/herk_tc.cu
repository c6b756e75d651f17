// herk_tc.cu -- stage 1 for large arrays (M = 64): the sample covariance as a tensor-core complex HERK, 3xTF32 split.
//
// Replaces the cgemm of autocorrelate_impl::general_work (gr-doa lib/autocorrelate_impl.cc:106) where the Hermitian update
// costs 32.5 flop per input byte -- far on the FP32 side of the CUDA-core ridge.
//
// Formulation.  Read channel m's interleaved samples (re0, im0, re1, im1, ...) as a REAL row Z_m of length 2N, and let W_m
// be the same row with every (re, im) pair replaced by (im, -re).  Then
//     Re R[r][c] = sum_t re_r re_c + im_r im_c = (Z Z^T)[r][c]        Im R[r][c] = sum_t im_r re_c - re_r im_c = (W Z^T)[r][c]
// so ONE real product  D (128 x 64) = [Z; W] (128 x 2N) * Z^T  gives the whole covariance: a tcgen05 kind::tf32 UMMA with
// M = 128, N = 64, K = 8, accumulator = 128 TMEM lanes x 64 columns, both operands K-major in 128-byte-swizzled shared
// memory, the B operand aliasing rows 0..63 of A.  fp32 accuracy comes from the 3xTF32 split x = hi + lo
// (hi = tf32(x), lo = tf32(x - hi)):  D += A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T  (the dropped lo*lo term is 2^-22 relative).
//
// Accumulation.  The tensor core truncates the fp32 accumulator on every MMA (measured: relative error of the diagonal
// grows as #MMAs * 2^-24: 3e-5 at N = 2048, 2e-4 at N = 16384 with a single accumulator).  So the big terms A_hi B_hi^T go
// into a ping-pong pair of TMEM accumulators that is folded into fp32 REGISTERS (properly rounded adds) every TC_CHUNK
// stages = 16 MMAs, and the two cross terms, 2^-11 smaller, into a third accumulator folded once per frame.
//
// One persistent CTA per SM, 13 warps:
//   warp 0, one lane : MMA issuer.  Per 128-byte stage: wait full[s]; 4 K-steps x (hi*hi -> D_big[b], hi*lo and lo*hi ->
//                      D_small); tcgen05.commit -> empty[s]; every TC_CHUNK stages commit -> chunk_full[b] and switch b;
//                      after a frame's last stage commit -> small_full.
//   warps 1..8       : converters.  Thread t owns 4 complex samples of channel t/4 per stage: cp.async them into a private raw
//                      ring (RAW_STAGES deep: HBM bytes in flight do not depend on registers), split hi/lo with integer
//                      round-to-nearest on the tf32 boundary, store the four operand rows (Z_hi, Z_lo, W_hi, W_lo) with the
//                      swizzle applied, fence.proxy.async, arrive on full[s].  They run ahead across frame boundaries.
//   warps 9..12      : adders + epilogue.  Thread = one TMEM lane (row of [Re R; Im R]), 64 fp32 register accumulators:
//                      wait chunk_full[b], tcgen05.ld, add, arrive chunk_empty[b]; at the frame end add D_small, stage the
//                      128 x 64 result in shared memory, combine Re/Im, scale, forward-backward term, store R.
#include "doa_internal.h"

#include <algorithm>

namespace doa {
namespace {

constexpr int TC_M = 64;                 // channels
constexpr int TC_ROWS = 128;             // rows of [Z; W]
constexpr int TC_OP_STAGES = 4;          // operand ring depth (32 KB per stage: hi + lo tiles of 128 x 128 B)
constexpr int TC_RAW_STAGES = 6;         // raw fp32 ring depth (8 KB per stage)
constexpr int TC_CHUNK = 4;              // stages (= 16 hi*hi MMAs) per big-accumulator chunk
constexpr int TC_TILE_BYTES = TC_ROWS * 128;
constexpr int TC_CONV_WARPS = 8, TC_ADD_WARPS = 4;
constexpr int TC_CONV_THREADS = TC_CONV_WARPS * 32;
constexpr int TC_THREADS = (1 + TC_CONV_WARPS + TC_ADD_WARPS) * 32;
constexpr int TC_TMEM_COLS = 256;        // D_big[0] @0, D_big[1] @64, D_small[0] @128, D_small[1] @192

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(b)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: start >> 4, LBO = 16 B, SBO = 1024 B (8 rows x 128 B), version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// round to nearest (ties away) on the tf32 boundary: add half an ulp of the 10-bit mantissa, clear the low 13 bits
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void cp_async16z(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
// byte offset of the 16-byte chunk `chunk` (0..7) of row r inside a [128][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_chunk(int r, int chunk) {
  return ((uint32_t)r >> 3) * 1024u + ((uint32_t)r & 7u) * 128u + (((uint32_t)chunk ^ ((uint32_t)r & 7u)) << 4);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
herk_tc64_kernel(const float2* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
                 float2* __restrict__ out, float scale, float bscale, int avg_method) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* op = smem;                                                                  // [OP_STAGES][hi tile | lo tile]
  float4* raw = reinterpret_cast<float4*>(smem + TC_OP_STAGES * 2 * TC_TILE_BYTES);      // [RAW_STAGES][256 threads][2]
  float* stg = reinterpret_cast<float*>(smem + TC_OP_STAGES * 2 * TC_TILE_BYTES + TC_RAW_STAGES * TC_CONV_THREADS * 2 * 16);   // [128][65]
  __shared__ uint64_t full_bar[TC_OP_STAGES], empty_bar[TC_OP_STAGES], chunk_full[2], chunk_empty[2], small_full[2], small_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < TC_OP_STAGES; ++s) { mbar_init(&full_bar[s], TC_CONV_THREADS); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&chunk_full[b], 1); mbar_init(&chunk_empty[b], TC_ADD_WARPS * 32); }
    for (int b = 0; b < 2; ++b) { mbar_init(&small_full[b], 1); mbar_init(&small_empty[b], TC_ADD_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  const int spf = (N + 15) / 16;                                             // stages per frame (16 complex samples each)
  const int my_frames = (nframes > (int)blockIdx.x) ? (nframes - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total = (long long)my_frames * spf;

  if (warp == 0) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_M >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      int s = 0; uint32_t ph = 0; int sf = 0;
      int cb = 0, in_chunk = 0; uint32_t ce_ph = 0u;   // parity bits, one per buffer (bit b)
      long long chunks = 0;
      uint32_t se_ph = 0u; int frames_done = 0;
      for (long long q = 0; q < total; ++q) {
        if (in_chunk == 0) {                                        // first stage of a chunk: D_big[cb] must have been folded
          if (chunks >= 2) { mbar_wait(&chunk_empty[cb], (ce_ph >> cb) & 1u); ce_ph ^= 1u << cb; }
          ++chunks;
        }
        const int fb = frames_done & 1;                             // D_small buffer of this frame
        if (sf == 0 && frames_done >= 2) { mbar_wait(&small_empty[fb], (se_ph >> fb) & 1u); se_ph ^= 1u << fb; }   // folded by the epilogue
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t hi = smem_u32(op + (size_t)s * 2 * TC_TILE_BYTES), lo = hi + TC_TILE_BYTES;
        const uint32_t dbig = tmem_d + (uint32_t)cb * 64u, dsmall = tmem_d + 128u + (uint32_t)fb * 64u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dhi = umma_desc(hi + k * 32), dlo = umma_desc(lo + k * 32);
          umma_tf32(dbig, dhi, dhi, idesc, (in_chunk | k) != 0);   // A_hi B_hi^T   (B = rows 0..63 of the same tile)
          umma_tf32(dsmall, dhi, dlo, idesc, (sf | k) != 0);       // A_hi B_lo^T
          umma_tf32(dsmall, dlo, dhi, idesc, 1u);                  // A_lo B_hi^T
        }
        umma_commit(&empty_bar[s]);                                 // stage s reusable once these MMAs retire
        ++sf; ++in_chunk;
        const bool frame_end = (sf == spf);
        if (in_chunk == TC_CHUNK || frame_end) { umma_commit(&chunk_full[cb]); cb ^= 1; in_chunk = 0; }
        if (frame_end) { sf = 0; ++frames_done; umma_commit(&small_full[fb]); }
        if (++s == TC_OP_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp <= TC_CONV_WARPS) {
    // ================================ converters ================================
    const int ct = tid - 32;                       // 0..255
    const int ch = ct >> 2, qt = ct & 3;           // channel, which 4 of the stage's 16 complex samples
    float4* myraw = raw + ct;                      // raw ring laid out [stage][piece j][thread]: conflict-free LDS.128
    long long fi = blockIdx.x;
    const float2* ibase = in + fi * frame_stride + (long long)ch * chan_stride;
    int isf = 0, irs = 0; long long issued = 0;
    auto issue = [&]() {
      const int t = isf * 16 + qt * 4;             // first complex sample of this thread's 4
      float4* dst = myraw + (size_t)irs * TC_CONV_THREADS * 2;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int tj = t + 2 * j;
        const int nb = (tj < N) ? 16 : 0;          // N is even (launcher): a 16-byte piece is all in or all out
        cp_async16z(dst + j * TC_CONV_THREADS, ibase + (nb ? tj : 0), nb);
      }
      if (++isf == spf) { isf = 0; fi += gridDim.x; ibase = in + fi * frame_stride + (long long)ch * chan_stride; }
      if (++irs == TC_RAW_STAGES) irs = 0;
      ++issued;
    };
#pragma unroll
    for (int p = 0; p < TC_RAW_STAGES - 1; ++p) { if (issued < total) issue(); asm volatile("cp.async.commit_group;" ::: "memory"); }
    int s = 0; uint32_t ph = 0; int rs = 0;
    for (long long q = 0; q < total; ++q) {
      if (issued < total) issue();
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" :: "n"(TC_RAW_STAGES - 1) : "memory");
      if (q >= TC_OP_STAGES) mbar_wait(&empty_bar[s], ph ^ 1u);     // MMAs of the previous use of this stage have retired
      const float4* src = myraw + (size_t)rs * TC_CONV_THREADS * 2;
      uint8_t* thi = op + (size_t)s * 2 * TC_TILE_BYTES;
      uint8_t* tlo = thi + TC_TILE_BYTES;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 v = src[j * TC_CONV_THREADS];   // (re0, im0, re1, im1)
        float4 zh, zl;
        zh.x = to_tf32(v.x); zh.y = to_tf32(v.y); zh.z = to_tf32(v.z); zh.w = to_tf32(v.w);
        zl.x = to_tf32(v.x - zh.x); zl.y = to_tf32(v.y - zh.y); zl.z = to_tf32(v.z - zh.z); zl.w = to_tf32(v.w - zh.w);
        const float4 wh = make_float4(zh.y, -zh.x, zh.w, -zh.z);     // (im, -re)
        const float4 wl = make_float4(zl.y, -zl.x, zl.w, -zl.z);
        const int chunk = qt * 2 + j;              // 16-byte chunk within the 128-byte row
        const uint32_t oz = sw128_chunk(ch, chunk), ow = sw128_chunk(TC_M + ch, chunk);
        *reinterpret_cast<float4*>(thi + oz) = zh;
        *reinterpret_cast<float4*>(tlo + oz) = zl;
        *reinterpret_cast<float4*>(thi + ow) = wh;
        *reinterpret_cast<float4*>(tlo + ow) = wl;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      mbar_arrive(&full_bar[s]);
      if (++rs == TC_RAW_STAGES) rs = 0;
      if (++s == TC_OP_STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    // ============================ adders + epilogue ============================
    const int at = tid - (1 + TC_CONV_WARPS) * 32;                 // 0..127
    const int q4 = warp & 3;                                       // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;                                // row of [Re R; Im R]
    const uint32_t lane_addr = tmem_d + ((uint32_t)(q4 * 32) << 16);
    uint32_t cf_ph = 0u, sf_ph = 0u;                               // parity bits, one per buffer
    int cb = 0;
    long long fcur = blockIdx.x;
    for (int fr = 0; fr < my_frames; ++fr) {
      float acc[TC_M];
#pragma unroll
      for (int c = 0; c < TC_M; ++c) acc[c] = 0.0f;
      const int nchunks = (spf + TC_CHUNK - 1) / TC_CHUNK;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&chunk_full[cb], (cf_ph >> cb) & 1u); cf_ph ^= 1u << cb;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < TC_M; c0 += 8) {
          float v[8];
          tmem_ld8(lane_addr + (uint32_t)cb * 64u + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c0 + j] += v[j];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&chunk_empty[cb]);
        cb ^= 1;
      }
      const int fb = fr & 1;
      mbar_wait(&small_full[fb], (sf_ph >> fb) & 1u); sf_ph ^= 1u << fb;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < TC_M; c0 += 8) {
        float v[8];
        tmem_ld8(lane_addr + 128u + (uint32_t)fb * 64u + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) stg[row * 65 + c0 + j] = acc[c0 + j] + v[j];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&small_empty[fb]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float2* o = out + fcur * (long long)(TC_M * TC_M);
      for (int e = at; e < TC_M * TC_M; e += 128) {
        const int r = e % TC_M, c = e / TC_M;
        float2 v = make_float2(stg[r * 65 + c] * scale, stg[(TC_M + r) * 65 + c] * scale);
        if (avg_method == 1) {
          const int rr = TC_M - 1 - r, cc = TC_M - 1 - c;
          const float wx = stg[rr * 65 + cc] * scale, wy = stg[(TC_M + rr) * 65 + cc] * scale;
          v.x = __fadd_rn(__fmul_rn(0.5f, v.x), __fmul_rn(bscale, wx));
          v.y = __fadd_rn(__fmul_rn(0.5f, v.y), __fmul_rn(bscale, -wy));
        }
        o[e] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");               // staging area free again
      fcur += gridDim.x;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(TC_TMEM_COLS));
}

}  // namespace

// Returns 1 if launched, 0 if the shape is not covered (caller uses the CUDA-core kernels).
int launch_covariance_tc(const float2* in, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                         int avg_method, float2* out, cudaStream_t st) {
  if (M != TC_M || nframes <= 0) return 0;
  const bool aligned = (N % 2 == 0) && (frame_stride % 2 == 0) && (chan_stride % 2 == 0) &&
                       ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);
  if (!aligned) return 0;
  const size_t smem = (size_t)TC_OP_STAGES * 2 * TC_TILE_BYTES + (size_t)TC_RAW_STAGES * TC_CONV_THREADS * 2 * sizeof(float4) +
                      (size_t)128 * 65 * sizeof(float) + 1024;
  cudaFuncSetAttribute(herk_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(nframes, sms);
  herk_tc64_kernel<<<grid, TC_THREADS, smem, st>>>(in, frame_stride, chan_stride, N, nframes, out, (float)(1.0 / N),
                                                   (float)(0.5 / N), avg_method);
  return 1;
}

}  // namespace doa
