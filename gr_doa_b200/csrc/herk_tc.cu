// herk_tc.cu -- stage 1 for large arrays (M = 64): the sample covariance as a tensor-core complex HERK, 3xTF32 split.
//
// Replaces the cgemm of autocorrelate_impl::general_work (gr-doa lib/autocorrelate_impl.cc:106) where the Hermitian update
// costs 32.5 flop per input byte -- far on the FP32 side of the CUDA-core ridge.
//
// Formulation.  Read channel m's interleaved samples (re0, im0, re1, im1, ...) as a REAL row Z_m of length 2N, and let W_m
// be the same row with every (re, im) pair replaced by (im, -re).  Then
//     Re R[r][c] = sum_t re_r re_c + im_r im_c = (Z Z^T)[r][c]        Im R[r][c] = sum_t im_r re_c - re_r im_c = (W Z^T)[r][c]
// so ONE real product  D (128 x 64) = A (128 x 2N) * Z^T  with  A = [Z; W]  gives the whole covariance.  fp32 accuracy comes
// from the 3xTF32 split x = hi + lo (hi = x rounded to tf32, lo = x - hi, which the tensor core truncates to tf32):
//     D += A_hi Z_hi^T + A_hi Z_lo^T + A_lo Z_hi^T          (the dropped lo*lo term is 2^-22 relative)
// issued per K-step of 8 as TWO tcgen05 kind::tf32 UMMAs: M = 128, N = 128 for A_hi [Z_hi; Z_lo]^T (hi*hi into accumulator
// columns 0..63, hi*lo into 64..127; the Z_hi and Z_lo tiles are adjacent in shared memory, i.e. one 128-row K-major operand)
// and M = 128, N = 64 for A_lo Z_hi^T into columns 64..127.  A is written straight into TMEM (tcgen05.st; lane = row,
// column = k) and only the B operand Z lives in 128-byte-swizzled shared memory: with A in shared memory as well the
// tensor core's operand reads plus the converters' stores saturate the 128 B/clk shared-memory port (measured: 48 clk per
// M128 N64 K8 UMMA from shared memory, against 32 for the math).
//
// Accumulation.  The tensor core TRUNCATES the fp32 accumulator on every MMA (measured: the relative error of the diagonal
// grows as #MMAs * 2^-24 -- 3e-5 at N = 2048, 7e-5 at N = 4096 with one accumulator per frame).  So the 128-column
// accumulator is a ping-pong pair, and every TC_CHUNK stages (16 MMAs per column block) the finished one is folded into fp32
// REGISTERS with properly rounded adds: rel. Frobenius error 4e-7 at every N (tests/test_gpu_parity.py).
//
// One persistent CTA per SM, 17 warps; a "stage" is 16 complex samples = 32 K-values = 128 bytes per row:
//   warp 0           : MMA issuer.  The whole warp runs the loop (warp-uniform operands, so ptxas emits bare UTCHMMA) and one
//                      elected lane issues: wait full[s]; 4 K-steps x 2 MMAs; tcgen05.commit -> empty[s]; every TC_CHUNK
//                      stages and at the end of a frame commit -> chunk_full[b] and switch accumulators.
//   warps 1..12      : converters, three independent groups of 4 warps taking the stages round-robin.  Loader role: cp.async 16-byte pieces,
//                      eight lanes per 128-byte line, into the group's raw ring (HBM bytes in flight do not depend on registers).  Converter role:
//                      a thread IS one row of A (TMEM lanes belong to warp % 4): read the row's 32 raw values, split hi/lo
//                      (integer round-to-nearest on the tf32 boundary), (im, -re) for W rows, tcgen05.st into the A ring;
//                      Z rows also store the swizzled B tiles; fences, arrive on full[s].  They run ahead across frames.
//   warps 13..16     : adders + epilogue.  Thread = one TMEM lane (row of [Re R; Im R]), 64 fp32 register accumulators:
//                      wait chunk_full[b], tcgen05.ld both column blocks, add, arrive chunk_empty[b]; at the end of a SEGMENT
//                      (at most eight per frame, a stage count that depends on N alone) add the registers to the frame's
//                      running total in the staging area; at the frame end combine Re/Im, scale, forward-backward term, store R.
// Frames that do not fill a round of the grid are shared between CTAs by segment ranges, the partial sums folded in segment
// order by the last CTA to finish (launch_covariance_tc below): the association of a frame's sum never depends on the batch.
// TMEM (512 columns): accumulators @0 and @128, A ring @256 + 64 s (32 hi | 32 lo columns per stage, 4 stages).
//
// What bounds it (end of round 2): the shared-memory port.  The MMA warp's issue sequence is back-pressured by the tensor pipe (a second
// issuer warp changes nothing), the pipe needs ~800 clk per stage against 384 nominal and ncu shows it active 44 % of the time: it
// waits for its B operand behind the converters' traffic (65 KB per stage through a 128 B/clk port = 510 clk before arbitration).
// A timing-only build without the A_lo Z_hi^T MMAs (8 KB fewer operand reads per stage) is 8 % faster; that product is the
// (negated) transpose of A_hi Z_lo^T, so it could be dropped for real at the price of a per-segment exchange of the cross term.
// Measured (B200, 512 frames of 64 x 16384): 1.93 ms = 2.2 TB/s of input (2.10 ms before the split tail; tensor pipe 41 % busy
// under ncu then), against 5.30 ms for the CUDA-core tiled kernel.  What the round-2 profile (profiles/r02_ncu_herk_tc.txt) says about the rest: the converters never wait
// for the tensor core (1.6 % of their samples at empty[s]) and the MMA warp waits for them 29 % of its time -- a converter
// iteration is ~1900 clk of dependent latency for ~250 instructions (group barrier skew between Z rows, which also store the B
// tiles, and W rows: 22 %; raw-ring loads: 16 %; fences and the TMEM store drain: ~10 %).  Tried and measured: converting before
// the wait on empty[s] (no change), a fully unrolled MMA-issue loop with compile-time operands (6 % SLOWER: tighter MMA issue
// takes shared-memory cycles from the converters), line-contiguous cp.async with an XOR-swizzled raw ring (8x fewer
// shared-memory wavefronts per copy, 1 % faster: kept), a third converter group (3 % faster: kept), the B tiles written half by the
// Z rows and half by the W rows (10 % SLOWER: every thread then pays the generic->async proxy fence, which turns out to be the
// expensive part of a Z-row thread's iteration, not its sixteen stores).
#include "cov_device.cuh"
#include "herk_geometry.h"

#include <algorithm>

namespace doa {
namespace {

constexpr int TC_M = 64;                 // channels
constexpr int TC_ROWS = 128;             // rows of [Z; W]
constexpr int TC_OP_STAGES = 4;          // operand ring depth: A = 64 TMEM columns (hi | lo), B = 16 KB smem (hi | lo tiles of 64 x 128 B)
constexpr int TC_RAW_STAGES = 5;         // raw fp32 ring depth PER converter group (8 KB per stage)
constexpr int TC_CHUNK = HERK_CHUNK;     // stages (= 16 hi*hi MMAs) per big-accumulator chunk
constexpr int TC_MAX_SEGS = HERK_MAX_SEGS;   // segments per frame (the canonical association of a frame's sum; also the widest split)
constexpr int TC_TILE_BYTES = TC_M * 128; // one B tile (Z_hi or Z_lo): 64 rows x 128 B
#ifndef DOA_HERK_EXP
#define DOA_HERK_EXP 0      // timing experiments only (tools/herk_bound_exp.sh): bit 0 no A_lo Z_hi^T MMA, bit 1 no A_lo store, bit 2 no B tile stores
#endif
#ifndef DOA_HERK_GROUPS
#define DOA_HERK_GROUPS 3   // measured: 2 groups 2.15 ms, 3 groups 2.08 ms per 512 frames (the converters are latency-bound: more of them in flight)
#endif
constexpr int TC_GROUPS = DOA_HERK_GROUPS;       // converter groups of 4 warps, taking stages round-robin
constexpr int TC_CONV_WARPS = 4 * TC_GROUPS, TC_ADD_WARPS = 4;
constexpr int TC_RAW_STAGE_BYTES = 8 * TC_M * 16;   // one raw stage: 8 pieces x 64 channels x 16 B
constexpr int TC_CONV_THREADS = TC_CONV_WARPS * 32;
#ifndef DOA_HERK_PAIRED
#define DOA_HERK_PAIRED 0   // 1: the Z row and the W row of a channel sit in ONE warp (TMEM lanes 32k + j and 32k + 16 + j: channel 16k + j), so the
                            //    two threads' loads of the raw row are one shared-memory broadcast (8 KB less port traffic per stage) and the
                            //    (re, im) -> (im, -re) map is done by selects ahead of one uniform hi/lo split (same bits; tested).  Measured
                            //    SLOWER, 2.30 against 2.055 ms per 592 frames: all four warps of a group then store B tiles and pay the
                            //    generic->async proxy fence.  0 (shipped): Z rows in lanes 0..63, W rows in 64..127
#endif
#ifndef DOA_HERK_ISSUERS
#define DOA_HERK_ISSUERS 1  // MMA-issuing warps (1: warp 0 alone; 2: warp 0 and the last warp take the stages alternately)
#endif
constexpr int TC_ISSUERS = DOA_HERK_ISSUERS;
constexpr int TC_THREADS = (TC_ISSUERS + TC_CONV_WARPS + TC_ADD_WARPS) * 32;
constexpr int TC_ISSUER2_WARP = (TC_ISSUERS == 2) ? 1 + TC_CONV_WARPS + TC_ADD_WARPS : -1;   // the second issuer is the LAST warp: the others keep their TMEM lane quarters
constexpr int TC_TMEM_COLS = 512;        // D[0] @0, D[1] @128 (64 hi*hi columns | 64 cross-term columns), A ring @256 + 64 s (hi | lo)
constexpr uint32_t TC_A_COL = 256;
constexpr int TC_STG_IM = TC_M * 65 + 16;  // staging area: Re rows at r * 65, Im rows at TC_STG_IM + r * 65 (the 16 keeps the two halves of a warp on different banks)
// staging row of TMEM lane `row` (= the row of [Re R; Im R] it accumulates)
__device__ __forceinline__ int stg_row_off(int row) {
#if DOA_HERK_PAIRED
  const int j = row & 31, ch = (row >> 5) * 16 + (j & 15);
  return (j >= 16) ? TC_STG_IM + ch * 65 : ch * 65;
#else
  return (row >= TC_M) ? TC_STG_IM + (row - TC_M) * 65 : row * 65;
#endif
}

#ifndef DOA_HERK_TRACE
#define DOA_HERK_TRACE 0    // 1: CTA 0 records clock64 at the ring's hand-offs for its first 1024 stages (tools/herk_trace.py)
#endif
#if DOA_HERK_TRACE
__device__ long long g_herk_trace[5][1024];   // [full seen | MMAs issued | before empty wait | empty seen | full arrived][stage]
#define HERK_TRACE(ev, q, cond) do { if (blockIdx.x == 0 && (q) < 1024 && (cond)) g_herk_trace[ev][q] = clock64(); } while (0)
#else
#define HERK_TRACE(ev, q, cond) do { } while (0)
#endif
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
// A operand from TMEM (lane = row, column = k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                  "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                  "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                  "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
// round to nearest (ties away) on the tf32 boundary: add half an ulp of the 10-bit mantissa, clear the low 13 bits
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void cp_async16z(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// 16 consecutive columns of this thread's TMEM lane; no wait (the caller batches loads, then tmem_ld_wait())
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__global__ void __launch_bounds__(TC_THREADS, 1)
herk_tc64_kernel(const float2* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
                 float2* __restrict__ out, float scale, float bscale, int avg_method, const float2* __restrict__ gains,
                 int nfull, int seg_len, int tail_S, int tail_sps, float* __restrict__ ws_part, unsigned* __restrict__ ws_cnt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by OFFSETTING the shared array (an integer round trip would turn every access into a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* op = smem;                                                                  // B ring [OP_STAGES][Z_hi tile | Z_lo tile]
  float4* raw = reinterpret_cast<float4*>(smem + TC_OP_STAGES * 2 * TC_TILE_BYTES);      // [groups][RAW_STAGES][8 pieces][64 channels ^ piece]
  float* stg = reinterpret_cast<float*>(smem + TC_OP_STAGES * 2 * TC_TILE_BYTES + TC_GROUPS * TC_RAW_STAGES * TC_RAW_STAGE_BYTES);   // [128][65]
  __shared__ uint64_t full_bar[TC_OP_STAGES], empty_bar[TC_OP_STAGES], chunk_full[2], chunk_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int tail_last_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < TC_OP_STAGES; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&chunk_full[b], 1); mbar_init(&chunk_empty[b], TC_ADD_WARPS * 32); }
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
  // Work units of this CTA: whole frames blockIdx.x + r * gridDim.x < nfull, then at most one TAIL unit -- a range of whole
  // segments of one of the nframes - nfull frames that do not fill a round of the grid (split-K, see launch_covariance_tc).
  const int my_frames = herk_whole_frames((int)blockIdx.x, (int)gridDim.x, nfull);
  const int nseg = (spf + seg_len - 1) / seg_len;
  const HerkTail tl_ = herk_tail((int)blockIdx.x, nframes, nfull, tail_S, tail_sps, seg_len, spf);
  const bool has_tail = tl_.has != 0;
  const int tail_idx = tl_.idx, tail_seg0 = tl_.seg0, tail_start = tl_.start, tail_count = tl_.count;
  const long long total = (long long)my_frames * spf + tail_count;

  if (warp == 0 || warp == TC_ISSUER2_WARP) {
    // ================================ MMA issuers ================================
    // A whole warp runs the loop (warp-uniform control flow and operands: ptxas emits bare UTCHMMA with uniform registers);
    // one elected lane issues.  A descriptor differs from the ring's base descriptor only in its low word (start address).
    // -DDOA_HERK_ISSUERS=2: two issuer warps take the stages alternately -- both walk the same stage sequence, each does the waits
    // and the set-up of its own stages, and a token (named barriers 5 / 6) serialises the issue itself, so the MMAs still enter
    // the pipe in stage order (same accumulation order, same bits; tested).  Built because the trace (tools/herk_trace.py,
    // profiles/r02_herk_trace.log) shows the single issuer ~850 clk per stage behind operands that have been ready for 1900 clk.
    // Measured: NO gain (2.09 against 2.085 ms per 592 frames) -- the issue sequence is back-pressured by the tensor pipe, which
    // takes ~800 clk for a stage's eight MMAs against 384 nominal: it starves on its B operand (shared-memory port: 24 KB of
    // operand reads + 33 KB of converter loads / stores + 8 KB of ring fills per stage).  Default: one issuer.
    {
      const int iw = (warp == 0) ? 0 : 1;
      const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_M >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      const uint32_t idesc128 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * TC_M >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      const uint64_t desc0 = umma_desc(smem_u32(op));
      const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
      int s = 0; uint32_t ph = 0; int sf = 0;
      int fr = 0, cur = (my_frames > 0) ? spf : tail_count;   // unit index, stages of the current unit
      int cb = 0, in_chunk = 0; uint32_t ce_ph = 0u;   // parity bits, one per buffer (bit b)
      int chunks = 0;                                  // saturating: only "< 2" matters
      const int ntot = (int)total;                     // stages of this CTA (the launcher keeps it below 2^31)
      for (int q = 0; q < ntot; ++q) {
        // bookkeeping of stage q, identical in both warps
        bool ce_wait = false; uint32_t ce_par = 0u;
        if (in_chunk == 0) {                                        // first stage of a chunk: D[cb] must have been folded
          if (chunks >= 2) { ce_wait = true; ce_par = (ce_ph >> cb) & 1u; ce_ph ^= 1u << cb; }
          else ++chunks;
        }
        ++sf; ++in_chunk;
        const bool frame_end = (sf == cur);
        const bool chunk_end = (in_chunk == TC_CHUNK) || frame_end;   // segments are whole chunks: no chunk straddles one
        if (TC_ISSUERS == 1 || (q & 1) == iw) {
          if (ce_wait) mbar_wait(&chunk_empty[cb], ce_par);
          mbar_wait(&full_bar[s], ph);
          HERK_TRACE(0, q, lane == 0);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b32 = desc_lo0 + (uint32_t)s * (2u * TC_TILE_BYTES >> 4);
          const uint32_t ahi = tmem_d + TC_A_COL + (uint32_t)s * 64u, alo = ahi + 32u;
          const uint32_t dacc = tmem_d + (uint32_t)cb * 128u;
          if (TC_ISSUERS == 2 && q > 0) {                           // the other warp has issued stage q - 1
            if (iw) asm volatile("bar.sync 5, 64;" ::: "memory"); else asm volatile("bar.sync 6, 64;" ::: "memory");
          }
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // the Z_hi and Z_lo tiles are contiguous: one 128-row B operand
              const uint64_t db = ((uint64_t)desc_hi << 32) | (b32 + (uint32_t)k * 2u);
              umma_tf32_ts(dacc, ahi + k * 8, db, idesc128, k == 0 ? (uint32_t)(in_chunk != 1) : 1u);   // A_hi [Z_hi; Z_lo]^T
#if !(DOA_HERK_EXP & 1)
              umma_tf32_ts(dacc + 64u, alo + k * 8, db, idesc64, 1u);                                   // A_lo Z_hi^T
#endif
            }
            umma_commit(&empty_bar[s]);                             // stage s reusable once these MMAs retire
            if (chunk_end) umma_commit(&chunk_full[cb]);            // the pipe is in order: every MMA of the chunk has retired then
          }
          __syncwarp();
          if (TC_ISSUERS == 2 && q + 1 < ntot) {
            if (iw) asm volatile("bar.arrive 6, 64;" ::: "memory"); else asm volatile("bar.arrive 5, 64;" ::: "memory");
          }
          HERK_TRACE(1, q, lane == 0);
        }
        if (chunk_end) { cb ^= 1; in_chunk = 0; }
        if (frame_end) { sf = 0; ++fr; cur = (fr < my_frames) ? spf : tail_count; }
        if (++s == TC_OP_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp <= TC_CONV_WARPS) {
    // ================================ converters ================================
    // TC_GROUPS independent groups of 4 warps take the stages round-robin (group g: stages q = g, g + TC_GROUPS, ...), each with its own raw ring
    // and named barrier, so one group's latencies (barrier, ring waits, TMEM store drain) hide behind the other's work.
    const int g = (warp - 1) >> 2;
    const int gt = ((warp - 1) & 3) * 32 + lane;   // 0..127 within the group
    // Loader role (a piece = 2 complex samples = 16 B, a stage = 8 pieces per channel): EIGHT consecutive lanes copy the eight 16-byte pieces of ONE channel's 128-byte line (copy j of a thread:
    // channel 16 j + gt / 8, piece gt % 8), so a warp-wide cp.async reads four whole lines.  With two lanes per line and four
    // copies each (round 1), every lane's 16 bytes came back from L2 in a sector of their own and were written to shared memory as a
    // separate wavefront: 32 wavefronts per warp-wide copy instead of 4, 511 of the ~830 shared-memory wavefronts of a stage
    // (ncu source page: 67 M wavefronts per LDGSTS against 8.4 M ideal) -- the port, not the tensor core, set the stage time.
    // Raw ring layout [stage][piece][channel ^ piece]: the XOR keeps both the loaders' stores (fixed channel, 8 pieces) and
    // the converters' loads (fixed piece, 32 consecutive channels) conflict-free.
    const int lpc = gt & 7, lch0 = gt >> 3;         // piece, channel of copy 0
    float4* graw = raw + (size_t)g * TC_RAW_STAGES * (8 * TC_M);
    float4* lraw = graw + lpc * TC_M;
    // converter role: this thread IS row (warp & 3) * 32 + lane of A = [Z; W] (TMEM lanes are per-warp quarters)
#if DOA_HERK_PAIRED
    const int cch = (warp & 3) * 16 + (lane & 15);
    const bool is_w = lane >= 16;
#else
    const int row = (warp & 3) * 32 + lane, cch = row & (TC_M - 1);
    const bool is_w = row >= TC_M;
#endif
    const uint32_t a_lane = tmem_d + TC_A_COL + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t row_off = ((uint32_t)cch >> 3) * 1024u + ((uint32_t)cch & 7u) * 128u, row_x = (uint32_t)cch & 7u;
    constexpr int NG = TC_GROUPS;
    const int nunits = my_frames + (has_tail ? 1 : 0);
    // loader position: unit iu, stage isf of its icnt stages (isf + ioff = stage of the frame)
    int iu = 0, isf = g, icnt = 0, ioff = 0;
    const float2* ibase = in;
    auto set_unit = [&]() {
      const bool tl = iu >= my_frames;
      icnt = tl ? tail_count : spf;
      ioff = tl ? tail_start : 0;
      const long long fi = tl ? (long long)nfull + tail_idx : (long long)blockIdx.x + (long long)iu * gridDim.x;
      ibase = in + fi * frame_stride + (long long)lch0 * chan_stride;
    };
    auto normalize = [&]() {
      while (isf >= icnt) { isf -= icnt; if (++iu >= nunits) break; set_unit(); }
    };
    if (nunits > 0) { set_unit(); normalize(); }
    const long long cstep = 16 * chan_stride;      // copy j reads channel lch0 + 16 j
    int irs = 0; long long qi = g;
    auto issue = [&]() {
      const int tj = (ioff + isf) * 16 + 2 * lpc;  // this thread's two complex samples of the stage
      const int nb = (tj < N) ? 16 : 0;            // N is even (launcher): a 16-byte piece is all in or all out
      float4* dst = lraw + (size_t)irs * (8 * TC_M);
      const float2* srcp = ibase + (nb ? tj : 0);
#pragma unroll
      for (int j = 0; j < 4; ++j) cp_async16z(dst + ((lch0 + 16 * j) ^ lpc), srcp + j * cstep, nb);
      isf += NG;
      normalize();
      if (++irs == TC_RAW_STAGES) irs = 0;
      qi += NG;
    };
#pragma unroll
    for (int p = 0; p < TC_RAW_STAGES - 2; ++p) { if (qi < total) issue(); asm volatile("cp.async.commit_group;" ::: "memory"); }
    int rs = 0;
    for (long long q = g; q < total; q += NG) {
      // the slot refilled here was read two iterations ago at the latest, before the barrier every thread passed last time
      if (qi < total) issue();
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" :: "n"(TC_RAW_STAGES - 2) : "memory");
      asm volatile("bar.sync %0, 128;" :: "r"(2 + g) : "memory");   // every loader's pieces of this raw stage have landed
      const int s = (int)(q & (TC_OP_STAGES - 1));
      HERK_TRACE(2, q, gt == 0);
      if (q >= TC_OP_STAGES) {                                      // MMAs of the previous use of this stage have retired
        mbar_wait(&empty_bar[s], (uint32_t)((q >> 2) + 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      HERK_TRACE(3, q, gt == 0);
      const float4* src = graw + (size_t)rs * (8 * TC_M);
      uint8_t* thi = op + (size_t)s * 2 * TC_TILE_BYTES + row_off;
#pragma unroll
      for (int h = 0; h < 2; ++h) {                 // 16 K-values (4 pieces) at a time
        float hi[16], lo[16];
        float4 v[4];                                // (re0, im0, re1, im1) each
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = src[(4 * h + j) * TC_M + (cch ^ (4 * h + j))];
#if DOA_HERK_PAIRED
        // W row: every (re, im) pair becomes (im, -re) BEFORE the split (the split is odd-symmetric: the same bits as
        // splitting first), by selects, so that the two halves of the warp run one instruction stream
        {
          const uint32_t sgn = is_w ? 0x80000000u : 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a0 = is_w ? v[j].y : v[j].x, b0 = __uint_as_float(__float_as_uint(is_w ? v[j].x : v[j].y) ^ sgn);
            const float a1 = is_w ? v[j].w : v[j].z, b1 = __uint_as_float(__float_as_uint(is_w ? v[j].z : v[j].w) ^ sgn);
            const float x[4] = {a0, b0, a1, b1};
#pragma unroll
            for (int e = 0; e < 4; ++e) { hi[4 * j + e] = to_tf32(x[e]); lo[4 * j + e] = x[e] - hi[4 * j + e]; }
          }
        }
#else
        if (!is_w) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { hi[4 * j + e] = to_tf32(x[e]); lo[4 * j + e] = x[e] - hi[4 * j + e]; }
          }
        } else {                                    // W row: every (re, im) pair becomes (im, -re)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const float hre = to_tf32(x[e]), him = to_tf32(x[e + 1]);
              hi[4 * j + e] = him;                 lo[4 * j + e] = x[e + 1] - him;
              hi[4 * j + e + 1] = __uint_as_float(__float_as_uint(hre) ^ 0x80000000u);   lo[4 * j + e + 1] = hre - x[e];
            }
          }
        }
#endif
        if (!is_w && !(DOA_HERK_EXP & 4)) {             // Z rows are also the B operand: swizzled K-major tiles in shared memory
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t o = (((uint32_t)(4 * h + j)) ^ row_x) << 4;
            *reinterpret_cast<float4*>(thi + o) = make_float4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<float4*>(thi + TC_TILE_BYTES + o) = make_float4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
        }
        tmem_st16(a_lane + (uint32_t)s * 64u + (uint32_t)h * 16u, hi);
#if !(DOA_HERK_EXP & 2)
        tmem_st16(a_lane + (uint32_t)s * 64u + 32u + (uint32_t)h * 16u, lo);
#endif
      }
      if (!is_w) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&full_bar[s]);
      HERK_TRACE(4, q, gt == 0);
      if (++rs == TC_RAW_STAGES) rs = 0;
    }
  } else {
    // ============================ adders + epilogue ============================
    const int at = tid - (1 + TC_CONV_WARPS) * 32;                 // 0..127
    const int q4 = warp & 3;                                       // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;                                // row of [Re R; Im R]
    const uint32_t lane_addr = tmem_d + ((uint32_t)(q4 * 32) << 16);
    const int srow = stg_row_off(row);                             // this row's place in the staging area
    uint32_t cf_ph = 0u;                                           // parity bits, one per buffer
    int cb = 0;
    const int nunits = my_frames + (has_tail ? 1 : 0);
    for (int un = 0; un < nunits; ++un) {
      const bool tl = un >= my_frames;
      const int cnt = tl ? tail_count : spf;
      const long long fcur = tl ? (long long)nfull + tail_idx : (long long)blockIdx.x + (long long)un * gridDim.x;
      const int segs = (cnt + seg_len - 1) / seg_len;
      // A frame is the sum of its SEGMENTS (seg_len stages each, a function of N alone) taken in order, each segment the
      // in-order sum of its chunks: the association is the same whether one CTA walks the frame or several share it.
      for (int sg = 0; sg < segs; ++sg) {
        float acc[TC_M];
#pragma unroll
        for (int c = 0; c < TC_M; ++c) acc[c] = 0.0f;
        const int nchunks = (min(seg_len, cnt - sg * seg_len) + TC_CHUNK - 1) / TC_CHUNK;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&chunk_full[cb], (cf_ph >> cb) & 1u); cf_ph ^= 1u << cb;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int h = 0; h < 4; ++h) {               // columns 16h.. of the hi*hi block and of the cross-term block
            uint32_t v0[16], v1[16];
            tmem_ld16(lane_addr + (uint32_t)cb * 128u + (uint32_t)h * 16u, v0);
            tmem_ld16(lane_addr + (uint32_t)cb * 128u + 64u + (uint32_t)h * 16u, v1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[16 * h + j] += __uint_as_float(v0[j]) + __uint_as_float(v1[j]);
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&chunk_empty[cb]);
          cb ^= 1;
        }
        if (!tl) {                                    // running total of the frame in the staging area (own row: no barrier)
          if (sg == 0) {
#pragma unroll
            for (int c = 0; c < TC_M; ++c) stg[srow + c] = acc[c];
          } else {
#pragma unroll
            for (int c = 0; c < TC_M; ++c) stg[srow + c] += acc[c];
          }
        } else {                                      // shared frame: the segment goes to the workspace
          float4* p = reinterpret_cast<float4*>(ws_part + (((size_t)tail_idx * nseg + tail_seg0 + sg) * TC_ROWS + row) * TC_M);
#pragma unroll
          for (int c = 0; c < TC_M / 4; ++c) __stcg(p + c, make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]));
        }
      }
      if (tl) {
        // The CTA that delivers the frame's last range folds all segments, in segment order, and emits the frame.
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (at == 0) {
          const unsigned ticket = atomicAdd(&ws_cnt[tail_idx], 1u);
          const int last = (ticket == (unsigned)(tail_S - 1));
          if (last) ws_cnt[tail_idx] = 0u;           // ready for the next launch (nobody else touches it in this one)
          tail_last_s = last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (!tail_last_s) continue;
        __threadfence();
        const float4* p = reinterpret_cast<const float4*>(ws_part + (((size_t)tail_idx * nseg) * TC_ROWS + row) * TC_M);
#pragma unroll 2
        for (int q = 0; q < TC_M / 4; ++q) {          // four columns at a time, every segment's load in flight together
          float4 v[TC_MAX_SEGS];
#pragma unroll
          for (int sg = 0; sg < TC_MAX_SEGS; ++sg)
            v[sg] = (sg < nseg) ? __ldcg(p + (size_t)sg * (TC_ROWS * TC_M / 4) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 a = v[0];
#pragma unroll
          for (int sg = 1; sg < TC_MAX_SEGS; ++sg)
            if (sg < nseg) { a.x += v[sg].x; a.y += v[sg].y; a.z += v[sg].z; a.w += v[sg].w; }
          float* d = stg + srow + 4 * q;
          d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float2* o = out + fcur * (long long)(TC_M * TC_M);
      for (int e = at; e < TC_M * TC_M; e += 128) {
        const int r = e % TC_M, c = e / TC_M;
        float2 v = apply_gain(make_float2(stg[r * 65 + c] * scale, stg[TC_STG_IM + r * 65 + c] * scale), gains, r, c);
        if (avg_method == 1) {
          const int rr = TC_M - 1 - r, cc = TC_M - 1 - c;
          const float2 w = apply_gain(make_float2(stg[rr * 65 + cc] * scale, stg[TC_STG_IM + rr * 65 + cc] * scale), gains, rr, cc);
          v.x = __fadd_rn(__fmul_rn(0.5f, v.x), __fmul_rn(bscale, w.x));
          v.y = __fadd_rn(__fmul_rn(0.5f, v.y), __fmul_rn(bscale, -w.y));
        }
        o[e] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");               // staging area free again
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(TC_TMEM_COLS));
}

}  // namespace

// Workspace of the split tail (below): 256 frame counters, then per shareable frame TC_MAX_SEGS partial sums of 128 x 64 floats.
constexpr int TC_WS_FRAMES = HERK_WS_FRAMES;   // frames a launch may share between CTAs (at most half the grid)
constexpr size_t TC_WS_CNT_BYTES = 1024;
size_t covariance_tc_workspace_bytes() {
  return TC_WS_CNT_BYTES + (size_t)TC_WS_FRAMES * TC_MAX_SEGS * TC_ROWS * TC_M * sizeof(float);
}

// Returns 1 if launched, 0 if the shape is not covered (caller uses the CUDA-core kernels).
//
// Split tail (ws != null: covariance_tc_workspace_bytes() of zero-initialised device memory owned by the caller's handle and
// used by one stream at a time).  A persistent CTA takes whole frames, so nframes = r * SMs + t leaves the last round to t
// SMs (512 frames on 148 SMs: a fourth round on 68 of them, 13 % of the kernel).  When t <= SMs / 2 those t frames are
// shared instead: S CTAs take consecutive ranges of whole segments of one frame, park the per-segment partial sums in the
// workspace, and the CTA whose ticket is the last folds ALL segments of the frame in segment order and emits it.  Whole
// frames are accumulated with the same association (segments in order), so a frame's bits do not depend on whether it was
// shared, on the batch size or on its position in the batch.  A call of a few frames (GNU Radio hands a block a handful per
// work()) runs on up to 8 SMs per frame instead of one.
int launch_covariance_tc(const float2* in, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                         int avg_method, float2* out, cudaStream_t st, const float2* gains, void* ws) {
  if (M != TC_M || nframes <= 0) return 0;
  const bool aligned = (N % 2 == 0) && (frame_stride % 2 == 0) && (chan_stride % 2 == 0) &&
                       ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);
  if (!aligned) return 0;
  const size_t smem = (size_t)TC_OP_STAGES * 2 * TC_TILE_BYTES + (size_t)TC_GROUPS * TC_RAW_STAGES * TC_RAW_STAGE_BYTES +
                      (size_t)(128 * 65 + 16) * sizeof(float) + 1024;
  cudaFuncSetAttribute(herk_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const HerkGeometry g = herk_geometry(nframes, N, sms, ws != nullptr && dev_option(OPT_HERK_SPLIT, 1));   // herk_geometry.h
  if (((long long)nframes / sms + 2) * g.spf > 0x7fffffffLL) return 0;        // a CTA counts its stages in an int
  const int grid = g.grid, nfull = g.nfull, seg_len = g.seg_len, S = g.S, sps = g.sps;
  herk_tc64_kernel<<<grid, TC_THREADS, smem, st>>>(in, frame_stride, chan_stride, N, nframes, out, (float)(1.0 / N),
                                                   (float)(0.5 / N), avg_method, gains, nfull, seg_len, S, sps,
                                                   ws ? reinterpret_cast<float*>(static_cast<char*>(ws) + TC_WS_CNT_BYTES) : nullptr,
                                                   static_cast<unsigned*>(ws));
  return 1;
}

}  // namespace doa

#if DOA_HERK_TRACE
extern "C" int doa_herk_trace_dump(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, doa::g_herk_trace, sizeof(long long) * 5 * 1024);
}
#endif
