// fused.cu -- the whole chain (autocorrelate -> MUSIC_lin_array -> find_local_max) in ONE persistent, warp-specialised kernel.
//
// Run back to back, the three stage kernels leave the machine half idle twice: the covariance is HBM-bound and uses ~40 %
// of the issue slots, the eigendecomposition and the scan are issue-bound and move no data.  Here one CTA of 16 warps per SM
// owns a contiguous range of frames and keeps every intermediate (R, G, u) in shared memory:
//   producer warps: stream their frames through per-lane cp.async rings (WS_STAGES x 4 KB per warp, so the bytes in flight
//       depend neither on registers nor on what the other warps do), accumulate the covariance (cov_device.cuh), emit R into
//       one of WS_NBUF tile buffers and signal "full"; they wait for the consumers only when every tile buffer is taken.
//   consumer warps: wait for a full tile (TILE = 32/M matrices per consumer warp), run Jacobi on their own matrices
//       (eig_device.cuh), release the tile buffer, scan + pick + refine their frames (scan_device.cuh) and write K peaks.
// Hand-off with named barriers (bar.arrive / bar.sync): FULL0/1 and EMPTY0/1 between the two groups; a consumer warp owns
// its matrices end to end, so the consumers need no barrier among themselves.
// The device code is the stage kernels' own and per-entry operation order is unchanged, so the fused path is bit-identical to
// the three-kernel path (tested).  Measured on B200 at cfg3: 2.21 ms (three kernels) -> 1.64 ms; a phase-structured variant
// (two CTAs/SM alternating stream / Jacobi / scan phases behind __syncthreads) measured 2.02 ms and was dropped.
#include "cov_device.cuh"
#include "eig_os_device.cuh"
#include "scan_device.cuh"

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link dependency)

#include <algorithm>
#include <cstring>

namespace doa {
namespace {

constexpr int BAR_FULL = 1;                     // named barriers: FULL b = 1 + b, EMPTY b = 1 + WS_NBUF + b

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }
// TMA = true (N a multiple of 64): a ring stage is filled by M bulk async copies of 512 contiguous bytes (one per channel),
// issued by one lane and tracked by an mbarrier with complete_tx -- no per-lane addresses, no LSU instructions in the
// other 31 lanes; the ring layout ([stage][channel][32 lanes] float4) is the same as with per-lane cp.async.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_addr(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// FILL 3: ONE tensor-map copy per stage -- cp.async.bulk.tensor.3d of the box [1 frame][M channels][64 samples] (M x 512 bytes),
// coordinates (sample, channel, frame); samples beyond the frame are zero-filled by the copy engine.  SASS: UTMALDG.3D.
__device__ __forceinline__ void tma_box_g2s(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               :: "r"(smem_addr(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar)) : "memory");
}

// S = float2: fc32 samples, a ring slot is one float4 (two samples) per lane and channel; S = unsigned: sc16 samples, a ring
// slot is one uint2 (the same two samples in 8 bytes).  Either way lane l of chunk c holds samples 64 c + 2 l and + 1, so the
// two formats accumulate in the same order.
template <typename S> struct RingSlot { typedef float4 type; };
template <> struct RingSlot<unsigned> { typedef uint2 type; };

// FILL: how a producer warp fills a ring stage (the stage's layout is the same in all three).
//   0  lane l copies samples 2l, 2l+1 of the chunk for each of the M channels: M cp.async per lane, each with its own 64-bit
//      source address (k * chan_stride is a run-time stride);
//   1  bulk (TMA) copies issued by one lane (fc32 only);
//   3  one tensor-map TMA copy per stage (fc32, dense [B][M][N] batches only);
//   2  lane l copies for channel l / (32/M) only: its M cp.async are 64 bytes apart (sc16: 32) in that channel, i.e. ONE 64-bit
//      address and immediates -- about 30 fewer address instructions per chunk in an issue-bound loop.  A warp-wide copy
//      then touches M segments of 64 bytes instead of 512 contiguous bytes (the same 16 sectors).  Measured at cfg3
//      (tools/ws_exp.py 80824 80824c): 1.705 against 1.688 ms -- the 12 instructions it saves per 228-instruction chunk do
//      not show, so FILL 0 stays the default (dev knob ws_fill = 2 selects this one; bit-identical, tested).
template <int M, int KL, int WS_P, int WS_C, int WS_STAGES, int WS_NBUF, int FILL, typename S>
__global__ void __launch_bounds__((WS_P + WS_C) * 32, 1)
chain_ws_kernel(const S* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
                int avg_method, float scale, float bscale, int T, int max_sweeps, const float* __restrict__ zpair,
                const float2* __restrict__ zplain, const float2* __restrict__ Vtab, const float* __restrict__ xaxis, int P, int K,
                float* __restrict__ out_val, float* __restrict__ out_loc, int* __restrict__ out_bin,
                const float2* __restrict__ gains, float2* __restrict__ G_out, float2* __restrict__ u_out,
                const __grid_constant__ CUtensorMap tmap) {
  static_assert(M == 8 || M == 4, "instantiated for 8 and 4 lanes per matrix");
  constexpr bool TMA = FILL == 1 || FILL == 3;
  static_assert(!TMA || sizeof(S) == 8, "bulk ring fills are an fc32 variant");
  typedef typename RingSlot<S>::type Slot;
  constexpr bool SC16 = sizeof(S) == 4;
  constexpr int TILE = WS_C * 32 / M;          // frames per tile: every consumer warp owns 32/M of them
  constexpr int BAR_EMPTY = BAR_FULL + WS_NBUF; // WS_NBUF tile buffers between producers and consumers
  static_assert(BAR_EMPTY + WS_NBUF <= 16, "named barriers");
  constexpr int MM = M * M;
  constexpr int NTHREADS = (WS_P + WS_C) * 32;
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const ZTab zt = (zpair != nullptr) ? ztab_fill(smem, zpair, P) : ZTab();   // no scan table: the split form without a scan (Root-MUSIC chain)
  float2* Rbuf = reinterpret_cast<float2*>(smem + ((ztab_floats(P) + 3) & ~(size_t)3));   // [WS_NBUF][TILE][MM]
  float2* Gs = Rbuf + WS_NBUF * TILE * MM;                                 // [TILE][MM]
  float2* us = Gs + TILE * MM;                                       // [TILE][M]
  float* red = reinterpret_cast<float*>(us + TILE * M);              // [WS_P][MM]
  Slot* ring = reinterpret_cast<Slot*>(red + WS_P * MM);             // [WS_P][WS_STAGES][M][32]
  if constexpr (FILL == 3) {   // a tensor-map copy wants a 128-byte aligned destination (offset, not an integer round trip: keeps LDS)
    char* rb = reinterpret_cast<char*>(ring);
    ring = reinterpret_cast<Slot*>(rb + ((128u - (smem_addr(rb) & 127u)) & 127u));
  }
  __shared__ unsigned long long ring_bar[TMA ? WS_P * WS_STAGES : 1];
  if constexpr (TMA) {
    if (threadIdx.x < WS_P * WS_STAGES) mbar_init(&ring_bar[threadIdx.x], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < WS_NBUF * TILE * MM; i += blockDim.x) Rbuf[i] = make_float2(0.f, 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  const long long per = nframes / gridDim.x, rem = nframes % gridDim.x;
  const long long lo = blockIdx.x * per + min((long long)blockIdx.x, rem);
  const int nf = (int)(per + (blockIdx.x < rem ? 1 : 0));            // frames of this CTA
  const int ntiles = (nf + TILE - 1) / TILE;

  // More than 16 warps do not fit at the producers' 128 registers: the kernel is then compiled for (and launched with) 96
  // per thread, the consumer warpgroups hand registers back and the producer warpgroup takes them (setmaxnreg works on
  // aligned groups of 4 warps, hence WS_P == 4).
  constexpr bool REALLOC = (WS_P + WS_C) > 16;
  constexpr int REG_P = (WS_P == 4) ? 128 : 120, REG_C = (WS_P == 4) ? 88 : 80;
  static_assert(!REALLOC || (WS_P % 4 == 0 && WS_C % 4 == 0 && (WS_P * REG_P + WS_C * REG_C) * 32 <= 65536 / NTHREADS / 8 * 8 * NTHREADS),
                "register budget of the re-allocated configuration");
  if (warp < WS_P) {
    // ================================ producer ================================
    if constexpr (REALLOC) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(REG_P));
    const int w = warp;
    const int NCH = (N + 63) / 64;                                   // 64-sample chunks per frame (2 samples per lane)
    const int nfw = (w < nf) ? (nf - w + WS_P - 1) / WS_P : 0;       // frames of this warp: w, w+P, w+2P, ...
    const int total = nfw * NCH;                                     // chunks of this warp, frame-major
    Slot* myring = ring + (size_t)w * WS_STAGES * M * 32;
    // issue cursor (frame base pointer, chunk in frame, ring stage) advances incrementally: no divisions in the loop
    constexpr int LPC = 32 / M;                                      // FILL 2: lanes per channel
    const S* ibase = in + (lo + w) * frame_stride;
    long long iframe = lo + w;                                       // FILL 3: the frame coordinate of the tensor map
    if constexpr (FILL == 2) ibase += (long long)(lane / LPC) * chan_stride + (lane % LPC) * 2;
    int ic = 0, istage = 0, issued = 0;
    auto issue = [&]() {
      if constexpr (FILL == 3) {
        if (lane == 0) {
          unsigned long long* bar = &ring_bar[w * WS_STAGES + istage];
          mbar_expect_tx(bar, M * 512u);
          tma_box_g2s(myring + (size_t)istage * M * 32, &tmap, ic * 64, 0, (int)iframe, bar);
        }
      } else if constexpr (TMA) {
        if (lane == 0) {
          unsigned long long* bar = &ring_bar[w * WS_STAGES + istage];
          mbar_expect_tx(bar, M * 512u);
          Slot* dst = myring + (size_t)istage * M * 32;
#pragma unroll
          for (int k = 0; k < M; ++k) bulk_g2s(dst + k * 32, ibase + (long long)k * chan_stride + ic * 64, 512u, bar);
        }
      } else if constexpr (FILL == 2) {
        const S* src = ibase + ic * 64;
        Slot* dst = myring + (size_t)istage * M * 32 + (lane / LPC) * 32 + (lane % LPC);
        if (ic * 64 + 64 <= N) {                                       // whole chunk inside the frame (warp-uniform)
#pragma unroll
          for (int j = 0; j < M; ++j) {
            if constexpr (SC16) cp_async8(dst + j * LPC, src + j * LPC * 2, 8);
            else cp_async16(dst + j * LPC, src + j * LPC * 2, 16);
          }
        } else {
#pragma unroll
          for (int j = 0; j < M; ++j) {
            const int nbytes = (ic * 64 + (j * LPC + lane % LPC) * 2 < N) ? (int)sizeof(Slot) : 0;   // beyond the frame: zero fill
            if constexpr (SC16) cp_async8(dst + j * LPC, src + (nbytes ? j * LPC * 2 : 0), nbytes);
            else cp_async16(dst + j * LPC, src + (nbytes ? j * LPC * 2 : 0), nbytes);
          }
        }
      } else {
        const int t = ic * 64 + lane * 2;
        const int nbytes = (t < N) ? (int)sizeof(Slot) : 0;            // beyond the frame: zero fill (sc16 zero = 0.0f)
        Slot* dst = myring + (size_t)istage * M * 32 + lane;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          if constexpr (SC16) cp_async8(dst + k * 32, ibase + (long long)k * chan_stride + (nbytes ? t : 0), nbytes);
          else cp_async16(dst + k * 32, ibase + (long long)k * chan_stride + (nbytes ? t : 0), nbytes);
        }
      }
      if (++ic == NCH) { ic = 0; ibase += (long long)WS_P * frame_stride; iframe += WS_P; }
      if (++istage == WS_STAGES) istage = 0;
      ++issued;
    };
#pragma unroll
    for (int q = 0; q < WS_STAGES - 1; ++q) { if (issued < total) issue(); if constexpr (!TMA) cp_async_commit(); }
    CovAcc<M> acc;
    acc.clear();
    int cur_tile = 0; bool opened = false;
    int c = 0, m = 0, rstage = 0;
    unsigned rphase = 0;
    for (int q = 0; q < total; ++q) {
      // (the stage refilled here was read in the previous iteration; its loads have returned: their FFMAs were issued)
      if (issued < total) issue();
      if constexpr (TMA) {
        mbar_wait(&ring_bar[w * WS_STAGES + rstage], rphase);
      } else {
        cp_async_commit();
        cp_async_wait<WS_STAGES - 1>();
      }
      const Slot* src = myring + (size_t)rstage * M * 32 + lane;
      if (++rstage == WS_STAGES) { rstage = 0; rphase ^= 1u; }
      float2 x0[M], x1[M];
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const Slot v = src[k * 32];
        if constexpr (SC16) { x0[k] = sc16_to_c64(v.x); x1[k] = sc16_to_c64(v.y); }
        else { x0[k] = make_float2(v.x, v.y); x1[k] = make_float2(v.z, v.w); }
      }
      acc.add(x0);
      acc.add(x1);
      if (++c == NCH) {
        // frame complete: fold, then emit into the tile buffer (waiting for it only now)
        c = 0;
        const int g = w + m * WS_P; ++m;
        const int tf = g / TILE, slot = g - tf * TILE;
        acc.fold((unsigned)lane, red + w * MM);
        while (cur_tile < tf) {                                       // close tiles this warp is done with
          if (!opened && cur_tile >= WS_NBUF) bar_sync(BAR_EMPTY + (cur_tile % WS_NBUF), NTHREADS);
          __threadfence_block();
          bar_arrive(BAR_FULL + (cur_tile % WS_NBUF), NTHREADS);
          ++cur_tile; opened = false;
        }
        if (!opened) { if (cur_tile >= WS_NBUF) bar_sync(BAR_EMPTY + (cur_tile % WS_NBUF), NTHREADS); opened = true; }
        cov_warp_emit<M>(red + w * MM, scale, bscale, avg_method, (unsigned)lane, Rbuf + ((size_t)(tf % WS_NBUF) * TILE + slot) * MM, gains);
        acc.clear();
      }
    }
    while (cur_tile < ntiles) {
      if (!opened && cur_tile >= WS_NBUF) bar_sync(BAR_EMPTY + (cur_tile % WS_NBUF), NTHREADS);
      __threadfence_block();
      bar_arrive(BAR_FULL + (cur_tile % WS_NBUF), NTHREADS);
      ++cur_tile; opened = false;
    }
  } else {
    // ================================ consumer ================================
    if constexpr (REALLOC) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(REG_C));
    const int ct = threadIdx.x - WS_P * 32, cw = warp - WS_P;
    for (int t = 0; t < ntiles; ++t) {
      const int b = t % WS_NBUF;
      const int nt = min(TILE, nf - t * TILE);
      bar_sync(BAR_FULL + b, NTHREADS);
      {
        const int g = ct / M, j = ct % M;
        if (G_out != nullptr || u_out != nullptr) {
          // split form: the noise projector and / or its diagonal sums go to global memory; the scan (scan_tc.cu, dev builds) or
          // the Root-MUSIC root finder (root.cu: doa_cuda_rootchain_*) runs as its own kernel
          const long long f = lo + (long long)t * TILE + min(g, nt - 1);
          noise_subspace_solve<M>(Rbuf + ((size_t)b * TILE + g) * MM, j, T, max_sweeps, g < nt, G_out ? G_out + f * MM : nullptr,
                                  u_out ? u_out + f * M : nullptr, nullptr);
        } else {
          noise_subspace_solve<M>(Rbuf + ((size_t)b * TILE + g) * MM, j, T, max_sweeps, g < nt, Gs + g * MM, us + g * M, nullptr);
        }
      }
      // A consumer warp owns its 32/M matrices end to end (Jacobi -> G/u -> scan), so nothing but the tile buffer is shared:
      // no consumer-wide barrier, a warp whose matrices converge early starts scanning early.
      __syncwarp();
      if (t + WS_NBUF < ntiles) { __threadfence_block(); bar_arrive(BAR_EMPTY + b, NTHREADS); }
      if (G_out != nullptr || u_out != nullptr) continue;
      for (int i = cw * (32 / M); i < min(nt, (cw + 1) * (32 / M)); ++i) {
        const long long f = lo + (long long)t * TILE + i;
        if (K == 1)   // index_max: the global arg-max (find_local_max_impl.h:53-56), not a local-peak search
          scan_frame_argmax<M>(us + i * M, Gs + i * MM, zplain, nullptr, Vtab, xaxis, M, P, lane, out_val + f, out_loc + f,
                               out_bin ? out_bin + f : nullptr);
        else
          scan_frame_peaks<M, KL, true>(us + i * M, Gs + i * MM, zt, nullptr, Vtab, xaxis, M, P, K, lane, out_val + f * K,
                                        out_loc + f * K, out_bin ? out_bin + f * K : nullptr);
      }
      __syncwarp();
    }
  }
}

template <int M, int WS_P, int WS_C, int WS_STAGES, int WS_NBUF, int FILL, typename S>
int launch_ws_cfg2(const S* in, long long fs, long long cs, int N, int nframes, int avg, int T, const ScanTables& tb,
                  int K, float* out_val, float* out_loc, int* out_bin, cudaStream_t st, const float2* gains, float in_scale2,
                  float2* G_out = nullptr, float2* u_out = nullptr) {
  constexpr int TILE = WS_C * 32 / M;
  const size_t smem = ((ztab_floats(tb.P) + 3) & ~(size_t)3) * sizeof(float) + ((size_t)(WS_NBUF + 1) * TILE * M * M + (size_t)TILE * M) * sizeof(float2) +
                      (size_t)WS_P * M * M * sizeof(float) + (size_t)WS_P * WS_STAGES * M * 32 * sizeof(typename RingSlot<S>::type) + (FILL == 3 ? 128 : 0);
  if (smem > 225 * 1024) return 0;
  auto kern = chain_ws_kernel<M, 4, WS_P, WS_C, WS_STAGES, WS_NBUF, FILL, S>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // Multi-GPU runs leave a few SMs to the peak gather's NCCL kernel so that it overlaps the next step's chain kernel
  // (a persistent CTA owns a whole SM's registers and shared memory: nothing else can co-reside).
  sms = std::max(1, sms - std::max(0, dev_option(OPT_SMS_RESERVE, 0)));
  // small batches (a GNU Radio work() call): spread over the SMs down to one consumer warp's worth of frames per CTA
  constexpr int GRP = 32 / M;
  const int grid = std::max(1, std::min(sms, (nframes + GRP - 1) / GRP));
  const float scale = (float)(1.0 / N) * in_scale2, bscale = (float)(0.5 / N);   // in_scale2: sc16 converter scale squared (cov.cu)
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  if constexpr (FILL == 3) {
    // [B][M][N] complex64 as a rank-3 tensor of 8-byte elements (sample, channel, frame); box = 64 samples x M channels x 1 frame
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = nullptr;
    if (encode == nullptr) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qr;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || fn == nullptr) return 0;
      encode = (encode_t)fn;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)nframes};
    const cuuint64_t strides[2] = {(cuuint64_t)cs * 8u, (cuuint64_t)fs * 8u};
    const cuuint32_t box[3] = {64u, (cuuint32_t)M, 1u}, estr[3] = {1u, 1u, 1u};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<S*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 0;
  }
  kern<<<grid, (WS_P + WS_C) * 32, smem, st>>>(in, fs, cs, N, nframes, avg, scale, bscale, T, eig_sweeps_arg(M), tb.zpair, tb.z, tb.V, tb.xaxis, tb.P,
                                               K, out_val, out_loc, out_bin, gains, G_out, u_out, tmap);
  return 1;
}

template <int M, int WS_P, int WS_C, int WS_STAGES, int WS_NBUF = 3>
int launch_ws_cfg(const float2* in, long long fs, long long cs, int N, int nframes, int avg, int T, const ScanTables& tb,
                  int K, float* out_val, float* out_loc, int* out_bin, cudaStream_t st, const float2* gains,
                  float2* G_out = nullptr, float2* u_out = nullptr) {
  const float in_scale2 = 1.0f;
  // Bulk (TMA) ring fills, measured at cfg3: 1.99 ms against 1.67 ms with per-lane cp.async -- a 512-byte copy per channel and
  // chunk is too small for the bulk-copy engine (16.8 M copies per launch) and larger ones do not fit per-warp rings.  Kept
  // selectable (dev knob ws_tma) for the default configuration only.
  // Tensor-map TMA ring fills (FILL 3, SASS UTMALDG.3D): one cp.async.bulk.tensor.3d box [1 frame][M channels][64 samples] per stage
  // and producer warp, issued by one lane -- no per-lane addresses and no LSU instructions in the streaming loop, tail samples
  // zero-filled by the copy engine.  Needs the dense [B][M][N] layout (a rank-3 tensor map); measured at cfg3 against per-lane
  // cp.async, interleaved (tools/ws_exp.py 80824 80824m, profiles/r02_ws_tma_map*.log): 1.648-1.657 against 1.672-1.677 ms,
  // same bits.  (The round-1 bulk variant -- M separate 512-byte UBLKCP copies per stage -- measured 2.09 ms: eight times the
  // copies.)  Shipped for the 8-element configuration on batches of at least 1024 frames (a tensor map is encoded per call); at 4
  // elements (2 KB boxes, six stages) it measured SLOWER (2.76 against 2.46 ms per 262,144 frames) and is not used; streaming
  // layouts (overlapping frames) and sc16 samples keep cp.async.  Option "tma" = 0 switches it off.
  if constexpr (M == 8 && WS_P == 8 && WS_C == 8 && WS_STAGES == 2 && WS_NBUF == 4) {
    if (dev_option(OPT_TMA, 1) && dev_option(OPT_WS_TMA, 0) == 0 && dev_option(OPT_WS_FILL, 0) == 0 && nframes >= 1024 && cs == N &&
        fs == (long long)M * N && (N % 2) == 0) {
      const int r3 = launch_ws_cfg2<M, WS_P, WS_C, WS_STAGES, WS_NBUF, 3>(in, fs, cs, N, nframes, avg, T, tb, K, out_val, out_loc, out_bin, st, gains, in_scale2, G_out, u_out);
      if (r3) return r3;
    }
  }
#ifdef DOA_DEV_KNOBS
  // the same fills for a handful of other configurations, for the sweep (tools/ws_exp.py, suffix 'm')
  if constexpr (WS_P + WS_C <= 16 && (WS_STAGES == 2 || WS_STAGES == 3) && WS_NBUF >= 2 && WS_NBUF <= 4) {
    if (dev_option(OPT_WS_TMA, 0) == 2 && cs == N && fs == (long long)M * N && (cs % 2) == 0) {
      const int r3 = launch_ws_cfg2<M, WS_P, WS_C, WS_STAGES, WS_NBUF, 3>(in, fs, cs, N, nframes, avg, T, tb, K, out_val, out_loc, out_bin, st, gains, in_scale2, G_out, u_out);
      if (r3) return r3;
    }
  }
  if constexpr (WS_P == 8 && WS_C == 8 && WS_STAGES == 2 && WS_NBUF == 4) {
    if (N % 64 == 0 && dev_option(OPT_WS_TMA, 0) == 1)
      return launch_ws_cfg2<M, WS_P, WS_C, WS_STAGES, WS_NBUF, 1>(in, fs, cs, N, nframes, avg, T, tb, K, out_val, out_loc, out_bin, st, gains, in_scale2);
  }
  // channel-major fills (FILL 2): instantiated for the shipped configurations only
  if constexpr ((M == 8 && WS_P == 8 && WS_C == 8 && WS_STAGES == 2 && WS_NBUF == 4) || (M == 4 && WS_P == 8 && WS_C == 8 && WS_STAGES == 6 && WS_NBUF == 4)) {
    if (dev_option(OPT_WS_FILL, 0) == 2)
      return launch_ws_cfg2<M, WS_P, WS_C, WS_STAGES, WS_NBUF, 2>(in, fs, cs, N, nframes, avg, T, tb, K, out_val, out_loc, out_bin, st, gains, in_scale2);
  }
#endif
  return launch_ws_cfg2<M, WS_P, WS_C, WS_STAGES, WS_NBUF, 0>(in, fs, cs, N, nframes, avg, T, tb, K, out_val, out_loc, out_bin, st, gains, in_scale2, G_out, u_out);
}

}  // namespace

// Returns 1 if the fused kernel was launched, 0 if this shape is not covered (caller falls back to the three kernels).
// Covered: M = 8 or 4, K <= 4 (K = 1: global arg-max), 16-byte aligned even strides, scan table + tiles + rings within one
// SM's shared memory (P <= ~6000).  M = 4 (cfg1 / cfg2 shapes): 3.15 / 3.33 ms unfused -> 2.47 / 2.40 ms, ~7 TB/s of input.
int launch_chain_fused(const void* in_v, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                       int avg_method, int T, const ScanTables& tb, int K, float* out_val, float* out_loc, int* out_bin,
                       cudaStream_t st, const float2* gains, InputFormat fmt, float2* G_out, float2* u_out) {
  if (nframes <= 0 || (M != 8 && M != 4)) return 0;
  if (K < 1 || K > 4) return 0;                       // K > 4: the wide candidate lists
  const bool split = G_out != nullptr || u_out != nullptr;
  if (!split && tb.zpair == nullptr) return 0;        // a scan needs its table
  const bool vec2 = (N % 2 == 0) && (frame_stride % 2 == 0) && (chan_stride % 2 == 0) &&
                    ((reinterpret_cast<uintptr_t>(in_v) & (fmt.sc16 ? 7u : 15u)) == 0);
  if (!vec2) return 0;
  if (fmt.sc16) {
    // sc16 samples: same warp split as fc32; the ring slots are half as wide, which buys a third stage (M = 8) in the same
    // shared memory.  Not tuned separately yet.
    const unsigned* in = static_cast<const unsigned*>(in_v);
    const float s2 = fmt.scale * fmt.scale;
#ifdef DOA_DEV_KNOBS
    if (dev_option(OPT_WS_FILL, 0) == 2) {
      if (M == 4) return launch_ws_cfg2<4, 8, 8, 6, 4, 2>(in, frame_stride, chan_stride, N, nframes, avg_method, T, tb, K, out_val, out_loc, out_bin, st, gains, s2, G_out, u_out);
      return launch_ws_cfg2<8, 8, 8, 3, 4, 2>(in, frame_stride, chan_stride, N, nframes, avg_method, T, tb, K, out_val, out_loc, out_bin, st, gains, s2, G_out, u_out);
    }
#endif
    if (M == 4) return launch_ws_cfg2<4, 8, 8, 6, 4, 0>(in, frame_stride, chan_stride, N, nframes, avg_method, T, tb, K, out_val, out_loc, out_bin, st, gains, s2, G_out, u_out);
    return launch_ws_cfg2<8, 8, 8, 3, 4, 0>(in, frame_stride, chan_stride, N, nframes, avg_method, T, tb, K, out_val, out_loc, out_bin, st, gains, s2, G_out, u_out);
  }
  const float2* in = static_cast<const float2*>(in_v);
  // Producer/consumer split, ring depth and tile buffers, measured at cfg3 with the packed (FFMA2) covariance: 8+8 warps,
  // 2 stages x 4 tile buffers or 3 x 3: 1.59-1.64 ms (equal within box-to-box noise; 2 x 4 needs 201 KB of shared memory);
  // 4+12: 1.88-1.99, 5+11: 1.91, 6+10: 1.85, 7+9: 1.84, 9+7: 1.94, 10+6: 2.03, 12+4: 2.39; 4+16 (setmaxnreg re-allocation,
  // 96-register launch): 1.97.  Warp-stall sampling (profiles/) shows why: with 4 producers the consumers idle at the FULL
  // barrier 41 % of the time while each producer warp, alone on its scheduler, issues at 0.25 IPC.  At 8+8 the tile barriers
  // still hold 12 % of the warp samples, but that is slack, not lost throughput: a variant with pairwise hand-off (producer w
  // feeds consumer w through a private ring of 4-frame slots and mbarriers, no CTA-wide barrier) measured the same 1.67 ms.
#define WS_ARGS in, frame_stride, chan_stride, N, nframes, avg_method, T, tb, K, out_val, out_loc, out_bin, st, gains, G_out, u_out
  if (M == 4) {   // covariance-dominated (80 % of the step): the consumers only have to hide 0.1 + 0.7 ms under 2.6 ms of streaming
#ifndef DOA_DEV_KNOBS
    return launch_ws_cfg<4, 8, 8, 6, 4>(WS_ARGS);
#else
    switch (dev_option(OPT_WS4, 5)) {
      case 0: return 0;
      case 1: return launch_ws_cfg<4, 8, 8, 2, 4>(WS_ARGS);
      case 2: return launch_ws_cfg<4, 12, 4, 2, 4>(WS_ARGS);
      case 3: return launch_ws_cfg<4, 10, 6, 3, 4>(WS_ARGS);
      case 4: return launch_ws_cfg<4, 8, 8, 4, 4>(WS_ARGS);
      case 5: return launch_ws_cfg<4, 8, 8, 6, 4>(WS_ARGS);
      case 6: return launch_ws_cfg<4, 10, 6, 4, 4>(WS_ARGS);
      case 7: return launch_ws_cfg<4, 12, 4, 4, 4>(WS_ARGS);
      default: return launch_ws_cfg<4, 6, 10, 4, 4>(WS_ARGS);
    }
#endif
  }
#ifndef DOA_DEV_KNOBS
  return launch_ws_cfg<8, 8, 8, 2, 4>(WS_ARGS);
#else
  // -DDOA_DEV_KNOBS (tools/ws_exp.py): ws_split = producers * 100 + consumers, ws_stages = cp.async ring depth, ws_nbuf = tile buffers
  if (G_out != nullptr) {   // split form (covariance + Jacobi here, scan as its own kernel): the consumers have less to do
    switch (dev_option(OPT_WS_SPLIT, 808)) {
      case 1204: return launch_ws_cfg<8, 12, 4, 2, 4>(WS_ARGS);
      case 1006: return launch_ws_cfg<8, 10, 6, 2, 4>(WS_ARGS);
      case 1204 + 1: return launch_ws_cfg<8, 12, 4, 3, 4>(WS_ARGS);
      default: return launch_ws_cfg<8, 8, 8, 2, 4>(WS_ARGS);
    }
  }
  switch ((dev_option(OPT_WS_SPLIT, 808) * 10 + dev_option(OPT_WS_STAGES, 2)) * 10 + dev_option(OPT_WS_NBUF, 4)) {
    case 41232: return launch_ws_cfg<8, 4, 12, 3, 2>(WS_ARGS);      // the first fused configuration (1.91 ms)
    case 41252: return launch_ws_cfg<8, 4, 12, 5, 2>(WS_ARGS);
    case 41642: return launch_ws_cfg<8, 4, 16, 4, 2>(WS_ARGS);      // setmaxnreg re-allocated
    case 61032: return launch_ws_cfg<8, 6, 10, 3, 2>(WS_ARGS);
    case 81232: return launch_ws_cfg<8, 8, 12, 3, 2>(WS_ARGS);      // setmaxnreg re-allocated
    case 80832: return launch_ws_cfg<8, 8, 8, 3, 2>(WS_ARGS);
    case 80833: return launch_ws_cfg<8, 8, 8, 3, 3>(WS_ARGS);
    case 80823: return launch_ws_cfg<8, 8, 8, 2, 3>(WS_ARGS);
    case 80825: return launch_ws_cfg<8, 8, 8, 2, 5>(WS_ARGS);
    case 80826: return launch_ws_cfg<8, 8, 8, 2, 6>(WS_ARGS);
    case 70924: return launch_ws_cfg<8, 7, 9, 2, 4>(WS_ARGS);
    case 61024: return launch_ws_cfg<8, 6, 10, 2, 4>(WS_ARGS);
    case 61034: return launch_ws_cfg<8, 6, 10, 3, 4>(WS_ARGS);
    case 90724: return launch_ws_cfg<8, 9, 7, 2, 4>(WS_ARGS);
    case 100624: return launch_ws_cfg<8, 10, 6, 2, 4>(WS_ARGS);
    case 90734: return launch_ws_cfg<8, 9, 7, 3, 4>(WS_ARGS);
    default: return launch_ws_cfg<8, 8, 8, 2, 4>(WS_ARGS);
  }
#endif
#undef WS_ARGS
}

}  // namespace doa
