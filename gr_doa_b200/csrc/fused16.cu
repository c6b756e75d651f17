#ifdef DOA_DEV_KNOBS
// fused16.cu -- 16-element arrays: covariance + eigendecomposition in ONE persistent, warp-specialised kernel.
// AN EXPERIMENT WITH A NEGATIVE RESULT, compiled only with -DDOA_DEV_KNOBS (option "fused16"): measured on B200 at the cfg5 shape
// (65,536 frames): 4.9 ms (2 producer pairs + 4 consumer warps), 5.8 ms (1 + 6), 7.6 ms (1 + 8) against 1.7 + 2.3 ms for
// cov16_ring_kernel followed by jacobi_group_kernel<16> -- both sides need registers (255 / 125 per thread) more than they need
// each other's idle issue slots, and eight warps per SM hide neither the Jacobi's shuffle chains nor the ring waits.
//
// BASELINE configs[4] (1,048,576 frames of 16 channels x 1024 snapshots) is bound by the FP32 pipe, not by HBM: the Hermitian
// covariance costs 8.5 flop per input byte (4.1 k SM-cycles of FFMA2 per frame against 5.8 k cycles of HBM time) and the
// 16 x 16 Jacobi another 4.3 k.  Run back to back, cov16_ring_kernel (FFMA2-bound, but idle at every ring wait and pair
// barrier) and jacobi_group_kernel<16> (shuffle -> FMA chains, half of the issue slots empty) each leave the pipe half
// unused; here they share every SM:
//   warps 0..3  producers: two pairs; a pair streams one frame at a time through its cp.async ring and accumulates the
//               packed Hermitian half (cov_device.cuh: cov16_ring_role, the stage kernel's own code), folds it and emits R into
//               one of NBUF tile buffers in shared memory;
//   warps 4..7  consumers: each owns two matrices of a tile (16 lanes per matrix): Jacobi (eig_device.cuh:
//               jacobi_group_solve<16>), then G = U_N U_N^H and the diagonal sums u straight to global memory.
// Hand-off with named barriers FULL b / EMPTY b per tile buffer, as in fused.cu.  R never leaves the SM.  The scan + peak
// picking then runs on the tensor cores (scan_tc.cu) from G and u: two launches for the whole chain.
// Device code and per-entry operation order are the stage kernels': same bits as the three-kernel path (tested).
#include "cov_device.cuh"
#include "eig_os_device.cuh"

#include <algorithm>

namespace doa {
namespace {

constexpr int F16_NBUF = 3;                           // tile buffers between producers and consumers
constexpr int F16_BAR_PAIR = 1;                       // named barriers: pair p = 1 + p, FULL b = 3 + b, EMPTY b = 3 + NBUF + b
constexpr int F16_BAR_FULL = 3, F16_BAR_EMPTY = F16_BAR_FULL + F16_NBUF;
static_assert(F16_BAR_EMPTY + F16_NBUF <= 16, "named barriers");

__device__ __forceinline__ void bar_sync16(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive16(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }

// NPAIR producer pairs (2 warps each), NCONS consumer warps (two matrices each)
template <int NPAIR, int NCONS, typename S>
__global__ void __launch_bounds__((2 * NPAIR + NCONS) * 32, 1)
chain16_kernel(const S* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes, int avg_method,
               float scale, float bscale, int T, int max_sweeps, float2* __restrict__ G_out, float2* __restrict__ u_out,
               const float2* __restrict__ gains) {
  typedef typename Ring16Slot<S>::type Slot;
  constexpr int MM = 256;
  constexpr int F16_P = 2 * NPAIR, F16_C = NCONS, F16_THREADS = (F16_P + F16_C) * 32, F16_TILE = F16_C * 2;
  extern __shared__ float4 smem4[];
  float2* Rbuf = reinterpret_cast<float2*>(smem4);                                   // [NBUF][TILE][256]
  float* red = reinterpret_cast<float*>(Rbuf + F16_NBUF * F16_TILE * MM);             // [NPAIR][256]
  Slot* ring = reinterpret_cast<Slot*>(red + NPAIR * 256);                            // [NPAIR][C16_STAGES][16][32]
  for (int i = threadIdx.x; i < F16_NBUF * F16_TILE * MM; i += blockDim.x) Rbuf[i] = make_float2(0.f, 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  const long long per = nframes / gridDim.x, rem = nframes % gridDim.x;
  const long long lo = blockIdx.x * per + min((long long)blockIdx.x, rem);
  const int nf = (int)(per + (blockIdx.x < rem ? 1 : 0));                             // frames of this CTA
  const int ntiles = (nf + F16_TILE - 1) / F16_TILE;

  if (warp < F16_P) {
    // ================================ producers ================================
    const int pair = warp >> 1, role = warp & 1;
    const int nfw = (pair < nf) ? (nf - pair + NPAIR - 1) / NPAIR : 0;                 // this pair's frames: pair, pair + NPAIR, ...
    int cur_tile = 0; bool opened = false;
    auto emit = [&](int k, const float* rd) {
      const int g = pair + NPAIR * k;                                                 // frame within the CTA's range
      const int tf = g / F16_TILE, slot = g - tf * F16_TILE;
      while (cur_tile < tf) {                                                         // close the tiles this warp is done with
        if (!opened && cur_tile >= F16_NBUF) bar_sync16(F16_BAR_EMPTY + (cur_tile % F16_NBUF), F16_THREADS);
        __threadfence_block();
        bar_arrive16(F16_BAR_FULL + (cur_tile % F16_NBUF), F16_THREADS);
        ++cur_tile; opened = false;
      }
      if (!opened) { if (cur_tile >= F16_NBUF) bar_sync16(F16_BAR_EMPTY + (cur_tile % F16_NBUF), F16_THREADS); opened = true; }
      float2* o = Rbuf + ((size_t)(tf % F16_NBUF) * F16_TILE + slot) * MM;
      if (role == 0) cov16_pair_emit<0>(rd, scale, bscale, avg_method, lane, o, gains);
      else cov16_pair_emit<1>(rd, scale, bscale, avg_method, lane, o, gains);
    };
    Slot* myring = ring + (size_t)pair * C16_STAGES * 16 * 32;
    if (role == 0) cov16_ring_role<0, S>(in, frame_stride, chan_stride, N, lo + pair, NPAIR, nfw, myring, red + pair * 256, F16_BAR_PAIR + pair, lane, emit);
    else cov16_ring_role<1, S>(in, frame_stride, chan_stride, N, lo + pair, NPAIR, nfw, myring, red + pair * 256, F16_BAR_PAIR + pair, lane, emit);
    while (cur_tile < ntiles) {
      if (!opened && cur_tile >= F16_NBUF) bar_sync16(F16_BAR_EMPTY + (cur_tile % F16_NBUF), F16_THREADS);
      __threadfence_block();
      bar_arrive16(F16_BAR_FULL + (cur_tile % F16_NBUF), F16_THREADS);
      ++cur_tile; opened = false;
    }
  } else {
    // ================================ consumers ================================
    const int cw = warp - F16_P;
    const int g = lane >> 4, j = lane & 15;
    for (int t = 0; t < ntiles; ++t) {
      const int b = t % F16_NBUF;
      const int nt = min(F16_TILE, nf - t * F16_TILE);
      bar_sync16(F16_BAR_FULL + b, F16_THREADS);
      const int slot = cw * 2 + g;
      const bool live = slot < nt;
      const long long f = lo + (long long)t * F16_TILE + (live ? slot : 0);
      noise_subspace_solve<16>(Rbuf + ((size_t)b * F16_TILE + slot) * MM, j, T, max_sweeps, live, G_out + f * MM, u_out + f * 16, nullptr);
      __syncwarp();
      if (t + F16_NBUF < ntiles) { __threadfence_block(); bar_arrive16(F16_BAR_EMPTY + b, F16_THREADS); }
    }
  }
}

template <int NPAIR, int NCONS, typename S>
int launch16(const S* in, long long fs, long long cs, int N, int nframes, int avg, int T, float2* G, float2* u, cudaStream_t st,
             const float2* gains, float in_scale2) {
  constexpr int F16_TILE = NCONS * 2, F16_THREADS = (2 * NPAIR + NCONS) * 32;
  const size_t smem = (size_t)F16_NBUF * F16_TILE * 256 * sizeof(float2) + NPAIR * 256 * sizeof(float) +
                      (size_t)NPAIR * C16_STAGES * 16 * 32 * sizeof(typename Ring16Slot<S>::type);
  auto kern = chain16_kernel<NPAIR, NCONS, S>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  sms = std::max(1, sms - std::max(0, dev_option(OPT_SMS_RESERVE, 0)));
  const int grid = std::max(1, std::min(sms, (nframes + NPAIR - 1) / NPAIR));
  const float scale = (float)(1.0 / N) * in_scale2, bscale = (float)(0.5 / N);
  kern<<<grid, F16_THREADS, smem, st>>>(in, fs, cs, N, nframes, avg, scale, bscale, T, eig_sweeps_arg(16), G, u, gains);
  return 1;
}

}  // namespace

// Returns 1 if launched (G [nframes][256] and u [nframes][16] then hold the noise projector and its diagonal sums), 0 if the
// shape is not covered (M != 16, unaligned input): the caller runs the covariance and eigendecomposition kernels instead.
int launch_cov_eig_fused16(const void* in_v, long long frame_stride, long long chan_stride, int M, int N, int nframes, int avg_method,
                           int T, float2* G, float2* u, cudaStream_t st, const float2* gains, InputFormat fmt) {
  if (M != 16 || nframes <= 0) return 0;
  const bool vec2 = (N % 2 == 0) && (frame_stride % 2 == 0) && (chan_stride % 2 == 0) &&
                    ((reinterpret_cast<uintptr_t>(in_v) & (fmt.sc16 ? 7u : 15u)) == 0);
  if (!vec2) return 0;
  if (fmt.sc16)
    return launch16<2, 4>(static_cast<const unsigned*>(in_v), frame_stride, chan_stride, N, nframes, avg_method, T, G, u, st, gains, fmt.scale * fmt.scale);
#ifdef DOA_DEV_KNOBS
  if (dev_option(OPT_WS_SPLIT, 0) == 108) return launch16<1, 8>(static_cast<const float2*>(in_v), frame_stride, chan_stride, N, nframes, avg_method, T, G, u, st, gains, 1.0f);
  if (dev_option(OPT_WS_SPLIT, 0) == 106) return launch16<1, 6>(static_cast<const float2*>(in_v), frame_stride, chan_stride, N, nframes, avg_method, T, G, u, st, gains, 1.0f);
#endif
  return launch16<2, 4>(static_cast<const float2*>(in_v), frame_stride, chan_stride, N, nframes, avg_method, T, G, u, st, gains, 1.0f);
}

}  // namespace doa
#endif  // DOA_DEV_KNOBS
