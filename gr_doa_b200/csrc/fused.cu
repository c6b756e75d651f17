// fused.cu -- the whole chain (autocorrelate -> MUSIC_lin_array -> find_local_max) in ONE persistent kernel.
//
// The three stage kernels run back to back leave the machine half idle twice: the covariance is HBM-bound and uses ~40 %
// of the issue slots, the eigendecomposition and the scan are issue-bound and move no data.  Here a CTA owns a contiguous
// range of frames and walks it in tiles of TILE = 256/M frames through three phases, with every intermediate (R, G, u)
// in shared memory:
//     phase 1  each warp streams TILE/8 frames from HBM and folds their covariance        (cov_device.cuh)
//     phase 2  256 threads = TILE matrices x M lanes: Jacobi, noise projector, diagonal sums (eig_device.cuh)
//     phase 3  each warp scans TILE/8 frames, picks, refines and writes K peaks             (scan_device.cuh)
// Two CTAs share an SM (<= 128 registers, ~86 KB shared memory each); the second half of the grid starts with a
// half-size tile, so the two CTAs of an SM run in antiphase and one streams while the other computes.
// The device code is the stage kernels' own, so the fused path is bit-identical to the three-kernel path (tested).
#include "cov_device.cuh"
#include "eig_device.cuh"
#include "scan_device.cuh"

#include <algorithm>

namespace doa {
namespace {

constexpr int FU_WARPS = 8;

template <int M, int KL>
__global__ void __launch_bounds__(FU_WARPS * 32, 2)
chain_fused_kernel(const float2* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
                   int avg_method, float scale, float bscale, int T, int max_sweeps, const float* __restrict__ zpair,
                   const float2* __restrict__ Vtab, const float* __restrict__ xaxis, int P, int K,
                   float* __restrict__ out_val, float* __restrict__ out_loc, int* __restrict__ out_bin, int stagger) {
  constexpr int TILE = FU_WARPS * 32 / M;      // frames per tile = matrices the CTA's threads cover in phase 2
  constexpr int MM = M * M;
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const ZTab zt = ztab_fill(smem, zpair, P);
  float2* Rs = reinterpret_cast<float2*>(smem + ztab_floats(P));   // [TILE][M*M] covariance, then eigenvector staging
  float2* Gs = Rs + TILE * MM;                                     // [TILE][M*M] noise projector
  float2* us = Gs + TILE * MM;                                     // [TILE][M]   diagonal sums
  float* red = reinterpret_cast<float*>(us + TILE * M);            // [FU_WARPS][M*M] covariance fold scratch
  for (int i = threadIdx.x; i < TILE * MM; i += blockDim.x) Rs[i] = make_float2(0.f, 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned ulane = (unsigned)lane;

  // contiguous, balanced frame range of this CTA
  const long long per = nframes / gridDim.x, rem = nframes % gridDim.x;
  const long long lo = blockIdx.x * per + min((long long)blockIdx.x, rem);
  const long long hi = lo + per + (blockIdx.x < rem ? 1 : 0);
  long long f0 = lo;
  int tile = (stagger && blockIdx.x >= gridDim.x / 2) ? TILE / 2 : TILE;   // antiphase start for the SM's second CTA

  while (f0 < hi) {
    const int nt = (int)min((long long)tile, hi - f0);
    // ---- phase 1: covariance of nt frames, warp-strided -------------------------------------------------------------
    for (int i = warp; i < nt; i += FU_WARPS) {
      cov_warp_frame<M, 2, 1>(in + (f0 + i) * frame_stride, chan_stride, N, ulane, red + warp * MM);
      cov_warp_emit<M>(red + warp * MM, scale, bscale, avg_method, ulane, Rs + i * MM);
    }
    __syncthreads();
    // ---- phase 2: Jacobi on TILE matrices at once (groups beyond nt carry stale finite data and store nothing) ----------
    {
      const int g = threadIdx.x / M, j = threadIdx.x % M;
      jacobi_group_solve<M>(Rs + g * MM, j, T, max_sweeps, g < nt, Gs + g * MM, us + g * M, nullptr);
    }
    __syncthreads();
    // ---- phase 3: scan + peaks, warp-strided; results go straight to global memory ---------------------------------------
    for (int i = warp; i < nt; i += FU_WARPS) {
      const long long f = f0 + i;
      scan_frame_peaks<M, KL>(us + i * M, Gs + i * MM, zt, nullptr, Vtab, xaxis, M, P, K, lane, out_val + f * K,
                              out_loc + f * K, out_bin ? out_bin + f * K : nullptr);
    }
    // no barrier needed here: the next tile's phase 1 only writes Rs/red, and its closing barrier orders every warp's
    // phase 3 reads of Gs/us before the next phase 2 overwrites them
    f0 += nt;
    tile = TILE;
  }
}

template <int M>
int launch_fused_m(const float2* in, long long fs, long long cs, int N, int nframes, int avg, int T, const ScanTables& tb,
                   int K, float* out_val, float* out_loc, int* out_bin, cudaStream_t st) {
  constexpr int TILE = FU_WARPS * 32 / M;
  const size_t smem = ztab_floats(tb.P) * sizeof(float) + ((size_t)2 * TILE * M * M + (size_t)TILE * M) * sizeof(float2) +
                      (size_t)FU_WARPS * M * M * sizeof(float);
  if (smem > 110 * 1024) return 0;   // two CTAs per SM must fit; larger scans use the three-kernel path
  auto kern = chain_fused_kernel<M, 4>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::max(1, std::min(2 * sms, (nframes + TILE - 1) / TILE));
  const float scale = (float)(1.0 / N), bscale = (float)(0.5 / N);
  kern<<<grid, FU_WARPS * 32, smem, st>>>(in, fs, cs, N, nframes, avg, scale, bscale, T, 12, tb.zpair, tb.V, tb.xaxis, tb.P, K,
                                          out_val, out_loc, out_bin, dev_option("fused_stagger", 1));
  return 1;
}

}  // namespace

// Returns 1 if the fused kernel was launched, 0 if this shape is not covered (caller falls back to the three kernels).
int launch_chain_fused(const float2* in, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                       int avg_method, int T, const ScanTables& tb, int K, float* out_val, float* out_loc, int* out_bin,
                       cudaStream_t st) {
  if (nframes <= 0) return 0;
  if (K < 2 || K > 4) return 0;                       // K == 1 is the arg-max kernel, K > 4 the wide candidate lists
  const bool vec2 = (N % 2 == 0) && (frame_stride % 2 == 0) && (chan_stride % 2 == 0) &&
                    ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);
  if (!vec2) return 0;
  switch (M) {
    case 4: return launch_fused_m<4>(in, frame_stride, chan_stride, N, nframes, avg_method, T, tb, K, out_val, out_loc, out_bin, st);
    case 8: return launch_fused_m<8>(in, frame_stride, chan_stride, N, nframes, avg_method, T, tb, K, out_val, out_loc, out_bin, st);
    default: return 0;
  }
}

}  // namespace doa
