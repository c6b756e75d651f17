// cov.cu -- stage 1: sample covariance R = X X^H / N per frame (+ the reference's forward-backward term).
//
// Replaces the body of autocorrelate_impl::general_work (gr-doa lib/autocorrelate_impl.cc:92-110): the per-frame
// memcpy framing, the conj(X) temporary and the cgemm collapse into one pass over the samples.
//
// cov_small_kernel<M> (M = 2, 4, 8): one warp per frame.  Every lane owns the whole Hermitian lower half of R as
// register accumulators (M diagonal reals + M(M-1)/2 complex = M*M floats), walks the frame's time axis with
// 128-bit coalesced streaming loads (one LDG.128 = two complex samples per channel, a warp reads 512 contiguous
// bytes per channel per step) and the 32 partial matrices are folded with a reduce-scatter butterfly
// (M*M - 1 shuffles instead of 5*M*M).  Arithmetic intensity is (M+1)/2 complex MACs per 8-byte sample, so for
// M <= 8 the kernel is HBM-bound on B200.
//
// cov_tiled_kernel (any M <= 64): one CTA per frame, time tiles staged in shared memory, 4x4 complex register
// blocks over the lower block triangle, slices of the tile's time axis per thread, shared-memory fold at the end.
#include "cov_device.cuh"

namespace doa {
namespace {

constexpr int COV_WARPS = 8;

template <int M, int VEC, int G>
__global__ void __launch_bounds__(COV_WARPS * 32, (G == 1 && M >= 8) ? 2 : 1)
cov_small_kernel(const float2* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
                 float2* __restrict__ out, float scale, float bscale, int avg_method) {
  constexpr int CNT = M * M;
  __shared__ float red_s[COV_WARPS][CNT];
  const unsigned lane = threadIdx.x & 31u;
  const int warp = threadIdx.x >> 5;
  float* red = red_s[warp];
  for (int f = blockIdx.x * COV_WARPS + warp; f < nframes; f += gridDim.x * COV_WARPS) {
    cov_warp_frame<M, VEC, G>(in + (long long)f * frame_stride, chan_stride, N, lane, red);
    cov_warp_emit<M>(red, scale, bscale, avg_method, lane, out + (long long)f * CNT);
  }
}

// ---- generic M ---------------------------------------------------------------------------------------------
constexpr int CT_THREADS = 512;
constexpr int CT_TT = 64;   // time samples per shared-memory tile

__global__ void __launch_bounds__(CT_THREADS)
cov_tiled_kernel(const float2* __restrict__ in, long long frame_stride, long long chan_stride, int M, int N,
                 int nframes, float2* __restrict__ out, float scale, float bscale, int avg_method) {
  extern __shared__ float2 smem[];
  const int nb = (M + 3) / 4;           // 4-row blocks
  const int Mp = nb * 4;                // padded channel count
  const int nbp = nb * (nb + 1) / 2;    // lower block triangle
  const int TS = max(1, min(CT_TT, CT_THREADS / nbp));   // time slices per block pair
  const int LDT = CT_TT + 1;            // padded row stride (float2 units)
  float2* tile = smem;                  // [Mp][LDT]
  float* Racc = reinterpret_cast<float*>(smem + (size_t)Mp * LDT);   // [Mp*Mp*2] folded sums (re,im), row r col c at (r*Mp+c)*2

  const int tid = threadIdx.x;
  const int item_bp = tid / TS, ts = tid % TS;
  const bool active = item_bp < nbp;
  int bi = 0, bj = 0;
  if (active) {   // invert p = bi(bi+1)/2 + bj, bj <= bi
    int b = (int)((sqrtf(8.0f * item_bp + 1.0f) - 1.0f) * 0.5f);
    while ((b + 1) * (b + 2) / 2 <= item_bp) ++b;
    while (b * (b + 1) / 2 > item_bp) --b;
    bi = b; bj = item_bp - b * (b + 1) / 2;
  }

  for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
    const float2* base = in + (long long)f * frame_stride;
    float are[4][4], aim[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { are[i][j] = 0.f; aim[i][j] = 0.f; }
    for (int i = tid; i < Mp * Mp * 2; i += CT_THREADS) Racc[i] = 0.f;

    for (int t0 = 0; t0 < N; t0 += CT_TT) {
      __syncthreads();
      for (int i = tid; i < Mp * CT_TT; i += CT_THREADS) {
        const int r = i / CT_TT, t = i % CT_TT;
        float2 v = make_float2(0.f, 0.f);
        if (r < M && t0 + t < N) v = ldg_stream2(base + (long long)r * chan_stride + t0 + t);
        tile[r * LDT + t] = v;
      }
      __syncthreads();
      if (active) {
        for (int t = ts; t < CT_TT; t += TS) {
          float2 xr[4], xc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { xr[i] = tile[(bi * 4 + i) * LDT + t]; xc[i] = tile[(bj * 4 + i) * LDT + t]; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              are[i][j] = fmaf(xr[i].x, xc[j].x, are[i][j]);
              are[i][j] = fmaf(xr[i].y, xc[j].y, are[i][j]);
              aim[i][j] = fmaf(xr[i].y, xc[j].x, aim[i][j]);
              aim[i][j] = fmaf(-xr[i].x, xc[j].y, aim[i][j]);
            }
        }
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = bi * 4 + i, c = bj * 4 + j;
          atomicAdd(&Racc[(r * Mp + c) * 2], are[i][j]);
          atomicAdd(&Racc[(r * Mp + c) * 2 + 1], aim[i][j]);
        }
    }
    __syncthreads();
    // Racc holds R(r,c) = sum x_r conj(x_c) for block-lower entries (bi >= bj); mirror the rest.
    auto entry = [&](int r, int c) -> float2 {
      if (r == c) return make_float2(Racc[(r * Mp + c) * 2] * scale, 0.f);
      if ((r >> 2) > (c >> 2) || ((r >> 2) == (c >> 2) && r > c))
        return make_float2(Racc[(r * Mp + c) * 2] * scale, Racc[(r * Mp + c) * 2 + 1] * scale);
      return make_float2(Racc[(c * Mp + r) * 2] * scale, -(Racc[(c * Mp + r) * 2 + 1] * scale));
    };
    float2* o = out + (long long)f * M * M;
    for (int e = tid; e < M * M; e += CT_THREADS) {
      const int r = e % M, c = e / M;
      float2 v = entry(r, c);
      if (avg_method == 1) {
        const float2 w = entry(M - 1 - r, M - 1 - c);
        v.x = __fadd_rn(__fmul_rn(0.5f, v.x), __fmul_rn(bscale, w.x));
        v.y = __fadd_rn(__fmul_rn(0.5f, v.y), __fmul_rn(bscale, -w.y));
      }
      o[e] = v;
    }
    __syncthreads();
  }
}

int num_sms() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

template <int M>
int launch_small(const float2* in, long long fs, long long cs, int N, int nframes, float2* out, float scale,
                 float bscale, int avg, cudaStream_t st) {
  const bool vec2 = (N % 2 == 0) && (fs % 2 == 0) && (cs % 2 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);
  const int blocks = (nframes + COV_WARPS - 1) / COV_WARPS;
  const int variant = dev_option("cov_groups", 1);   // 1: 128 regs, 2 CTAs/SM (6.4 TB/s at M=8); 2: 167 regs, 1 CTA/SM (5.9 TB/s)
  if (vec2) {
    if (variant == 1) cov_small_kernel<M, 2, 1><<<blocks, COV_WARPS * 32, 0, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg);
    else cov_small_kernel<M, 2, 2><<<blocks, COV_WARPS * 32, 0, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg);
  } else {
    cov_small_kernel<M, 1, 2><<<blocks, COV_WARPS * 32, 0, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg);
  }
  return 1;
}

}  // namespace

int launch_covariance(const float2* in, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                      int avg_method, float2* out, cudaStream_t st) {
  if (nframes <= 0) return 0;
  const float scale = (float)(1.0 / N);     // (1.0/d_snapshot_size) narrowed to float, lib/autocorrelate_impl.cc:106
  const float bscale = (float)(0.5 / N);    // (0.5/d_snapshot_size), :108
  switch (M) {
    case 2: return launch_small<2>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st);
    case 4: return launch_small<4>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st);
    case 8: return launch_small<8>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st);
    default: break;
  }
  if (M > 64) return DOA_CUDA_EINVAL;
  if (M == 64 && dev_option("herk_tc", 1)) {   // tensor-core complex HERK (3xTF32) when alignment allows
    const int r = launch_covariance_tc(in, frame_stride, chan_stride, M, N, nframes, avg_method, out, st);
    if (r != 0) return r;
  }
  const int Mp = ((M + 3) / 4) * 4;
  const size_t smem = (size_t)Mp * (CT_TT + 1) * sizeof(float2) + (size_t)Mp * Mp * 2 * sizeof(float);
  cudaFuncSetAttribute(cov_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int blocks = min(nframes, num_sms() * 4);
  cov_tiled_kernel<<<blocks, CT_THREADS, smem, st>>>(in, frame_stride, chan_stride, M, N, nframes, out, scale, bscale,
                                                     avg_method);
  return 1;
}

}  // namespace doa
