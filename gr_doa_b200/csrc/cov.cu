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
#include "doa_internal.h"

namespace doa {
namespace {

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_stream2(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// Reduce-scatter over the warp: on entry every lane holds CNT partial sums a[0..CNT); on exit lane L holds the
// full sums of max(1, CNT/32) consecutive elements starting at rs_base<CNT>(L).
template <int CNT, int OFF>
__device__ __forceinline__ void warp_reduce_scatter(float* a, unsigned lane) {
  if constexpr (OFF >= 1) {
    if constexpr (CNT > 1) {
      constexpr int H = CNT / 2;
      const bool up = (lane & OFF) != 0;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float send = up ? a[i] : a[i + H];
        const float keep = up ? a[i + H] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
      }
      warp_reduce_scatter<H, OFF / 2>(a, lane);
    } else {
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], OFF);
      warp_reduce_scatter<1, OFF / 2>(a, lane);
    }
  }
}
template <int CNT>
__device__ __forceinline__ int rs_base(unsigned lane) {
  if constexpr (CNT >= 32) return (int)lane * (CNT / 32);
  else if constexpr (CNT == 16) return (int)(lane >> 1);
  else if constexpr (CNT == 8) return (int)(lane >> 2);
  else if constexpr (CNT == 4) return (int)(lane >> 3);
  else if constexpr (CNT == 2) return (int)(lane >> 4);
  else return 0;
}

// Scale + optional forward-backward average + Hermitian expansion of the folded sums in `red`
// (layout: offdiag pair p=(r>c): red[2p], red[2p+1] with p = r(r-1)/2 + c; diagonals at red[M*(M-1) + r]).
template <int M>
__device__ __forceinline__ float2 folded_entry(const float* red, int r, int c, float scale) {
  constexpr int NP = M * (M - 1) / 2;
  if (r == c) return make_float2(red[2 * NP + r] * scale, 0.0f);
  if (r > c) { const int p = r * (r - 1) / 2 + c; return make_float2(red[2 * p] * scale, red[2 * p + 1] * scale); }
  const int p = c * (c - 1) / 2 + r;
  return make_float2(red[2 * p] * scale, -(red[2 * p + 1] * scale));
}

constexpr int COV_WARPS = 8;

template <int M, int VEC, int G>
__global__ void __launch_bounds__(COV_WARPS * 32, (G == 1 && M >= 8) ? 2 : 1)
cov_small_kernel(const float2* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
                 float2* __restrict__ out, float scale, float bscale, int avg_method) {
  constexpr int NP = M * (M - 1) / 2;
  constexpr int CNT = M * M;
  __shared__ float red_s[COV_WARPS][CNT];
  const unsigned lane = threadIdx.x & 31u;
  const int warp = threadIdx.x >> 5;
  float* red = red_s[warp];

  for (int f = blockIdx.x * COV_WARPS + warp; f < nframes; f += gridDim.x * COV_WARPS) {
    const float2* base = in + (long long)f * frame_stride;
    float dg[M];
    float ore[NP > 0 ? NP : 1], oim[NP > 0 ? NP : 1];
#pragma unroll
    for (int r = 0; r < M; ++r) dg[r] = 0.0f;
#pragma unroll
    for (int p = 0; p < NP; ++p) { ore[p] = 0.0f; oim[p] = 0.0f; }

    // G independent load groups per iteration (G*M LDG.128 in flight per lane).
    for (int t0 = (int)lane * VEC; t0 < N; t0 += G * 32 * VEC) {
      float2 x[G][VEC][M];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int t = t0 + g * 32 * VEC;
        const bool ok = t < N;   // N % VEC == 0 is guaranteed by the launcher
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const float2* p = base + (long long)k * chan_stride + t;
          if constexpr (VEC == 2) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) v = ldg_stream4(reinterpret_cast<const float4*>(p));
            x[g][0][k] = make_float2(v.x, v.y);
            x[g][1][k] = make_float2(v.z, v.w);
          } else {
            float2 v = make_float2(0.f, 0.f);
            if (ok) v = ldg_stream2(p);
            x[g][0][k] = v;
          }
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int s = 0; s < VEC; ++s) {
#pragma unroll
          for (int r = 0; r < M; ++r) {
            const float2 xr = x[g][s][r];
            dg[r] = fmaf(xr.x, xr.x, dg[r]);
            dg[r] = fmaf(xr.y, xr.y, dg[r]);
#pragma unroll
            for (int c = 0; c < r; ++c) {
              const float2 xc = x[g][s][c];
              const int p = r * (r - 1) / 2 + c;
              ore[p] = fmaf(xr.x, xc.x, ore[p]);   // x_r conj(x_c)
              ore[p] = fmaf(xr.y, xc.y, ore[p]);
              oim[p] = fmaf(xr.y, xc.x, oim[p]);
              oim[p] = fmaf(-xr.x, xc.y, oim[p]);
            }
          }
        }
    }

    float a[CNT];
#pragma unroll
    for (int p = 0; p < NP; ++p) { a[2 * p] = ore[p]; a[2 * p + 1] = oim[p]; }
#pragma unroll
    for (int r = 0; r < M; ++r) a[2 * NP + r] = dg[r];
    warp_reduce_scatter<CNT, 16>(a, lane);
    {
      const int b = rs_base<CNT>(lane);
      constexpr int F = CNT >= 32 ? CNT / 32 : 1;
#pragma unroll
      for (int i = 0; i < F; ++i) red[b + i] = a[i];
    }
    __syncwarp();
    float2* o = out + (long long)f * CNT;
    for (int e = (int)lane; e < CNT; e += 32) {
      const int r = e % M, c = e / M;
      float2 v = folded_entry<M>(red, r, c, scale);
      if (avg_method == 1) {
        // 0.5*R + (0.5/N) * J conj(R) J : (J conj(R) J)(r,c) = conj(R(M-1-r, M-1-c))   lib/autocorrelate_impl.cc:108
        const float2 w = folded_entry<M>(red, M - 1 - r, M - 1 - c, scale);
        v.x = __fadd_rn(__fmul_rn(0.5f, v.x), __fmul_rn(bscale, w.x));
        v.y = __fadd_rn(__fmul_rn(0.5f, v.y), __fmul_rn(bscale, -w.y));
      }
      o[e] = v;
    }
    __syncwarp();
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
  const int Mp = ((M + 3) / 4) * 4;
  const size_t smem = (size_t)Mp * (CT_TT + 1) * sizeof(float2) + (size_t)Mp * Mp * 2 * sizeof(float);
  cudaFuncSetAttribute(cov_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int blocks = min(nframes, num_sms() * 4);
  cov_tiled_kernel<<<blocks, CT_THREADS, smem, st>>>(in, frame_stride, chan_stride, M, N, nframes, out, scale, bscale,
                                                     avg_method);
  return 1;
}

}  // namespace doa
