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
// blocks over the lower block triangle, slices of the tile's time axis per thread, fixed-order shared-memory fold at the end
// (no atomics: the result does not depend on the schedule).
#include "cov_device.cuh"

#include <algorithm>

namespace doa {
namespace {

constexpr int COV_WARPS = 8;

int num_sms() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}


template <int M, int VEC, int G, typename S>
__global__ void __launch_bounds__(COV_WARPS * 32, (G == 1 && M >= 8) ? 2 : 1)
cov_small_kernel(const S* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
                 float2* __restrict__ out, float scale, float bscale, int avg_method, const float2* __restrict__ gains) {
  constexpr int CNT = M * M;
  __shared__ float red_s[COV_WARPS][CNT];
  const unsigned lane = threadIdx.x & 31u;
  const int warp = threadIdx.x >> 5;
  float* red = red_s[warp];
  for (int f = blockIdx.x * COV_WARPS + warp; f < nframes; f += gridDim.x * COV_WARPS) {
    cov_warp_frame<M, VEC, G, S>(in + (long long)f * frame_stride, chan_stride, N, lane, red);
    cov_warp_emit<M>(red, scale, bscale, avg_method, lane, out + (long long)f * CNT, gains);
  }
}

// ---- M = 16 ----------------------------------------------------------------------------------------------------
// The Hermitian lower half of a 16 x 16 matrix is 256 floats: too many accumulators for one lane.  Two warps share a frame:
// role A keeps the two diagonal 8 x 8 blocks (two CovAcc<8> = 128 accumulators), role B the off-diagonal block R[8..15][0..7]
// (64 complex = 128 accumulators).  Both walk the frame's time axis the same way as the one-warp kernels (lanes split the
// time axis, LDG.128 = two samples per channel), fold with the reduce-scatter butterfly and scatter their sums into the
// frame's shared 256-float area in the layout folded_entry<16> reads; the pair then emits half of R each.
constexpr int C16_FRAMES = 4;   // frames in flight per CTA (8 warps)

template <int VEC, typename S>
__device__ __forceinline__ void cov16_load(const S* __restrict__ base, long long chan_stride, int t, bool ok,
                                           float2 (&x)[VEC][16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float2 v[VEC];
    load_samples<VEC, S>(base + (long long)k * chan_stride + t, ok, v);
#pragma unroll
    for (int s = 0; s < VEC; ++s) x[s][k] = v[s];
  }
}

template <int VEC, typename S>
__global__ void __launch_bounds__(C16_FRAMES * 64, 1)
cov16_kernel(const S* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
             float2* __restrict__ out, float scale, float bscale, int avg_method, const float2* __restrict__ gains) {
  constexpr int M = 16, CNT = 256, NP16 = 120;
  __shared__ float red_s[C16_FRAMES][CNT];
  const unsigned lane = threadIdx.x & 31u;
  const int warp = threadIdx.x >> 5, slot = warp >> 1, role = warp & 1;
  float* red = red_s[slot];
  const int nslots = gridDim.x * C16_FRAMES;
  // every warp of a pair runs the same number of iterations (the pair barrier needs both)
  for (int f = blockIdx.x * C16_FRAMES + slot; f < nframes; f += nslots) {
    const S* base = in + (long long)f * frame_stride;
    float a[128];
    if (role == 0) {
      CovAcc<8> d0, d1;
      d0.clear(); d1.clear();
      for (int t = (int)lane * VEC; t < N; t += 32 * VEC) {
        float2 x[VEC][16];
        cov16_load<VEC, S>(base, chan_stride, t, true, x);
#pragma unroll
        for (int s = 0; s < VEC; ++s) {
          float2 lo8[8], hi8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) { lo8[k] = x[s][k]; hi8[k] = x[s][8 + k]; }
          d0.add(lo8);
          d1.add(hi8);
        }
      }
#pragma unroll
      for (int p = 0; p < 28; ++p) { d0.get(p, a[2 * p], a[2 * p + 1]); d1.get(p, a[64 + 2 * p], a[64 + 2 * p + 1]); }
#pragma unroll
      for (int r = 0; r < 8; ++r) { a[56 + r] = d0.diag(r); a[120 + r] = d1.diag(r); }
    } else {
      f32x2 od[64];                                 // (re, im) of R[8 + i][j] at od[i * 8 + j]
#pragma unroll
      for (int i = 0; i < 64; ++i) od[i] = 0ull;
      for (int t = (int)lane * VEC; t < N; t += 32 * VEC) {
        float2 x[VEC][16];
        cov16_load<VEC, S>(base, chan_stride, t, true, x);
#pragma unroll
        for (int s = 0; s < VEC; ++s)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 xr = x[s][8 + i];
            const f32x2 xp = pk2(xr.x, xr.y), xs = pk2(xr.y, -xr.x);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 xc = x[s][j];
              od[i * 8 + j] = fma2(xp, pk2(xc.x, xc.x), od[i * 8 + j]);   // x_r conj(x_c), see CovAcc
              od[i * 8 + j] = fma2(xs, pk2(xc.y, xc.y), od[i * 8 + j]);
            }
          }
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) upk2(od[i], a[2 * i], a[2 * i + 1]);
    }
    warp_reduce_scatter<128, 16>(a, lane);          // lane L now holds the full sums of elements 4L .. 4L+3
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = (int)lane * 4 + i;
      int pos;
      if (role == 0) {
        const int blk = k >> 6, kk = k & 63;
        if (kk < 56) {
          const int r = 8 * blk + kTri8Row[kk >> 1], c = 8 * blk + kTri8Col[kk >> 1];
          pos = 2 * (r * (r - 1) / 2 + c) + (kk & 1);
        } else {
          pos = 2 * NP16 + 8 * blk + (kk - 56);
        }
      } else {
        const int r = 8 + (k >> 4), c = (k >> 1) & 7;
        pos = 2 * (r * (r - 1) / 2 + c) + (k & 1);
      }
      red[pos] = a[i];
    }
    asm volatile("bar.sync %0, 64;" :: "r"(1 + slot) : "memory");   // both roles' sums are in red
    float2* o = out + (long long)f * CNT;
    for (int e = role * 32 + (int)lane; e < CNT; e += 64) {
      const int r = e % M, c = e / M;
      float2 v = apply_gain(folded_entry<M>(red, r, c, scale), gains, r, c);
      if (avg_method == 1) {   // 0.5*R + (0.5/N) * J conj(R) J, lib/autocorrelate_impl.cc:108
        const float2 w = apply_gain(folded_entry<M>(red, M - 1 - r, M - 1 - c, scale), gains, M - 1 - r, M - 1 - c);
        v.x = __fadd_rn(__fmul_rn(0.5f, v.x), __fmul_rn(bscale, w.x));
        v.y = __fadd_rn(__fmul_rn(0.5f, v.y), __fmul_rn(bscale, -w.y));
      }
      o[e] = v;
    }
    asm volatile("bar.sync %0, 64;" :: "r"(1 + slot) : "memory");   // red may be overwritten
  }
}

template <typename S>
__global__ void __launch_bounds__(C16_FRAMES * 64, 1)
cov16_ring_kernel(const S* __restrict__ in, long long frame_stride, long long chan_stride, int N, int nframes,
                  float2* __restrict__ out, float scale, float bscale, int avg_method, const float2* __restrict__ gains) {
  typedef typename Ring16Slot<S>::type Slot;
  extern __shared__ float4 ring_s[];                 // [C16_FRAMES][C16_STAGES][16][32] slots
  __shared__ float red_s[C16_FRAMES][256];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5, slot = warp >> 1;
  Slot* ring = reinterpret_cast<Slot*>(ring_s) + (size_t)slot * C16_STAGES * 16 * 32;
  const int nslots = gridDim.x * C16_FRAMES;
  const int first = blockIdx.x * C16_FRAMES + slot;
  const int nfw = (first < nframes) ? (nframes - first + nslots - 1) / nslots : 0;    // frames of this pair
  if ((warp & 1) == 0) {
    cov16_ring_role<0, S>(in, frame_stride, chan_stride, N, first, nslots, nfw, ring, red_s[slot], 1 + slot, lane, [&](int k, const float* red) {
      cov16_pair_emit<0>(red, scale, bscale, avg_method, lane, out + ((long long)first + (long long)k * nslots) * 256, gains);
    });
  } else {
    cov16_ring_role<1, S>(in, frame_stride, chan_stride, N, first, nslots, nfw, ring, red_s[slot], 1 + slot, lane, [&](int k, const float* red) {
      cov16_pair_emit<1>(red, scale, bscale, avg_method, lane, out + ((long long)first + (long long)k * nslots) * 256, gains);
    });
  }
}

template <typename S>
int launch_cov16(const S* in, long long fs, long long cs, int N, int nframes, float2* out, float scale, float bscale,
                 int avg, cudaStream_t st, const float2* gains) {
  // two samples per ring slot / load: 16 bytes of fc32, 8 bytes of sc16
  const bool vec2 = (N % 2 == 0) && (fs % 2 == 0) && (cs % 2 == 0) && ((reinterpret_cast<uintptr_t>(in) & (2 * sizeof(S) - 1)) == 0);
  const int blocks = (nframes + C16_FRAMES - 1) / C16_FRAMES;
  if (vec2 && dev_option(OPT_COV16_RING, 1)) {
    const size_t smem = (size_t)C16_FRAMES * C16_STAGES * 16 * 32 * sizeof(typename Ring16Slot<S>::type);
    cudaFuncSetAttribute(cov16_ring_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = std::min(blocks, num_sms());
    cov16_ring_kernel<S><<<grid, C16_FRAMES * 64, smem, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg, gains);
    return 1;
  }
  if (vec2) cov16_kernel<2, S><<<blocks, C16_FRAMES * 64, 0, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg, gains);
  else cov16_kernel<1, S><<<blocks, C16_FRAMES * 64, 0, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg, gains);
  return 1;
}

// ---- generic M ---------------------------------------------------------------------------------------------
constexpr int CT_THREADS = 512;
constexpr int CT_TT = 64;   // time samples per shared-memory tile

template <typename S>
__global__ void __launch_bounds__(CT_THREADS)
cov_tiled_kernel(const S* __restrict__ in, long long frame_stride, long long chan_stride, int M, int N,
                 int nframes, float2* __restrict__ out, float scale, float bscale, int avg_method, const float2* __restrict__ gains) {
  extern __shared__ float2 smem[];
  const int nb = (M + 3) / 4;           // 4-row blocks
  const int Mp = nb * 4;                // padded channel count
  const int nbp = nb * (nb + 1) / 2;    // lower block triangle
  const int TS = max(1, min(CT_TT, CT_THREADS / nbp));   // time slices per block pair
  const int LDT = CT_TT + 1;            // padded row stride (float2 units)
  float2* tile = smem;                  // [Mp][LDT]; after the time loop the same memory holds the threads' partial blocks
  float* part = reinterpret_cast<float*>(smem);                       // [32][CT_THREADS]: float q of thread t at q*CT_THREADS + t
  const size_t tile_f2 = max((size_t)Mp * LDT, (size_t)16 * CT_THREADS);
  float* Racc = reinterpret_cast<float*>(smem + tile_f2);            // [Mp*Mp*2] folded sums (re,im), row r col c at (r*Mp+c)*2

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
    const S* base = in + (long long)f * frame_stride;
    float are[4][4], aim[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { are[i][j] = 0.f; aim[i][j] = 0.f; }

    for (int t0 = 0; t0 < N; t0 += CT_TT) {
      __syncthreads();
      for (int i = tid; i < Mp * CT_TT; i += CT_THREADS) {
        const int r = i / CT_TT, t = i % CT_TT;
        float2 v[1];
        load_samples<1, S>(base + (long long)r * chan_stride + t0 + t, r < M && t0 + t < N, v);
        tile[r * LDT + t] = v[0];
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
    __syncthreads();                      // the tile is dead: park every thread's 4 x 4 partial block in its place
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          part[((i * 4 + j) * 2) * CT_THREADS + tid] = are[i][j];
          part[((i * 4 + j) * 2 + 1) * CT_THREADS + tid] = aim[i][j];
        }
    }
    __syncthreads();
    // fold the TS time slices of every block pair in slice order (deterministic, unlike atomics)
    for (int e = tid; e < nbp * 32; e += CT_THREADS) {
      const int p = e >> 5, q = e & 31;
      float acc = 0.f;
      for (int s2 = 0; s2 < TS; ++s2) acc += part[q * CT_THREADS + p * TS + s2];
      int b = (int)((sqrtf(8.0f * p + 1.0f) - 1.0f) * 0.5f);
      while ((b + 1) * (b + 2) / 2 <= p) ++b;
      while (b * (b + 1) / 2 > p) --b;
      const int r = b * 4 + (q >> 3), c = (p - b * (b + 1) / 2) * 4 + ((q >> 1) & 3);
      Racc[(r * Mp + c) * 2 + (q & 1)] = acc;
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
      float2 v = apply_gain(entry(r, c), gains, r, c);
      if (avg_method == 1) {
        const float2 w = apply_gain(entry(M - 1 - r, M - 1 - c), gains, M - 1 - r, M - 1 - c);
        v.x = __fadd_rn(__fmul_rn(0.5f, v.x), __fmul_rn(bscale, w.x));
        v.y = __fadd_rn(__fmul_rn(0.5f, v.y), __fmul_rn(bscale, -w.y));
      }
      o[e] = v;
    }
    __syncthreads();
  }
}

template <int M, typename S>
int launch_small(const S* in, long long fs, long long cs, int N, int nframes, float2* out, float scale,
                 float bscale, int avg, cudaStream_t st, const float2* gains) {
  // two samples per load (LDG.128 for fc32, LDG.64 for sc16) when the layout is aligned for it
  const bool vec2 = (N % 2 == 0) && (fs % 2 == 0) && (cs % 2 == 0) && ((reinterpret_cast<uintptr_t>(in) & (2 * sizeof(S) - 1)) == 0);
  const int blocks = (nframes + COV_WARPS - 1) / COV_WARPS;
  const int variant = dev_option(OPT_COV_GROUPS, 1);   // 1: 128 regs, 2 CTAs/SM (6.4 TB/s at M=8); 2: 167 regs, 1 CTA/SM (5.9 TB/s)
  if (vec2) {
    if (variant == 1) cov_small_kernel<M, 2, 1, S><<<blocks, COV_WARPS * 32, 0, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg, gains);
    else cov_small_kernel<M, 2, 2, S><<<blocks, COV_WARPS * 32, 0, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg, gains);
  } else {
    cov_small_kernel<M, 1, 2, S><<<blocks, COV_WARPS * 32, 0, st>>>(in, fs, cs, N, nframes, out, scale, bscale, avg, gains);
  }
  return 1;
}

template <typename S>
int launch_tiled(const S* in, long long frame_stride, long long chan_stride, int M, int N, int nframes, float2* out,
                 float scale, float bscale, int avg_method, cudaStream_t st, const float2* gains) {
  const int Mp = ((M + 3) / 4) * 4;
  const size_t tile_f2 = std::max((size_t)Mp * (CT_TT + 1), (size_t)16 * CT_THREADS);   // time tile, later the partial blocks
  const size_t smem = tile_f2 * sizeof(float2) + (size_t)Mp * Mp * 2 * sizeof(float);
  cudaFuncSetAttribute(cov_tiled_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int blocks = min(nframes, num_sms() * 4);
  cov_tiled_kernel<S><<<blocks, CT_THREADS, smem, st>>>(in, frame_stride, chan_stride, M, N, nframes, out, scale, bscale,
                                                        avg_method, gains);
  return 1;
}

}  // namespace

int launch_covariance(const void* in_v, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                      int avg_method, float2* out, cudaStream_t st, const float2* gains, InputFormat fmt, void* tc_ws) {
  if (nframes <= 0) return 0;
  if (M > 64) return DOA_CUDA_EINVAL;
  const float bscale = (float)(0.5 / N);    // (0.5/d_snapshot_size), lib/autocorrelate_impl.cc:108
  if (fmt.sc16) {
    // The kernels accumulate the int16 values as exact floats; the converter's scale s enters once, squared, in the
    // emit factor.  For a power-of-two s that is bit-identical to the fc32 path fed float(i16) * s.
    const float scale = (float)(1.0 / N) * (fmt.scale * fmt.scale);
    const unsigned* in = static_cast<const unsigned*>(in_v);
    switch (M) {
      case 2: return launch_small<2>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st, gains);
      case 4: return launch_small<4>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st, gains);
      case 8: return launch_small<8>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st, gains);
      case 16: return launch_cov16(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st, gains);
      default: return launch_tiled(in, frame_stride, chan_stride, M, N, nframes, out, scale, bscale, avg_method, st, gains);
    }
  }
  const float2* in = static_cast<const float2*>(in_v);
  const float scale = (float)(1.0 / N);     // (1.0/d_snapshot_size) narrowed to float, lib/autocorrelate_impl.cc:106
  switch (M) {
    case 2: return launch_small<2>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st, gains);
    case 4: return launch_small<4>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st, gains);
    case 8: return launch_small<8>(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st, gains);
    case 16: return launch_cov16(in, frame_stride, chan_stride, N, nframes, out, scale, bscale, avg_method, st, gains);
    default: break;
  }
  if (M == 64 && dev_option(OPT_HERK_TC, 1)) {   // tensor-core complex HERK (3xTF32) when alignment allows
    const int r = launch_covariance_tc(in, frame_stride, chan_stride, M, N, nframes, avg_method, out, st, gains, tc_ws);
    if (r != 0) return r;
  }
  return launch_tiled(in, frame_stride, chan_stride, M, N, nframes, out, scale, bscale, avg_method, st, gains);
}

}  // namespace doa
