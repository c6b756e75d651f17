// cov_device.cuh -- device code of the covariance stage shared by cov.cu and fused.cu.
#pragma once
#include "doa_internal.h"
#include "f32x2.cuh"

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

// ---- sc16 input (UHD wire / cpu format "sc16": one little-endian 32-bit word per complex sample, I in the low half) ----
// The reference flowgraphs receive fc32 from UHD (python/twinrx_usrp_source.py:57), i.e. the host converts every int16 to
// float before the sample reaches autocorrelate; accepting the int16 pairs directly halves the bytes the whole chain is
// bound by.  Conversion is exact and stays off the conversion unit: bias the halves to
// unsigned (xor 0x8000), splice them under the exponent of 2^23 (PRMT) and subtract 2^23 + 32768 (one packed add).
__device__ __forceinline__ float2 sc16_to_c64(unsigned w) {
  const unsigned b = w ^ 0x80008000u;
  const unsigned lo = __byte_perm(b, 0x4B000000u, 0x7410);
  const unsigned hi = __byte_perm(b, 0x4B000000u, 0x7432);
  float2 r;   // one packed add for the pair (FADD2): 4 issue slots per sample
  upk2(add2(pk2(__uint_as_float(lo), __uint_as_float(hi)), pk2(-8421376.0f, -8421376.0f)), r.x, r.y);
  return r;
}
// VEC consecutive samples of one channel, fc32 (S = float2) or sc16 (S = unsigned), streaming loads; !ok: zeros.
// Both formats give a lane the same samples, so the two paths accumulate in the same order.
template <int VEC, typename S>
__device__ __forceinline__ void load_samples(const S* p, bool ok, float2 (&x)[VEC]) {
  static_assert(VEC == 1 || VEC == 2, "one or two samples per load");
  if constexpr (sizeof(S) == 8) {
    if constexpr (VEC == 2) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v = ldg_stream4(reinterpret_cast<const float4*>(p));
      x[0] = make_float2(v.x, v.y);
      x[1] = make_float2(v.z, v.w);
    } else {
      float2 v = make_float2(0.f, 0.f);
      if (ok) v = ldg_stream2(reinterpret_cast<const float2*>(p));
      x[0] = v;
    }
  } else {
    if constexpr (VEC == 2) {
      unsigned a = 0u, b = 0u;
      if (ok) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
      x[0] = sc16_to_c64(a);
      x[1] = sc16_to_c64(b);
    } else {
      unsigned a = 0u;
      if (ok) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(a) : "l"(p));
      x[0] = sc16_to_c64(a);
    }
  }
}

// cp.async (LDGSTS) helpers: 16-byte global -> shared copies that bypass registers; src_bytes = 0 zero-fills.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
// 8-byte variant (two sc16 samples); .cg exists for 16 bytes only
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NKEEP> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(NKEEP) : "memory"); }

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


// Per-lane accumulators of the Hermitian lower half (M diagonal pairs + M(M-1)/2 complex).  An off-diagonal
// entry is one packed pair (re, im) updated by two fma.rn.f32x2 per sample:
//     (re, im) += (x_r.re, x_r.im) * x_c.re          (re, im) += (x_r.im, -x_r.re) * x_c.im
// which is, per component and in this order, the scalar form  re = fma(xr.re, xc.re, re); re = fma(xr.im, xc.im, re);
// im = fma(xr.im, xc.re, im); im = fma(-xr.re, xc.im, im)  -- same bits, half the issue slots (M = 8: 56 FFMA2 + 16 FFMA
// per sample instead of 128 FFMA; the broadcast and the swapped/negated pair are operand modifiers in SASS).
template <int M>
struct CovAcc {
  static constexpr int NP = M * (M - 1) / 2;
  // M > 8: a diagonal entry as the packed pair (sum re^2, sum im^2), ONE fma.rn.f32x2 per sample instead of two dependent FFMA,
  // the halves added when the lanes' partial sums are folded (the 16-element covariance: 1.67 -> 1.55 ms per 65,536 frames).
  // M <= 8 keeps the scalar chain: measured 0.3 % (8 elements) and 0.7 % (4) slower packed, A/B on one box (tools/ab_libs.py).
  static constexpr bool PKD = M > 8;
  float dg[PKD ? 1 : M];
  f32x2 dg2[PKD ? M : 1];
  f32x2 od[NP > 0 ? NP : 1];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int r = 0; r < M; ++r) { if constexpr (PKD) dg2[r] = 0ull; else dg[r] = 0.0f; }
#pragma unroll
    for (int p = 0; p < NP; ++p) od[p] = 0ull;
  }
  // one time sample of all M channels
  __device__ __forceinline__ void add(const float2 (&x)[M]) {
#pragma unroll
    for (int r = 0; r < M; ++r) {
      const float2 xr = x[r];
      const f32x2 xp = pk2(xr.x, xr.y), xs = pk2(xr.y, -xr.x);
      if constexpr (PKD) dg2[r] = fma2(xp, xp, dg2[r]);
      else { dg[r] = fmaf(xr.x, xr.x, dg[r]); dg[r] = fmaf(xr.y, xr.y, dg[r]); }
#pragma unroll
      for (int c = 0; c < r; ++c) {
        const float2 xc = x[c];
        const int p = r * (r - 1) / 2 + c;
        od[p] = fma2(xp, pk2(xc.x, xc.x), od[p]);   // x_r conj(x_c)
        od[p] = fma2(xs, pk2(xc.y, xc.y), od[p]);
      }
    }
  }
  __device__ __forceinline__ void get(int p, float& re, float& im) const { upk2(od[p], re, im); }
  __device__ __forceinline__ float diag(int r) const {
    if constexpr (PKD) { float a, b; upk2(dg2[r], a, b); return a + b; }
    else return dg[r];
  }
  // fold the 32 lanes' partial matrices; `red` (M*M floats, shared, this warp's) receives the raw sums
  __device__ __forceinline__ void fold(unsigned lane, float* red) {
    constexpr int CNT = M * M;
    float a[CNT];
#pragma unroll
    for (int p = 0; p < NP; ++p) upk2(od[p], a[2 * p], a[2 * p + 1]);
#pragma unroll
    for (int r = 0; r < M; ++r) a[2 * NP + r] = diag(r);
    warp_reduce_scatter<CNT, 16>(a, lane);
    const int b = rs_base<CNT>(lane);
    constexpr int F = CNT >= 32 ? CNT / 32 : 1;
#pragma unroll
    for (int i = 0; i < F; ++i) red[b + i] = a[i];
    __syncwarp();
  }
};

// One frame by one warp: accumulate the Hermitian lower half over the frame's time axis and fold the 32 partial
// matrices; on return `red` (M*M floats, shared memory, this warp's) holds the raw sums (layout: see folded_entry).
template <int M, int VEC, int G, typename S>
__device__ __forceinline__ void cov_warp_frame(const S* __restrict__ base, long long chan_stride, int N, unsigned lane,
                                               float* red) {
  CovAcc<M> acc;
  acc.clear();
  // G independent load groups per iteration (G*M LDG.128 in flight per lane; LDG.64 for sc16).
  for (int t0 = (int)lane * VEC; t0 < N; t0 += G * 32 * VEC) {
    float2 x[G][VEC][M];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int t = t0 + g * 32 * VEC;
      const bool ok = t < N;   // N % VEC == 0 is guaranteed by the launcher
#pragma unroll
      for (int k = 0; k < M; ++k) {
        float2 v[VEC];
        load_samples<VEC, S>(base + (long long)k * chan_stride + t, ok, v);
#pragma unroll
        for (int s = 0; s < VEC; ++s) x[g][s][k] = v[s];
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int s = 0; s < VEC; ++s) acc.add(x[g][s]);
  }
  acc.fold(lane, red);
}

// Per-channel complex gains applied in front of the covariance (the reference's antenna_correction block,
// lib/antenna_correction_impl.cc:65-70,90-96, and phase_correct_hier: x'_k = g_k x_k) folded into the emit stage:
// R'(r, c) = g_r conj(g_c) R(r, c).  g = nullptr: none.  The diagonal factor |g_r|^2 is kept exactly real.
__device__ __forceinline__ float2 apply_gain(float2 v, const float2* __restrict__ g, int r, int c) {
  if (g == nullptr) return v;
  const float2 a = g[r], b = g[c];
  const float gx = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
  const float gy = (r == c) ? 0.0f : __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y));
  return make_float2(__fsub_rn(__fmul_rn(v.x, gx), __fmul_rn(v.y, gy)), __fadd_rn(__fmul_rn(v.x, gy), __fmul_rn(v.y, gx)));
}

// Scale, optional channel gains, optional forward-backward term, Hermitian expansion; writes the M x M column-major matrix
// to `o` (global or shared).
template <int M>
__device__ __forceinline__ void cov_warp_emit(const float* red, float scale, float bscale, int avg_method, unsigned lane,
                                              float2* o, const float2* __restrict__ gains = nullptr) {
  constexpr int CNT = M * M;
  for (int e = (int)lane; e < CNT; e += 32) {
    const int r = e % M, c = e / M;
    float2 v = apply_gain(folded_entry<M>(red, r, c, scale), gains, r, c);
    if (avg_method == 1) {
      // 0.5*R + (0.5/N) * J conj(R) J : (J conj(R) J)(r,c) = conj(R(M-1-r, M-1-c))   lib/autocorrelate_impl.cc:108
      const float2 w = apply_gain(folded_entry<M>(red, M - 1 - r, M - 1 - c, scale), gains, M - 1 - r, M - 1 - c);
      v.x = __fadd_rn(__fmul_rn(0.5f, v.x), __fmul_rn(bscale, w.x));
      v.y = __fadd_rn(__fmul_rn(0.5f, v.y), __fmul_rn(bscale, -w.y));
    }
    o[e] = v;
  }
  __syncwarp();
}

// ---- M = 16: two warps per frame sharing one cp.async ring (cov.cu: cov16_ring_kernel; fused16.cu) ------------------------------
__device__ const unsigned char kTri8Row[28] = {1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7};
__device__ const unsigned char kTri8Col[28] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6};


// Ring-fed variant (16-byte aligned, even N): the two warps of a pair share ONE cp.async ring -- each loads 8 of the 16
// channels, both read all 16 after the pair's barrier -- so a frame crosses L2 -> SM once instead of twice and the loads of
// chunk q + STAGES - 2 are in flight while chunk q is accumulated (the LDG version stalls a 255-register warp on every load).
constexpr int C16_STAGES = 5;

// One role's whole loop (the two roles are separate instantiations so that each only carries its own 128 accumulators).
// S = float2: fc32 samples, float4 ring slots; S = unsigned: sc16 samples, uint2 ring slots (the same two samples per lane).
template <typename S> struct Ring16Slot { typedef float4 type; };
template <> struct Ring16Slot<unsigned> { typedef uint2 type; };

// The pair works through frames first, first + fstep, ... (nfw of them); after a frame's sums are complete in `red` (layout of
// folded_entry<16>) both warps call emit(k, red) with k the frame's position in that sequence: cov.cu writes R to global
// memory there, the fused M = 16 kernel (fused16.cu) into a shared-memory tile.  bar_id: the pair's named barrier (64 threads).
template <int ROLE, typename S, typename Emit>
__device__ __forceinline__ void cov16_ring_role(const S* __restrict__ in, long long frame_stride, long long chan_stride, int N,
                                                long long first, long long fstep, int nfw, typename Ring16Slot<S>::type* ring,
                                                float* red, int bar_id, int lane, Emit&& emit) {
  typedef typename Ring16Slot<S>::type Slot;
  constexpr bool SC16 = sizeof(S) == 4;
  constexpr int NP16 = 120;
  const int NCH = (N + 63) / 64;                      // 64-sample chunks per frame (2 samples per lane)
  const int total = nfw * NCH;
  // issue cursor: this warp loads channels 8 ROLE .. 8 ROLE + 7 of every chunk
  const S* ibase = in + first * frame_stride + (long long)(8 * ROLE) * chan_stride;
  int ic = 0, istage = 0, issued = 0;
  auto issue = [&]() {
    const int t = ic * 64 + lane * 2;
    const int nbytes = (t < N) ? (int)sizeof(Slot) : 0;
    Slot* dst = ring + ((size_t)istage * 16 + 8 * ROLE) * 32 + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if constexpr (SC16) cp_async8(dst + k * 32, ibase + (long long)k * chan_stride + (nbytes ? t : 0), nbytes);
      else cp_async16(dst + k * 32, ibase + (long long)k * chan_stride + (nbytes ? t : 0), nbytes);
    }
    if (++ic == NCH) { ic = 0; ibase += fstep * frame_stride; }
    if (++istage == C16_STAGES) istage = 0;
    ++issued;
  };
#pragma unroll
  for (int q = 0; q < C16_STAGES - 2; ++q) { if (issued < total) issue(); cp_async_commit(); }

  f32x2 od[64];                                       // ROLE 0: two CovAcc<8>-style diagonal blocks; ROLE 1: R[8 + i][j] at od[i * 8 + j]
  f32x2 dg2[ROLE == 0 ? 16 : 1];                      // (sum re^2, sum im^2) per diagonal entry, see CovAcc
  auto clear = [&]() {
#pragma unroll
    for (int i = 0; i < 64; ++i) od[i] = 0ull;
    if constexpr (ROLE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dg2[i] = 0ull;
    }
  };
  clear();
  int c = 0, rstage = 0, kf = 0;
  for (int q = 0; q < total; ++q) {
    // the stage refilled here was read two iterations ago at the latest, before the barrier both warps passed last time
    if (issued < total) issue();
    cp_async_commit();
    cp_async_wait<C16_STAGES - 2>();
    asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory");      // both halves of this chunk have landed
    const Slot* src = ring + (size_t)rstage * 16 * 32 + lane;
    if (++rstage == C16_STAGES) rstage = 0;
    float2 x[2][16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const Slot v = src[k * 32];
      if constexpr (SC16) { x[0][k] = sc16_to_c64(v.x); x[1][k] = sc16_to_c64(v.y); }
      else { x[0][k] = make_float2(v.x, v.y); x[1][k] = make_float2(v.z, v.w); }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if constexpr (ROLE == 0) {
        // diagonal blocks b = 0, 1: entry (r > c) of block b at od[32 b + r (r - 1) / 2 + c]  (28 of 32 used), diagonals in dg
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float2 xr = x[s][8 * b + r];
            const f32x2 xp = pk2(xr.x, xr.y), xs = pk2(xr.y, -xr.x);
            dg2[8 * b + r] = fma2(xp, xp, dg2[8 * b + r]);
#pragma unroll
            for (int cc = 0; cc < r; ++cc) {
              const float2 xc = x[s][8 * b + cc];
              const int p = 32 * b + r * (r - 1) / 2 + cc;
              od[p] = fma2(xp, pk2(xc.x, xc.x), od[p]);   // x_r conj(x_c), see CovAcc
              od[p] = fma2(xs, pk2(xc.y, xc.y), od[p]);
            }
          }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 xr = x[s][8 + i];
          const f32x2 xp = pk2(xr.x, xr.y), xs = pk2(xr.y, -xr.x);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 xc = x[s][j];
            od[i * 8 + j] = fma2(xp, pk2(xc.x, xc.x), od[i * 8 + j]);
            od[i * 8 + j] = fma2(xs, pk2(xc.y, xc.y), od[i * 8 + j]);
          }
        }
      }
    }
    if (++c == NCH) {
      c = 0;
      float a[128];
      if constexpr (ROLE == 0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
          for (int p = 0; p < 28; ++p) upk2(od[32 * b + p], a[64 * b + 2 * p], a[64 * b + 2 * p + 1]);
#pragma unroll
          for (int r = 0; r < 8; ++r) { float d0, d1; upk2(dg2[8 * b + r], d0, d1); a[64 * b + 56 + r] = d0 + d1; }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) upk2(od[i], a[2 * i], a[2 * i + 1]);
      }
      clear();
      warp_reduce_scatter<128, 16>(a, (unsigned)lane);   // lane L now holds the full sums of elements 4L .. 4L+3
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = lane * 4 + i;
        int pos;
        if constexpr (ROLE == 0) {
          const int blk = k >> 6, kk = k & 63;
          if (kk < 56) {
            const int r = 8 * blk + kTri8Row[kk >> 1], cc = 8 * blk + kTri8Col[kk >> 1];
            pos = 2 * (r * (r - 1) / 2 + cc) + (kk & 1);
          } else {
            pos = 2 * NP16 + 8 * blk + (kk - 56);
          }
        } else {
          const int r = 8 + (k >> 4), cc = (k >> 1) & 7;
          pos = 2 * (r * (r - 1) / 2 + cc) + (k & 1);
        }
        red[pos] = a[i];
      }
      asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory");   // both roles' sums are in red
      emit(kf, red);
      ++kf;
      // red is rewritten only after the next frame's chunk barriers, which both warps pass after finishing this loop
    }
  }
}


// Half of the 16 x 16 matrix per role from the pair's folded sums: scale, channel gains, forward-backward term, Hermitian
// expansion (the emit stage of cov_warp_emit for two warps); `o` may be global or shared memory.
template <int ROLE>
__device__ __forceinline__ void cov16_pair_emit(const float* red, float scale, float bscale, int avg_method, int lane, float2* o,
                                                const float2* __restrict__ gains) {
  constexpr int M = 16, CNT = 256;
  for (int e = ROLE * 32 + lane; e < CNT; e += 64) {
    const int r = e % M, cc = e / M;
    float2 v = apply_gain(folded_entry<M>(red, r, cc, scale), gains, r, cc);
    if (avg_method == 1) {   // 0.5*R + (0.5/N) * J conj(R) J, lib/autocorrelate_impl.cc:108
      const float2 wv = apply_gain(folded_entry<M>(red, M - 1 - r, M - 1 - cc, scale), gains, M - 1 - r, M - 1 - cc);
      v.x = __fadd_rn(__fmul_rn(0.5f, v.x), __fmul_rn(bscale, wv.x));
      v.y = __fadd_rn(__fmul_rn(0.5f, v.y), __fmul_rn(bscale, -wv.y));
    }
    o[e] = v;
  }
}

}  // namespace
}  // namespace doa
