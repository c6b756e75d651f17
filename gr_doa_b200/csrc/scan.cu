// scan.cu -- stage 2b (pseudo-spectrum scan) and stage 4 (peak picking), fused and standalone.
//
// Replaces the angle loop + normalisation of MUSIC_lin_array_impl::work (gr-doa lib/MUSIC_lin_array_impl.cc:137-142)
// and find_local_max_impl (lib/find_local_max_impl.cc:80-165,179-190).
//
// Null spectrum.  For a ULA the steering vector is a(theta)_r = e^{-j psi (M-1-2r)/2}, psi = 2 pi d cos(theta), so
//     Q(theta) = a^H G a = u_0 + 2 Re sum_{l=1}^{M-1} u_l z^l,   z = e^{j psi},  u_l = sum_r G[r][r+l]
// (the polynomial Root-MUSIC builds, lib/rootMUSIC_linear_array_impl.cc:74-79): M-1 complex MACs per angle instead
// of the M*(M+1) of the literal v^H G v.  The scan evaluates it by Horner with z read from a per-plan table
// (the grid is uniform in theta, not in psi, so z has no recurrence).  The bins the chain reports are then
// re-evaluated with the reference's own arithmetic -- v^H G v on the steering table the constructor builds, same
// operation order as a plain C++ loop (row = v^H G, then row . v) -- in a +-2 bin window, so peak bins and peak heights follow the reference's rounding.
//
// Mapping: one warp per frame, lane L owns the contiguous bins [L*S, (L+1)*S), S = ceil(P/32).  Peak picking is a
// sequential state machine per lane (rise .. plateau .. fall, the plateau rule of find_local_max_impl.cc:92-107),
// stitched across lanes with one ballot; per-lane top-K lists are merged with K warp arg-reductions.  Nothing but
// K (value, location, bin) triples per frame leaves the chip.
#include "scan_device.cuh"
#include <algorithm>

namespace doa {
namespace {

constexpr int SCAN_WARPS = 8;

template <int MT, int KL>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_peaks_kernel(const float2* __restrict__ u, const float2* __restrict__ G, const float* __restrict__ zpair,
                  const float2* __restrict__ Vtab, const float* __restrict__ xaxis, int M, int P, int nframes, int K,
                  float* __restrict__ out_val, float* __restrict__ out_loc, int* __restrict__ out_bin, int z_in_smem) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  // z table: staged in shared memory when it fits (conflict-free LDS.128), read in place from global memory otherwise
  const ZTab zt = z_in_smem ? ztab_fill(smem, zpair, P) : ztab_view(zpair, P);
  float2* us_all = reinterpret_cast<float2*>(smem + (z_in_smem ? ztab_floats(P) : 0));   // [SCAN_WARPS][M] (runtime-M path only)
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* us = us_all + warp * (MT > 0 ? 0 : M);
  for (int f = blockIdx.x * SCAN_WARPS + warp; f < nframes; f += gridDim.x * SCAN_WARPS)
    scan_frame_peaks<MT, KL>(u + (size_t)f * M, G + (size_t)f * M * M, zt, us, Vtab, xaxis, M, P, K, lane,
                             out_val + (size_t)f * K, out_loc + (size_t)f * K, out_bin ? out_bin + (size_t)f * K : nullptr);
}

// Generic M (large arrays), few frames: one warp per frame leaves the machine empty and walks P x (M-1) Horner steps
// serially (1.34 ms for 256 frames of 64 elements x 16,384 bins).  Here a CTA owns a frame: all its warps evaluate the
// coarse spectrum into shared memory (same arithmetic, same z values: same bits), then one warp runs the unchanged
// walker / merge / refinement on that table.
constexpr int SCAN_WIDE_THREADS = 256;

// The refinement (v^H G v with the reference's operation order at the +-2 bins around each of the K winners and around the
// global minimum) is M^2 complex multiply-adds per bin: at 64 elements a lane that evaluates it alone walks 4,096 dependent
// steps through global memory, and with one warp at work the CTA spent 90 % of its time there (1.02 ms for 512 frames).  The
// reference's order only chains the additions WITHIN a column of G (row sum over r, then the sum of the M column results over
// c), so the M column sums of a bin are independent: warp 0 posts its 32 (slot, bin) requests of a round to shared memory, all
// eight warps take (request, column) pairs -- G staged once per frame in shared memory with an odd leading dimension -- and
// the requesting lane adds the M column results in column order.  Operation for operation q_faithful: identical bits.
__device__ __forceinline__ void wide_bar(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(SCAN_WIDE_THREADS) : "memory"); }

struct CoopEval {
  int* req;            // [32] bin requested by lane l of warp 0, or -1
  float* pxs;          // [32][M + 1] column results of request l
  const float2* Gs;    // [M][M + 1] the frame's projector, column c at Gs + c * (M + 1)
  const float2* Vtab;
  int M;
  // one round by every thread of the CTA (warp 0 between its two barriers, the others in their helper loop)
  __device__ __forceinline__ void round_work() const {
    const int LD = M + 1;
    for (int id = threadIdx.x; id < 32 * M; id += SCAN_WIDE_THREADS) {
      const int rq = id / M, c = id - rq * M;
      const int b = req[rq];
      if (b < 0) continue;
      const float2* v = Vtab + (size_t)b * M;
      const float2* Gc = Gs + (size_t)c * LD;
      float rx = 0.0f, ry = 0.0f;
#pragma unroll 8
      for (int r = 0; r < M; ++r) {
        const float2 g = Gc[r]; const float2 vr = v[r];
        const float px = __fsub_rn(__fmul_rn(vr.x, g.x), __fmul_rn(-vr.y, g.y));
        const float py = __fadd_rn(__fmul_rn(vr.x, g.y), __fmul_rn(-vr.y, g.x));
        rx = __fadd_rn(rx, px); ry = __fadd_rn(ry, py);
      }
      const float2 vc = v[c];
      pxs[rq * LD + c] = __fsub_rn(__fmul_rn(rx, vc.x), __fmul_rn(ry, vc.y));
    }
  }
  __device__ __forceinline__ float operator()(bool valid, int b, int lane) const {
    req[lane] = valid ? b : -1;
    wide_bar(1);                                  // requests posted
    round_work();
    wide_bar(2);                                  // column results ready
    float qx = 0.0f;
    if (valid) { const float* p = pxs + lane * (M + 1); for (int c = 0; c < M; ++c) qx = __fadd_rn(qx, p[c]); }
    return qx;
  }
};

template <int KL>
__global__ void __launch_bounds__(SCAN_WIDE_THREADS)
scan_peaks_wide_kernel(const float2* __restrict__ u, const float2* __restrict__ G, const float2* __restrict__ z,
                       const float* __restrict__ zpair, const float2* __restrict__ Vtab, const float* __restrict__ xaxis, int M,
                       int P, int nframes, int K, float* __restrict__ out_val, float* __restrict__ out_loc,
                       int* __restrict__ out_bin) {
  extern __shared__ float4 smem4[];
  float* qtab = reinterpret_cast<float*>(smem4);                           // [P]
  float2* us = reinterpret_cast<float2*>(qtab + ((P + 3) & ~3));            // [M]
  float2* Gs = us + M;                                                     // [M][M + 1]
  float* pxs = reinterpret_cast<float*>(Gs + (size_t)M * (M + 1));         // [32][M + 1]
  int* req = reinterpret_cast<int*>(pxs + 32 * (M + 1));                   // [32]
  const ZTab zt = ztab_view(zpair, P);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const CoopEval ev{req, pxs, Gs, Vtab, M};
  const int rounds = K / 4 + 1;                                            // peaks_refine_emit: for (base = 0; base <= K; base += 4)
  for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
    __syncthreads();                                                       // the previous frame is done with qtab / us / Gs
    for (int l = threadIdx.x; l < M; l += SCAN_WIDE_THREADS) us[l] = u[(size_t)f * M + l];
    for (int e = threadIdx.x; e < M * M; e += SCAN_WIDE_THREADS) { const int c = e / M, r = e - c * M; Gs[c * (M + 1) + r] = G[(size_t)f * M * M + e]; }
    __syncthreads();
    // coarse spectrum: four adjacent bins per thread as two packed pairs (fma.rn.f32x2 is two scalar FMAs: the bits of
    // q_coarse<0>), so a coefficient costs one shared-memory load per four bins
    for (int i0 = 4 * threadIdx.x; i0 < P; i0 += 4 * SCAN_WIDE_THREADS) {
      if (i0 + 4 <= P) {
        const float4 za = *reinterpret_cast<const float4*>(z + i0), zb = *reinterpret_cast<const float4*>(z + i0 + 2);   // (x0,y0,x1,y1), (x2,y2,x3,y3)
        const f32x2 zx0 = pk2(za.x, za.z), zy0 = pk2(za.y, za.w), nzy0 = pk2(-za.y, -za.w);
        const f32x2 zx1 = pk2(zb.x, zb.z), zy1 = pk2(zb.y, zb.w), nzy1 = pk2(-zb.y, -zb.w);
        const float2 top = us[M - 1];
        f32x2 ax0 = pk2(top.x, top.x), ay0 = pk2(top.y, top.y), ax1 = ax0, ay1 = ay0;
        for (int l = M - 2; l >= 1; --l) {
          const float2 c = us[l];
          const f32x2 cx = pk2(c.x, c.x), cy = pk2(c.y, c.y);
          const f32x2 nx0 = fma2(ax0, zx0, fma2(ay0, nzy0, cx)), ny0 = fma2(ax0, zy0, fma2(ay0, zx0, cy));
          const f32x2 nx1 = fma2(ax1, zx1, fma2(ay1, nzy1, cx)), ny1 = fma2(ax1, zy1, fma2(ay1, zx1, cy));
          ax0 = nx0; ay0 = ny0; ax1 = nx1; ay1 = ny1;
        }
        // re = fmaf(ax, z.x, -ay * z.y);  q = fmaf(2, re, u0)
        const f32x2 re0 = fma2(ax0, zx0, mul2(ay0, nzy0)), re1 = fma2(ax1, zx1, mul2(ay1, nzy1));
        const f32x2 two = pk2(2.0f, 2.0f), u0 = pk2(us[0].x, us[0].x);
        float4 q;
        upk2(fma2(two, re0, u0), q.x, q.y); upk2(fma2(two, re1, u0), q.z, q.w);
        *reinterpret_cast<float4*>(qtab + i0) = q;
      } else {
        const float2 dummy[1] = {make_float2(0.f, 0.f)};
        for (int i = i0; i < P; ++i) qtab[i] = q_coarse<0>(dummy, us, M, z[i]);
      }
    }
    __syncthreads();
    if (warp == 0) {
      scan_frame_peaks_ev<0, KL>(u + (size_t)f * M, ev, zt, us, xaxis, M, P, K, lane, out_val + (size_t)f * K,
                                 out_loc + (size_t)f * K, out_bin ? out_bin + (size_t)f * K : nullptr, qtab);
    } else {
      for (int rd = 0; rd < rounds; ++rd) { wide_bar(1); ev.round_work(); wide_bar(2); }
    }
  }
}

// K == 1 uses index_max (find_local_max_impl.h:53-56): the global arg-max, no local-peak logic.
template <int MT>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_argmax_kernel(const float2* __restrict__ u, const float2* __restrict__ G, const float2* __restrict__ ztab,
                   const float2* __restrict__ Vtab, const float* __restrict__ xaxis, int M, int P, int nframes,
                   float* __restrict__ out_val, float* __restrict__ out_loc, int* __restrict__ out_bin) {
  extern __shared__ float2 smem[];
  float2* us_all = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* us = us_all + warp * (MT > 0 ? 0 : M);
  for (int f = blockIdx.x * SCAN_WARPS + warp; f < nframes; f += gridDim.x * SCAN_WARPS)
    scan_frame_argmax<MT>(u + (size_t)f * M, G + (size_t)f * M * M, ztab, us, Vtab, xaxis, M, P, lane, out_val + f, out_loc + f,
                          out_bin ? out_bin + f : nullptr);
}

// ---------------------------------------------------------------------------------------------------------------
// Full dB spectrum: one warp per frame, bins interleaved across lanes (coalesced stores), two passes over the
// Horner form (minimum first, then 10*log10(y/ymax)); no intermediate spectrum is stored.
template <int MT>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_spectrum_kernel(const float2* __restrict__ u, const float2* __restrict__ ztab, int M, int P, int nframes,
                     float* __restrict__ out) {
  extern __shared__ float2 smem[];
  float2* us_all = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* us = us_all + warp * (MT > 0 ? 0 : M);
  for (int f = blockIdx.x * SCAN_WARPS + warp; f < nframes; f += gridDim.x * SCAN_WARPS) {
    float2 uc[MT > 0 ? MT : 1];
    const float2* uf = u + (size_t)f * M;
    if constexpr (MT > 0) {
#pragma unroll
      for (int l = 0; l < MT; ++l) uc[l] = uf[l];
    } else {
      __syncwarp();
      for (int l = lane; l < M; l += 32) us[l] = uf[l];
      __syncwarp();
    }
    // max over bins of y = 1/Q taken on y itself so that negative/zero Q behave like the reference's float max
    float ymax = -INFINITY;
    for (int i = lane; i < P; i += 32) {
      const float y = __fdiv_rn(1.0f, q_coarse<MT>(uc, us, M, ztab[i]));
      ymax = fmaxf(ymax, y);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
    float* of = out + (size_t)f * P;
    for (int i = lane; i < P; i += 32) {
      const float y = __fdiv_rn(1.0f, q_coarse<MT>(uc, us, M, ztab[i]));
      of[i] = __fmul_rn(10.0f, log10f(__fdiv_rn(y, ymax)));
    }
  }
}

// Compile-time M, P floats of shared memory per warp: ONE Horner pass, over pairs of bins (i, i + 32) in fma.rn.f32x2 (the
// same operations as the scalar form), lg2(1/Q) parked in shared memory, then the dB pass.
// 10 log10(y / ymax) = 10 log10(2) (lg2 y - lg2 ymax): MUFU reciprocal and log2 (abs. error ~1e-6 dB against the 1e-3 dB
// tolerance) instead of two IEEE divisions and log10f, which were 45 of the two-pass kernel's ~120 instructions per bin.
template <int MT>
__global__ void __launch_bounds__(128)
scan_spectrum_smem_kernel(const float2* __restrict__ u, const float2* __restrict__ ztab, int P, int nframes, float* __restrict__ out) {
  extern __shared__ float ysm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* ys = ysm + (size_t)warp * P;
  for (int f = blockIdx.x * nwarps + warp; f < nframes; f += gridDim.x * nwarps) {
    float2 uc[MT];
    const float2* uf = u + (size_t)f * MT;
#pragma unroll
    for (int l = 0; l < MT; ++l) uc[l] = uf[l];
    f32x2 ux2[MT], muy2[MT];
#pragma unroll
    for (int l = 0; l < MT; ++l) { ux2[l] = pk2(uc[l].x, uc[l].x); muy2[l] = pk2(-uc[l].y, -uc[l].y); }
    const f32x2 u0_2 = pk2(uc[0].x, uc[0].x), two2 = pk2(2.0f, 2.0f);
    // max over bins of y = 1/Q taken on y itself so that negative/zero Q behave like the reference's float max
    float ymax = -INFINITY;
    __syncwarp();
    int i = lane;
    for (; i + 32 < P; i += 64) {
      const float2 za = ztab[i], zb = ztab[i + 32];
      float qa, qb;
      upk2(q_coarse_pair<MT>(ux2, muy2, u0_2, two2, pk2(za.x, zb.x), pk2(za.y, zb.y), pk2(-za.y, -zb.y)), qa, qb);
      const float ya = __fdividef(1.0f, qa), yb = __fdividef(1.0f, qb);
      ys[i] = __log2f(ya); ys[i + 32] = __log2f(yb);
      ymax = fmaxf(ymax, fmaxf(ya, yb));
    }
    if (i < P) {
      const float y = __fdividef(1.0f, q_coarse<MT>(uc, nullptr, MT, ztab[i]));
      ys[i] = __log2f(y);
      ymax = fmaxf(ymax, y);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
    const float lmax = __log2f(ymax);
    float* of = out + (size_t)f * P;
    for (int k = lane; k < P; k += 32) of[k] = 3.0102999566f * (ys[k] - lmax);   // own entries: no barrier needed
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Standalone find_local_max on arbitrary float vectors: one warp per vector, vector staged in shared memory
// (coalesced in, padded per-lane segments out), then the same walker in MAXIMA mode.  Bit-exact by construction:
// only comparisons and copies of the input floats.
constexpr int FLM_WARPS = 4;

template <int KL>
__global__ void __launch_bounds__(FLM_WARPS * 32)
find_local_max_kernel(const float* __restrict__ in, int len, int nframes, int K, const float* __restrict__ xaxis,
                      float* __restrict__ out_val, float* __restrict__ out_loc, int* __restrict__ out_bin) {
  extern __shared__ float fsm[];
  const int S = (len + 31) / 32, SP = S | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* vs = fsm + (size_t)warp * 32 * SP;
  const int s0 = lane * S, s1 = min(len, s0 + S);
  const bool pow2 = (S & (S - 1)) == 0;
  const int lgS = 31 - __clz(S);
  // first occurrence of the maximum over interleaved bins (index_max semantics: NaNs never win, ties -> lowest index)
  auto arg_max = [&](const float* src) -> int {
    float bv = -INFINITY; int bi = 0x7fffffff;
    constexpr int U = 8;
    int i = lane;
    for (; i + (U - 1) * 32 < len; i += U * 32) {
      float t[U];
#pragma unroll
      for (int j = 0; j < U; ++j) t[j] = __ldg(src + i + j * 32);
#pragma unroll
      for (int j = 0; j < U; ++j) if (t[j] > bv) { bv = t[j]; bi = i + j * 32; }
    }
    for (; i < len; i += 32) { const float c = src[i]; if (c > bv) { bv = c; bi = i; } }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL, bv, o); const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    return (bi == 0x7fffffff) ? 0 : bi;   // nothing compared greater than -inf (all -inf / NaN): index 0 like index_max
  };
  for (int f = blockIdx.x * nwarps + warp; f < nframes; f += gridDim.x * nwarps) {
    const float* src = in + (size_t)f * len;
    int bin; float val;
    if (K == 1) {   // index_max (find_local_max_impl.h:53-56): a pure stream, nothing is staged
      bin = arg_max(src); val = src[bin];
    } else {
      __syncwarp();
      {   // coalesced in, padded per-lane segments out.  Sixteen independent loads are issued before the first store: with
          // one load in flight per warp the kernel streamed at 1.1 TB/s (12 warps per SM x 128 B against ~1 us of latency)
        constexpr int U = 16;
        int i = lane;
        for (; i + (U - 1) * 32 < len; i += U * 32) {
          float t[U];
#pragma unroll
          for (int j = 0; j < U; ++j) t[j] = __ldg(src + i + j * 32);
          if (pow2) {
#pragma unroll
            for (int j = 0; j < U; ++j) { const int e = i + j * 32; vs[(e >> lgS) * SP + (e & (S - 1))] = t[j]; }
          } else {
#pragma unroll
            for (int j = 0; j < U; ++j) { const int e = i + j * 32; vs[(e / S) * SP + (e % S)] = t[j]; }
          }
        }
        for (; i < len; i += 32) vs[(i / S) * SP + (i % S)] = src[i];
      }
      __syncwarp();
      const float* vl = vs + lane * SP;
      auto v_at = [&](int b) -> float { return vs[(b / S) * SP + (b % S)]; };
      Walker<KL, true> w; w.init(s0 > 0);
      if (s0 < s1) {
        float prev; int k = 0;
        if (s0 > 0) prev = vs[(lane - 1) * SP + (S - 1)];
        else { prev = vl[0]; k = 1; }
        for (; k < s1 - s0; ++k) {
          const float c = vl[k];
          w.step(prev, c, s0 + k, v_at);
          prev = c;
        }
      }
      Merged m = stitch_and_merge<KL, true>(w, K, lane, v_at);
      const int nref = min(K, m.nvalid);
      // fill-in rule incl. the reference's index bug (:145-163); the arg-max is only needed when there is no peak at all
      const int pad = (m.nvalid == 0) ? arg_max(src) : m.best_ord;
      bin = (lane < nref) ? m.bin : pad;
      val = (lane < K) ? src[min(bin, len - 1)] : 0.f;
    }
    const float loc = (lane < K) ? xaxis[min(bin, len - 1)] : -INFINITY;
    const int lrank = rank_desc(loc, K, lane);
    if (lane < K) {
      out_val[(size_t)f * K + lane] = val;
      out_loc[(size_t)f * K + lrank] = loc;
      if (out_bin) out_bin[(size_t)f * K + lane] = bin;
    }
  }
}

int sm_count() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

template <int MT>
int launch_peaks_mt(const float2* u, const float2* G, const ScanTables& tb, int nframes, int K, float* out_val,
                    float* out_loc, int* out_bin, cudaStream_t st) {
  const int M = tb.M, P = tb.P;
  const int blocks_needed = (nframes + SCAN_WARPS - 1) / SCAN_WARPS;
  if (K == 1) {
    const size_t smem = (MT > 0 ? 0 : (size_t)SCAN_WARPS * M) * sizeof(float2);
    const int blocks = min(blocks_needed, sm_count() * 8);
    scan_argmax_kernel<MT><<<blocks, SCAN_WARPS * 32, smem, st>>>(u, G, tb.z, tb.V, tb.xaxis, M, P, nframes, out_val,
                                                                  out_loc, out_bin);
    return 1;
  }
  if constexpr (MT == 0) {
    const size_t wsmem = (size_t)((P + 3) & ~3) * sizeof(float) + (size_t)M * sizeof(float2) + (size_t)M * (M + 1) * sizeof(float2) +
                         (size_t)32 * (M + 1) * sizeof(float) + 32 * sizeof(int);
    if (wsmem <= 200 * 1024 && nframes <= 8 * sm_count() && dev_option(OPT_SCAN_WIDE, 1)) {
      const int per = (int)std::max<size_t>(1, std::min<size_t>(4, (222 * 1024) / (wsmem + 1024)));   // 227 KB per SM, 1 KB reserved per CTA
      const int grid = min(nframes, sm_count() * per);
      if (K <= 4) {
        auto kern = scan_peaks_wide_kernel<4>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        kern<<<grid, SCAN_WIDE_THREADS, wsmem, st>>>(u, G, tb.z, tb.zpair, tb.V, tb.xaxis, M, P, nframes, K, out_val, out_loc, out_bin);
      } else {
        auto kern = scan_peaks_wide_kernel<16>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        kern<<<grid, SCAN_WIDE_THREADS, wsmem, st>>>(u, G, tb.z, tb.zpair, tb.V, tb.xaxis, M, P, nframes, K, out_val, out_loc, out_bin);
      }
      return 1;
    }
  }
  const size_t us_bytes = (MT > 0 ? 0 : (size_t)SCAN_WARPS * M) * sizeof(float2);
  const bool z_in_smem = ztab_floats(P) * sizeof(float) + us_bytes <= 200 * 1024;
  const size_t smem = (z_in_smem ? ztab_floats(P) * sizeof(float) : 0) + us_bytes;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem, 1)));
  const int blocks = min(blocks_needed, sm_count() * per_sm);
  if (K <= 4) {
    auto kern = scan_peaks_kernel<MT, 4>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, SCAN_WARPS * 32, smem, st>>>(u, G, tb.zpair, tb.V, tb.xaxis, M, P, nframes, K, out_val, out_loc, out_bin, (int)z_in_smem);
  } else {
    auto kern = scan_peaks_kernel<MT, 16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, SCAN_WARPS * 32, smem, st>>>(u, G, tb.zpair, tb.V, tb.xaxis, M, P, nframes, K, out_val, out_loc, out_bin, (int)z_in_smem);
  }
  return 1;
}

template <int MT>
int launch_spectrum_mt(const float2* u, const ScanTables& tb, int nframes, float* out, cudaStream_t st) {
  if constexpr (MT > 0) {
    const size_t per_warp = (size_t)tb.P * sizeof(float);
    if (per_warp <= 48 * 1024 && dev_option(OPT_SPECTRUM_SMEM, 1)) {
      const int warps = (int)std::max<size_t>(1, std::min<size_t>(4, (64 * 1024) / per_warp));
      const size_t smem = per_warp * warps;
      auto kern = scan_spectrum_smem_kernel<MT>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
      const int blocks = min((nframes + warps - 1) / warps, sm_count() * per_sm);
      kern<<<blocks, warps * 32, smem, st>>>(u, tb.z, tb.P, nframes, out);
      return 1;
    }
  }
  const size_t smem = (MT > 0 ? 0 : (size_t)SCAN_WARPS * tb.M) * sizeof(float2);
  const int blocks = min((nframes + SCAN_WARPS - 1) / SCAN_WARPS, sm_count() * 8);
  scan_spectrum_kernel<MT><<<blocks, SCAN_WARPS * 32, smem, st>>>(u, tb.z, tb.M, tb.P, nframes, out);
  return 1;
}

}  // namespace

void build_zpair_table(const std::vector<float2>& z, std::vector<float>& out) {
  const int P = (int)z.size();
  out.assign(ztab_floats(P), 0.0f);
  const ZTab zt = ztab_view(out.data(), P);
  float* za = out.data(); float* zb = out.data() + 32 * zt.LA;
  for (int i = 0; i < 32 * zt.S; ++i) {
    const float2 v = z[std::min(i, P - 1)];
    const int L = i / zt.S, k = i - L * zt.S;
    za[L * zt.LA + (k >> 1) * 4 + (k & 1)] = v.x;
    za[L * zt.LA + (k >> 1) * 4 + 2 + (k & 1)] = v.y;
    zb[L * zt.LB + (k >> 1) * 2 + (k & 1)] = -v.y;
  }
}

int launch_scan_peaks(const float2* u, const float2* G, const ScanTables& tb, int nframes, int K, float* out_val,
                      float* out_loc, int* out_bin, cudaStream_t st) {
  if (nframes <= 0) return 0;
  if (K < 1 || K > 16) return DOA_CUDA_EINVAL;
  switch (tb.M) {
    case 2: return launch_peaks_mt<2>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
    case 4: return launch_peaks_mt<4>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
    case 8: return launch_peaks_mt<8>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
    case 16: return launch_peaks_mt<16>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
    default: return launch_peaks_mt<0>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
  }
}

int launch_scan_spectrum(const float2* u, const float2* G, const ScanTables& tb, int nframes, float* out,
                         cudaStream_t st) {
  (void)G;
  if (nframes <= 0) return 0;
  switch (tb.M) {
    case 2: return launch_spectrum_mt<2>(u, tb, nframes, out, st);
    case 4: return launch_spectrum_mt<4>(u, tb, nframes, out, st);
    case 8: return launch_spectrum_mt<8>(u, tb, nframes, out, st);
    case 16: return launch_spectrum_mt<16>(u, tb, nframes, out, st);
    default: return launch_spectrum_mt<0>(u, tb, nframes, out, st);
  }
}

int launch_find_local_max(const float* in, int len, int nframes, int K, const float* xaxis, float* out_val,
                          float* out_loc, int* out_bin, cudaStream_t st) {
  if (nframes <= 0) return 0;
  if (K < 1 || K > 16 || len < 1) return DOA_CUDA_EINVAL;
  if (K == 1) {   // index_max streams the vectors: no staging area, full occupancy
    auto kern = find_local_max_kernel<4>;
    kern<<<min((nframes + FLM_WARPS - 1) / FLM_WARPS, sm_count() * 16), FLM_WARPS * 32, 0, st>>>(in, len, nframes, K, xaxis, out_val, out_loc, out_bin);
    return 1;
  }
  const int S = (len + 31) / 32, SP = S | 1;
  int warps = FLM_WARPS;                                        // fewer warps per CTA for very long vectors
  while (warps > 1 && (size_t)warps * 32 * SP * sizeof(float) > 200 * 1024) warps >>= 1;
  const size_t smem = (size_t)warps * 32 * SP * sizeof(float);
  if (smem > 200 * 1024) return DOA_CUDA_EINVAL;                // vector_len beyond ~51k
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
  const int blocks = min((nframes + warps - 1) / warps, sm_count() * per_sm);
  if (K <= 4) {
    auto kern = find_local_max_kernel<4>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, warps * 32, smem, st>>>(in, len, nframes, K, xaxis, out_val, out_loc, out_bin);
  } else {
    auto kern = find_local_max_kernel<16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, warps * 32, smem, st>>>(in, len, nframes, K, xaxis, out_val, out_loc, out_bin);
  }
  return 1;
}

}  // namespace doa
