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
#include "doa_internal.h"
#include <algorithm>
#include <cfloat>

namespace doa {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SCAN_WARPS = 8;
constexpr int REFINE_W = 2;   // refinement half-window in bins

// ---------------------------------------------------------------------------------------------------------------
// Per-lane sorted candidate list (best first).  MAXIMA: best = largest value; otherwise best = smallest.
template <int KL, bool MAXIMA>
struct CandList {
  float val[KL];
  int idx[KL];
  int ord[KL];
  __device__ __forceinline__ static bool better(float a, float b) { return MAXIMA ? (a > b) : (a < b); }
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < KL; ++k) { val[k] = MAXIMA ? -INFINITY : INFINITY; idx[k] = 0x7fffffff; ord[k] = 0; }
  }
  // Insert keeping (value best-first, then lower index first); the worst entry falls off.
  __device__ __forceinline__ void insert(float v, int i, int o) {
#pragma unroll
    for (int k = 0; k < KL; ++k) {
      const bool sw = better(v, val[k]) || (v == val[k] && i < idx[k]);
      if (sw) {
        const float tv = val[k]; const int ti = idx[k]; const int to = ord[k];
        val[k] = v; idx[k] = i; ord[k] = o; v = tv; i = ti; o = to;
      }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int k = 0; k + 1 < KL; ++k) { val[k] = val[k + 1]; idx[k] = idx[k + 1]; ord[k] = ord[k + 1]; }
    val[KL - 1] = MAXIMA ? -INFINITY : INFINITY; idx[KL - 1] = 0x7fffffff; ord[KL - 1] = 0;
  }
};

// Sequential peak walker over one lane's segment.  A peak is the first bin of a (possibly one-bin) plateau that was
// entered by a strict move towards "better" and is left by a strict move away from it; flats inherit the direction
// of the next strict move to their right, a trailing flat counts as "towards" (no peak)  -- find_local_max_impl.cc:89-114.
// State is one integer: cand >= 0 = bin of the pending plateau start, NONE = nothing pending, INCOMING = no strict move
// seen yet in this segment (whatever the previous lanes left pending is still pending).  The hot path per bin is two
// compares and two predicated moves; emitting a candidate (a handful of times per vector) is the only branch.
constexpr int W_NONE = -2, W_INCOMING = -1;
template <int KL, bool MAXIMA>
struct Walker {
  CandList<KL, MAXIMA> list;
  int cand;            // see above
  bool first_away;     // the first strict move of the segment was "away": an incoming pending plateau is a peak
  int n_emit;          // candidates this lane emitted (local ordinals start at 1; 0 is reserved for a stitched one)
  __device__ __forceinline__ void init(bool has_incoming) {
    list.init(); cand = has_incoming ? W_INCOMING : W_NONE; first_away = false; n_emit = 0;
  }
  // move from bin i-1 (value prev) to bin i (value cur); value_of(bin) re-reads / re-evaluates a bin on the rare emit
  template <typename F>
  __device__ __forceinline__ void step(float prev, float cur, int i, F&& value_of) {
    const bool toward = CandList<KL, MAXIMA>::better(cur, prev);
    const bool away = CandList<KL, MAXIMA>::better(prev, cur);
    if (away && cand != W_NONE) {
      if (cand == W_INCOMING) first_away = true;
      else { ++n_emit; list.insert(value_of(cand), cand, n_emit); }
    }
    cand = toward ? i : (away ? W_NONE : cand);
  }
};

// Result of the cross-lane merge, distributed: lane r < K holds final entry r.
struct Merged {
  float val; int bin;   // this lane's final entry (lane < K)
  int nvalid;           // number of peaks found in the whole vector
  int best_ord;         // position of the best peak in the index-ordered peak list (the reference's fill-in "bin")
};

template <int KL, bool MAXIMA, typename F>
__device__ __forceinline__ Merged stitch_and_merge(Walker<KL, MAXIMA>& w, int K, int lane, F&& value_of) {
  // 1. stitch the lane boundaries: incoming state = state of the nearest lower lane that saw a strict move
  const unsigned strict_mask = __ballot_sync(FULL, w.cand != W_INCOMING);
  const unsigned below = strict_mask & ((1u << lane) - 1u);
  const int src = below ? (31 - __clz(below)) : 0;
  const int in_cand = __shfl_sync(FULL, w.cand, src);
  const bool stitched = w.first_away && below != 0 && in_cand >= 0;
  if (stitched) w.list.insert(value_of(in_cand), in_cand, 0);
  const int my_count = w.n_emit + (stitched ? 1 : 0);
  // 2. ordinals: exclusive prefix of counts over lanes
  int incl = my_count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
  const int excl = incl - my_count;
  const int nvalid = __shfl_sync(FULL, incl, 31);
  const int ord_shift = excl - (stitched ? 0 : 1);   // global ordinal = local ordinal + ord_shift
  // 3. K rounds of arg-best over the list heads
  Merged m; m.val = 0.f; m.bin = 0; m.nvalid = nvalid; m.best_ord = 0;
  const int rounds = min(K, nvalid);
  for (int r = 0; r < rounds; ++r) {
    float hv = w.list.val[0]; int hi = w.list.idx[0]; int ho = w.list.ord[0] + ord_shift; int hl = lane;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL, hv, o); const int oi = __shfl_xor_sync(FULL, hi, o);
      const int oo = __shfl_xor_sync(FULL, ho, o); const int ol = __shfl_xor_sync(FULL, hl, o);
      if (CandList<KL, MAXIMA>::better(ov, hv) || (ov == hv && oi < hi)) { hv = ov; hi = oi; ho = oo; hl = ol; }
    }
    if (lane == hl) w.list.pop();
    if (lane == r) { m.val = hv; m.bin = hi; }
    if (r == 0) m.best_ord = ho;
  }
  return m;
}

// ---------------------------------------------------------------------------------------------------------------
// Horner evaluation of the ULA null spectrum at z.  MT > 0: coefficients in registers (compile-time M).
template <int MT>
__device__ __forceinline__ float q_coarse(const float2 (&u)[MT > 0 ? MT : 1], const float2* us, int M, float2 z) {
  if constexpr (MT > 0) {
    float ax = u[MT - 1].x, ay = u[MT - 1].y;
#pragma unroll
    for (int l = MT - 2; l >= 1; --l) {
      const float nx = fmaf(ax, z.x, fmaf(-ay, z.y, u[l].x));
      const float ny = fmaf(ax, z.y, fmaf(ay, z.x, u[l].y));
      ax = nx; ay = ny;
    }
    const float re = fmaf(ax, z.x, -ay * z.y);
    return fmaf(2.0f, re, u[0].x);
  } else {
    float ax = us[M - 1].x, ay = us[M - 1].y;
    for (int l = M - 2; l >= 1; --l) {
      const float2 c = us[l];
      const float nx = fmaf(ax, z.x, fmaf(-ay, z.y, c.x));
      const float ny = fmaf(ax, z.y, fmaf(ay, z.x, c.y));
      ax = nx; ay = ny;
    }
    const float re = fmaf(ax, z.x, -ay * z.y);
    return fmaf(2.0f, re, us[0].x);
  }
}

// v^H G v in the reference operation order (row = v^H G first, then row . v), plain fp32 multiplies and adds.
__device__ __forceinline__ float q_faithful(const float2* __restrict__ G, const float2* __restrict__ v, int M) {
  float qx = 0.0f, qy = 0.0f;
  for (int c = 0; c < M; ++c) {
    float rx = 0.0f, ry = 0.0f;
    const float2* Gc = G + (size_t)c * M;
    for (int r = 0; r < M; ++r) {
      const float2 g = Gc[r]; const float2 vr = v[r];
      // conj(v_r) * g
      const float px = __fsub_rn(__fmul_rn(vr.x, g.x), __fmul_rn(-vr.y, g.y));
      const float py = __fadd_rn(__fmul_rn(vr.x, g.y), __fmul_rn(-vr.y, g.x));
      rx = __fadd_rn(rx, px); ry = __fadd_rn(ry, py);
    }
    const float2 vc = v[c];
    const float px = __fsub_rn(__fmul_rn(rx, vc.x), __fmul_rn(ry, vc.y));
    const float py = __fadd_rn(__fmul_rn(rx, vc.y), __fmul_rn(ry, vc.x));
    qx = __fadd_rn(qx, px); qy = __fadd_rn(qy, py);
  }
  (void)qy;
  return qx;
}

__device__ __forceinline__ float db_value(float q, float qmin_global) {
  // out = 1.0/Q (double divide narrowed to float == correctly rounded float divide), out/max, 10*log10  (:140-142)
  const float y = __fdiv_rn(1.0f, q), ymax = __fdiv_rn(1.0f, qmin_global);
  return __fmul_rn(10.0f, log10f(__fdiv_rn(y, ymax)));
}

// Sort K values held by lanes 0..K-1 descending (ties: lower lane first) and return this lane's destination slot.
__device__ __forceinline__ int rank_desc(float v, int K, int lane) {
  int rank = 0;
  for (int r = 0; r < K; ++r) {
    const float o = __shfl_sync(FULL, v, r);
    rank += (o > v || (o == v && r < lane)) ? 1 : 0;
  }
  return rank;
}

// ---------------------------------------------------------------------------------------------------------------
// Packed pair arithmetic: Blackwell's fma.rn.f32x2 does two FMAs per lane per instruction (same FMA-pipe rate as FFMA,
// half the issue slots; measured, tools/microbench/ffma2.cu).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// Two adjacent bins at once.  State (A, B) = (Re acc, -Im acc); with the table holding zx, zy and -zy no negation is needed:
//   A' = A zx + (B zy + ux)        B' = A (-zy) + (B zx - uy)
// Operation for operation this is the scalar Horner of q_coarse (negations are exact), so both give identical bits.
template <int MT>
__device__ __forceinline__ f32x2 q_coarse_pair(const f32x2 (&ux2)[MT], const f32x2 (&muy2)[MT], f32x2 u0_2, f32x2 two2,
                                               f32x2 zx2, f32x2 zy2, f32x2 nzy2) {
  f32x2 A = ux2[MT - 1], B = muy2[MT - 1];
#pragma unroll
  for (int l = MT - 2; l >= 1; --l) {
    const f32x2 nA = fma2(A, zx2, fma2(B, zy2, ux2[l]));
    const f32x2 nB = fma2(A, nzy2, fma2(B, zx2, muy2[l]));
    A = nA; B = nB;
  }
  const f32x2 re = fma2(A, zx2, mul2(B, zy2));
  return fma2(two2, re, u0_2);
}

// Shared-memory table of z = e^{j psi}: lane L owns bins [L*S, (L+1)*S), S even, stored as pairs of adjacent bins:
//   za[L*LA + p*4 + {0,1,2,3}] = zx(2p), zx(2p+1), zy(2p), zy(2p+1)       (LDS.128, lane stride LA = 2S+4 floats)
//   zb[L*LB + p*2 + {0,1}]     = -zy(2p), -zy(2p+1)                       (LDS.64,  lane stride LB = S+2 floats)
// The +16 B / +8 B lane skew makes both loads bank-conflict free.
struct ZTab {
  const float* za; const float* zb; int S, LA, LB;
  __device__ __forceinline__ float2 at(int bin) const {
    const int L = bin / S, k = bin - L * S;
    const float* p = za + L * LA + (k >> 1) * 4 + (k & 1);
    return make_float2(p[0], p[2]);
  }
};

template <int MT, int KL>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_peaks_kernel(const float2* __restrict__ u, const float2* __restrict__ G, const float2* __restrict__ ztab,
                  const float2* __restrict__ Vtab, const float* __restrict__ xaxis, int M, int P, int nframes, int K,
                  float* __restrict__ out_val, float* __restrict__ out_loc, int* __restrict__ out_bin) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int S = 2 * ((P + 63) / 64);                        // bins per lane, even
  ZTab zt; zt.S = S; zt.LA = 2 * S + 4; zt.LB = S + 2;
  float* za = smem; float* zb = smem + 32 * zt.LA;
  zt.za = za; zt.zb = zb;
  float2* us_all = reinterpret_cast<float2*>(zb + 32 * zt.LB);   // [SCAN_WARPS][M] (runtime-M path only)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * S; i += blockDim.x) {
    const float2 z = ztab[min(i, P - 1)];
    const int L = i / S, k = i - L * S;
    za[L * zt.LA + (k >> 1) * 4 + (k & 1)] = z.x;
    za[L * zt.LA + (k >> 1) * 4 + 2 + (k & 1)] = z.y;
    zb[L * zt.LB + (k >> 1) * 2 + (k & 1)] = -z.y;
  }
  __syncthreads();
  const int s0 = lane * S, s1 = min(P, s0 + S);
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
    auto q_at = [&](int bin) -> float { return q_coarse<MT>(uc, us, M, zt.at(bin)); };
    Walker<KL, false> w; w.init(s0 > 0);
    bool exact = (MT == 0);        // generic M: always the exact walker
    if constexpr (MT > 0) {
      // Fast path.  Lane L decides the peaks whose plateau STARTS in [s0, s1) and looks at q[s0-1] .. q[s1]; there is no
      // cross-lane state.  The plateau rule (find_local_max_impl.cc:92-107) is applied inside the lane: a strict descent
      // arms `cand`, equal neighbours keep it armed, the next strict ascent emits it.  Only a plateau that is still
      // unresolved at the lane's right edge (equal values across a lane boundary) needs the stitched walker below.
      bool tie = false;
      w.cand = W_NONE;
      if (s0 < s1) {
        f32x2 ux2[MT], muy2[MT];
#pragma unroll
        for (int l = 0; l < MT; ++l) { ux2[l] = pk2(uc[l].x, uc[l].x); muy2[l] = pk2(-uc[l].y, -uc[l].y); }
        const f32x2 u0_2 = pk2(uc[0].x, uc[0].x), two2 = pk2(2.0f, 2.0f);
        float p1; bool d1 = false; int cand = 0; int k = 0;
        if (s0 > 0) p1 = q_at(s0 - 1);
        else { p1 = q_at(0); k = 1; }
        auto feed = [&](float q, int bin) {   // q = value of `bin`
          const bool up = q > p1, down = q < p1;
          if (up && d1 && cand < s1) { ++w.n_emit; w.list.insert(p1, cand, w.n_emit); }
          d1 = down || (d1 && !up);
          cand = down ? bin : cand;
          p1 = q;
        };
        const int len = s1 - s0;
        if (k == 1) { feed(q_at(1), 1); k = 2; }          // lane 0: bins 0,1 handled, continue pair-aligned
        const float4* pa = reinterpret_cast<const float4*>(za + lane * zt.LA);
        const float2* pb = reinterpret_cast<const float2*>(zb + lane * zt.LB);
        for (; k + 4 <= len; k += 4) {
          const float4 a0 = pa[k >> 1], a1 = pa[(k >> 1) + 1];
          const float2 b0 = pb[k >> 1], b1 = pb[(k >> 1) + 1];
          const f32x2 Q0 = q_coarse_pair<MT>(ux2, muy2, u0_2, two2, pk2(a0.x, a0.y), pk2(a0.z, a0.w), pk2(b0.x, b0.y));
          const f32x2 Q1 = q_coarse_pair<MT>(ux2, muy2, u0_2, two2, pk2(a1.x, a1.y), pk2(a1.z, a1.w), pk2(b1.x, b1.y));
          float q0, q1, q2, q3;
          upk2(Q0, q0, q1); upk2(Q1, q2, q3);
          feed(q0, s0 + k); feed(q1, s0 + k + 1); feed(q2, s0 + k + 2); feed(q3, s0 + k + 3);
        }
        for (; k < len; ++k) feed(q_at(s0 + k), s0 + k);
        if (s1 < P) {
          feed(q_at(s1), s1);                              // right neighbour resolves a peak at bin s1-1
          tie = d1 && cand < s1;                           // plateau runs across the lane boundary
        }
      }
      exact = __any_sync(FULL, tie);
      if (exact) w.init(s0 > 0);
    }
    if (exact && s0 < s1) {
      float prev; int k = 0;
      if (s0 > 0) prev = q_at(s0 - 1);
      else { prev = q_at(0); k = 1; }
      for (; k < s1 - s0; ++k) { const float q = q_at(s0 + k); w.step(prev, q, s0 + k, q_at); prev = q; }
    }
    Merged m = stitch_and_merge<KL, false>(w, K, lane, q_at);
    const int nref = min(K, m.nvalid);
    // Global minimum of the coarse spectrum (it sets the 0 dB level): the deepest local minimum or one of the two end
    // bins, which are never local peaks.  With no local minimum at all (a monotone spectrum) take the exact first arg-min.
    int gbest_bin;
    {
      const float q0 = q_at(0), qe = q_at(P - 1);
      float gv = q0; gbest_bin = 0;
      if (m.nvalid > 0) {
        const float bv = __shfl_sync(FULL, m.val, 0); const int bb = __shfl_sync(FULL, m.bin, 0);
        if (bv < gv) { gv = bv; gbest_bin = bb; }
      } else {
        float lv = INFINITY; int li = 0x7fffffff;
        for (int i = lane; i < P; i += 32) { const float q = q_at(i); if (q < lv) { lv = q; li = i; } }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          const float ov = __shfl_xor_sync(FULL, lv, o); const int oi = __shfl_xor_sync(FULL, li, o);
          if (ov < lv || (ov == lv && oi < li)) { lv = ov; li = oi; }
        }
        gv = lv; gbest_bin = li;
      }
      if (qe < gv) { gv = qe; gbest_bin = P - 1; }
    }
    // the reference's fill-in rule (find_local_max_impl.cc:145-163): global arg-max when no peak exists, otherwise
    // all_pks_sorted_indx(0) -- the POSITION of the best peak in the peak list, used as a bin (reference bug, kept)
    const int pad_bin = (m.nvalid == 0) ? gbest_bin : m.best_ord;

    // Refinement with the reference's arithmetic.  Slot 0 = the global minimum (it sets the 0 dB level), slot 1+r =
    // output entry r.  Eight lanes per slot, bin offset = sub-lane - REFINE_W (sub-lanes > 2W idle); entries that
    // are fill-ins (r >= nref) are evaluated at their single bin only.
    const float2* Gf = G + (size_t)f * M * M;
    float fin_q = 0.f; int fin_bin = 0;     // lane r: refined entry r
    float gmin_q = 0.f; int gmin_bin = 0;
    for (int base = 0; base <= K; base += 4) {
      const int slot = base + (lane >> 3), sub = lane & 7;
      const int entry = slot - 1;
      const int cb = __shfl_sync(FULL, m.bin, max(0, min(entry, 31)));
      int centre = 0; bool refine = false, used = false;
      if (slot == 0) { used = true; refine = true; centre = gbest_bin; }
      else if (entry < K) { used = true; refine = entry < nref; centre = refine ? cb : pad_bin; }
      const int b = centre + (refine ? sub - REFINE_W : 0);
      const bool valid = used && b >= 0 && b < P && (refine ? sub <= 2 * REFINE_W : sub == 0);
      float qf = INFINITY; int qb = 0x7fffffff;
      if (valid) { qf = q_faithful(Gf, Vtab + (size_t)b * M, M); qb = b; }
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, qf, o); const int ob = __shfl_xor_sync(FULL, qb, o);
        if (ov < qf || (ov == qf && ob < qb)) { qf = ov; qb = ob; }
      }
      // hand slot results to their owner lanes
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float sv = __shfl_sync(FULL, qf, g * 8); const int sb = __shfl_sync(FULL, qb, g * 8);
        const int sl = base + g;
        if (sl == 0) { gmin_q = sv; gmin_bin = sb; }
        else if (sl - 1 < K && lane == sl - 1) { fin_q = sv; fin_bin = sb; }
      }
    }
    if (m.nvalid == 0) { fin_q = gmin_q; fin_bin = gmin_bin; }   // no local peak at all: every entry is the arg-max (:149-150)
    fin_bin = min(fin_bin, P - 1);
    {   // the 0 dB level is the smallest refined value anywhere (two nulls of near-equal depth can swap order on refinement)
      float mq = (lane < K) ? fin_q : INFINITY;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) mq = fminf(mq, __shfl_xor_sync(FULL, mq, o));
      gmin_q = fminf(gmin_q, mq);
    }
    float val = (lane < K) ? db_value(fin_q, gmin_q) : -INFINITY;
    // entries 0..nref-1 are real peaks: order them by height like sort_index(..., "descend"); fill-ins stay behind
    int slot = lane;
    {
      const float key = (lane < nref) ? val : -INFINITY;
      const int rk = rank_desc(key, nref, lane);
      if (lane < nref) slot = rk;
    }
    const float loc = (lane < K) ? xaxis[fin_bin] : -INFINITY;
    const int lrank = rank_desc(loc, K, lane);      // sort(x_axis(pk), "descend")  find_local_max_impl.cc:188
    if (lane < K) {
      out_val[(size_t)f * K + slot] = val;
      out_loc[(size_t)f * K + lrank] = loc;
      if (out_bin) out_bin[(size_t)f * K + slot] = fin_bin;
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
    float bv = INFINITY; int bi = 0x7fffffff;
    for (int i = lane; i < P; i += 32) {   // interleaved bins: coalesced table reads, no ordering needed
      const float q = q_coarse<MT>(uc, us, M, ztab[i]);
      if (q < bv) { bv = q; bi = i; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL, bv, o); const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    // refine around the coarse arg-min with the reference arithmetic
    const float2* Gf = G + (size_t)f * M * M;
    const int b = bi + lane - REFINE_W;
    float qf = INFINITY; int qb = 0x7fffffff;
    if (lane <= 2 * REFINE_W && b >= 0 && b < P) { qf = q_faithful(Gf, Vtab + (size_t)b * M, M); qb = b; }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL, qf, o); const int ob = __shfl_xor_sync(FULL, qb, o);
      if (ov < qf || (ov == qf && ob < qb)) { qf = ov; qb = ob; }
    }
    if (lane == 0) {
      out_val[f] = db_value(qf, qf);
      out_loc[f] = xaxis[qb];
      if (out_bin) out_bin[f] = qb;
    }
  }
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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* vs = fsm + (size_t)warp * 32 * SP;
  const int s0 = lane * S, s1 = min(len, s0 + S);
  for (int f = blockIdx.x * FLM_WARPS + warp; f < nframes; f += gridDim.x * FLM_WARPS) {
    const float* src = in + (size_t)f * len;
    __syncwarp();
    for (int i = lane; i < len; i += 32) vs[(i / S) * SP + (i % S)] = src[i];
    __syncwarp();
    const float* vl = vs + lane * SP;
    auto v_at = [&](int b) -> float { return vs[(b / S) * SP + (b % S)]; };
    Walker<KL, true> w; w.init(s0 > 0);
    float bv = -INFINITY; int bi = 0x7fffffff;      // first occurrence of the maximum (index_max semantics)
    if (s0 < s1) {
      float prev; int k = 0;
      if (s0 > 0) prev = vs[(lane - 1) * SP + (S - 1)];
      else { prev = vl[0]; if (prev > bv) { bv = prev; bi = 0; } k = 1; }
      for (; k < s1 - s0; ++k) {
        const float c = vl[k];
        if (c > bv) { bv = c; bi = s0 + k; }
        if (K > 1) w.step(prev, c, s0 + k, v_at);
        prev = c;
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL, bv, o); const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;   // nothing compared greater than -inf (all -inf / NaN): index 0 like index_max
    int bin; float val;
    if (K == 1) {   // index_max (find_local_max_impl.h:53-56)
      bin = bi; val = src[bi];
    } else {
      Merged m = stitch_and_merge<KL, true>(w, K, lane, v_at);
      const int nref = min(K, m.nvalid);
      const int pad = (m.nvalid == 0) ? bi : m.best_ord;     // fill-in rule incl. the reference's index bug (:145-163)
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
  const int S = 2 * ((P + 63) / 64);
  const size_t smem = (size_t)32 * ((2 * S + 4) + (S + 2)) * sizeof(float) + (MT > 0 ? 0 : (size_t)SCAN_WARPS * M) * sizeof(float2);
  if (smem > 200 * 1024) return DOA_CUDA_EINVAL;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem, 1)));
  const int blocks = min(blocks_needed, sm_count() * per_sm);
  if (K <= 4) {
    auto kern = scan_peaks_kernel<MT, 4>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, SCAN_WARPS * 32, smem, st>>>(u, G, tb.z, tb.V, tb.xaxis, M, P, nframes, K, out_val, out_loc, out_bin);
  } else {
    auto kern = scan_peaks_kernel<MT, 16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, SCAN_WARPS * 32, smem, st>>>(u, G, tb.z, tb.V, tb.xaxis, M, P, nframes, K, out_val, out_loc, out_bin);
  }
  return 1;
}

template <int MT>
int launch_spectrum_mt(const float2* u, const ScanTables& tb, int nframes, float* out, cudaStream_t st) {
  const size_t smem = (MT > 0 ? 0 : (size_t)SCAN_WARPS * tb.M) * sizeof(float2);
  const int blocks = min((nframes + SCAN_WARPS - 1) / SCAN_WARPS, sm_count() * 8);
  scan_spectrum_kernel<MT><<<blocks, SCAN_WARPS * 32, smem, st>>>(u, tb.z, tb.M, tb.P, nframes, out);
  return 1;
}

}  // namespace

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
  const int S = (len + 31) / 32, SP = S | 1;
  const size_t smem = (size_t)FLM_WARPS * 32 * SP * sizeof(float);
  if (smem > 200 * 1024) return DOA_CUDA_EINVAL;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
  const int blocks = min((nframes + FLM_WARPS - 1) / FLM_WARPS, sm_count() * per_sm);
  if (K <= 4) {
    auto kern = find_local_max_kernel<4>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, FLM_WARPS * 32, smem, st>>>(in, len, nframes, K, xaxis, out_val, out_loc, out_bin);
  } else {
    auto kern = find_local_max_kernel<16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, FLM_WARPS * 32, smem, st>>>(in, len, nframes, K, xaxis, out_val, out_loc, out_bin);
  }
  return 1;
}

}  // namespace doa
