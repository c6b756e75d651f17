// scan_device.cuh -- device code of the pseudo-spectrum scan + peak picking shared by scan.cu and fused.cu.
#pragma once
#include "doa_internal.h"
#include "f32x2.cuh"
#include <cfloat>

namespace doa {
namespace {

constexpr unsigned FULL = 0xffffffffu;
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

// The same arithmetic with compile-time M: the steering row is loaded once into registers and the loops are unrolled, so the
// M columns' row sums are independent instruction streams; GS = the projector is in shared memory (explicit LDS).  Operation
// for operation q_faithful: identical bits.
template <int MT, bool GS>
__device__ __forceinline__ float q_faithful_t(const float2* __restrict__ G, const float2* __restrict__ v, int M) {
  if constexpr (MT == 0) {
    return q_faithful(G, v, M);
  } else {
    float2 vr[MT];
#pragma unroll
    for (int r = 0; r < MT; ++r) vr[r] = v[r];
    const unsigned gs = GS ? (unsigned)__cvta_generic_to_shared(G) : 0u;
    float qx = 0.0f;
#pragma unroll
    for (int c = 0; c < MT; ++c) {
      float rx = 0.0f, ry = 0.0f;
#pragma unroll
      for (int r = 0; r < MT; ++r) {
        float2 g;
        if constexpr (GS) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(g.x), "=f"(g.y) : "r"(gs + (unsigned)(c * MT + r) * 8u));
        else g = G[c * MT + r];
        const float px = __fsub_rn(__fmul_rn(vr[r].x, g.x), __fmul_rn(-vr[r].y, g.y));
        const float py = __fadd_rn(__fmul_rn(vr[r].x, g.y), __fmul_rn(-vr[r].y, g.x));
        rx = __fadd_rn(rx, px); ry = __fadd_rn(ry, py);
      }
      const float px = __fsub_rn(__fmul_rn(rx, vr[c].x), __fmul_rn(ry, vr[c].y));
      qx = __fadd_rn(qx, px);
    }
    return qx;
  }
}

__device__ __forceinline__ float db_value(float q, float qmin_global) {
  // out = 1.0/Q (double divide narrowed to float == correctly rounded float divide), out/max, 10*log10  (:140-142)
  const float y = __fdiv_rn(1.0f, q), ymax = __fdiv_rn(1.0f, qmin_global);
  return __fmul_rn(10.0f, log10f(__fdiv_rn(y, ymax)));
}

// Sort K values held by lanes 0..K-1 descending (ties: lower lane first) and return this lane's destination slot.  A total
// order also with NaNs among the values (they go last, by lane), so that the K destinations are always a permutation of
// 0..K-1 and every output slot is written: a peak value is NaN when the null spectrum rounds to a non-positive number at
// that bin (num_targets = num_ant_ele - 1 puts exact zeros of Q on the grid's doorstep), exactly as the reference's
// 10*log10((1/Q)/max) is (MUSIC_lin_array_impl.cc:140-142).
__device__ __forceinline__ int rank_desc(float v, int K, int lane) {
  int rank = 0;
  const bool vn = v != v;
  for (int r = 0; r < K; ++r) {
    const float o = __shfl_sync(FULL, v, r);
    const bool before = (o != o) ? (vn && r < lane) : (vn || o > v || (o == v && r < lane));
    rank += before ? 1 : 0;
  }
  return rank;
}

// ---------------------------------------------------------------------------------------------------------------
// Packed pair arithmetic (f32x2.cuh).
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


// The z table in ZTab layout is built once per plan on the host (build_zpair_table) and lives in global memory; kernels
// copy it linearly into shared memory when it fits and otherwise read it in place (any P, just slower).
__host__ __device__ inline size_t ztab_floats(int P) { const int S = 2 * ((P + 63) / 64); return (size_t)32 * ((2 * S + 4) + (S + 2)); }
__host__ __device__ inline ZTab ztab_view(const float* base, int P) {
  const int S = 2 * ((P + 63) / 64);                        // bins per lane, even
  ZTab zt; zt.S = S; zt.LA = 2 * S + 4; zt.LB = S + 2;
  zt.za = base; zt.zb = base + 32 * zt.LA;
  return zt;
}
// All threads of the CTA; caller issues __syncthreads() afterwards.
__device__ __forceinline__ ZTab ztab_fill(float* smem, const float* __restrict__ zpair, int P) {
  const int n = (int)ztab_floats(P);
  for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = zpair[i];
  return ztab_view(smem, P);
}

// K == 1, second half: refine around the coarse arg-min `bi` with the reference arithmetic and write the outputs.
template <int MT = 0, bool GS = false>
__device__ __forceinline__ void argmax_refine_emit(int bi, const float2* __restrict__ Gf, const float2* __restrict__ Vtab,
                                                   const float* __restrict__ xaxis, int M, int P, int lane,
                                                   float* __restrict__ o_val, float* __restrict__ o_loc, int* __restrict__ o_bin) {
  constexpr unsigned FULLM = 0xffffffffu;
  const int b = bi + lane - REFINE_W;
  float qf = INFINITY; int qb = 0x7fffffff;
  if (lane <= 2 * REFINE_W && b >= 0 && b < P) { qf = q_faithful_t<MT, GS>(Gf, Vtab + (size_t)b * M, M); qb = b; }
#pragma unroll
  for (int o = 4; o >= 1; o >>= 1) {
    const float ov = __shfl_xor_sync(FULLM, qf, o); const int ob = __shfl_xor_sync(FULLM, qb, o);
    if (ov < qf || (ov == qf && ob < qb)) { qf = ov; qb = ob; }
  }
  if (lane == 0) {
    o_val[0] = db_value(qf, qf);
    o_loc[0] = xaxis[qb];
    if (o_bin) o_bin[0] = qb;
  }
}

// K == 1 (index_max, find_local_max_impl.h:53-56): the global arg-max of one frame by one warp -- coarse arg-min of Q over
// interleaved bins, refinement around it with the reference arithmetic, outputs.  ztab: plain z[P] table (global memory).
template <int MT>
__device__ __forceinline__ void scan_frame_argmax(const float2* __restrict__ uf, const float2* __restrict__ Gf,
                                                  const float2* __restrict__ ztab, float2* us, const float2* __restrict__ Vtab,
                                                  const float* __restrict__ xaxis, int M, int P, int lane,
                                                  float* __restrict__ o_val, float* __restrict__ o_loc, int* __restrict__ o_bin) {
  constexpr unsigned FULLM = 0xffffffffu;
  float2 uc[MT > 0 ? MT : 1];
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
    const float ov = __shfl_xor_sync(FULLM, bv, o); const int oi = __shfl_xor_sync(FULLM, bi, o);
    if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  argmax_refine_emit<MT, false>(bi, Gf, Vtab, xaxis, M, P, lane, o_val, o_loc, o_bin);
}

// Second half of the peak search, shared by the Horner scan below and the tensor-core scan (scan_tc.cu): from the merged
// coarse candidates (lane r < K holds entry r) to the K outputs.  q0 / qe: coarse values of the two end bins (never local
// peaks, but either may hold the global minimum); q_at(bin): coarse value of any bin, used only when the frame has no local
// minimum at all.
// How a refinement value is obtained: every lane of the warp calls the evaluator once per round of the refinement loop
// (K/4 + 1 rounds), convergently, with its own (valid, bin).  DirectEval: the lane evaluates v^H G v itself.  The wide kernel's
// evaluator (scan.cu: CoopEval) posts the warp's 32 requests to shared memory and lets the whole CTA evaluate them.
template <int MT, bool GS>
struct DirectEval {
  const float2* Gf; const float2* Vtab; int M;
  __device__ __forceinline__ float operator()(bool valid, int b, int) const {
    return valid ? q_faithful_t<MT, GS>(Gf, Vtab + (size_t)b * M, M) : INFINITY;
  }
};

template <int MT = 0, bool GS = false, typename QAT, typename EV>
__device__ __forceinline__ void peaks_refine_emit(const Merged& m, float q0, float qe, QAT&& q_at, const EV& ev,
                                                  const float* __restrict__ xaxis, int M, int P,
                                                  int K, int lane, float* __restrict__ o_val, float* __restrict__ o_loc,
                                                  int* __restrict__ o_bin) {
  const int nref = min(K, m.nvalid);
  // Global minimum of the coarse spectrum (it sets the 0 dB level): the deepest local minimum or one of the two end
  // bins, which are never local peaks.  With no local minimum at all (a monotone spectrum) take the exact first arg-min.
  int gbest_bin;
  {
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
    const float qf0 = ev(valid, b, lane);
    float qf = valid ? qf0 : INFINITY; int qb = valid ? b : 0x7fffffff;
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
    // ... among the POSITIVE ones: the level is max(1/Q), and 1/Q of a value that rounded to below zero is negative
    float mq = (lane < K && fin_q > 0.f) ? fin_q : INFINITY;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mq = fminf(mq, __shfl_xor_sync(FULL, mq, o));
    gmin_q = fminf(gmin_q > 0.f ? gmin_q : INFINITY, mq);
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
    o_val[slot] = val;
    o_loc[lrank] = loc;
    if (o_bin) o_bin[slot] = fin_bin;
  }
}

// One frame by one warp: coarse scan over all P bins, peak picking, refinement with the reference arithmetic, dB
// conversion, sorted outputs.  uf: the frame's M diagonal sums, Gf: its M x M projector (global or shared memory);
// us: M float2 of per-warp shared scratch (runtime-M path only); o_*: this frame's K output slots.
// ZS: the z table is known to live in shared memory (fused kernel): the hot loop then uses LDS instead of generic loads.
template <int MT, int KL, bool ZS = false, typename EV>
__device__ __forceinline__ void scan_frame_peaks_ev(const float2* __restrict__ uf, const EV& ev, const ZTab& zt,
                                                    float2* us,
                                                    const float* __restrict__ xaxis, int M, int P, int K, int lane,
                                                    float* __restrict__ o_val, float* __restrict__ o_loc,
                                                    int* __restrict__ o_bin, const float* qtab = nullptr) {
  const int S = zt.S;
  const float* za = zt.za; const float* zb = zt.zb;
  const int s0 = lane * S, s1 = min(P, s0 + S);
  float2 uc[MT > 0 ? MT : 1];
  if constexpr (MT > 0) {
#pragma unroll
    for (int l = 0; l < MT; ++l) uc[l] = uf[l];
  } else {
    __syncwarp();
    for (int l = lane; l < M; l += 32) us[l] = uf[l];
    __syncwarp();
  }
  // qtab (generic M only): the frame's coarse spectrum already evaluated by the whole CTA (scan_peaks_wide_kernel)
  auto q_at = [&](int bin) -> float {
    if constexpr (MT == 0) { if (qtab != nullptr) return qtab[bin]; }
    return q_coarse<MT>(uc, us, M, zt.at(bin));
  };
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
      // Four bins at a time.  Inside [s0, s1) an armed candidate is always a bin of this lane (cand < s1), so an emission in
      // the group is  up_i && d1_i  for some i; Q is a degree-(M-1) trigonometric polynomial with at most M-1 minima per
      // frame, so that is rare: the common path only advances the detector state, the rare one replays the four steps.
      auto feed4 = [&](float q0, float q1, float q2, float q3, int bin0) {
        const bool up0 = q0 > p1, dn0 = q0 < p1, up1 = q1 > q0, dn1 = q1 < q0, up2 = q2 > q1, dn2 = q2 < q1, up3 = q3 > q2, dn3 = q3 < q2;
        const bool da = dn0 || (d1 && !up0), db = dn1 || (da && !up1), dc = dn2 || (db && !up2), dd = dn3 || (dc && !up3);
        if ((up0 && d1) || (up1 && da) || (up2 && db) || (up3 && dc)) {
          feed(q0, bin0); feed(q1, bin0 + 1); feed(q2, bin0 + 2); feed(q3, bin0 + 3);
        } else {
          cand = dn3 ? bin0 + 3 : dn2 ? bin0 + 2 : dn1 ? bin0 + 1 : dn0 ? bin0 : cand;
          d1 = dd; p1 = q3;
        }
      };
      const int len = s1 - s0;
      if (k == 1) { feed(q_at(1), 1); k = 2; }          // lane 0: bins 0,1 handled, continue pair-aligned
      const float4* pa = reinterpret_cast<const float4*>(za + lane * zt.LA);
      const float2* pb = reinterpret_cast<const float2*>(zb + lane * zt.LB);
      const unsigned pa_s = ZS ? (unsigned)__cvta_generic_to_shared(pa) : 0u, pb_s = ZS ? (unsigned)__cvta_generic_to_shared(pb) : 0u;
      for (; k + 4 <= len; k += 4) {
        float4 a0, a1; float2 b0, b1;
        if constexpr (ZS) {
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a0.x), "=f"(a0.y), "=f"(a0.z), "=f"(a0.w) : "r"(pa_s + (unsigned)(k >> 1) * 16u));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(a1.x), "=f"(a1.y), "=f"(a1.z), "=f"(a1.w) : "r"(pa_s + (unsigned)(k >> 1) * 16u));
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(b0.x), "=f"(b0.y) : "r"(pb_s + (unsigned)(k >> 1) * 8u));
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+8];" : "=f"(b1.x), "=f"(b1.y) : "r"(pb_s + (unsigned)(k >> 1) * 8u));
        } else {
          a0 = pa[k >> 1]; a1 = pa[(k >> 1) + 1];
          b0 = pb[k >> 1]; b1 = pb[(k >> 1) + 1];
        }
        const f32x2 Q0 = q_coarse_pair<MT>(ux2, muy2, u0_2, two2, pk2(a0.x, a0.y), pk2(a0.z, a0.w), pk2(b0.x, b0.y));
        const f32x2 Q1 = q_coarse_pair<MT>(ux2, muy2, u0_2, two2, pk2(a1.x, a1.y), pk2(a1.z, a1.w), pk2(b1.x, b1.y));
        float q0, q1, q2, q3;
        upk2(Q0, q0, q1); upk2(Q1, q2, q3);
        feed4(q0, q1, q2, q3, s0 + k);
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
  // compile-time M: steering rows in registers, unrolled row sums; ZS (fused kernel): the projector is in shared memory too
  peaks_refine_emit<MT, ZS>(m, q_at(0), q_at(P - 1), q_at, ev, xaxis, M, P, K, lane, o_val, o_loc, o_bin);
}

template <int MT, int KL, bool ZS = false>
__device__ __forceinline__ void scan_frame_peaks(const float2* __restrict__ uf, const float2* __restrict__ Gf, const ZTab& zt,
                                                 float2* us, const float2* __restrict__ Vtab,
                                                 const float* __restrict__ xaxis, int M, int P, int K, int lane,
                                                 float* __restrict__ o_val, float* __restrict__ o_loc,
                                                 int* __restrict__ o_bin, const float* qtab = nullptr) {
  const DirectEval<MT, ZS> ev{Gf, Vtab, M};
  scan_frame_peaks_ev<MT, KL, ZS>(uf, ev, zt, us, xaxis, M, P, K, lane, o_val, o_loc, o_bin, qtab);
}

}  // namespace
}  // namespace doa
