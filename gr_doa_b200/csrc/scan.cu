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
template <int KL, bool MAXIMA>
struct Walker {
  CandList<KL, MAXIMA> list;
  bool pending, has_strict, first_away;
  int cand_idx; float cand_val;
  int n_emit;          // candidates this lane emitted (local ordinals start at 1; 0 is reserved for a stitched one)
  float best_val; int best_idx;   // first occurrence of the extreme over the segment
  __device__ __forceinline__ void init() {
    list.init(); pending = false; has_strict = false; first_away = false; cand_idx = 0; cand_val = 0.f; n_emit = 0;
    best_val = MAXIMA ? -INFINITY : INFINITY; best_idx = 0x7fffffff;
  }
  __device__ __forceinline__ void see(float cur, int i) {   // extreme tracking (every bin of the segment)
    if (CandList<KL, MAXIMA>::better(cur, best_val)) { best_val = cur; best_idx = i; }
  }
  __device__ __forceinline__ void step(float prev, float cur, int i) {   // move from bin i-1 to bin i
    if (CandList<KL, MAXIMA>::better(cur, prev)) {
      if (!has_strict) { has_strict = true; first_away = false; }
      pending = true; cand_idx = i; cand_val = cur;
    } else if (CandList<KL, MAXIMA>::better(prev, cur)) {
      if (!has_strict) { has_strict = true; first_away = true; }
      if (pending) { ++n_emit; list.insert(cand_val, cand_idx, n_emit); pending = false; }
    }
  }
};

// Result of the cross-lane merge, distributed: lane r < K holds final entry r.
struct Merged {
  float val; int bin;   // this lane's final entry (lane < K)
  int nvalid;           // number of peaks found in the whole vector
  int pad_bin;          // reference's fill-in bin when nvalid < K
  float gbest_val; int gbest_bin;   // global extreme, first occurrence
};

template <int KL, bool MAXIMA>
__device__ __forceinline__ Merged stitch_and_merge(Walker<KL, MAXIMA>& w, int K, int lane) {
  // 1. stitch the lane boundaries: incoming state = state of the nearest lower lane that saw a strict move
  const unsigned strict_mask = __ballot_sync(FULL, w.has_strict);
  const unsigned below = strict_mask & ((1u << lane) - 1u);
  const int src = below ? (31 - __clz(below)) : 0;
  const bool in_pending = __shfl_sync(FULL, (int)w.pending, src) != 0 && below != 0;
  const int in_idx = __shfl_sync(FULL, w.cand_idx, src);
  const float in_val = __shfl_sync(FULL, w.cand_val, src);
  const bool stitched = w.has_strict && w.first_away && in_pending;
  if (stitched) w.list.insert(in_val, in_idx, 0);
  const int my_count = w.n_emit + (stitched ? 1 : 0);
  // 2. ordinals: exclusive prefix of counts over lanes
  int incl = my_count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
  const int excl = incl - my_count;
  const int nvalid = __shfl_sync(FULL, incl, 31);
  const int ord_shift = excl - (stitched ? 0 : 1);   // global ordinal = local ordinal + ord_shift
  // 3. global extreme (first occurrence)
  float gv = w.best_val; int gi = w.best_idx;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float ov = __shfl_xor_sync(FULL, gv, o); const int oi = __shfl_xor_sync(FULL, gi, o);
    if (CandList<KL, MAXIMA>::better(ov, gv) || (ov == gv && oi < gi)) { gv = ov; gi = oi; }
  }
  // 4. K rounds of arg-best over the list heads
  Merged m; m.val = 0.f; m.bin = 0; m.nvalid = nvalid; m.gbest_val = gv; m.gbest_bin = gi; m.pad_bin = 0;
  int best_ord = 0;
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
    if (r == 0) best_ord = ho;
  }
  // 5. the reference's fill-in rule (find_local_max_impl.cc:145-163): global arg-max when no peak exists, otherwise
  //    all_pks_sorted_indx(0) -- the POSITION of the best peak in the peak list, used as a bin (reference bug, kept).
  m.pad_bin = (nvalid == 0) ? gi : best_ord;
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
template <int MT, int KL>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_peaks_kernel(const float2* __restrict__ u, const float2* __restrict__ G, const float2* __restrict__ ztab,
                  const float2* __restrict__ Vtab, const float* __restrict__ xaxis, int M, int P, int nframes, int K,
                  float* __restrict__ out_val, float* __restrict__ out_loc, int* __restrict__ out_bin) {
  extern __shared__ float2 smem[];
  const int S = (P + 31) / 32, SP = S | 1;
  float2* zs = smem;                                        // [32][SP] padded per-lane segments of z
  float2* us_all = smem + 32 * SP;                          // [SCAN_WARPS][M] (runtime-M path only)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < P; i += blockDim.x) zs[(i / S) * SP + (i % S)] = ztab[i];
  __syncthreads();
  const int s0 = lane * S, s1 = min(P, s0 + S);
  const float2* zl = zs + lane * SP;
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
    Walker<KL, false> w; w.init();
    if (s0 < s1) {
      float prev;
      if (s0 > 0) prev = q_coarse<MT>(uc, us, M, zs[(lane - 1) * SP + (S - 1)]);   // bin s0-1 (previous lane's last)
      int k = 0;
      if (s0 == 0) { prev = q_coarse<MT>(uc, us, M, zl[0]); w.see(prev, 0); k = 1; }
      const int len = s1 - s0;
      for (; k + 4 <= len; k += 4) {
        float q[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = q_coarse<MT>(uc, us, M, zl[k + e]);
#pragma unroll
        for (int e = 0; e < 4; ++e) { w.see(q[e], s0 + k + e); w.step(prev, q[e], s0 + k + e); prev = q[e]; }
      }
      for (; k < len; ++k) {
        const float q = q_coarse<MT>(uc, us, M, zl[k]);
        w.see(q, s0 + k); w.step(prev, q, s0 + k); prev = q;
      }
    }
    Merged m = stitch_and_merge<KL, false>(w, K, lane);
    const int nref = min(K, m.nvalid);

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
      if (slot == 0) { used = true; refine = true; centre = m.gbest_bin; }
      else if (entry < K) { used = true; refine = entry < nref; centre = refine ? cb : m.pad_bin; }
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
    Walker<KL, true> w; w.init();
    if (s0 < s1) {
      float prev; int k = 0;
      if (s0 > 0) prev = vs[(lane - 1) * SP + (S - 1)];
      else { prev = vl[0]; w.see(prev, 0); k = 1; }
      for (; k < s1 - s0; ++k) { const float c = vl[k]; w.see(c, s0 + k); w.step(prev, c, s0 + k); prev = c; }
    }
    int bin; float val;
    if (K == 1) {   // index_max
      float gv = w.best_val; int gi = w.best_idx;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, gv, o); const int oi = __shfl_xor_sync(FULL, gi, o);
        if (ov > gv || (ov == gv && oi < gi)) { gv = ov; gi = oi; }
      }
      if (gi == 0x7fffffff) gi = 0;   // nothing compared greater than -inf (all -inf / NaN): index 0 like index_max
      bin = gi; val = src[gi];
    } else {
      Merged m = stitch_and_merge<KL, true>(w, K, lane);
      const int nref = min(K, m.nvalid);
      int pad = m.pad_bin; if (pad == 0x7fffffff) pad = 0;
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
  const int S = (P + 31) / 32, SP = S | 1;
  const size_t smem = ((size_t)32 * SP + (MT > 0 ? 0 : (size_t)SCAN_WARPS * M)) * sizeof(float2);
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
