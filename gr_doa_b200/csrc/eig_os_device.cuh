// eig_os_device.cuh -- one-sided (Hestenes) Jacobi on the Cholesky factor: the eigensolver of the 8- and 16-element paths.
//
// Replaces the same reference lines as eig_device.cuh (eig_sym + U_N U_N^H, gr-doa lib/MUSIC_lin_array_impl.cc:128-133,
// lib/rootMUSIC_linear_array_impl.cc:112-116).  For a positive definite R = L L^H, rotating the COLUMNS of L until they are
// mutually orthogonal, L J_1 J_2 ... = U Sigma, gives R = U Sigma^2 U^H: the eigenvectors are the normalised columns
// themselves and the eigenvalues their squared norms (Veselic & Hari).  Against the two-sided iteration of eig_device.cuh:
//   * no eigenvector accumulator and no row rotations -- a step is one column fetch (2M shuffles), one complex dot product
//     and one column update, all in packed f32x2 FMAs: 169 warp-instructions per step instead of 466 at M = 16;
//   * 6-7 sweeps, like the two-sided iteration;
//   * the Gram matrix being diagonalised is L^H L, whose condition is that of R, not of R^2, and Jacobi on a factor is
//     relatively accurate: the projector is as close to the float64 one as LAPACK's (measured, tests).
// Mapping: lane j of an M-lane group holds column j (M complex numbers as two planes of row PAIRS: wr[h] = rows 2h, 2h+1 of
// the real part), columns never move; the round-robin tournament is the one of eig_device.cuh.  Both lanes of a pair compute
// the pair's dot product from the same products in the same order (the imaginary part as a difference of two separately
// accumulated sums, so that one lane's value is exactly the negative of the other's) and therefore the SAME rotation, with
// no exchange of rotation parameters.  Squared column norms are recomputed at the start of a sweep and carried through it by
// the rotation's own update (app - t|apq|, aqq + t|apq|).
//
// R / 2^e + delta I is what gets factored (2^e the trace's power of two, delta = 2^-14 of the scaled trace: far above the fp32
// rounding of a rank-deficient covariance, so that one still factors; noise eigenvalues below delta merely cluster at delta,
// which does not move the noise SUBSPACE -- tested at 40 dB SNR).  Eigenvectors are unchanged, eigenvalues are reported as
// 2^e (|column|^2 - delta).  A matrix whose factorisation meets a non-positive pivot (not a covariance: indefinite, zero or
// non-finite input) is reported back to the caller with its staging area restored, and the caller runs the two-sided solver on
// it (noise_subspace_solve).
#pragma once
#include "eig_device.cuh"
#include "f32x2.cuh"

namespace doa {
namespace {

__device__ __forceinline__ f32x2 shfl2(f32x2 v, int src, int width) {
  return __shfl_sync(0xffffffffu, v, src, width);
}

// Bare MUFU ops: rsqrtf() / __frcp_rn() wrap the unit in denormal scaling and a Newton step with a slow-path branch (12 extra
// instructions on the rotation's serial chain); the arguments here are normal numbers and 1-ulp results are all the rotation
// needs (see make_rotation).
__device__ __forceinline__ float mufu_rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Rotation for the pair's Gram 2x2 [[app, apq], [conj(apq), aqq]] (the one of make_rotation: t = sgn(d) 2b / (|d| + sqrt(d^2 + 4 b^2)),
// d = aqq - app, b = |apq|; c = 1/sqrt(1 + t^2); sigma = t c apq / b), also returning t b for the norm update and whether the
// pair is still visibly non-orthogonal (b^2 > tau^2 app aqq).  Written for a short dependent chain: three special-function ops
// in series, no division, no clamping -- the caller's matrix is scaled to unit trace, so d^2 + 4 b^2 <= 20.
__device__ __forceinline__ Rot make_rotation_os(float app, float aqq, float2 apq, bool frozen, float& tb, bool& dirty) {
  const float b2 = fmaf(apq.x, apq.x, apq.y * apq.y);
  const bool act = b2 > 1e-30f && !frozen;                   // branch-free: an inactive pair gets the exact identity
  dirty = act && b2 > 1e-10f * (app * aqq);                  // tau = 1e-5, see the stopping rule below
  const float b2s = act ? b2 : 1.0f;
  const float inv_b = mufu_rsq(b2s);
  const float d = aqq - app;
  const float r2 = fmaf(d, d, 4.0f * b2s);
  const float den = fmaf(r2, mufu_rsq(r2), fabsf(d));        // |d| + sqrt(d^2 + 4 b^2) >= 2 b > 0
  const float tbm = (2.0f * b2s) * mufu_rcp(den);            // |t| b
  tb = act ? ((d < 0.0f) ? -tbm : tbm) : 0.0f;
  const float t = tb * inv_b;
  const float c = mufu_rsq(fmaf(t, t, 1.0f));                // t = 0 -> exactly 1
  const float sb = t * c * inv_b;
  Rot r; r.c = c; r.sx = sb * apq.x; r.sy = sb * apq.y;
  return r;
}

// One M x M Hermitian positive definite matrix by M lanes.  Same contract as jacobi_group_solve (S: the matrix's M*M float2
// staging area in shared memory, column-major on entry, upper triangle read, overwritten; outputs may be null, in shared or
// global memory; all lanes of the warp call together).  Returns false -- storing nothing and leaving the upper triangle of
// S as it was -- for a matrix that is not numerically positive definite; the caller then owes it a jacobi_group_solve.
template <int M>
__device__ __forceinline__ bool jacobi_os_solve(float2* S, const int j, const int T, const int max_sweeps, const bool live,
                                                float2* __restrict__ Gdst, float2* __restrict__ udst,
                                                float* __restrict__ wdst) {
  static_assert(M % 2 == 0 && M >= 4 && M <= 32, "lane-group solver: even M up to a warp");
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int H = M / 2, R = M - 1;

  // ---- shift ---------------------------------------------------------------------------------------------------------
  float tr = S[j + j * M].x;
#pragma unroll
  for (int o = M / 2; o >= 1; o >>= 1) tr += __shfl_xor_sync(FULL, tr, o, M);
  bool ok = tr > 0.0f && tr < 3.0e38f;                        // false for NaN as well
  // exact power-of-two scaling to a trace in [1, 2): every later quantity is bounded (column norms^2 <= 2 + delta), the
  // results are those of the unscaled matrix bit for bit
  const int ex = min(max((__float_as_int(tr) >> 23) & 0xff, 1), 253);
  const float scl = __int_as_float((254 - ex) << 23), unscl = __int_as_float(ex << 23);
  const float delta = ok ? (tr * scl) * (1.0f / 16384.0f) : 1.0f;

  // ---- Cholesky, lane j = row j of L.  Slot S[c + j*M] (c <= j) holds A(c, j) = conj(A(j, c)) until this lane replaces it
  // with L(j, c): the lower factor is written row-major over the upper triangle it was computed from, and only lane j ever
  // touches row j's slots before the barrier that publishes them.
  float2 lrow[M], a0[M];
#pragma unroll
  for (int c = 0; c < M; ++c) {
    float2 acc = make_float2(0.0f, 0.0f);
    a0[c] = acc;
    if (c <= j) {
      const float2 a = S[c + j * M];
      a0[c] = a;
      acc = make_float2(c == j ? fmaf(a.x, scl, delta) : a.x * scl, c == j ? 0.0f : -a.y * scl);
    }
#pragma unroll
    for (int k = 0; k < c; ++k) {                             // acc -= L(j, k) conj(L(c, k)); row c is complete up to k < c
      const float2 lc = S[k + c * M];
      acc.x = fmaf(-lrow[k].x, lc.x, acc.x); acc.x = fmaf(-lrow[k].y, lc.y, acc.x);
      acc.y = fmaf(-lrow[k].y, lc.x, acc.y); acc.y = fmaf(lrow[k].x, lc.y, acc.y);
    }
    const float piv = __shfl_sync(FULL, acc.x, c, M);
    const bool good = piv > 0.0f && piv < 3.0e38f;
    ok = ok && good;
    // 1 / sqrt(pivot) from the special-function unit plus one Newton step (relative error ~1e-7) instead of an IEEE square root
    // and division on this serial chain: column c of L is then L(:, c) (1 + e) with |e| ~ 1e-7, a backward error of 2e |A(:, c)|
    // in R -- the size of the rounding of R itself
    const float pv = good ? piv : 1.0f;
    float inv = mufu_rsq(pv);
    inv = inv * fmaf(-0.5f * pv, inv * inv, 1.5f);
    const float d = pv * inv;
    lrow[c] = (c == j) ? make_float2(d, 0.0f) : make_float2(acc.x * inv, acc.y * inv);
    if (c <= j) S[c + j * M] = lrow[c];
    __syncwarp();
  }

  // ---- column j of L in packed row pairs --------------------------------------------------------------------------------
  f32x2 wr[H], wi[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const int i0 = 2 * h, i1 = 2 * h + 1;
    const float2 e0 = (i0 >= j) ? S[j + i0 * M] : make_float2(0.0f, 0.0f);
    const float2 e1 = (i1 >= j) ? S[j + i1 * M] : make_float2(0.0f, 0.0f);
    wr[h] = pk2(e0.x, e1.x); wi[h] = pk2(e0.y, e1.y);
  }
  __syncwarp();
  if (!ok) {                                                  // give the caller its matrix back (ok is the same in all lanes of a group)
#pragma unroll
    for (int c = 0; c < M; ++c) if (c <= j) S[c + j * M] = a0[c];
  }

  // ---- sweeps --------------------------------------------------------------------------------------------------------
  const unsigned gmask = (M == 32) ? FULL : (((1u << M) - 1u) << ((threadIdx.x & 31) / M * M));
  bool done = !ok;                                            // a failed matrix idles (identity rotations) while its warp-mates finish
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (__all_sync(FULL, done)) break;
    float n;
    {
      f32x2 s = pk2(0.0f, 0.0f);
#pragma unroll
      for (int h = 0; h < H; ++h) { s = fma2(wr[h], wr[h], s); s = fma2(wi[h], wi[h], s); }
      float lo, hi; upk2(s, lo, hi); n = lo + hi;
    }
    bool dirty_any = false;
#pragma unroll 1
    for (int st = 0; st < R; ++st) {
      // round-robin tournament on the ring Z_R (index R sits out): in step st ring index r meets (2 st - r) mod R, and the
      // index that would meet itself meets R.  The lower index of a pair plays the a-role (the "p" of the rotation formulas).
      int partner = 2 * st - j;
      partner += (partner < 0) ? R : 0;
      partner -= (partner >= R) ? R : 0;
      partner = (partner == j) ? R : partner;
      partner = (j == R) ? st : partner;
      const bool is_a = j < partner;
      f32x2 xr[H], xi[H];
#pragma unroll
      for (int h = 0; h < H; ++h) { xr[h] = shfl2(wr[h], partner, M); xi[h] = shfl2(wi[h], partner, M); }
      const float np = __shfl_sync(FULL, n, partner, M);
      f32x2 rr = pk2(0.0f, 0.0f), ii = rr, ri = rr, ir = rr;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        rr = fma2(wr[h], xr[h], rr); ii = fma2(wi[h], xi[h], ii);
        ri = fma2(wr[h], xi[h], ri); ir = fma2(wi[h], xr[h], ir);
      }
      float a0, a1, b0, b1, c0, c1, d0, d1;
      upk2(rr, a0, a1); upk2(ii, b0, b1); upk2(ri, c0, c1); upk2(ir, d0, d1);
      const float dx = (a0 + a1) + (b0 + b1);                 // Re w^H x: the same number in both lanes of the pair
      const float dy = (c0 + c1) - (d0 + d1);                 // Im w^H x: exactly minus the partner's
      const float2 apq = make_float2(dx, is_a ? dy : -dy);    // w_p^H w_q with p the a-role column
      float tb; bool dirty;
      const Rot rot = make_rotation_os(is_a ? n : np, is_a ? np : n, apq, done, tb, dirty);
      dirty_any = dirty_any || dirty;
      // col_p' = c col_p - conj(sigma) col_q ; col_q' = sigma col_p + c col_q  ->  w' = c w + g x, g = (-sx, sy) | (sx, sy)
      const float gx = is_a ? -rot.sx : rot.sx;
      const f32x2 c2 = pk2(rot.c, rot.c), gx2 = pk2(gx, gx), gy2 = pk2(rot.sy, rot.sy), ngy2 = pk2(-rot.sy, -rot.sy);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const f32x2 nr = fma2(ngy2, xi[h], fma2(gx2, xr[h], mul2(c2, wr[h])));
        const f32x2 ni = fma2(gy2, xr[h], fma2(gx2, xi[h], mul2(c2, wi[h])));
        wr[h] = nr; wi[h] = ni;
      }
      n = is_a ? fmaxf(n - tb, 0.0f) : n + tb;
    }
    // Stopping rule: a matrix is done after a sweep in which every pair was already orthogonal to tau = 1e-5 BEFORE its
    // rotation.  A laxer tau looks sufficient (a rotation leaves ~tau^2 behind) but is not: the noise eigenvalues are nearly
    // degenerate, so noise-noise pairs rotate by large angles in every sweep, the last one included, and carry a not yet
    // re-orthogonalised overlap of order tau with a signal column into a pair that was already clean (at tau = 3e-4 and
    // 40 dB SNR: projector error 5e-5 in the tail, prototype and GPU agree; at 1e-5: 3e-7, LAPACK 2e-7).
    const unsigned still = __ballot_sync(FULL, dirty_any) & gmask;   // every lane votes: not inside a short-circuit
    done = done || still == 0u;
  }

  // ---- eigenpairs: |column|^2 - delta and the normalised column ---------------------------------------------------------
  float2 v[M];
  float n2;
  {
    f32x2 s = pk2(0.0f, 0.0f);
#pragma unroll
    for (int h = 0; h < H; ++h) { s = fma2(wr[h], wr[h], s); s = fma2(wi[h], wi[h], s); }
    float lo, hi; upk2(s, lo, hi); n2 = lo + hi;
  }
  const float sc = 1.0f / sqrtf(fmaxf(n2, 1e-37f));
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float r0, r1, i0, i1;
    upk2(wr[h], r0, r1); upk2(wi[h], i0, i1);
    v[2 * h] = make_float2(r0 * sc, i0 * sc); v[2 * h + 1] = make_float2(r1 * sc, i1 * sc);
  }
  subspace_outputs<M>(S, j, T, live && ok, v, (n2 - delta) * unscl, Gdst, udst, wdst, ok);
  return ok;
}

// The eigensolver of a lane group as the kernels call it.  sweeps > 0: one-sided solver with that sweep limit (8 and 16
// elements), the two-sided one only for the matrices it rejects; sweeps < 0: two-sided solver with -sweeps (option
// "eig_onesided" = 0, and always below 8 elements).  A rejected matrix is redone alone: its warp-mates' results are already
// stored and do not depend on it.
template <int M>
__device__ __forceinline__ void noise_subspace_solve(float2* S, const int j, const int T, const int sweeps, const bool live,
                                                     float2* __restrict__ Gdst, float2* __restrict__ udst,
                                                     float* __restrict__ wdst) {
  if constexpr (M >= 8) {
    if (sweeps > 0) {
      const bool ok = jacobi_os_solve<M>(S, j, T, sweeps, live, Gdst, udst, wdst);
      if (__any_sync(0xffffffffu, !ok)) {
        __syncwarp();
        jacobi_group_solve<M>(S, j, T, 16, live && !ok, Gdst, udst, wdst);
      }
      return;
    }
  }
  jacobi_group_solve<M>(S, j, T, sweeps < 0 ? -sweeps : sweeps, live, Gdst, udst, wdst);
}

// Host side: the `sweeps` argument above from the handle's options.
inline int eig_sweeps_arg(int M) {
  const bool os = M >= 8 && dev_option(OPT_EIG_ONESIDED, 1) != 0;
  const int s = dev_option(OPT_JACOBI_SWEEPS, os ? 16 : (M <= 8 ? 12 : 16));
  return os ? s : -s;
}

}  // namespace
}  // namespace doa
