// eig_device.cuh -- device code of the batched Jacobi eigensolver shared by eig.cu and fused.cu.
#pragma once
#include "doa_internal.h"

namespace doa {
namespace {

struct Rot { float c; float sx, sy; };   // J_pp = J_qq = c, J_pq = sigma = (sx, sy), J_qp = -conj(sigma)

// Rotation annihilating the (p,q) entry of a Hermitian 2x2 [[app, apq],[conj(apq), aqq]]:
//   zeta = (aqq-app)/(2|apq|), t = sgn(zeta)/(|zeta|+sqrt(zeta^2+1)), c = 1/sqrt(1+t^2), s = t c, sigma = s apq/|apq|.
// Built from MUFU rsqrt/rcp (4 special-function ops) instead of IEEE sqrt/div sequences (~100 instructions): c and s
// share one relative error e, so J = (1+e) * (exact unitary) -- orthogonality of V is untouched, only its column norms
// drift by O(1e-7) per rotation, and the columns are renormalised once at the end.  A t that is 1 ulp off just leaves
// a pivot residue of 1e-7 |apq| for the next sweep.
__device__ __forceinline__ Rot make_rotation(float app, float aqq, float2 apq) {
  Rot r; r.c = 1.0f; r.sx = 0.0f; r.sy = 0.0f;
  const float b2 = fmaf(apq.x, apq.x, apq.y * apq.y);
  if (b2 > 1e-36f) {
    const float inv_b = rsqrtf(b2);
    float zeta = 0.5f * (aqq - app) * inv_b;
    zeta = fminf(fmaxf(zeta, -1e18f), 1e18f);               // keep zeta^2 finite; |t| ~ 1/(2|zeta|) either way
    const float az = fabsf(zeta);
    const float w = fmaf(zeta, zeta, 1.0f);
    float t = __frcp_rn(az + w * rsqrtf(w));                 // sqrt(w) = w * rsqrt(w)
    t = (zeta < 0.0f) ? -t : t;
    const float c = rsqrtf(fmaf(t, t, 1.0f));
    const float sb = t * c * inv_b;
    r.c = c; r.sx = sb * apq.x; r.sy = sb * apq.y;
  }
  return r;
}

template <int M> __host__ __device__ constexpr int pair_a(int s, int k) { return k == 0 ? s : (s + k) % (M - 1); }
template <int M> __host__ __device__ constexpr int pair_b(int s, int k) { return k == 0 ? (M - 1) : (s - k + (M - 1)) % (M - 1); }
template <int M> __host__ __device__ constexpr int pair_p(int s, int k) { return pair_a<M>(s, k) < pair_b<M>(s, k) ? pair_a<M>(s, k) : pair_b<M>(s, k); }
template <int M> __host__ __device__ constexpr int pair_q(int s, int k) { return pair_a<M>(s, k) < pair_b<M>(s, k) ? pair_b<M>(s, k) : pair_a<M>(s, k); }

template <int M, int S>
__device__ __forceinline__ void jacobi_step(float2 (&a)[M], float2 (&v)[M], const int j, const bool frozen) {
  constexpr int HP = M / 2;
  constexpr unsigned FULL = 0xffffffffu;
  // my diagonal entry, my partner, my role and my copy of the pivot
  float dj = 0.0f;
#pragma unroll
  for (int i = 0; i < M; ++i) dj = (i == j) ? a[i].x : dj;
  int partner = 0; bool is_p = false; float2 piv = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < HP; ++k) {
    constexpr int dummy = 0; (void)dummy;
    const int p = pair_p<M>(S, k), q = pair_q<M>(S, k);
    if (j == p) { partner = q; is_p = true; piv = make_float2(a[q].x, -a[q].y); }   // A[p][q] = conj(A[q][p])
    if (j == q) { partner = p; is_p = false; piv = a[p]; }
  }
  const float dpart = __shfl_sync(FULL, dj, partner, M);
  Rot mine = make_rotation(is_p ? dj : dpart, is_p ? dpart : dj, piv);
  if (frozen) { mine.c = 1.0f; mine.sx = 0.0f; mine.sy = 0.0f; }
  // the M/2 rotations of this step, as lane p_k computed them
  float ck[HP], sxk[HP], syk[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) {
    const int p = pair_p<M>(S, k);
    ck[k] = __shfl_sync(FULL, mine.c, p, M);
    sxk[k] = __shfl_sync(FULL, mine.sx, p, M);
    syk[k] = __shfl_sync(FULL, mine.sy, p, M);
  }
  float cm = 1.0f, wx = 0.0f, wy = 0.0f;
#pragma unroll
  for (int k = 0; k < HP; ++k) {
    const int p = pair_p<M>(S, k), q = pair_q<M>(S, k);
    if (j == p) { cm = ck[k]; wx = -sxk[k]; wy = syk[k]; }   // col_p' = c col_p - conj(sigma) col_q
    if (j == q) { cm = ck[k]; wx = sxk[k]; wy = syk[k]; }    // col_q' = sigma col_p + c col_q
  }
  // columns: A <- A J, V <- V J
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const float px = __shfl_sync(FULL, a[i].x, partner, M), py = __shfl_sync(FULL, a[i].y, partner, M);
    const float nx = fmaf(wx, px, fmaf(-wy, py, cm * a[i].x));
    const float ny = fmaf(wx, py, fmaf(wy, px, cm * a[i].y));
    a[i] = make_float2(nx, ny);
    const float qx = __shfl_sync(FULL, v[i].x, partner, M), qy = __shfl_sync(FULL, v[i].y, partner, M);
    const float mx = fmaf(wx, qx, fmaf(-wy, qy, cm * v[i].x));
    const float my = fmaf(wx, qy, fmaf(wy, qx, cm * v[i].y));
    v[i] = make_float2(mx, my);
  }
  // rows: A <- J^H A on my column: row_p' = c row_p - sigma row_q ; row_q' = conj(sigma) row_p + c row_q
#pragma unroll
  for (int k = 0; k < HP; ++k) {
    const int p = pair_p<M>(S, k), q = pair_q<M>(S, k);
    const float2 x = a[p], y = a[q];
    const float c = ck[k], sx = sxk[k], sy = syk[k];
    a[p] = make_float2(fmaf(-sx, y.x, fmaf(sy, y.y, c * x.x)), fmaf(-sx, y.y, fmaf(-sy, y.x, c * x.y)));
    a[q] = make_float2(fmaf(sx, x.x, fmaf(sy, x.y, c * y.x)), fmaf(sx, x.y, fmaf(-sy, x.x, c * y.y)));
  }
}

template <int M, int S>
__device__ __forceinline__ void jacobi_sweep(float2 (&a)[M], float2 (&v)[M], const int j, const bool frozen) {
  if constexpr (S < M - 1) {
    jacobi_step<M, S>(a, v, j, frozen);
    jacobi_sweep<M, S + 1>(a, v, j, frozen);
  }
}


// One M x M Hermitian matrix by M lanes (lane j = column j).  `S` is this matrix's M*M float2 staging area in shared
// memory; on entry it holds the input column-major (only the upper triangle is used, like cheevd 'U'), it is
// overwritten.  Outputs (any may be null): G column-major M x M, u[M] diagonal sums, w[M] eigenvalues ascending; they
// may live in shared or global memory.  `live` = false suppresses all stores (padding groups of a partial warp).
// All M lanes of every group in the warp must call this together (full-warp shuffles inside).
template <int M>
__device__ __forceinline__ void jacobi_group_solve(float2* S, const int j, const int T, const int max_sweeps, const bool live,
                                                   float2* __restrict__ Gdst, float2* __restrict__ udst,
                                                   float* __restrict__ wdst) {
  constexpr unsigned FULL = 0xffffffffu;
  float2 a[M], v[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    float2 e;
    if (i < j) e = S[i + j * M];
    else if (i == j) e = make_float2(S[i + j * M].x, 0.0f);
    else { const float2 t = S[j + i * M]; e = make_float2(t.x, -t.y); }
    a[i] = e;
    v[i] = make_float2(i == j ? 1.0f : 0.0f, 0.0f);
  }
  __syncwarp();

  // Convergence is decided PER MATRIX and latched: a converged matrix only sees identity rotations (exact no-ops) while
  // its warp-mates finish, so a frame's result never depends on which other frames share its warp.
  bool done = false;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    float off = 0.0f, dg = 0.0f;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const float m2 = a[i].x * a[i].x + a[i].y * a[i].y;
      if (i == j) dg += m2; else off += m2;
    }
#pragma unroll
    for (int o = M / 2; o >= 1; o >>= 1) {
      off += __shfl_xor_sync(FULL, off, o, M);
      dg += __shfl_xor_sync(FULL, dg, o, M);
    }
    // fp32 rotations leave off-diagonal mass of order M^2 * eps^2 * dg; once within ~4x of that floor the next sweep
    // (quadratic convergence) cannot improve the subspace any further
    done = done || (off <= dg * (1.5e-14f * M * M));
    if (__all_sync(FULL, done)) break;
    jacobi_sweep<M, 0>(a, v, j, done);
  }

  // undo the accumulated norm drift of the fast rotations: unit eigenvectors
  {
    float n2 = 0.0f;
#pragma unroll
    for (int i = 0; i < M; ++i) n2 = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, n2));
    const float sc = 1.0f / sqrtf(n2);
#pragma unroll
    for (int i = 0; i < M; ++i) { v[i].x *= sc; v[i].y *= sc; }
  }
  // eigenvalue of this lane's column, its ascending rank (ties by column index)
  float lam = 0.0f;
#pragma unroll
  for (int i = 0; i < M; ++i) lam = (i == j) ? a[i].x : lam;
  int rank = 0;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const float li = __shfl_sync(FULL, lam, i, M);
    rank += (li < lam || (li == lam && i < j)) ? 1 : 0;
  }
  const int nn = M - T;
  const bool noise = rank < nn;
  if (wdst != nullptr && live) wdst[rank] = lam;

  if (udst != nullptr) {
    // u_l = sum_{noise n} sum_r e_n[r] conj(e_n[r+l])
    float ux[M], uy[M];
#pragma unroll
    for (int l = 0; l < M; ++l) {
      float sx = 0.0f, sy = 0.0f;
#pragma unroll
      for (int r = 0; r + l < M; ++r) {
        sx = fmaf(v[r].x, v[r + l].x, sx); sx = fmaf(v[r].y, v[r + l].y, sx);
        sy = fmaf(v[r].y, v[r + l].x, sy); sy = fmaf(-v[r].x, v[r + l].y, sy);
      }
      ux[l] = noise ? sx : 0.0f; uy[l] = noise ? sy : 0.0f;
    }
#pragma unroll
    for (int o = M / 2; o >= 1; o >>= 1)
#pragma unroll
      for (int l = 0; l < M; ++l) {
        ux[l] += __shfl_xor_sync(FULL, ux[l], o, M);
        uy[l] += __shfl_xor_sync(FULL, uy[l], o, M);
      }
    float2 mine = make_float2(0.f, 0.f);
#pragma unroll
    for (int l = 0; l < M; ++l) if (l == j) mine = make_float2(ux[l], l == 0 ? 0.0f : uy[l]);
    if (live) udst[j] = mine;
  }

  if (Gdst != nullptr) {
    // eigenvectors to shared memory in ascending-eigenvalue order, then G(:, j) = sum_{n<nn} E(:, n) conj(E(j, n))
#pragma unroll
    for (int i = 0; i < M; ++i) S[i + rank * M] = v[i];
    __syncwarp();
    float2 gc[M];
#pragma unroll
    for (int i = 0; i < M; ++i) gc[i] = make_float2(0.f, 0.f);
    for (int n = 0; n < nn; ++n) {
      const float2 ej = S[j + n * M];
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const float2 ei = S[i + n * M];
        gc[i].x = fmaf(ei.x, ej.x, gc[i].x); gc[i].x = fmaf(ei.y, ej.y, gc[i].x);
        gc[i].y = fmaf(ei.y, ej.x, gc[i].y); gc[i].y = fmaf(-ei.x, ej.y, gc[i].y);
      }
    }
    if (live) {
      float2* dst = Gdst + (size_t)j * M;
#pragma unroll
      for (int i = 0; i < M; ++i) dst[i] = gc[i];
    }
  }
}

}  // namespace
}  // namespace doa
