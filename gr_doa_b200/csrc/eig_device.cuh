// eig_device.cuh -- device code of the batched Jacobi eigensolver shared by eig.cu and fused.cu.
#pragma once
#include "doa_internal.h"
#include "f32x2.cuh"

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

// One step of the round-robin tournament on the ring Z_R, R = M - 1 (index R sits out): in step S ring index r meets
// (2S - r) mod R and S itself meets R.  Pair k of step S is (a_k, b_k) = ((S + k) mod R, (S - k) mod R), pair 0 is (S, R).
// To keep ONE step body for all S (a fully unrolled tournament is 48 KB of code at M = 8 and 166 KB at M = 16, and the
// eigensolver was instruction-fetch bound: `no_instruction` was its second largest stall), every lane stores the rows of
// its column in a frame that rotates with S: original row i < R lives in register position (i - S) mod R, row R in
// position R.  In that frame the row pairs are always (0, R) and (k, R - k), compile-time register indices; partners,
// roles and source lanes are run-time arithmetic on (j, S); after each step the R ring rows move down one position
// (register moves), and after a whole sweep (R steps) the frame is back where it started.  Columns never move between
// lanes, so the shuffle count is unchanged.
template <int M>
__device__ __forceinline__ void jacobi_step(float2 (&a)[M], float2 (&v)[M], const int j, const int S, const bool frozen) {
  constexpr int HP = M / 2, R = M - 1;
  constexpr unsigned FULL = 0xffffffffu;
  // my diagonal's position, my partner (lane = original column index), my role (a-role = the "p" of the rotation formulas)
  // and the position of the pivot A[p][q] or its conjugate in my column
  int jr, partner, pp; bool is_a;
  if (j == R) { jr = R; partner = S; pp = 0; is_a = false; }
  else {
    jr = j - S; jr += (jr < 0) ? R : 0;
    if (jr == 0) { partner = R; pp = R; is_a = true; }
    else {
      pp = R - jr;
      is_a = jr < HP;
      partner = is_a ? (S - jr) : (S + pp);                 // b_k = S - k (k = jr)   |   a_k = S + k (k = R - jr)
      partner += (partner < 0) ? R : 0;
      partner -= (partner >= R) ? R : 0;
    }
  }
  float dj = 0.0f; float2 piv = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < M; ++i) {
    dj = (i == jr) ? a[i].x : dj;
    piv = (i == pp) ? a[i] : piv;
  }
  if (is_a) piv.y = -piv.y;                                  // column p holds A[q][p] = conj(A[p][q])
  const float dpart = __shfl_sync(FULL, dj, partner, M);
  Rot mine = make_rotation(is_a ? dj : dpart, is_a ? dpart : dj, piv);
  if (frozen) { mine.c = 1.0f; mine.sx = 0.0f; mine.sy = 0.0f; }
  // the M/2 rotations of this step, as the a-role lane of each pair computed them
  float ck[HP], sxk[HP], syk[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) {
    int ak = S + k; ak -= (ak >= R) ? R : 0;
    ck[k] = __shfl_sync(FULL, mine.c, ak, M);
    sxk[k] = __shfl_sync(FULL, mine.sx, ak, M);
    syk[k] = __shfl_sync(FULL, mine.sy, ak, M);
  }
  // my pair's rotation: col_p' = c col_p - conj(sigma) col_q ; col_q' = sigma col_p + c col_q
  const int asrc = is_a ? j : partner;
  const float cm = __shfl_sync(FULL, mine.c, asrc, M);
  const float sxm = __shfl_sync(FULL, mine.sx, asrc, M), wy = __shfl_sync(FULL, mine.sy, asrc, M);
  const float wx = is_a ? -sxm : sxm;
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
    constexpr int dummy = 0; (void)dummy;
    const int ip = k, iq = R - k;                            // register positions of rows a_k, b_k (pair 0: S and R)
    const float2 x = a[ip], y = a[iq];
    const float c = ck[k], sx = sxk[k], sy = syk[k];
    a[ip] = make_float2(fmaf(-sx, y.x, fmaf(sy, y.y, c * x.x)), fmaf(-sx, y.y, fmaf(-sy, y.x, c * x.y)));
    a[iq] = make_float2(fmaf(sx, x.x, fmaf(sy, x.y, c * y.x)), fmaf(sx, x.y, fmaf(-sy, x.x, c * y.y)));
  }
  // next step's frame: ring rows move down one position
  if constexpr (R > 1) {
    const float2 t0 = a[0];
#pragma unroll
    for (int i = 0; i + 1 < R; ++i) a[i] = a[i + 1];
    a[R - 1] = t0;
  }
}

template <int M>
__device__ __forceinline__ void jacobi_sweep(float2 (&a)[M], float2 (&v)[M], const int j, const bool frozen) {
#pragma unroll 1
  for (int S = 0; S < M - 1; ++S) jacobi_step<M>(a, v, j, S, frozen);
}


// Shared tail of the eigensolvers: lane j holds the unit eigenvector v of eigenvalue lam (any column order).  Ranks the
// eigenvalues ascending (ties by column index), then writes (any may be null) the eigenvalues, the diagonal sums u and the
// noise projector G = sum over the M - T smallest eigenvalues of e e^H.  `S` (M*M float2 of shared memory) is overwritten
// unless use_S is false (a matrix whose results are not wanted: live must be false then).
template <int M>
__device__ __forceinline__ void subspace_outputs(float2* S, const int j, const int T, const bool live, const float2 (&v)[M],
                                                 const float lam, float2* __restrict__ Gdst, float2* __restrict__ udst,
                                                 float* __restrict__ wdst, const bool use_S = true) {
  constexpr unsigned FULL = 0xffffffffu;
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
    for (int i = 0; i < M; ++i) if (use_S) S[i + rank * M] = v[i];     // use_S false: S must survive (nothing is stored either)
    __syncwarp();
    // packed: (gx, gy) += (ei.x, ei.y) * ej.x, then += (ei.y, ei.x) * (ej.y, -ej.y) -- the scalar form's operations in its order
    f32x2 gc[M];
#pragma unroll
    for (int i = 0; i < M; ++i) gc[i] = pk2(0.f, 0.f);
    for (int n = 0; n < nn; ++n) {
      const float2 ej = S[j + n * M];
      const f32x2 ejx = pk2(ej.x, ej.x), ejy = pk2(ej.y, -ej.y);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const float2 ei = S[i + n * M];
        gc[i] = fma2(pk2(ei.x, ei.y), ejx, gc[i]);
        gc[i] = fma2(pk2(ei.y, ei.x), ejy, gc[i]);
      }
    }
    if (live) {
      float2* dst = Gdst + (size_t)j * M;
#pragma unroll
      for (int i = 0; i < M; ++i) { float2 g; upk2(gc[i], g.x, g.y); dst[i] = g; }
    }
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
    jacobi_sweep<M>(a, v, j, done);
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
  // eigenvalue of this lane's column
  float lam = 0.0f;
#pragma unroll
  for (int i = 0; i < M; ++i) lam = (i == j) ? a[i].x : lam;
  subspace_outputs<M>(S, j, T, live, v, lam, Gdst, udst, wdst);
}

}  // namespace
}  // namespace doa
