// eig.cu -- stage 2a: batched Hermitian eigendecomposition (cyclic Jacobi) + noise subspace.
//
// Replaces eig_sym() + U_N = eig_vec.cols(0, M-T-1) + U_N_sq = U_N*trans(U_N) of
// gr-doa lib/MUSIC_lin_array_impl.cc:128-133 and lib/rootMUSIC_linear_array_impl.cc:112-116 (LAPACK cheevd there).
//
// jacobi_group_kernel<M> (M = 2, 4, 8, 16): M lanes per matrix, 32/M matrices per warp, everything in registers.
// Lane j holds column j of the working matrix A and of the accumulated eigenvector matrix V.  One sweep is M-1
// steps of a round-robin tournament; the M/2 disjoint rotations of a step are applied together:
//   columns (A J, V J): lane exchanges its column with its partner's through shuffles,
//   rows    (J^H A)   : every lane rotates the element pairs (p_k, q_k) of its own column, indices static.
// The tournament is unrolled at compile time so no register array is indexed dynamically.
// Like cheevd('U') only the upper triangle of the input is read.
//
// jacobi_block_kernel (any M <= 64): one CTA per matrix in shared memory, same tournament, threads over (pair,row).
//
// Outputs per frame: G = sum_{n<M-T} e_n e_n^H (eigenvalues ascending, column-major), the diagonal sums
// u_l = sum_r G[r][r+l] (the Root-MUSIC polynomial / ULA null-spectrum coefficients, cf.
// lib/rootMUSIC_linear_array_impl.cc:74-79), and the sorted eigenvalues.
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

constexpr int JG_WARPS = 4;

template <int M>
__global__ void __launch_bounds__(JG_WARPS * 32)
jacobi_group_kernel(const float2* __restrict__ R, int T, int nframes, float2* __restrict__ G, float2* __restrict__ u,
                    float* __restrict__ w, int max_sweeps) {
  constexpr int GPW = 32 / M;                 // matrices per warp
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ float2 stage_s[JG_WARPS * GPW][M * M];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = lane % M, g = lane / M;
  const long long mat_raw = ((long long)blockIdx.x * JG_WARPS + warp) * GPW + g;
  const bool live = mat_raw < nframes;
  const long long mat = live ? mat_raw : (long long)nframes - 1;
  float2* S = stage_s[warp * GPW + g];

  // load column-major R through shared memory; build column j from the upper triangle only
  {
    const float2* src = R + mat * M * M;
#pragma unroll
    for (int i = 0; i < M; ++i) S[i + j * M] = src[i + j * M];
  }
  __syncwarp();
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
  if (w != nullptr && live) w[mat * M + rank] = lam;

  if (u != nullptr) {
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
    if (live) u[mat * M + j] = mine;
  }

  if (G != nullptr) {
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
      float2* dst = G + mat * M * M + (long long)j * M;
#pragma unroll
      for (int i = 0; i < M; ++i) dst[i] = gc[i];
    }
  }
}

// ---- generic M: one CTA per matrix ----------------------------------------------------------------------------
constexpr int JB_THREADS = 256;

__global__ void __launch_bounds__(JB_THREADS)
jacobi_block_kernel(const float2* __restrict__ R, int M, int T, int nframes, float2* __restrict__ G,
                    float2* __restrict__ u, float* __restrict__ w, int max_sweeps) {
  extern __shared__ float2 sm[];
  float2* A = sm;                    // [M][M] column-major: A[i + j*M]
  float2* V = A + M * M;             // [M][M]
  float* rc = reinterpret_cast<float*>(V + M * M);   // rotation params: c[Mp/2], sx[Mp/2], sy[Mp/2]
  int* pp = reinterpret_cast<int*>(rc + 3 * 32);     // p[Mp/2], q[Mp/2]
  float* lam = reinterpret_cast<float*>(pp + 2 * 32);   // [M]
  int* rk = reinterpret_cast<int*>(lam + 64);           // [M]
  __shared__ float red[2];
  const int tid = threadIdx.x;
  const int Mp = (M + 1) & ~1, HP = Mp / 2;

  for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
    const float2* src = R + (long long)f * M * M;
    for (int e = tid; e < M * M; e += JB_THREADS) {
      const int i = e % M, j = e / M;
      float2 x;
      if (i < j) x = src[i + j * M];
      else if (i == j) x = make_float2(src[e].x, 0.f);
      else { const float2 t = src[j + i * M]; x = make_float2(t.x, -t.y); }
      A[e] = x;
      V[e] = make_float2(i == j ? 1.f : 0.f, 0.f);
    }
    __syncthreads();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
      if (tid == 0) { red[0] = 0.f; red[1] = 0.f; }
      __syncthreads();
      float off = 0.f, dg = 0.f;
      for (int e = tid; e < M * M; e += JB_THREADS) {
        const float m2 = A[e].x * A[e].x + A[e].y * A[e].y;
        if (e % M == e / M) dg += m2; else off += m2;
      }
      for (int o = 16; o >= 1; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
      if ((tid & 31) == 0) { atomicAdd(&red[0], off); atomicAdd(&red[1], dg); }
      __syncthreads();
      const bool conv = red[0] <= red[1] * (1.5e-14f * M * M);
      __syncthreads();
      if (conv) break;
      for (int s = 0; s < Mp - 1; ++s) {
        if (tid < HP) {
          int a_ = (tid == 0) ? s : (s + tid) % (Mp - 1);
          int b_ = (tid == 0) ? (Mp - 1) : (s - tid + (Mp - 1)) % (Mp - 1);
          int p = min(a_, b_), q = max(a_, b_);
          Rot r; r.c = 1.f; r.sx = 0.f; r.sy = 0.f;
          if (q < M) r = make_rotation(A[p + p * M].x, A[q + q * M].x, A[p + q * M]);
          else { p = -1; }   // pair with the padding index: skip
          rc[tid] = r.c; rc[32 + tid] = r.sx; rc[64 + tid] = r.sy; pp[tid] = p; pp[32 + tid] = q;
        }
        __syncthreads();
        for (int it = tid; it < HP * M; it += JB_THREADS) {   // columns of A and V
          const int k = it / M, i = it % M;
          const int p = pp[k], q = pp[32 + k];
          if (p < 0) continue;
          const float c = rc[k], sx = rc[32 + k], sy = rc[64 + k];
          {
            const float2 x = A[i + p * M], y = A[i + q * M];
            A[i + p * M] = make_float2(c * x.x - (sx * y.x + sy * y.y), c * x.y - (sx * y.y - sy * y.x));
            A[i + q * M] = make_float2(sx * x.x - sy * x.y + c * y.x, sx * x.y + sy * x.x + c * y.y);
          }
          {
            const float2 x = V[i + p * M], y = V[i + q * M];
            V[i + p * M] = make_float2(c * x.x - (sx * y.x + sy * y.y), c * x.y - (sx * y.y - sy * y.x));
            V[i + q * M] = make_float2(sx * x.x - sy * x.y + c * y.x, sx * x.y + sy * x.x + c * y.y);
          }
        }
        __syncthreads();
        for (int it = tid; it < HP * M; it += JB_THREADS) {   // rows of A
          const int k = it / M, i = it % M;
          const int p = pp[k], q = pp[32 + k];
          if (p < 0) continue;
          const float c = rc[k], sx = rc[32 + k], sy = rc[64 + k];
          const float2 x = A[p + i * M], y = A[q + i * M];
          A[p + i * M] = make_float2(c * x.x - (sx * y.x - sy * y.y), c * x.y - (sx * y.y + sy * y.x));
          A[q + i * M] = make_float2(sx * x.x + sy * x.y + c * y.x, sx * x.y - sy * x.x + c * y.y);
        }
        __syncthreads();
      }
    }
    // ranks
    for (int j = tid; j < M; j += JB_THREADS) lam[j] = A[j + j * M].x;
    __syncthreads();
    for (int j = tid; j < M; j += JB_THREADS) {
      int r = 0;
      for (int i = 0; i < M; ++i) r += (lam[i] < lam[j] || (lam[i] == lam[j] && i < j)) ? 1 : 0;
      rk[r] = j;   // rk[rank] = column holding that eigenvalue
      if (w) w[(long long)f * M + r] = lam[j];
    }
    __syncthreads();
    const int nn = M - T;
    // G into A's storage (A no longer needed)
    for (int e = tid; e < M * M; e += JB_THREADS) {
      const int i = e % M, j = e / M;
      float gx = 0.f, gy = 0.f;
      for (int n = 0; n < nn; ++n) {
        const float2 ei = V[i + rk[n] * M], ej = V[j + rk[n] * M];
        gx = fmaf(ei.x, ej.x, gx); gx = fmaf(ei.y, ej.y, gx);
        gy = fmaf(ei.y, ej.x, gy); gy = fmaf(-ei.x, ej.y, gy);
      }
      A[e] = make_float2(gx, gy);
      if (G) G[(long long)f * M * M + e] = make_float2(gx, gy);
    }
    __syncthreads();
    if (u) {
      for (int l = tid; l < M; l += JB_THREADS) {
        float sx = 0.f, sy = 0.f;
        for (int r = 0; r + l < M; ++r) { sx += A[r + (r + l) * M].x; sy += A[r + (r + l) * M].y; }
        u[(long long)f * M + l] = make_float2(sx, l == 0 ? 0.f : sy);
      }
    }
    __syncthreads();
  }
}

template <int M>
int launch_group(const float2* R, int T, int nframes, float2* G, float2* u, float* w, cudaStream_t st) {
  constexpr int per_block = JG_WARPS * (32 / M);
  const int blocks = (nframes + per_block - 1) / per_block;
  jacobi_group_kernel<M><<<blocks, JG_WARPS * 32, 0, st>>>(R, T, nframes, G, u, w, M <= 8 ? 12 : 16);
  return 1;
}

}  // namespace

int launch_noise_subspace(const float2* R, int M, int T, int nframes, float2* G, float2* u, float* w, cudaStream_t st) {
  if (nframes <= 0) return 0;
  switch (M) {
    case 2: return launch_group<2>(R, T, nframes, G, u, w, st);
    case 4: return launch_group<4>(R, T, nframes, G, u, w, st);
    case 8: return launch_group<8>(R, T, nframes, G, u, w, st);
    case 16: return launch_group<16>(R, T, nframes, G, u, w, st);
    default: break;
  }
  if (M > 64 || M < 2) return DOA_CUDA_EINVAL;
  const size_t smem = (size_t)2 * M * M * sizeof(float2) + (3 * 32) * sizeof(float) + (2 * 32) * sizeof(int) +
                      64 * sizeof(float) + 64 * sizeof(int);
  cudaFuncSetAttribute(jacobi_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = min(nframes, sms * 4);
  jacobi_block_kernel<<<blocks, JB_THREADS, smem, st>>>(R, M, T, nframes, G, u, w, 20);
  return 1;
}

}  // namespace doa
