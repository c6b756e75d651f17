// eig.cu -- stage 2a: batched Hermitian eigendecomposition (cyclic Jacobi) + noise subspace.
//
// Replaces eig_sym() + U_N = eig_vec.cols(0, M-T-1) + U_N_sq = U_N*trans(U_N) of
// gr-doa lib/MUSIC_lin_array_impl.cc:128-133 and lib/rootMUSIC_linear_array_impl.cc:112-116 (LAPACK cheevd there).
//
// jacobi_group_kernel<M> (M = 2, 4, 8, 16): M lanes per matrix, 32/M matrices per warp, everything in registers.
// 8 and 16 elements: one-sided Jacobi on the Cholesky factor (eig_os_device.cuh).  2 and 4 elements, and any matrix the
// factorisation rejects: the two-sided iteration of eig_device.cuh --
// lane j holds column j of the working matrix A and of the accumulated eigenvector matrix V.  One sweep is M-1
// steps of a round-robin tournament; the M/2 disjoint rotations of a step are applied together:
//   columns (A J, V J): lane exchanges its column with its partner's through shuffles,
//   rows    (J^H A)   : every lane rotates the element pairs (p_k, q_k) of its own column, indices static.
// The tournament is unrolled at compile time so no register array is indexed dynamically.
// Like cheevd('U') only the upper triangle of the input is read.
//
// Any other M <= 64: one CTA per matrix in shared memory (eig_block.cu).
//
// Outputs per frame: G = sum_{n<M-T} e_n e_n^H (eigenvalues ascending, column-major), the diagonal sums
// u_l = sum_r G[r][r+l] (the Root-MUSIC polynomial / ULA null-spectrum coefficients, cf.
// lib/rootMUSIC_linear_array_impl.cc:74-79), and the sorted eigenvalues.
#include "eig_os_device.cuh"

namespace doa {
namespace {

constexpr int JG_WARPS = 4;
#ifndef DOA_JG_MINBLOCKS
#define DOA_JG_MINBLOCKS 5   // 96 registers at 16 elements: 20 warps per SM instead of 16 (1.22 -> 1.12 ms per 65,536; 6 blocks: no further gain)
#endif

template <int M>
__global__ void __launch_bounds__(JG_WARPS * 32, DOA_JG_MINBLOCKS)
jacobi_group_kernel(const float2* __restrict__ R, int T, int nframes, float2* __restrict__ G, float2* __restrict__ u,
                    float* __restrict__ w, int max_sweeps) {
  constexpr int GPW = 32 / M;                 // matrices per warp
  __shared__ float2 stage_s[JG_WARPS * GPW][M * M];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = lane % M, g = lane / M;
  const long long mat_raw = ((long long)blockIdx.x * JG_WARPS + warp) * GPW + g;
  const bool live = mat_raw < nframes;
  const long long mat = live ? mat_raw : (long long)nframes - 1;
  float2* S = stage_s[warp * GPW + g];
  {
    const float2* src = R + mat * M * M;
#pragma unroll
    for (int i = 0; i < M; ++i) S[i + j * M] = src[i + j * M];
  }
  __syncwarp();
  noise_subspace_solve<M>(S, j, T, max_sweeps, live, G ? G + mat * M * M : nullptr, u ? u + mat * M : nullptr,
                          w ? w + mat * M : nullptr);
}

// calibrate_lin_array (gr-doa lib/calibrate_lin_array_impl.cc:112-126).  With ONE source the noise
// projector is G = I - u_S u_S^H, so U_S U_S^H = I - G and W = diag(conj v) U_S U_S^H diag(v) = w w^H with
// w = conj(v) o u_S: the eigenvector of W for its unit eigenvalue is w itself (up to the unit-modulus factor LAPACK leaves
// arbitrary in the reference).  u_S is read off the column of I - G with the largest diagonal entry (the best conditioned
// one), polished by one power-iteration step with R, normalised, and its phase fixed so that that entry is real and
// positive.  One warp per frame.
__global__ void __launch_bounds__(128)
calibrate_emit_kernel(const float2* __restrict__ R, const float2* __restrict__ G, const float2* __restrict__ v, int M, int nframes,
                      float2* __restrict__ out) {
  __shared__ float2 us_s[4][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (f >= nframes) return;
  float2* us = us_s[warp];
  const float2* Gf = G + f * M * M;
  const float2* Rf = R + f * M * M;
  float best = -INFINITY; int bc = 0;
  for (int r = lane; r < M; r += 32) { const float dgn = 1.0f - Gf[r + (size_t)r * M].x; if (dgn > best) { best = dgn; bc = r; } }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o); const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
    if (ob > best || (ob == best && oc < bc)) { best = ob; bc = oc; }
  }
  // column bc of I - G = u_S conj(u_S[bc]): u_S up to a positive factor, with entry bc real and positive
  for (int r = lane; r < M; r += 32) {
    const float2 g = Gf[r + (size_t)bc * M];
    us[r] = make_float2(((r == bc) ? 1.0f : 0.0f) - g.x, -g.y);
  }
  __syncwarp();
  // one power-iteration step with R itself (upper triangle, like cheevd 'U'): contracts the error of the projector-derived
  // vector by lambda_2 / lambda_1 and removes the dependence on how well the noise eigenvalues were separated
  float2 y[2]; float n2 = 0.0f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int r = lane + 32 * k;
    float ax = 0.0f, ay = 0.0f;
    if (r < M) {
      for (int c = 0; c < M; ++c) {
        float2 a;
        if (r < c) a = Rf[r + (size_t)c * M];
        else if (r == c) a = make_float2(Rf[r + (size_t)c * M].x, 0.0f);
        else { const float2 t = Rf[c + (size_t)r * M]; a = make_float2(t.x, -t.y); }
        const float2 u = us[c];
        ax = fmaf(a.x, u.x, fmaf(-a.y, u.y, ax));
        ay = fmaf(a.x, u.y, fmaf(a.y, u.x, ay));
      }
    }
    y[k] = make_float2(ax, ay);
    n2 = fmaf(ax, ax, fmaf(ay, ay, n2));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  // phase reference: entry bc real and positive; unit norm
  const int owner = bc & 31, kk = bc >> 5;
  const float pbx = __shfl_sync(0xffffffffu, kk ? y[1].x : y[0].x, owner), pby = __shfl_sync(0xffffffffu, kk ? y[1].y : y[0].y, owner);
  const float pm = rsqrtf(fmaxf(pbx * pbx + pby * pby, 1e-37f)), inv = rsqrtf(fmaxf(n2, 1e-37f));
  const float rx = pbx * pm * inv, ry = -pby * pm * inv;                    // conj(y_bc)/|y_bc| / ||y||
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int r = lane + 32 * k;
    if (r < M) {
      const float ux = y[k].x * rx - y[k].y * ry, uy = y[k].x * ry + y[k].y * rx;
      const float2 vr = v[r];
      out[f * M + r] = make_float2(vr.x * ux + vr.y * uy, vr.x * uy - vr.y * ux);     // conj(v_r) * u_r
    }
  }
}

template <int M>
int launch_group(const float2* R, int T, int nframes, float2* G, float2* u, float* w, cudaStream_t st) {
  constexpr int per_block = JG_WARPS * (32 / M);
  const int blocks = (nframes + per_block - 1) / per_block;
  jacobi_group_kernel<M><<<blocks, JG_WARPS * 32, 0, st>>>(R, T, nframes, G, u, w, eig_sweeps_arg(M));
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
  return launch_noise_subspace_block(R, M, T, nframes, G, u, w, st);
}

int launch_calibrate_emit(const float2* R, const float2* G, const float2* v, int M, int nframes, float2* out, cudaStream_t st) {
  if (nframes <= 0) return 0;
  if (M > 64) return DOA_CUDA_EINVAL;
  calibrate_emit_kernel<<<(nframes + 3) / 4, 128, 0, st>>>(R, G, v, M, nframes, out);
  return 1;
}

}  // namespace doa
