// eig_block.cu -- stage 2a for arrays of 17..64 elements: one CTA per covariance matrix, the matrix in shared memory.
//
// Replaces the same reference lines as eig.cu (eig_sym + U_N U_N^H, gr-doa lib/MUSIC_lin_array_impl.cc:128-133,
// lib/rootMUSIC_linear_array_impl.cc:112-116) at BASELINE configs[3] (64 elements, 8 sources).
//
// jacobi_os_block_kernel<MP, BLK> (MP = 32 or 64, the matrix padded with zero columns): one-sided Jacobi on the Cholesky factor,
// the algorithm of eig_os_device.cuh laid out for a CTA.  The factor's columns live in two shared-memory planes (real, imaginary;
// column-major, MP floats per column), factored in place (right-looking Cholesky).  EIGHT lanes own a pair: lane r holds rows
// 4r..4r+3 (and 32+4r.. at MP = 64) of the pair's columns, so every 8-lane phase of a 128-bit shared-memory access touches 128
// contiguous bytes (conflict-free whatever the pairing), the pair's dot product is a 3-level butterfly among the eight lanes (no
// shared memory, no barrier), all eight derive the same rotation and rotate their rows in place.  ONE CTA barrier per step.
// Squared column norms are carried in shared memory by the rotations' own update and recomputed at the start of a sweep.
//   BLK (shipped): the columns are grouped once and for all into blocks of two and a step pairs BLOCKS -- the eight lanes hold
//   four columns in registers and rotate the four cross pairs, the two pairs inside the blocks once per sweep.  Same rotations
//   per sweep, same 8 sweeps, half the steps: half the shared-memory traffic (the column scheme moves 64 KB per step and is
//   bound by it) and half the barriers.  512 matrices of 64 x 64: 0.87 ms against 1.08 ms (same box) and 3.5 ms for the
//   two-sided kernel below.
//   !BLK (option "eig_onesided" = 2, comparison): a step pairs columns, MP/2 pairs, MP - 1 steps.
//
// A matrix the Cholesky factorisation rejects (indefinite, zero: not a covariance) is handled in the same planes without a
// factor: H = R + ||R||_F I is positive semidefinite whatever R is and has R's eigenvectors, and for a Hermitian positive
// semidefinite H the same column rotations give H J_1 J_2 ... = U (Lambda + delta): eigenvalue = |column| - delta.  (Slightly
// less accurate than the factored path -- the Gram matrix is H^2 -- which is why it is only the fallback.)  No second
// shared-memory layout is needed, so the kernel runs six CTAs per SM at 64 elements.
//
// jacobi_block_kernel: the two-sided iteration on A and V in shared memory (the round-1 kernel), behind option
// "eig_onesided" = 0 as the comparison.
#include "eig_os_device.cuh"

namespace doa {
namespace {

// ---- two-sided iteration (fallback) --------------------------------------------------------------------------------------
// A and V live in shared memory with an odd leading dimension (M+1 float2) so that both the column phase (threads walk a
// column) and the row phase (threads walk a row, stride LD) are bank-conflict free.  All threads of the CTA call together.
__device__ void jacobi_block_body(const float2* __restrict__ src, const int M, const int T, float2* __restrict__ Gf,
                                  float2* __restrict__ uf, float* __restrict__ wf, const int max_sweeps, float2* sm, float* red) {
  const int LD = M | 1;              // odd leading dimension
  float2* A = sm;                    // element (i, j) at A[i + j*LD]
  float2* V = A + (size_t)M * LD;
  float* rc = reinterpret_cast<float*>(V + (size_t)M * LD);   // rotation params: c[32], sx[32], sy[32]
  int* pp = reinterpret_cast<int*>(rc + 3 * 32);              // p[32], q[32]
  float* lam = reinterpret_cast<float*>(pp + 2 * 32);         // [64]
  int* rk = reinterpret_cast<int*>(lam + 64);                 // [64]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int Mp = (M + 1) & ~1, HP = Mp / 2;

  for (int e = tid; e < M * M; e += nthr) {
    const int i = e % M, j = e / M;
    float2 x;
    if (i < j) x = src[i + j * M];                      // upper triangle only, like cheevd 'U'
    else if (i == j) x = make_float2(src[e].x, 0.f);
    else { const float2 t = src[j + i * M]; x = make_float2(t.x, -t.y); }
    A[i + j * LD] = x;
    V[i + j * LD] = make_float2(i == j ? 1.f : 0.f, 0.f);
  }
  __syncthreads();
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) { red[0] = 0.f; red[1] = 0.f; }
    __syncthreads();
    float off = 0.f, dg = 0.f;
    for (int e = tid; e < M * M; e += nthr) {
      const int i = e % M, j = e / M;
      const float2 a = A[i + j * LD];
      const float m2 = a.x * a.x + a.y * a.y;
      if (i == j) dg += m2; else off += m2;
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
        if (q < M) r = make_rotation(A[p + p * LD].x, A[q + q * LD].x, A[p + q * LD]);
        else { p = -1; }   // pair with the padding index: skip
        rc[tid] = r.c; rc[32 + tid] = r.sx; rc[64 + tid] = r.sy; pp[tid] = p; pp[32 + tid] = q;
      }
      __syncthreads();
      for (int it = tid; it < HP * M; it += nthr) {   // columns of A and V: A <- A J, V <- V J
        const int k = it / M, i = it % M;
        const int p = pp[k], q = pp[32 + k];
        if (p < 0) continue;
        const float c = rc[k], sx = rc[32 + k], sy = rc[64 + k];
        {
          const float2 x = A[i + p * LD], y = A[i + q * LD];
          A[i + p * LD] = make_float2(c * x.x - (sx * y.x + sy * y.y), c * x.y - (sx * y.y - sy * y.x));
          A[i + q * LD] = make_float2(sx * x.x - sy * x.y + c * y.x, sx * x.y + sy * x.x + c * y.y);
        }
        {
          const float2 x = V[i + p * LD], y = V[i + q * LD];
          V[i + p * LD] = make_float2(c * x.x - (sx * y.x + sy * y.y), c * x.y - (sx * y.y - sy * y.x));
          V[i + q * LD] = make_float2(sx * x.x - sy * x.y + c * y.x, sx * x.y + sy * x.x + c * y.y);
        }
      }
      __syncthreads();
      for (int it = tid; it < HP * M; it += nthr) {   // rows of A: A <- J^H A
        const int k = it / M, i = it % M;
        const int p = pp[k], q = pp[32 + k];
        if (p < 0) continue;
        const float c = rc[k], sx = rc[32 + k], sy = rc[64 + k];
        const float2 x = A[p + i * LD], y = A[q + i * LD];
        A[p + i * LD] = make_float2(c * x.x - (sx * y.x - sy * y.y), c * x.y - (sx * y.y + sy * y.x));
        A[q + i * LD] = make_float2(sx * x.x + sy * x.y + c * y.x, sx * x.y - sy * x.x + c * y.y);
      }
      __syncthreads();
    }
  }
  // unit eigenvectors (the MUFU rotations let column norms drift by O(1e-7) per rotation), eigenvalues, ranks
  for (int j = tid; j < M; j += nthr) {
    float n2 = 0.f;
    for (int i = 0; i < M; ++i) { const float2 v = V[i + j * LD]; n2 = fmaf(v.x, v.x, fmaf(v.y, v.y, n2)); }
    const float sc = 1.0f / sqrtf(n2);
    for (int i = 0; i < M; ++i) { V[i + j * LD].x *= sc; V[i + j * LD].y *= sc; }
    lam[j] = A[j + j * LD].x;
  }
  __syncthreads();
  for (int j = tid; j < M; j += nthr) {
    int r = 0;
    for (int i = 0; i < M; ++i) r += (lam[i] < lam[j] || (lam[i] == lam[j] && i < j)) ? 1 : 0;
    rk[r] = j;   // rk[rank] = column holding that eigenvalue
    if (wf) wf[r] = lam[j];
  }
  __syncthreads();
  const int nn = M - T;
  // G into A's storage (A no longer needed)
  for (int e = tid; e < M * M; e += nthr) {
    const int i = e % M, j = e / M;
    float gx = 0.f, gy = 0.f;
    for (int n = 0; n < nn; ++n) {
      const float2 ei = V[i + rk[n] * LD], ej = V[j + rk[n] * LD];
      gx = fmaf(ei.x, ej.x, gx); gx = fmaf(ei.y, ej.y, gx);
      gy = fmaf(ei.y, ej.x, gy); gy = fmaf(-ei.x, ej.y, gy);
    }
    A[i + j * LD] = make_float2(gx, gy);
    if (Gf) Gf[e] = make_float2(gx, gy);
  }
  __syncthreads();
  if (uf) {
    for (int l = tid; l < M; l += nthr) {
      float sx = 0.f, sy = 0.f;
      for (int r = 0; r + l < M; ++r) { sx += A[r + (r + l) * LD].x; sy += A[r + (r + l) * LD].y; }
      uf[l] = make_float2(sx, l == 0 ? 0.f : sy);
    }
  }
  __syncthreads();
}

constexpr size_t block_body_smem(int M) {
  return (size_t)2 * M * (M | 1) * sizeof(float2) + (3 * 32) * sizeof(float) + (2 * 32) * sizeof(int) + 64 * sizeof(float) + 64 * sizeof(int);
}

constexpr int JB_THREADS = 512;

__global__ void __launch_bounds__(JB_THREADS)
jacobi_block_kernel(const float2* __restrict__ R, int M, int T, int nframes, float2* __restrict__ G,
                    float2* __restrict__ u, float* __restrict__ w, int max_sweeps) {
  extern __shared__ __align__(16) float2 sm[];
  __shared__ float red[2];
  for (int f = blockIdx.x; f < nframes; f += gridDim.x)
    jacobi_block_body(R + (long long)f * M * M, M, T, G ? G + (long long)f * M * M : nullptr, u ? u + (long long)f * M : nullptr,
                      w ? w + (long long)f * M : nullptr, max_sweeps, sm, red);
}

// ---- one-sided Jacobi on the Cholesky factor, CTA form ----------------------------------------------------------------------
#ifndef DOA_EIGBLK_MINBLOCKS
#define DOA_EIGBLK_MINBLOCKS 4   // 128 registers, 16 warps per SM: 0.87 ms per 512 matrices of 64 x 64 (3 blocks / 164 registers: 0.99, 5 / 96: 0.91)
#endif
template <int MP, bool BLK>
struct OsBlock {
  static constexpr int THREADS = BLK ? MP * 2 : MP * 4;   // MP/4 block pairs (BLK) or MP/2 column pairs, 8 lanes each
  static constexpr int NC = MP / 32;              // 4-row chunks per lane and column
  static constexpr size_t PLANE = (size_t)MP * MP * sizeof(float);
  // planes | nrm[MP] | lam[MP] | rk[MP] | red[THREADS/32]
  static constexpr size_t SMEM_OS = 2 * PLANE + 3 * MP * sizeof(float) + (THREADS / 32) * sizeof(float);
};

// Sum of one float per thread over the CTA in a fixed order (warp butterflies, then the warps' partials in warp order).
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                            // scratch free
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.0f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) s += scratch[w];
  return s;
}

template <int MP, bool BLK>
__global__ void __launch_bounds__(OsBlock<MP, BLK>::THREADS, BLK ? DOA_EIGBLK_MINBLOCKS : 1)
jacobi_os_block_kernel(const float2* __restrict__ R, const int M, const int T, const int nframes, float2* __restrict__ G,
                       float2* __restrict__ u, float* __restrict__ w, const int max_sweeps) {
  using C = OsBlock<MP, BLK>;
  constexpr int NT = C::THREADS, NC = C::NC, RR = MP - 1;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float2 sm[];
  float* Wr = reinterpret_cast<float*>(sm);                   // column c of the real plane at Wr + c*MP
  float* Wi = Wr + MP * MP;
  float* nrm = Wi + MP * MP;                                  // squared column norms (tracked)
  float* lam = nrm + MP;
  int* rk = reinterpret_cast<int*>(lam + MP);
  float* red = reinterpret_cast<float*>(rk + MP);             // [NT / 32]
  const int tid = threadIdx.x;
  const int grp = tid >> 3, r8 = tid & 7;                     // pair slot, lane of the slot

  for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
    const float2* src = R + (long long)f * M * M;
    // ---- trace, exact power-of-two scale, shift (see eig_os_device.cuh) ----
    float tr = block_sum<NT>(tid < M ? src[tid + (size_t)tid * M].x : 0.0f, red);
    const bool tr_ok = tr > 0.0f && tr < 3.0e38f;
    int ex = min(max((__float_as_int(tr) >> 23) & 0xff, 1), 253);
    float scl = __int_as_float((254 - ex) << 23), unscl = __int_as_float(ex << 23);
    float delta = tr_ok ? (tr * scl) * (1.0f / 16384.0f) : 1.0f;
    // ---- lower triangle of (R scl + delta I) from the upper triangle of R (like cheevd 'U'); the rest zero ----
    for (int e = tid; e < MP * MP; e += NT) {
      const int i = e % MP, k = e / MP;                       // row, column
      float re = 0.0f, im = 0.0f;
      if (i < M && k < M && i >= k) {
        const float2 a = src[k + (size_t)i * M];              // A(k, i), k <= i;  A(i, k) = conj
        re = (i == k) ? fmaf(a.x, scl, delta) : a.x * scl;
        im = (i == k) ? 0.0f : -a.y * scl;
      }
      Wr[e] = re; Wi[e] = im;
    }
    __syncthreads();
    // ---- Cholesky, right-looking, in place ----
    bool ok = tr_ok;
    for (int j = 0; j < M && ok; ++j) {
      const float piv = Wr[j * MP + j];
      ok = piv > 0.0f && piv < 3.0e38f;                       // the same in every thread
      const float d = sqrtf(ok ? piv : 1.0f), inv = 1.0f / d;
      __syncthreads();                                        // everyone has read the pivot
      if (!ok) break;
      for (int i = j + tid; i < M; i += NT) {
        if (i == j) { Wr[j * MP + j] = d; Wi[j * MP + j] = 0.0f; }
        else { Wr[j * MP + i] *= inv; Wi[j * MP + i] *= inv; }
      }
      __syncthreads();
      // trailing columns k > j, rows i >= k: A(i, k) -= L(i, j) conj(L(k, j)); a warp per column, lanes along rows
      for (int k = j + 1 + (tid >> 5); k < M; k += NT / 32) {
        const float lkx = Wr[j * MP + k], lky = Wi[j * MP + k];
        for (int i = k + (tid & 31); i < M; i += 32) {
          const float lix = Wr[j * MP + i], liy = Wi[j * MP + i];
          Wr[k * MP + i] = fmaf(-lix, lkx, fmaf(-liy, lky, Wr[k * MP + i]));
          Wi[k * MP + i] = fmaf(-liy, lkx, fmaf(lix, lky, Wi[k * MP + i]));
        }
      }
      __syncthreads();
    }
    const bool factored = ok;
    if (!factored) {
      // ---- not a covariance: H = R / 2^e + delta I with delta = ||R / 2^e||_F in [1, 2), all of it, no factor ----
      float part = 0.0f;
      for (int e = tid; e < M * M; e += NT) {
        const int i = e % M, k = e / M;
        if (i <= k) { const float2 a = src[e]; const float m2 = (i == k) ? a.x * a.x : 2.0f * fmaf(a.x, a.x, a.y * a.y); part += m2; }
      }
      // squares of huge entries overflow to +inf, of tiny ones vanish: scale by the largest-exponent estimate first is not
      // worth its code here -- such a matrix gets delta from the clamped exponent below and still converges
      const float fro = sqrtf(block_sum<NT>(part, red));
      const bool any = fro > 0.0f && fro < 3.0e38f;
      ex = min(max((__float_as_int(fro) >> 23) & 0xff, 1), 253);
      scl = any ? __int_as_float((254 - ex) << 23) : 1.0f; unscl = any ? __int_as_float(ex << 23) : 1.0f;
      delta = any ? fro * scl : 1.0f;                         // the zero matrix becomes the identity: eigenvalues 1 - 1
      for (int e = tid; e < MP * MP; e += NT) {
        const int i = e % MP, k = e / MP;
        float re = 0.0f, im = 0.0f;
        if (i < M && k < M) {
          if (i == k) re = fmaf(src[i + (size_t)i * M].x, scl, delta);
          else if (i < k) { const float2 a = src[i + (size_t)k * M]; re = a.x * scl; im = a.y * scl; }
          else { const float2 a = src[k + (size_t)i * M]; re = a.x * scl; im = -a.y * scl; }
        }
        Wr[e] = re; Wi[e] = im;
      }
      __syncthreads();
    }

    // ---- sweeps ----
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
      // squared column norms: four lanes per column
      for (int c4 = tid; c4 < 4 * MP; c4 += NT) {
        const int c = c4 >> 2, sub = c4 & 3;
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < MP / 16; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(Wr + c * MP + sub * (MP / 4) + 4 * i);
          const float4 b = *reinterpret_cast<const float4*>(Wi + c * MP + sub * (MP / 4) + 4 * i);
          s = fmaf(a.x, a.x, s); s = fmaf(a.y, a.y, s); s = fmaf(a.z, a.z, s); s = fmaf(a.w, a.w, s);
          s = fmaf(b.x, b.x, s); s = fmaf(b.y, b.y, s); s = fmaf(b.z, b.z, s); s = fmaf(b.w, b.w, s);
        }
        s += __shfl_xor_sync(FULL, s, 1);
        s += __shfl_xor_sync(FULL, s, 2);
        if (sub == 0) nrm[c] = s;
      }
      __syncthreads();
      bool dirty_any = false;
      if constexpr (BLK) {
        // Block scheme: the columns are paired up once and for all into MP/2 blocks of two; a step pairs BLOCKS (round-robin over
        // MP/2 block indices: MP/2 - 1 steps, MP/4 block pairs each) and the eight lanes of a block pair hold their rows of all four
        // columns in registers and rotate the four cross pairs (0,2) (1,3) (0,3) (1,2); the two pairs inside the blocks are rotated
        // once per sweep (step 0).  The same number of rotations per sweep as pairing columns (MP (MP - 1) / 2), the same 8
        // sweeps (prototype and GPU), HALF the shared-memory traffic and half the CTA barriers: the column scheme is bound by
        // the shared-memory port (64 KB per step, profiles/r02_ncu_cfg4_eig_scan.txt).
        constexpr int NB = MP / 2, RB = NB - 1;
#pragma unroll 1
        for (int st = 0; st < RB; ++st) {
          int bp = st + grp; bp -= (bp >= RB) ? RB : 0;
          int bq = st - grp; bq += (bq < 0) ? RB : 0;
          bq = (grp == 0) ? RB : bq;
          const int col[4] = {2 * bp, 2 * bp + 1, 2 * bq, 2 * bq + 1};
          f32x2 xr[4][2 * NC], xi[4][2 * NC];     // [column][row pair]: rows 4 r8 + {0,1}, {2,3} (+ 32 per chunk)
          float nv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(Wr + col[k] * MP + 4 * r8 + 32 * c);
              const ulonglong2 b2_ = *reinterpret_cast<const ulonglong2*>(Wi + col[k] * MP + 4 * r8 + 32 * c);
              xr[k][2 * c] = a.x; xr[k][2 * c + 1] = a.y; xi[k][2 * c] = b2_.x; xi[k][2 * c + 1] = b2_.y;
            }
            nv[k] = nrm[col[k]];
          }
          auto rotate = [&](const int a, const int b) {
            f32x2 rr = pk2(0.f, 0.f), ii = rr, ri = rr, ir = rr;
#pragma unroll
            for (int h = 0; h < 2 * NC; ++h) {
              rr = fma2(xr[a][h], xr[b][h], rr); ii = fma2(xi[a][h], xi[b][h], ii);
              ri = fma2(xr[a][h], xi[b][h], ri); ir = fma2(xi[a][h], xr[b][h], ir);
            }
            float a0, a1, b0, b1, c0, c1, d0, d1;
            upk2(rr, a0, a1); upk2(ii, b0, b1); upk2(ri, c0, c1); upk2(ir, d0, d1);
            float dx = (a0 + a1) + (b0 + b1), dy = (c0 + c1) - (d0 + d1);
#pragma unroll
            for (int o = 1; o <= 4; o <<= 1) { dx += __shfl_xor_sync(FULL, dx, o); dy += __shfl_xor_sync(FULL, dy, o); }
            float tb; bool dirty;
            const Rot rot = make_rotation_os(nv[a], nv[b], make_float2(dx, dy), false, tb, dirty);
            dirty_any = dirty_any || dirty;
            const f32x2 c2 = pk2(rot.c, rot.c), sx2 = pk2(rot.sx, rot.sx), nsx2 = pk2(-rot.sx, -rot.sx), sy2 = pk2(rot.sy, rot.sy),
                        nsy2 = pk2(-rot.sy, -rot.sy);
#pragma unroll
            for (int h = 0; h < 2 * NC; ++h) {
              const f32x2 pr_ = xr[a][h], pi_ = xi[a][h], qr_ = xr[b][h], qi_ = xi[b][h];
              xr[a][h] = fma2(nsy2, qi_, fma2(nsx2, qr_, mul2(c2, pr_)));      // c pr - sx qr - sy qi
              xi[a][h] = fma2(sy2, qr_, fma2(nsx2, qi_, mul2(c2, pi_)));       // c pi - sx qi + sy qr
              xr[b][h] = fma2(nsy2, pi_, fma2(sx2, pr_, mul2(c2, qr_)));       // c qr + sx pr - sy pi
              xi[b][h] = fma2(sy2, pr_, fma2(sx2, pi_, mul2(c2, qi_)));        // c qi + sx pi + sy pr
            }
            nv[a] = fmaxf(nv[a] - tb, 0.0f); nv[b] = nv[b] + tb;
          };
          if (st == 0) { rotate(0, 1); rotate(2, 3); }
          rotate(0, 2); rotate(1, 3); rotate(0, 3); rotate(1, 2);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              ulonglong2 a, b2_;
              a.x = xr[k][2 * c]; a.y = xr[k][2 * c + 1]; b2_.x = xi[k][2 * c]; b2_.y = xi[k][2 * c + 1];
              *reinterpret_cast<ulonglong2*>(Wr + col[k] * MP + 4 * r8 + 32 * c) = a;
              *reinterpret_cast<ulonglong2*>(Wi + col[k] * MP + 4 * r8 + 32 * c) = b2_;
            }
            if (r8 == 0) nrm[col[k]] = nv[k];
          }
          __syncthreads();
        }
      } else {
#pragma unroll 1
      for (int st = 0; st < RR; ++st) {
        // pair of this slot: slot 0 is (st, RR), slot k is ((st + k) mod RR, (st - k) mod RR)
        int p = st + grp; p -= (p >= RR) ? RR : 0;
        int q = st - grp; q += (q < 0) ? RR : 0;
        q = (grp == 0) ? RR : q;
        float* pr_ = Wr + p * MP + 4 * r8; float* pi_ = Wi + p * MP + 4 * r8;
        float* qr_ = Wr + q * MP + 4 * r8; float* qi_ = Wi + q * MP + 4 * r8;
        float4 pr[NC], pi[NC], qr[NC], qi[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          pr[c] = *reinterpret_cast<const float4*>(pr_ + 32 * c); pi[c] = *reinterpret_cast<const float4*>(pi_ + 32 * c);
          qr[c] = *reinterpret_cast<const float4*>(qr_ + 32 * c); qi[c] = *reinterpret_cast<const float4*>(qi_ + 32 * c);
        }
        const float app = nrm[p], aqq = nrm[q];
        // w_p^H w_q over my rows: Re = sum pr qr + pi qi, Im = sum pr qi - pi qr
        f32x2 rr = pk2(0.f, 0.f), ii = rr, ri = rr, ir = rr;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const f32x2 pr0 = pk2(pr[c].x, pr[c].y), pr1 = pk2(pr[c].z, pr[c].w), pi0 = pk2(pi[c].x, pi[c].y), pi1 = pk2(pi[c].z, pi[c].w);
          const f32x2 qr0 = pk2(qr[c].x, qr[c].y), qr1 = pk2(qr[c].z, qr[c].w), qi0 = pk2(qi[c].x, qi[c].y), qi1 = pk2(qi[c].z, qi[c].w);
          rr = fma2(pr0, qr0, rr); ii = fma2(pi0, qi0, ii); ri = fma2(pr0, qi0, ri); ir = fma2(pi0, qr0, ir);
          rr = fma2(pr1, qr1, rr); ii = fma2(pi1, qi1, ii); ri = fma2(pr1, qi1, ri); ir = fma2(pi1, qr1, ir);
        }
        float a0, a1, b0, b1, c0, c1, d0, d1;
        upk2(rr, a0, a1); upk2(ii, b0, b1); upk2(ri, c0, c1); upk2(ir, d0, d1);
        float dx = (a0 + a1) + (b0 + b1), dy = (c0 + c1) - (d0 + d1);
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) { dx += __shfl_xor_sync(FULL, dx, o); dy += __shfl_xor_sync(FULL, dy, o); }
        float tb; bool dirty;
        const Rot rot = make_rotation_os(app, aqq, make_float2(dx, dy), false, tb, dirty);
        dirty_any = dirty_any || dirty;
        // col_p' = c col_p - conj(sigma) col_q ; col_q' = sigma col_p + c col_q
        const f32x2 c2 = pk2(rot.c, rot.c), sx2 = pk2(rot.sx, rot.sx), nsx2 = pk2(-rot.sx, -rot.sx), sy2 = pk2(rot.sy, rot.sy),
                    nsy2 = pk2(-rot.sy, -rot.sy);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float4 npr, npi, nqr, nqi;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const f32x2 xr = hh ? pk2(pr[c].z, pr[c].w) : pk2(pr[c].x, pr[c].y), xi = hh ? pk2(pi[c].z, pi[c].w) : pk2(pi[c].x, pi[c].y);
            const f32x2 yr = hh ? pk2(qr[c].z, qr[c].w) : pk2(qr[c].x, qr[c].y), yi = hh ? pk2(qi[c].z, qi[c].w) : pk2(qi[c].x, qi[c].y);
            const f32x2 tpr = fma2(nsy2, yi, fma2(nsx2, yr, mul2(c2, xr)));      // c pr - sx qr - sy qi
            const f32x2 tpi = fma2(sy2, yr, fma2(nsx2, yi, mul2(c2, xi)));       // c pi - sx qi + sy qr
            const f32x2 tqr = fma2(nsy2, xi, fma2(sx2, xr, mul2(c2, yr)));       // c qr + sx pr - sy pi
            const f32x2 tqi = fma2(sy2, xr, fma2(sx2, xi, mul2(c2, yi)));        // c qi + sx pi + sy pr
            if (hh) { upk2(tpr, npr.z, npr.w); upk2(tpi, npi.z, npi.w); upk2(tqr, nqr.z, nqr.w); upk2(tqi, nqi.z, nqi.w); }
            else { upk2(tpr, npr.x, npr.y); upk2(tpi, npi.x, npi.y); upk2(tqr, nqr.x, nqr.y); upk2(tqi, nqi.x, nqi.y); }
          }
          *reinterpret_cast<float4*>(pr_ + 32 * c) = npr; *reinterpret_cast<float4*>(pi_ + 32 * c) = npi;
          *reinterpret_cast<float4*>(qr_ + 32 * c) = nqr; *reinterpret_cast<float4*>(qi_ + 32 * c) = nqi;
        }
        if (r8 == 0) { nrm[p] = fmaxf(app - tb, 0.0f); nrm[q] = aqq + tb; }
        __syncthreads();
      }
      }
      // stopping rule of eig_os_device.cuh: done after a sweep in which every pair was orthogonal to 1e-5 before its rotation
      if (!__syncthreads_or(dirty_any ? 1 : 0)) break;
    }

    // ---- eigenpairs: the normalised column and 2^e (|column|^2 - delta), or 2^e (|column| - delta) without a factor ----
    for (int c4 = tid; c4 < 4 * MP; c4 += NT) {
      const int c = c4 >> 2, sub = c4 & 3;
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < MP / 16; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(Wr + c * MP + sub * (MP / 4) + 4 * i);
        const float4 b = *reinterpret_cast<const float4*>(Wi + c * MP + sub * (MP / 4) + 4 * i);
        s = fmaf(a.x, a.x, s); s = fmaf(a.y, a.y, s); s = fmaf(a.z, a.z, s); s = fmaf(a.w, a.w, s);
        s = fmaf(b.x, b.x, s); s = fmaf(b.y, b.y, s); s = fmaf(b.z, b.z, s); s = fmaf(b.w, b.w, s);
      }
      s += __shfl_xor_sync(FULL, s, 1);
      s += __shfl_xor_sync(FULL, s, 2);
      const float nr = sqrtf(fmaxf(s, 1e-37f)), sc = 1.0f / nr;
#pragma unroll
      for (int i = 0; i < MP / 16; ++i) {
        float4* a = reinterpret_cast<float4*>(Wr + c * MP + sub * (MP / 4) + 4 * i);
        float4* b = reinterpret_cast<float4*>(Wi + c * MP + sub * (MP / 4) + 4 * i);
        float4 x = *a, y = *b;
        x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; y.x *= sc; y.y *= sc; y.z *= sc; y.w *= sc;
        *a = x; *b = y;
      }
      if (sub == 0 && c < M) lam[c] = ((factored ? s : nr) - delta) * unscl;
    }
    __syncthreads();
    for (int c = tid; c < M; c += NT) {
      int rnk = 0;
      const float lc = lam[c];
      for (int i = 0; i < M; ++i) rnk += (lam[i] < lc || (lam[i] == lc && i < c)) ? 1 : 0;
      rk[rnk] = c;                                            // rk[rank] = column holding that eigenvalue
      if (w) w[(long long)f * M + rnk] = lc;
    }
    __syncthreads();
    const int nn = M - T;
    if (G) {
      for (int e = tid; e < M * M; e += NT) {
        const int i = e % M, jj = e / M;
        float gx = 0.f, gy = 0.f;
        for (int n = 0; n < nn; ++n) {
          const int c = rk[n];
          const float eix = Wr[c * MP + i], eiy = Wi[c * MP + i], ejx = Wr[c * MP + jj], ejy = Wi[c * MP + jj];
          gx = fmaf(eix, ejx, gx); gx = fmaf(eiy, ejy, gx);
          gy = fmaf(eiy, ejx, gy); gy = fmaf(-eix, ejy, gy);
        }
        G[(long long)f * M * M + e] = make_float2(gx, gy);
      }
    }
    if (u) {
      // u_l = sum over noise eigenvectors of sum_r e[r] conj(e[r + l]): four lanes per l, rows r = part, part + 4, ...
      for (int l4 = tid; l4 < 4 * MP; l4 += NT) {
      const int l = l4 >> 2, part = l4 & 3;
      float sx = 0.f, sy = 0.f;
      if (l < M) {
        for (int n = 0; n < nn; ++n) {
          const float* er = Wr + rk[n] * MP; const float* ei = Wi + rk[n] * MP;
          for (int r = part; r + l < M; r += 4) {
            sx = fmaf(er[r], er[r + l], sx); sx = fmaf(ei[r], ei[r + l], sx);
            sy = fmaf(ei[r], er[r + l], sy); sy = fmaf(-er[r], ei[r + l], sy);
          }
        }
      }
      sx += __shfl_xor_sync(FULL, sx, 1); sy += __shfl_xor_sync(FULL, sy, 1);
      sx += __shfl_xor_sync(FULL, sx, 2); sy += __shfl_xor_sync(FULL, sy, 2);
      if (part == 0 && l < M) u[(long long)f * M + l] = make_float2(sx, l == 0 ? 0.f : sy);
      }
    }
    __syncthreads();
  }
}

template <int MP, bool BLK>
int launch_os_block(const float2* R, int M, int T, int nframes, float2* G, float2* u, float* w, int sweeps, cudaStream_t st) {
  using C = OsBlock<MP, BLK>;
  const size_t smem = C::SMEM_OS;
  auto kern = jacobi_os_block_kernel<MP, BLK>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DOA_CUDA_ECUDA;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS, smem);
  const int blocks = min(nframes, sms * max(per_sm, 1));
  kern<<<blocks, C::THREADS, smem, st>>>(R, M, T, nframes, G, u, w, sweeps);
  return 1;
}

}  // namespace

int launch_noise_subspace_block(const float2* R, int M, int T, int nframes, float2* G, float2* u, float* w, cudaStream_t st) {
  if (M > 64 || M < 2) return DOA_CUDA_EINVAL;
  if (dev_option(OPT_EIG_ONESIDED, 1) != 0) {
    const int sweeps = dev_option(OPT_JACOBI_SWEEPS, 20);
    if (dev_option(OPT_EIG_ONESIDED, 1) == 2)   // 2: the column-pair scheme (comparison; twice the shared-memory traffic)
      return M <= 32 ? launch_os_block<32, false>(R, M, T, nframes, G, u, w, sweeps, st) : launch_os_block<64, false>(R, M, T, nframes, G, u, w, sweeps, st);
    return M <= 32 ? launch_os_block<32, true>(R, M, T, nframes, G, u, w, sweeps, st) : launch_os_block<64, true>(R, M, T, nframes, G, u, w, sweeps, st);
  }
  const size_t smem = block_body_smem(M);
  cudaFuncSetAttribute(jacobi_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = min(nframes, sms * 4);
  jacobi_block_kernel<<<blocks, JB_THREADS, smem, st>>>(R, M, T, nframes, G, u, w, dev_option(OPT_JACOBI_SWEEPS, 20));
  return 1;
}

}  // namespace doa
