// root.cu -- stage 3: Root-MUSIC polynomial roots as eigenvalues of the Frobenius companion matrix.
//
// Replaces get_roots_polynomial() + the root selection of rootMUSIC_linear_array_impl::work
// (gr-doa lib/rootMUSIC_linear_array_impl.cc:68-87,119-145; LAPACK cgeev there).
//
// The polynomial sum_k a_k x^k, a_{M-1+l} = u_l, a_{M-1-l} = conj(u_l) (u_l = l-th diagonal sum of U_N U_N^H, :74-79)
// is normalised by -1/a_n (:80) into the last column of the (2M-2)x(2M-2) companion matrix (ones on the first
// sub-diagonal, :55-58,83).  The matrix is already upper Hessenberg, so its eigenvalues come from a shifted complex QR
// iteration with deflation, one thread per frame, in float64: the reference's float32 cgeev is itself only good to
// ~1e-2 degree on these nearly-double roots, float64 roots agree with a float64 LAPACK twin to ~1e-6 degree (tests).
// The working matrix lives in a frame-interleaved global scratch (element (i,j) of frame f at ((i*n+j)*stride + f)),
// so neighbouring threads touch neighbouring addresses; it stays L1/L2 resident for the sizes gr-doa uses (n = 6..30).
// Fast path (n <= 30): the same roots from an Aberth-Ehrlich simultaneous iteration on the monic polynomial itself --
// coefficients and iterates of a frame live in a thread-interleaved shared-memory array, no global scratch, 8-14 iterations
// of n Horner evaluations + n(n-1) reciprocal differences in float64 -- 80x faster than the QR at n = 14 (the QR is
// latency-bound on its 205 MB scratch).  A frame whose iteration has not settled to a 1e-10 relative step in 40 rounds
// (about 1 in 10^4 at seven sources) is redone by the QR kernel.
// Root selection repeats the reference's float arithmetic (:122-141): dist = 1 - |z| in float, strictly inside only,
// the T closest to the circle, angle = 180*acos(arg(z)/(2 pi d))/pi with the double math of :136, ascending sort.
#include "doa_internal.h"

namespace doa {
namespace {

struct cplx { double x, y; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cconj(cplx a) { return {a.x, -a.y}; }
__device__ __forceinline__ cplx cneg(cplx a) { return {-a.x, -a.y}; }
__device__ __forceinline__ double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  const double d = b.x * b.x + b.y * b.y;
  return {(a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d};
}
__device__ __forceinline__ cplx csqrt_(cplx a) {
  const double m = hypot(a.x, a.y);
  if (m == 0.0) return {0.0, 0.0};
  double re = sqrt(0.5 * (m + fabs(a.x)));
  double im = 0.5 * a.y / re;
  if (a.x < 0.0) { const double t = re; re = fabs(im); im = (a.y < 0.0) ? -t : t; }
  return {re, im};
}

// Root selection with the float arithmetic of the reference (:122-141): dist = 1 - |z| in float, strictly inside only, the T
// closest to the circle, angle = 180*acos(arg(z)/(2 pi d))/pi with the double math of :136, slots beyond the roots found = 90
// degrees, ascending sort (NaNs -- no root inside at all -- last).
template <class GetRoot>
__device__ __forceinline__ void emit_angles(GetRoot root, int n, int T, bool failed, float norm_spacing, float* of) {
  unsigned long long used0 = 0ull, used1 = 0ull;
  const double two_pi_d = 2.0 * 3.14159265358979323846 * (double)norm_spacing;
  for (int ii = 0; ii < T; ++ii) {
    float best = INFINITY; int bk = -1; float bre = 0.f, bim = 0.f;
    if (!failed) {
      for (int k = 0; k < n; ++k) {
        const bool used = (k < 64) ? ((used0 >> k) & 1ull) : ((used1 >> (k - 64)) & 1ull);
        if (used) continue;
        const cplx z = root(k);
        const float re = (float)z.x, im = (float)z.y;
        const float dist = 1.0f - hypotf(re, im);
        if (dist > 0.0f && dist < best) { best = dist; bk = k; bre = re; bim = im; }
      }
    }
    // No root strictly inside the unit circle at all: the reference's index_min() runs on an empty vector (Armadillo throws) --
    // NaN.  Fewer than T inside: a consumed root is overwritten with (inf, 0) and its distance with inf (:139-140), index_min of
    // an all-inf vector returns entry 0, arg(inf + 0j) = 0 and acos(0) = pi/2: every further slot is 90 degrees (:136).
    float aoa = (ii == 0 || failed) ? __int_as_float(0x7fc00000) : 90.0f;
    if (bk < 0 && ii > 0 && of[0] != of[0]) aoa = __int_as_float(0x7fc00000);   // the empty set stays NaN in every slot
    if (bk >= 0) {
      if (bk < 64) used0 |= 1ull << bk; else used1 |= 1ull << (bk - 64);
      aoa = (float)(180.0 * acos((double)atan2f(bim, bre) / two_pi_d) / 3.14159265358979323846);
    }
    // insertion into the ascending prefix of[0..ii) (NaNs stay at the end)
    int pos = ii;
    while (pos > 0 && !(of[pos - 1] <= aoa) && !(aoa != aoa)) { of[pos] = of[pos - 1]; --pos; }
    of[pos] = aoa;
  }
}

struct Hmat {
  double2* base; long long stride; int n;
  __device__ __forceinline__ cplx get(int i, int j) const { const double2 v = base[((long long)i * n + j) * stride]; return {v.x, v.y}; }
  __device__ __forceinline__ void set(int i, int j, cplx v) const { base[((long long)i * n + j) * stride] = make_double2(v.x, v.y); }
};

__global__ void __launch_bounds__(128)
rootmusic_kernel(const float2* __restrict__ u, int M, int T, float norm_spacing, int nframes, double2* __restrict__ scratch,
                 long long stride, float* __restrict__ out, int only_flagged) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  if (only_flagged && out[f * T] != -INFINITY) return;      // the Aberth kernel marks the frames it gave up on
  const int n = 2 * M - 2;
  Hmat H{scratch + f, stride, n};
  const float2* uf = u + f * M;

  // companion matrix
  const cplx an = {(double)uf[M - 1].x, (double)uf[M - 1].y};       // a_n = u_{M-1}
  const cplx scale = cdiv({-1.0, 0.0}, an);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) H.set(i, j, {(i == j + 1) ? 1.0 : 0.0, 0.0});
  for (int k = 0; k < n; ++k) {
    const int l = k - (M - 1);
    cplx a;
    if (l >= 0) a = {(double)uf[l].x, l == 0 ? 0.0 : (double)uf[l].y};
    else a = {(double)uf[-l].x, -(double)uf[-l].y};
    H.set(k, n - 1, cmul(scale, a));
  }

  // shifted QR with deflation on the active window [l, hi]
  const double eps = 2.220446049250313e-16;
  bool failed = false;
  int hi = n - 1, iter = 0;
  while (hi >= 0) {
    int l = hi;
    while (l > 0) {
      double s = cabs1(H.get(l - 1, l - 1)) + cabs1(H.get(l, l));
      if (s == 0.0) s = 1.0;
      if (cabs1(H.get(l, l - 1)) < eps * s) { H.set(l, l - 1, {0.0, 0.0}); break; }
      --l;
    }
    if (l == hi) { --hi; iter = 0; continue; }
    if (iter >= 60) { failed = true; break; }
    cplx sigma;
    if (iter == 10 || iter == 20 || iter == 40) {
      sigma = {fabs(H.get(hi, hi - 1).x) + (hi >= 2 ? fabs(H.get(hi - 1, hi - 2).x) : 0.0), 0.0};
    } else {   // Wilkinson: eigenvalue of the trailing 2x2 closest to its last diagonal entry
      const cplx a = H.get(hi - 1, hi - 1), b = H.get(hi - 1, hi), c = H.get(hi, hi - 1), d = H.get(hi, hi);
      const cplx hd = {0.5 * (a.x - d.x), 0.5 * (a.y - d.y)};
      const cplx disc = csqrt_(cadd(cmul(hd, hd), cmul(b, c)));
      const cplx mid = {0.5 * (a.x + d.x), 0.5 * (a.y + d.y)};
      const cplx e1 = cadd(mid, disc), e2 = csub(mid, disc);
      sigma = (cabs1(csub(e1, d)) <= cabs1(csub(e2, d))) ? e1 : e2;
    }
    ++iter;
    for (int i = l; i <= hi; ++i) H.set(i, i, csub(H.get(i, i), sigma));
    cplx pc = {1.0, 0.0}, ps = {0.0, 0.0};   // previous rotation
    for (int k = l; k < hi; ++k) {
      const cplx x = H.get(k, k), y = H.get(k + 1, k);
      const double r = sqrt(x.x * x.x + x.y * x.y + y.x * y.x + y.y * y.y);
      cplx c = {1.0, 0.0}, s = {0.0, 0.0};
      if (r > 0.0) { c = {x.x / r, x.y / r}; s = {y.x / r, y.y / r}; }
      // rows k, k+1 <- [[conj(c), conj(s)], [-s, c]] * rows
      for (int j = k; j <= hi; ++j) {
        const cplx a = H.get(k, j), b = H.get(k + 1, j);
        H.set(k, j, cadd(cmul(cconj(c), a), cmul(cconj(s), b)));
        H.set(k + 1, j, cadd(cmul(cneg(s), a), cmul(c, b)));
      }
      if (k > l) {   // columns k-1, k <- cols * [[pc, -conj(ps)], [ps, conj(pc)]], rows l..k
        for (int i = l; i <= k; ++i) {
          const cplx a = H.get(i, k - 1), b = H.get(i, k);
          H.set(i, k - 1, cadd(cmul(a, pc), cmul(b, ps)));
          H.set(i, k, cadd(cmul(cneg(a), cconj(ps)), cmul(b, cconj(pc))));
        }
      }
      pc = c; ps = s;
    }
    for (int i = l; i <= hi; ++i) {
      const cplx a = H.get(i, hi - 1), b = H.get(i, hi);
      H.set(i, hi - 1, cadd(cmul(a, pc), cmul(b, ps)));
      H.set(i, hi, cadd(cmul(cneg(a), cconj(ps)), cmul(b, cconj(pc))));
    }
    for (int i = l; i <= hi; ++i) H.set(i, i, cadd(H.get(i, i), sigma));
  }

  emit_angles([&](int k) { return H.get(k, k); }, n, T, failed, norm_spacing, out + f * T);
}

// ---- Aberth-Ehrlich -------------------------------------------------------------------------------------------------------
constexpr int AB_THREADS = 128;
constexpr int AB_MAX_N = 30;
constexpr int AB_MAX_ITER = 40;

__global__ void __launch_bounds__(AB_THREADS)
rootmusic_aberth_kernel(const float2* __restrict__ u, int M, int T, float norm_spacing, int nframes, float* __restrict__ out) {
  extern __shared__ double2 ab_s[];                  // [2n][AB_THREADS]: coefficients c_0..c_{n-1} (monic), then the iterates
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  const int n = 2 * M - 2;
  double2* c = ab_s + threadIdx.x;                   // c[k] at c[k * AB_THREADS]
  double2* z = ab_s + (size_t)n * AB_THREADS + threadIdx.x;
  const float2* uf = u + f * M;
  {
    const cplx an = {(double)uf[M - 1].x, (double)uf[M - 1].y};     // a_n = u_{M-1}
    const cplx inv = cdiv({1.0, 0.0}, an);
    for (int k = 0; k < n; ++k) {
      const int l = k - (M - 1);
      cplx a;
      if (l >= 0) a = {(double)uf[l].x, l == 0 ? 0.0 : (double)uf[l].y};
      else a = {(double)uf[-l].x, -(double)uf[-l].y};
      const cplx ck = cmul(inv, a);
      c[k * AB_THREADS] = make_double2(ck.x, ck.y);
    }
    // the roots come in pairs (z, 1/conj(z)) around the unit circle: start on two circles, irrational angle offset
    for (int k = 0; k < n; ++k) {
      double sn, cs;
      sincos(6.283185307179586 * (double)k / (double)n + 0.35, &sn, &cs);
      const double r = (k & 1) ? 1.15 : 0.85;
      z[k * AB_THREADS] = make_double2(r * cs, r * sn);
    }
  }
  bool settled = false;
  for (int it = 0; it < AB_MAX_ITER && !settled; ++it) {
    settled = true;
    for (int i = 0; i < n; ++i) {
      const double2 zi2 = z[i * AB_THREADS];
      const cplx zi = {zi2.x, zi2.y};
      cplx p = {1.0, 0.0}, dp = {0.0, 0.0};
      for (int k = n - 1; k >= 0; --k) {              // Horner: p(z_i), p'(z_i)
        dp = cadd(cmul(dp, zi), p);
        const double2 ck = c[k * AB_THREADS];
        p = cadd(cmul(p, zi), {ck.x, ck.y});
      }
      cplx sum = {0.0, 0.0};
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const double2 zj = z[j * AB_THREADS];
        const double dx = zi.x - zj.x, dy = zi.y - zj.y;
        const double inv = 1.0 / (dx * dx + dy * dy);
        sum.x += dx * inv; sum.y -= dy * inv;         // 1 / (z_i - z_j)
      }
      const cplx r = cdiv(p, dp);
      const cplx w = cdiv(r, csub({1.0, 0.0}, cmul(r, sum)));
      const cplx zn = csub(zi, w);
      z[i * AB_THREADS] = make_double2(zn.x, zn.y);   // Gauss-Seidel: later roots of this round already see it
      const double w2 = w.x * w.x + w.y * w.y, z2 = zn.x * zn.x + zn.y * zn.y;
      if (!(w2 <= 1e-20 * fmax(1.0, z2))) settled = false;     // also catches NaN
    }
  }
  if (!settled) { out[f * T] = -INFINITY; return; }  // left to the QR kernel
  emit_angles([&](int k) { const double2 v = z[k * AB_THREADS]; return cplx{v.x, v.y}; }, n, T, false, norm_spacing, out + f * T);
}

}  // namespace

// scratch: caller-provided, at least (2M-2)^2 * stride double2, stride >= nframes.
int launch_rootmusic_scratch(const float2* u, int M, int T, float norm_spacing, int nframes, double2* scratch,
                             long long stride, float* out, cudaStream_t st) {
  if (nframes <= 0) return 0;
  const int threads = 128;
  const int blocks = (nframes + threads - 1) / threads;
  const int n = 2 * M - 2;
  int only_flagged = 0;
  if (n >= 1 && n <= AB_MAX_N && dev_option(OPT_ROOT_ABERTH, 1)) {
    const size_t smem = (size_t)2 * n * AB_THREADS * sizeof(double2);
    cudaFuncSetAttribute(rootmusic_aberth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rootmusic_aberth_kernel<<<(nframes + AB_THREADS - 1) / AB_THREADS, AB_THREADS, smem, st>>>(u, M, T, norm_spacing, nframes, out);
    only_flagged = 1;
  }
  rootmusic_kernel<<<blocks, threads, 0, st>>>(u, M, T, norm_spacing, nframes, scratch, stride, out, only_flagged);
  return 1;
}

}  // namespace doa
