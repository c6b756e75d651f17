// oracle/doa_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  This file is the CPU oracle for the gr-doa DoA hot path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
// product (libdoa_cuda) never links, loads or calls anything in here.
//
// It restates, in plain C++ on top of the same BLAS/LAPACK entry points Armadillo dispatches to
// (cgemm, cheevd, cgeev; Armadillo itself is not vendored by the reference and is absent from this
// image -- reference pin: find_package(Armadillo "7.300"), CMakeLists.txt:107), the arithmetic of
//   lib/autocorrelate_impl.cc:47-118          -> oracle_autocorrelate*
//   lib/MUSIC_lin_array_impl.cc:47-150        -> oracle_music_tables, oracle_music
//   lib/rootMUSIC_linear_array_impl.cc:46-152 -> oracle_rootmusic
//   lib/find_local_max_impl.cc:47-194         -> oracle_find_local_max
// with the float/double mixing of the reference kept statement by statement.
//
// PARITY PINNING: the reference holds no golden vectors; its QA (python/qa_*.py) needs a live Octave
// and unseeded randn.  What is pinned (tests/test_oracle_pinning.py): the Octave model's statement of
// the covariance incl. the forward-backward 1/N quirk (examples/@wpi_twinrx_doa_testbench/
// autocorrelate.m:38-45), the QA known-answer cases (23/121/52 degrees within +-2 degrees) and the
// deterministic find-peaks cases (python/test00{1,2}_findpeaks.m).  Beyond those loose bounds parity
// is UNPINNED by the reference; the float64 twins below (*_f64) measure the oracle's own fp32 noise.
//
// LAPACK comes from the OpenBLAS bundled with scipy in this image, loaded with dlopen at run time.

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef std::complex<float> cf;
typedef std::complex<double> cd;

namespace {

typedef void (*cgemm_t)(const char*, const char*, const int*, const int*, const int*, const cf*, const cf*,
                        const int*, const cf*, const int*, const cf*, cf*, const int*);
typedef void (*cheevd_t)(const char*, const char*, const int*, cf*, const int*, float*, cf*, const int*, float*,
                         const int*, int*, const int*, int*);
typedef void (*zheevd_t)(const char*, const char*, const int*, cd*, const int*, double*, cd*, const int*, double*,
                         const int*, int*, const int*, int*);
typedef void (*cgeev_t)(const char*, const char*, const int*, cf*, const int*, cf*, cf*, const int*, cf*, const int*,
                        cf*, const int*, float*, int*);
typedef void (*zgeev_t)(const char*, const char*, const int*, cd*, const int*, cd*, cd*, const int*, cd*, const int*,
                        cd*, const int*, double*, int*);
typedef void (*setthreads_t)(int);

struct Lapack {
  void* handle = nullptr;
  cgemm_t cgemm = nullptr;
  cheevd_t cheevd = nullptr;
  zheevd_t zheevd = nullptr;
  cgeev_t cgeev = nullptr;
  zgeev_t zgeev = nullptr;
  setthreads_t set_threads = nullptr;
} L;

const double kPi = 3.14159265358979323846;  // arma::datum::pi

// Workspace-owning wrappers -------------------------------------------------------------------------
struct HeevdWork {
  std::vector<cf> work; std::vector<float> rwork; std::vector<int> iwork;
  std::vector<cd> zwork; std::vector<double> drwork;
};

// eig_sym(eigval, eigvec, X) for cx_fmat: Armadillo's default "dc" method -> cheevd, jobz 'V', uplo 'U'.
int heevd_f32(int M, cf* A, float* w, HeevdWork& ws) {
  int info = 0, lwork = -1, lrwork = -1, liwork = -1;
  cf wq; float rq; int iq;
  L.cheevd("V", "U", &M, A, &M, w, &wq, &lwork, &rq, &lrwork, &iq, &liwork, &info);
  lwork = (int)wq.real(); lrwork = (int)rq; liwork = iq;
  if ((int)ws.work.size() < lwork) ws.work.resize(lwork);
  if ((int)ws.rwork.size() < lrwork) ws.rwork.resize(lrwork);
  if ((int)ws.iwork.size() < liwork) ws.iwork.resize(liwork);
  L.cheevd("V", "U", &M, A, &M, w, ws.work.data(), &lwork, ws.rwork.data(), &lrwork, ws.iwork.data(), &liwork, &info);
  return info;
}
int heevd_f64(int M, cd* A, double* w, HeevdWork& ws) {
  int info = 0, lwork = -1, lrwork = -1, liwork = -1;
  cd wq; double rq; int iq;
  L.zheevd("V", "U", &M, A, &M, w, &wq, &lwork, &rq, &lrwork, &iq, &liwork, &info);
  lwork = (int)wq.real(); lrwork = (int)rq; liwork = iq;
  if ((int)ws.zwork.size() < lwork) ws.zwork.resize(lwork);
  if ((int)ws.drwork.size() < lrwork) ws.drwork.resize(lrwork);
  if ((int)ws.iwork.size() < liwork) ws.iwork.resize(liwork);
  L.zheevd("V", "U", &M, A, &M, w, ws.zwork.data(), &lwork, ws.drwork.data(), &lrwork, ws.iwork.data(), &liwork, &info);
  return info;
}

// eig_gen(eigval, X) for cx_fmat -> cgeev, values only.
int geev_f32(int n, cf* A, cf* w) {
  int info = 0, lwork = -1, one = 1;
  cf wq; std::vector<float> rwork(2 * n);
  L.cgeev("N", "N", &n, A, &n, w, nullptr, &one, nullptr, &one, &wq, &lwork, rwork.data(), &info);
  lwork = std::max(1, (int)wq.real());
  std::vector<cf> work(lwork);
  L.cgeev("N", "N", &n, A, &n, w, nullptr, &one, nullptr, &one, work.data(), &lwork, rwork.data(), &info);
  return info;
}
int geev_f64(int n, cd* A, cd* w) {
  int info = 0, lwork = -1, one = 1;
  cd wq; std::vector<double> rwork(2 * n);
  L.zgeev("N", "N", &n, A, &n, w, nullptr, &one, nullptr, &one, &wq, &lwork, rwork.data(), &info);
  lwork = std::max(1, (int)wq.real());
  std::vector<cd> work(lwork);
  L.zgeev("N", "N", &n, A, &n, w, nullptr, &one, nullptr, &one, work.data(), &lwork, rwork.data(), &info);
  return info;
}

// ---- stage 1 -------------------------------------------------------------------------------------
// One frame: X is N x M column-major (column k = channel k), out is M x M column-major.
// lib/autocorrelate_impl.cc:106-108.
void cov_one_frame(const cf* X, int M, int N, int avg_method, cf* out, std::vector<cf>& conjX, std::vector<cf>& tmp) {
  conjX.resize((size_t)N * M);
  for (size_t i = 0; i < (size_t)N * M; ++i) conjX[i] = std::conj(X[i]);           // conj(X) temporary
  const cf alpha((float)(1.0 / N), 0.0f), beta(0.0f, 0.0f);                          // scalar narrowed to float
  L.cgemm("T", "N", &M, &M, &N, &alpha, X, &N, conjX.data(), &N, &beta, out, &M);    // (1/N) * X.st() * conj(X)
  if (avg_method == 1) {
    // out = 0.5*out + (0.5/N) * J*conj(out)*J ; J = fliplr(eye).  J*A*J is the exact index reversal
    // (products with 0/1 are exact), so (J conj(out) J)(r,c) = conj(out(M-1-r, M-1-c)).  The extra 1/N
    // on the backward term is the reference's (and the Octave model's) quirk and is kept.
    tmp.assign(out, out + (size_t)M * M);
    const float half = 0.5f, hb = (float)(0.5 / N);
    for (int c = 0; c < M; ++c)
      for (int r = 0; r < M; ++r) {
        cf fwd = tmp[r + (size_t)c * M] * half;
        cf bwd = std::conj(tmp[(M - 1 - r) + (size_t)(M - 1 - c) * M]) * hb;
        out[r + (size_t)c * M] = fwd + bwd;
      }
  }
}

// ---- stage 2 tables (ctor of MUSIC_lin_array_impl, lib/MUSIC_lin_array_impl.cc:56-86) --------------
void music_tables(float norm_spacing, int M, int P, float* array_loc, float* theta_rad, cf* V /*M x P col-major*/) {
  for (int nn = 0; nn < M; ++nn) array_loc[nn] = (float)(norm_spacing * 0.5 * (M - 1 - 2 * nn));  // :60
  theta_rad[0] = 0.0f;
  float theta_prev = 0.0f, theta;
  for (int ii = 1; ii < P; ++ii) {
    theta = (float)(theta_prev + 180.0 / P);      // :69 double add narrowed to float every step
    theta_prev = theta;
    theta_rad[ii] = (float)(kPi * theta / 180.0);  // :71
  }
  for (int ii = 0; ii < P; ++ii) {
    // amv (:98-104): exp(i * (-1.0*2*pi*cos(theta) * array_loc)).  cos() on the float argument resolves to
    // the double overload in namespace gr::doa (no using-namespace-std) [ext]; the double scalar is narrowed
    // to float when it multiplies the fcolvec; exp(complex<float>(0, phi)) = (cosf(phi), sinf(phi)).
    const float s = (float)(-1.0 * 2 * kPi * std::cos((double)theta_rad[ii]));
    for (int nn = 0; nn < M; ++nn) {
      const float phi = s * array_loc[nn];
      V[nn + (size_t)ii * M] = cf(cosf(phi), sinf(phi));
    }
  }
}

// G = U_N * U_N^H from the M-T eigenvectors of smallest eigenvalue (eig_vec.cols(0, M-T-1)).
void noise_projector_f32(const cf* R, int M, int T, cf* G, std::vector<cf>& A, std::vector<float>& w, HeevdWork& ws) {
  A.assign(R, R + (size_t)M * M);
  w.resize(M);
  heevd_f32(M, A.data(), w.data(), ws);
  const int nn = M - T;
  const cf one(1.0f, 0.0f), zero(0.0f, 0.0f);
  L.cgemm("N", "C", &M, &M, &nn, &one, A.data(), &M, A.data(), &M, &zero, G, &M);
}

// ---- calibrate_lin_array (lib/calibrate_lin_array_impl.cc:46-134) ---------
// ctor (:57-74): array_loc as in MUSIC, v = amv(pi*pilot_angle/180) with the float theta argument of amv() (:84).
void calibrate_pilot_vector(float norm_spacing, int M, float pilot_angle, cf* v) {
  const float theta = (float)(kPi * pilot_angle / 180.0);                       // :70, narrowed by amv's float parameter
  const float s = (float)(-1.0 * 2 * kPi * std::cos((double)theta));           // :90, same overload note as music_tables
  for (int nn = 0; nn < M; ++nn) {
    const float loc = (float)(norm_spacing * 0.5 * (M - 1 - 2 * nn));          // :62
    const float phi = s * loc;
    v[nn] = cf(cosf(phi), sinf(phi));
  }
}
// work (:112-126): eig_sym(R) -> U_S = last column; W = diag(conj v) U_S U_S^H diag(v); eig_sym(W) -> last column.
void calibrate_one(const cf* R, int M, const cf* v, cf* out, std::vector<cf>& A, std::vector<cf>& W, std::vector<float>& w, HeevdWork& ws) {
  A.assign(R, R + (size_t)M * M);
  w.resize(M);
  heevd_f32(M, A.data(), w.data(), ws);
  const cf* us = A.data() + (size_t)(M - 1) * M;
  W.resize((size_t)M * M);
  for (int c = 0; c < M; ++c)
    for (int r = 0; r < M; ++r) W[r + (size_t)c * M] = std::conj(v[r]) * (us[r] * std::conj(us[c])) * v[c];
  heevd_f32(M, W.data(), w.data(), ws);
  std::copy(W.begin() + (size_t)(M - 1) * M, W.end(), out);
}

// ---- stage 4 (lib/find_local_max_impl.cc:80-165, lib/find_local_max_impl.h:53-56) -----------------
struct Packet { float val; unsigned idx; };

// index_max(): Armadillo's direct_max starts from the most negative float and keeps the FIRST strictly greater value.
unsigned index_max_first(const float* v, int len) {
  unsigned best = 0; float bv = -std::numeric_limits<float>::infinity();
  for (int i = 0; i < len; ++i) if (v[i] > bv) { bv = v[i]; best = i; }
  return best;
}

void local_peak_indices(const float* in, int len, int K, std::vector<unsigned>& pk) {
  pk.assign(K, 0u);
  if (K == 1) { pk[0] = index_max_first(in, len); return; }            // find_one_local_peak_indx
  const int nd = len - 1;
  std::vector<float> s(std::max(nd, 0));
  for (int i = 0; i < nd; ++i) {                                        // sign(diff(in)) :89
    const float d = in[i + 1] - in[i];
    s[i] = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : ((d == 0.0f) ? 0.0f : d));
  }
  std::vector<unsigned> flats;
  for (int i = 0; i < nd; ++i) if (s[i] == 0.0f) flats.push_back(i);    // :92
  for (int ii = (int)flats.size() - 1; ii >= 0; --ii) {                 // :94-107
    const unsigned nxt = std::min<unsigned>(flats[ii] + 1, (unsigned)(nd - 1));
    s[flats[ii]] = (s[nxt] >= 0.0f) ? 1.0f : -1.0f;
  }
  std::vector<unsigned> all_pk;                                         // find(diff(s) == -2) + 1  :114
  for (int i = 0; i + 1 < nd; ++i) if (s[i + 1] - s[i] == -2.0f) all_pk.push_back(i + 1);
  std::vector<Packet> pkts(all_pk.size());                              // sort_index(all_pks, "descend") :137
  for (size_t i = 0; i < all_pk.size(); ++i) { pkts[i].val = in[all_pk[i]]; pkts[i].idx = (unsigned)i; }
  std::sort(pkts.begin(), pkts.end(), [](const Packet& a, const Packet& b) { return a.val > b.val; });
  const unsigned nvalid = (unsigned)pkts.size();
  if (nvalid >= (unsigned)K) {
    for (int i = 0; i < K; ++i) pk[i] = all_pk[pkts[i].idx];            // :143
  } else {
    unsigned max_peak_ind;
    if (nvalid == 0) max_peak_ind = index_max_first(in, len);           // :149-150
    else max_peak_ind = pkts[0].idx;                                    // :152  (index into the PEAK LIST: reference bug, kept)
    for (unsigned ind = 0; ind < (unsigned)K; ++ind)
      pk[ind] = (ind < nvalid) ? all_pk[pkts[ind].idx] : max_peak_ind;  // :154-162
  }
}

void x_axis_table(int len, float x_min, float x_max, float* x) {        // :60-69, all float
  x[0] = x_min;
  float x_prev = x_min, xx;
  const float x_range = (x_max - x_min);
  for (int ii = 1; ii < len; ++ii) { xx = x_prev + x_range / len; x_prev = xx; x[ii] = xx; }
}

void find_local_max_one(const float* in, int len, int K, const float* xaxis, float* out_val, float* out_loc,
                        int* out_idx, std::vector<unsigned>& pk) {
  local_peak_indices(in, len, K, pk);
  for (int i = 0; i < K; ++i) { out_val[i] = in[pk[i]]; out_loc[i] = xaxis[pk[i]]; if (out_idx) out_idx[i] = (int)pk[i]; }
  std::sort(out_loc, out_loc + K, [](float a, float b) { return a > b; });   // sort(x_axis(pk), "descend") :188
}

// Per-frame MUSIC given tables (lib/MUSIC_lin_array_impl.cc:124-142).
void music_one(const cf* R, int M, int T, int P, const cf* V, float* out, std::vector<cf>& G, std::vector<cf>& A,
               std::vector<float>& w, HeevdWork& ws, std::vector<cf>& row) {
  G.resize((size_t)M * M); row.resize(M);
  noise_projector_f32(R, M, T, G.data(), A, w, ws);
  float vmax = -std::numeric_limits<float>::infinity();
  for (int ii = 0; ii < P; ++ii) {
    const cf* v = V + (size_t)ii * M;
    // (V_trans.row(ii) * U_N_sq) first (Armadillo's 3-term product keeps the cheaper-or-equal left pair), then * V.col(ii)
    for (int c = 0; c < M; ++c) {
      cf acc(0.0f, 0.0f);
      for (int r = 0; r < M; ++r) acc += std::conj(v[r]) * G[r + (size_t)c * M];
      row[c] = acc;
    }
    cf q(0.0f, 0.0f);
    for (int c = 0; c < M; ++c) q += row[c] * v[c];
    out[ii] = (float)(1.0 / q.real());            // :140 double division narrowed to float
    if (out[ii] > vmax) vmax = out[ii];
  }
  for (int ii = 0; ii < P; ++ii) out[ii] = 10.0f * log10f(out[ii] / vmax);   // :142
}

// Root-MUSIC per frame (lib/rootMUSIC_linear_array_impl.cc:68-87,105-145).
void rootmusic_one(const cf* R, int M, int T, float norm_spacing, float* out, cf* roots_out, std::vector<cf>& G,
                   std::vector<cf>& A, std::vector<float>& w, HeevdWork& ws) {
  G.resize((size_t)M * M);
  noise_projector_f32(R, M, T, G.data(), A, w, ws);
  const int n = 2 * M - 2;
  std::vector<cf> u(2 * M - 1);
  // sum(A.diag(ii)) is Armadillo's accu() of a vector-like view: two interleaved partial sums (even / odd positions), added at the end
  auto diag_sum = [&](int ii) {                    // ii <= 0: sub-diagonal (row = col - ii)
    const int n = M + ii;
    cf v1(0.0f, 0.0f), v2(0.0f, 0.0f);
    int i = 0, j = 1;
    for (; j < n; i += 2, j += 2) { v1 += G[(i - ii) + (size_t)i * M]; v2 += G[(j - ii) + (size_t)j * M]; }
    if (i < n) v1 += G[(i - ii) + (size_t)i * M];
    return v1 + v2;
  };
  for (int ii = -M + 1; ii < 0; ++ii) {          // :74-78
    const cf sacc = diag_sum(ii);
    u[ii + M - 1] = sacc;
    u[M - 1 - ii] = std::conj(sacc);
  }
  u[M - 1] = diag_sum(0);                          // :79
  // :80  gr_complex(-1, 0) / u(2M-2).  This file is built with -fcx-limited-range; the reference is not, so its complex<float>
  // division is libgcc's __divsc3, which (GCC >= 11) evaluates (ac + bd) / (cc + dd), (bc - ad) / (cc + dd) in double and narrows.
  cf scale;
  {
    const double a = -1.0, b = 0.0, c = u[2 * M - 2].real(), d = u[2 * M - 2].imag();
    const double denom = (c * c) + (d * d);
    scale = cf((float)(((a * c) + (b * d)) / denom), (float)(((b * c) - (a * d)) / denom));
  }
  for (auto& x : u) x = scale * x;
  std::vector<cf> comp((size_t)n * n, cf(0.0f, 0.0f));
  for (int i = 0; i + 1 < n; ++i) comp[(i + 1) + (size_t)i * n] = cf(1.0f, 0.0f);                       // :55-58
  for (int i = 0; i < n; ++i) comp[i + (size_t)(n - 1) * n] = u[i];                                     // :83
  std::vector<cf> roots(n);
  geev_f32(n, comp.data(), roots.data());                                                               // :86
  if (roots_out) std::copy(roots.begin(), roots.end(), roots_out);
  std::vector<cf> rin; std::vector<float> din;
  for (int i = 0; i < n; ++i) {                                                                        // :122-127
    const float dist = (float)(1.0 - (double)std::abs(roots[i]));   // 1.0 - abs(): double scalar minus float vec -> float
    if (dist > 0.0f) { rin.push_back(roots[i]); din.push_back(dist); }
  }
  std::vector<float> aoa(T, std::numeric_limits<float>::quiet_NaN());
  for (int ii = 0; ii < T; ++ii) {                                                                     // :131-141
    if (din.empty()) break;   // no root strictly inside the unit circle: the reference's index_min() throws (Armadillo: "object has no elements") -> NaN
    size_t mi = 0;            // index_min keeps the FIRST smallest value; once every entry is inf that is entry 0
    for (size_t k = 1; k < din.size(); ++k) if (din[k] < din[mi]) mi = k;
    // a consumed root is (inf, 0) (:140): arg = 0, acos(0) = pi/2 -> 90 degrees for every slot beyond the roots found
    aoa[ii] = (float)(180.0 * std::acos((double)std::arg(rin[mi]) / (2 * kPi * (double)norm_spacing)) / kPi);   // :136
    din[mi] = std::numeric_limits<float>::infinity();
    rin[mi] = cf(std::numeric_limits<float>::infinity(), 0.0f);
  }
  std::sort(aoa.begin(), aoa.end(), [](float a, float b) { return a < b; });                            // :144 (NaN unordered; only hit when undefined)
  for (int i = 0; i < T; ++i) out[i] = aoa[i];
}

}  // namespace

extern "C" {

// Load LAPACK. `prefix` is "" for stock OpenBLAS and "scipy_" for scipy's bundled build.
int oracle_init(const char* lapack_path, const char* prefix) {
  if (L.handle) return 0;
  L.handle = dlopen(lapack_path, RTLD_NOW | RTLD_LOCAL);
  if (!L.handle) { fprintf(stderr, "oracle_init: dlopen(%s) failed: %s\n", lapack_path, dlerror()); return -1; }
  const std::string p(prefix ? prefix : "");
  L.cgemm = (cgemm_t)dlsym(L.handle, (p + "cgemm_").c_str());
  L.cheevd = (cheevd_t)dlsym(L.handle, (p + "cheevd_").c_str());
  L.zheevd = (zheevd_t)dlsym(L.handle, (p + "zheevd_").c_str());
  L.cgeev = (cgeev_t)dlsym(L.handle, (p + "cgeev_").c_str());
  L.zgeev = (zgeev_t)dlsym(L.handle, (p + "zgeev_").c_str());
  L.set_threads = (setthreads_t)dlsym(L.handle, (p + "openblas_set_num_threads").c_str());
  if (!L.cgemm || !L.cheevd || !L.zheevd || !L.cgeev || !L.zgeev) { fprintf(stderr, "oracle_init: missing LAPACK symbols\n"); return -2; }
  if (L.set_threads) L.set_threads(1);   // parallelism is over frames (OpenMP), BLAS stays single-threaded
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Streaming form, exactly general_work(): M channel pointers, frame i of channel k starts at in[k] + i*hop.
int oracle_autocorrelate(const float* const* in, int M, int N, int overlap, int avg_method, int nframes, float* out,
                         int nthreads) {
  const int hop = N - overlap;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cf> X((size_t)N * M), conjX, tmp;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i) {
      for (int k = 0; k < M; ++k)   // memcpy framing :95-100
        std::memcpy(&X[(size_t)k * N], (const cf*)in[k] + (size_t)i * hop, sizeof(cf) * N);
      cov_one_frame(X.data(), M, N, avg_method, (cf*)out + (size_t)i * M * M, conjX, tmp);
    }
  }
  return 0;
}

// Independent frames [B][M][N] (each frame already in the d_input_matrix layout).
int oracle_autocorrelate_frames(const float* in, int M, int N, int avg_method, int nframes, float* out, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cf> conjX, tmp;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i)
      cov_one_frame((const cf*)in + (size_t)i * M * N, M, N, avg_method, (cf*)out + (size_t)i * M * M, conjX, tmp);
  }
  return 0;
}

int oracle_music_tables(float norm_spacing, int M, int P, float* array_loc, float* theta_rad, float* V) {
  music_tables(norm_spacing, M, P, array_loc, theta_rad, (cf*)V);
  return 0;
}

// R: [n][M*M] c64 col-major -> out: [n][P] f32 (dB, peak = 0).
int oracle_music(const float* R, int nframes, float norm_spacing, int T, int M, int P, float* out, int nthreads) {
  std::vector<float> loc(M), th(P); std::vector<cf> V((size_t)M * P);
  music_tables(norm_spacing, M, P, loc.data(), th.data(), V.data());
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cf> G, A, row; std::vector<float> w; HeevdWork ws;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i)
      music_one((const cf*)R + (size_t)i * M * M, M, T, P, V.data(), out + (size_t)i * P, G, A, w, ws, row);
  }
  return 0;
}

// float64 twin of stage 2: same fp32 R in, everything after in double with ideal steering at the ideal grid
// theta_i = i*180/P.  Outputs the un-normalised null spectrum Q (double) so deep nulls can be masked by the caller.
int oracle_music_f64(const float* R, int nframes, float norm_spacing, int T, int M, int P, double* Q, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cd> A((size_t)M * M), G((size_t)M * M); std::vector<double> w(M); HeevdWork ws;
    std::vector<cd> v(M);
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i) {
      const cf* Ri = (const cf*)R + (size_t)i * M * M;
      for (int k = 0; k < M * M; ++k) A[k] = cd(Ri[k].real(), Ri[k].imag());
      heevd_f64(M, A.data(), w.data(), ws);
      for (int c = 0; c < M; ++c)
        for (int r = 0; r < M; ++r) {
          cd acc(0, 0);
          for (int n = 0; n < M - T; ++n) acc += A[r + (size_t)n * M] * std::conj(A[c + (size_t)n * M]);
          G[r + (size_t)c * M] = acc;
        }
      for (int ii = 0; ii < P; ++ii) {
        const double theta = kPi * ((double)ii * 180.0 / P) / 180.0;
        const double s = -2.0 * kPi * std::cos(theta);
        for (int nn = 0; nn < M; ++nn) {
          const double phi = s * ((double)norm_spacing * 0.5 * (M - 1 - 2 * nn));
          v[nn] = cd(std::cos(phi), std::sin(phi));
        }
        cd q(0, 0);
        for (int c = 0; c < M; ++c) {
          cd acc(0, 0);
          for (int r = 0; r < M; ++r) acc += std::conj(v[r]) * G[r + (size_t)c * M];
          q += acc * v[c];
        }
        Q[(size_t)i * P + ii] = q.real();
      }
    }
  }
  return 0;
}

// The fp32 null spectrum Q itself as the reference forms it (before 1/Q, max, log10) -- for tolerance studies.
int oracle_music_q(const float* R, int nframes, float norm_spacing, int T, int M, int P, float* Q, int nthreads) {
  std::vector<float> loc(M), th(P); std::vector<cf> V((size_t)M * P);
  music_tables(norm_spacing, M, P, loc.data(), th.data(), V.data());
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cf> G((size_t)M * M), A, row(M); std::vector<float> w; HeevdWork ws;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i) {
      noise_projector_f32((const cf*)R + (size_t)i * M * M, M, T, G.data(), A, w, ws);
      for (int ii = 0; ii < P; ++ii) {
        const cf* v = V.data() + (size_t)ii * M;
        for (int c = 0; c < M; ++c) { cf acc(0, 0); for (int r = 0; r < M; ++r) acc += std::conj(v[r]) * G[r + (size_t)c * M]; row[c] = acc; }
        cf q(0, 0); for (int c = 0; c < M; ++c) q += row[c] * v[c];
        Q[(size_t)i * P + ii] = q.real();
      }
    }
  }
  return 0;
}

// G = U_N U_N^H (fp32 LAPACK) for subspace comparisons: [n][M*M] c64 col-major.
int oracle_noise_projector(const float* R, int nframes, int T, int M, float* G, float* eigvals, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cf> A; std::vector<float> w; HeevdWork ws;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i) {
      noise_projector_f32((const cf*)R + (size_t)i * M * M, M, T, (cf*)G + (size_t)i * M * M, A, w, ws);
      if (eigvals) std::copy(w.begin(), w.end(), eigvals + (size_t)i * M);
    }
  }
  return 0;
}
int oracle_noise_projector_f64(const float* R, int nframes, int T, int M, double* G, double* eigvals, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cd> A((size_t)M * M); std::vector<double> w(M); HeevdWork ws;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i) {
      const cf* Ri = (const cf*)R + (size_t)i * M * M;
      for (int k = 0; k < M * M; ++k) A[k] = cd(Ri[k].real(), Ri[k].imag());
      heevd_f64(M, A.data(), w.data(), ws);
      cd* Gi = (cd*)G + (size_t)i * M * M;
      for (int c = 0; c < M; ++c)
        for (int r = 0; r < M; ++r) {
          cd acc(0, 0);
          for (int n = 0; n < M - T; ++n) acc += A[r + (size_t)n * M] * std::conj(A[c + (size_t)n * M]);
          Gi[r + (size_t)c * M] = acc;
        }
      if (eigvals) std::copy(w.begin(), w.end(), eigvals + (size_t)i * M);
    }
  }
  return 0;
}

// R: [n][M*M] -> out [n][T] degrees ascending; roots (optional) [n][2M-2] c64 as cgeev returned them.
int oracle_rootmusic(const float* R, int nframes, float norm_spacing, int T, int M, float* out, float* roots, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cf> G, A; std::vector<float> w; HeevdWork ws;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i)
      rootmusic_one((const cf*)R + (size_t)i * M * M, M, T, norm_spacing, out + (size_t)i * T,
                    roots ? (cf*)roots + (size_t)i * (2 * M - 2) : nullptr, G, A, w, ws);
  }
  return 0;
}

// float64 twin of stage 3 (zheevd + zgeev on the same fp32 R; same selection rule evaluated in double).
// dist (optional) [n][T]: 1-|z| of the selected roots, in selection order -- the conditioning of the frame: a selected root
// closer to the circle than ~1e-3 is a nearly double root that float32 coefficients cannot place inside or outside reliably.
int oracle_rootmusic_f64(const float* R, int nframes, float norm_spacing, int T, int M, double* out, double* dist, int nthreads) {
  const int n = 2 * M - 2;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cd> A((size_t)M * M), G((size_t)M * M), u(2 * M - 1), comp((size_t)n * n), roots(n);
    std::vector<double> w(M); HeevdWork ws;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i) {
      const cf* Ri = (const cf*)R + (size_t)i * M * M;
      for (int k = 0; k < M * M; ++k) A[k] = cd(Ri[k].real(), Ri[k].imag());
      heevd_f64(M, A.data(), w.data(), ws);
      for (int c = 0; c < M; ++c)
        for (int r = 0; r < M; ++r) {
          cd acc(0, 0);
          for (int k = 0; k < M - T; ++k) acc += A[r + (size_t)k * M] * std::conj(A[c + (size_t)k * M]);
          G[r + (size_t)c * M] = acc;
        }
      for (int ii = -M + 1; ii < 0; ++ii) {
        cd sacc(0, 0);
        for (int c = 0; c < M + ii; ++c) sacc += G[(c - ii) + (size_t)c * M];
        u[ii + M - 1] = sacc; u[M - 1 - ii] = std::conj(sacc);
      }
      { cd sacc(0, 0); for (int c = 0; c < M; ++c) sacc += G[c + (size_t)c * M]; u[M - 1] = sacc; }
      const cd scale = cd(-1.0, 0.0) / u[2 * M - 2];
      for (auto& x : u) x = scale * x;
      std::fill(comp.begin(), comp.end(), cd(0, 0));
      for (int k = 0; k + 1 < n; ++k) comp[(k + 1) + (size_t)k * n] = cd(1, 0);
      for (int k = 0; k < n; ++k) comp[k + (size_t)(n - 1) * n] = u[k];
      geev_f64(n, comp.data(), roots.data());
      std::vector<double> din; std::vector<cd> rin;
      for (int k = 0; k < n; ++k) { const double dist = 1.0 - std::abs(roots[k]); if (dist > 0.0) { din.push_back(dist); rin.push_back(roots[k]); } }
      std::vector<double> aoa(T, std::numeric_limits<double>::quiet_NaN());
      for (int ii = 0; ii < T && !din.empty(); ++ii) {
        size_t mi = 0;
        for (size_t k = 1; k < din.size(); ++k) if (din[k] < din[mi]) mi = k;
        if (std::isinf(din[mi])) break;
        aoa[ii] = 180.0 * std::acos(std::arg(rin[mi]) / (2 * kPi * (double)norm_spacing)) / kPi;
        if (dist) dist[(size_t)i * T + ii] = din[mi];
        din[mi] = std::numeric_limits<double>::infinity();
      }
      std::sort(aoa.begin(), aoa.end(), [](double a, double b) { return a < b; });
      for (int k = 0; k < T; ++k) out[(size_t)i * T + k] = aoa[k];
    }
  }
  return 0;
}

// calibrate_lin_array: R [n][M*M] c64 col-major -> gain/phase estimate vectors [n][M] c64 (defined up to a unit-modulus factor).
int oracle_calibrate_lin_array(const float* R, int nframes, float norm_spacing, int M, float pilot_angle, float* out, int nthreads) {
  std::vector<cf> v(M);
  calibrate_pilot_vector(norm_spacing, M, pilot_angle, v.data());
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cf> A, W; std::vector<float> w; HeevdWork ws;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i)
      calibrate_one((const cf*)R + (size_t)i * M * M, M, v.data(), (cf*)out + (size_t)i * M, A, W, w, ws);
  }
  return 0;
}
int oracle_calibrate_pilot_vector(float norm_spacing, int M, float pilot_angle, float* v) {
  calibrate_pilot_vector(norm_spacing, M, pilot_angle, (cf*)v);
  return 0;
}
int oracle_x_axis(int len, float x_min, float x_max, float* x) { x_axis_table(len, x_min, x_max, x); return 0; }

// in [n][len] -> out_val [n][K] (descending by height), out_loc [n][K] (descending by x), out_idx [n][K] (optional, peak bins
// in out_val order -- not a block output, exported for bit-exact bin comparisons).
int oracle_find_local_max(const float* in, int nframes, int K, int len, float x_min, float x_max, float* out_val,
                          float* out_loc, int* out_idx, int nthreads) {
  std::vector<float> xaxis(len);
  x_axis_table(len, x_min, x_max, xaxis.data());
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<unsigned> pk;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i)
      find_local_max_one(in + (size_t)i * len, len, K, xaxis.data(), out_val + (size_t)i * K, out_loc + (size_t)i * K,
                         out_idx ? out_idx + (size_t)i * K : nullptr, pk);
  }
  return 0;
}

// The whole chain on independent frames [B][M][N]: autocorrelate -> MUSIC -> find_local_max, the way three chained
// blocks would run it (spectra are materialised per frame, as the blocks do).  Used as the timed CPU baseline.
int oracle_chain_frames(const float* in, int nframes, int M, int N, int avg_method, float norm_spacing, int T, int P,
                        int K, float* out_val, float* out_loc, int* out_idx, int nthreads) {
  std::vector<float> loc(M), th(P), xaxis(P); std::vector<cf> V((size_t)M * P);
  music_tables(norm_spacing, M, P, loc.data(), th.data(), V.data());
  x_axis_table(P, 0.0f, 180.0f, xaxis.data());
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<cf> conjX, tmp, R((size_t)M * M), G, A, row; std::vector<float> w, spec(P); HeevdWork ws;
    std::vector<unsigned> pk;
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i) {
      cov_one_frame((const cf*)in + (size_t)i * M * N, M, N, avg_method, R.data(), conjX, tmp);
      music_one(R.data(), M, T, P, V.data(), spec.data(), G, A, w, ws, row);
      find_local_max_one(spec.data(), P, K, xaxis.data(), out_val + (size_t)i * K, out_loc + (size_t)i * K,
                         out_idx ? out_idx + (size_t)i * K : nullptr, pk);
    }
  }
  return 0;
}

}  // extern "C"
