// doa_cuda.cu -- the C ABI of libdoa_cuda (include/doa_cuda.h): handles, constructor tables, staging, stage plumbing.
//
// One handle = one GNU Radio block instance: it owns its CUDA stream(s), device buffers sized for max_frames and the
// constructor tables.  Every entry point makes the handle's device current for the duration of the call and restores the
// caller's (GNU Radio may call work() from a thread other than the constructor's; torch or another library may own the thread's
// current device).  There is no global mutable state: options live in the handle, the error text of a failed *_create is
// thread-local.
#include "doa_internal.h"

#include <cmath>
#include <complex>
#include <fstream>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <climits>
#include <memory>
#include <mutex>
#include <thread>

namespace doa {

// ---- per-handle options ---------------------------------------------------------------------------------------------
// The calling thread's current option set: installed by Enter (below) for the duration of an ABI call.
static thread_local const Tuning* tl_tune = nullptr;
int dev_option(Opt key, int dflt) {
  const Tuning* t = tl_tune;
  return (t != nullptr && t->v[key] != OPT_UNSET) ? t->v[key] : dflt;
}

// ---- constructor tables (host) -----------------------------------------------------------------------------------
// MUSIC_lin_array_impl constructor, gr-doa lib/MUSIC_lin_array_impl.cc:56-86, with its float/double mixing:
//   array_loc[nn] = float(d*0.5*(M-1-2nn));  theta accumulates as float(theta_prev + 180.0/P);
//   theta_rad = float(pi*theta/180.0);  s = float(-1.0*2*pi*cos(theta_rad));  V[nn] = (cosf, sinf)(s*array_loc[nn]).
// z is not in the reference: it is the per-element phase step of that table, e^{-j s d}, used by the coarse scan.
void build_music_tables(float norm_spacing, int M, int P, std::vector<float>& array_loc, std::vector<float>& theta,
                        std::vector<float2>& V, std::vector<float2>& z) {
  const double pi = 3.14159265358979323846;
  array_loc.resize(M); theta.resize(P); V.resize((size_t)M * P); z.resize(P);
  for (int nn = 0; nn < M; ++nn) array_loc[nn] = (float)(norm_spacing * 0.5 * (M - 1 - 2 * nn));
  theta[0] = 0.0f;
  float prev = 0.0f;
  for (int ii = 1; ii < P; ++ii) {
    const float th = (float)(prev + 180.0 / P);
    prev = th;
    theta[ii] = (float)(pi * th / 180.0);
  }
  for (int ii = 0; ii < P; ++ii) {
    const float s = (float)(-1.0 * 2 * pi * std::cos((double)theta[ii]));
    for (int nn = 0; nn < M; ++nn) {
      const float phi = s * array_loc[nn];
      V[(size_t)ii * M + nn] = make_float2(cosf(phi), sinf(phi));
    }
    const double psi = -(double)s * (double)norm_spacing;
    z[ii] = make_float2((float)std::cos(psi), (float)std::sin(psi));
  }
}

// find_local_max_impl constructor, lib/find_local_max_impl.cc:60-69 (all float, accumulated).
void build_x_axis(int len, float x_min, float x_max, std::vector<float>& x) {
  x.resize(len);
  x[0] = x_min;
  float prev = x_min;
  const float range = x_max - x_min;
  for (int ii = 1; ii < len; ++ii) { const float v = prev + range / len; prev = v; x[ii] = v; }
}

}  // namespace doa

using namespace doa;

enum Kind { K_AUTOCORR = 1, K_MUSIC, K_ROOTMUSIC, K_FLM, K_CHAIN, K_CALIB, K_MULTI, K_ROOTCHAIN };
static bool takes_samples(int kind) { return kind == K_AUTOCORR || kind == K_CHAIN || kind == K_ROOTCHAIN; }

struct Lane {   // one stream's worth of buffers (the chain's host path double-buffers two of these)
  cudaStream_t stream = nullptr;
  float2* in = nullptr; size_t in_elems = 0;   // staged samples
  float2* R = nullptr; float2* G = nullptr; float2* u = nullptr;
  float* spec = nullptr; float* val = nullptr; float* loc = nullptr; int* bin = nullptr; float* aoa = nullptr;
  float* vecs = nullptr;
  double2* scratch = nullptr;
  int frames = 0;
  char* pin = nullptr; size_t pin_bytes = 0;   // page-locked host staging for small host-pointer calls (allocated on first use)
  void* tc_ws = nullptr;                       // 64 elements: workspace of the tensor-core HERK's split tail (allocated on first use)
};

// Host threads of a multi-device handle: device 0 is driven by the calling thread, every further device by one persistent
// worker (a GNU Radio work() call is short: spawning threads per call would cost as much as the call).  One run at a time,
// like every handle (a block's work() is called serially).
struct MultiPool {
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  const std::function<int(int)>* job = nullptr;
  unsigned long long epoch = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rcs;
  std::vector<std::thread> threads;
  explicit MultiPool(int G) : rcs(G, 0) {
    for (int g = 1; g < G; ++g) threads.emplace_back([this, g] { loop(g); });
  }
  void loop(int g) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<int(int)>* j = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return stop || epoch != seen; });
        if (stop) return;
        seen = epoch;
        j = job;
      }
      const int rc = (*j)(g);
      std::lock_guard<std::mutex> lk(mu);
      rcs[g] = rc;
      if (--pending == 0) cv_done.notify_one();
    }
  }
  void run(const std::function<int(int)>& j) {
    {
      std::lock_guard<std::mutex> lk(mu);
      job = &j;
      pending = (int)threads.size();
      ++epoch;
    }
    cv_job.notify_all();
    rcs[0] = j(0);
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  ~MultiPool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv_job.notify_all();
    for (auto& t : threads) t.join();
  }
};

struct doa_cuda_handle {
  Tuning tune;
  int kind = 0, device = 0, max_frames = 0;
  int M = 0, N = 0, overlap = 0, hop = 0, avg = 0, T = 0, P = 0, K = 0;
  float d = 0.f, x_min = 0.f, x_max = 0.f;
  Lane lane[2];
  int nlanes = 1;
  float2* d_z = nullptr; float2* d_V = nullptr; float* d_x = nullptr; float* d_zpair = nullptr; float* d_tctab = nullptr;
  float2* d_gains = nullptr;     // per-channel complex gains folded into the covariance (null: none)
  InputFormat fmt;               // sample format of the covariance input (fc32 unless doa_cuda_set_input_format said sc16)
  size_t sample_bytes() const { return fmt.sc16 ? 4 : sizeof(float2); }
  std::vector<float> h_loc, h_theta, h_x; std::vector<float2> h_V, h_z;
  std::string err;
  int launches = 0;
  bool profiling = false;
  std::vector<doa_cuda_handle*> children;   // K_MULTI: one chain handle per listed device
  std::unique_ptr<MultiPool> pool;          // K_MULTI: its host threads
  std::vector<cudaEvent_t> ev;   // profiling: 4 events per recorded chain call (ring of PROF_SETS calls)
  int prof_calls = 0;
};
static const int PROF_SETS = 256;

static thread_local std::string g_create_err;

// Scope of one ABI call on a handle: the handle's device becomes current (GNU Radio may call work() from a thread other than the
// constructor's) and is put back on return, so a call never changes the calling thread's device under torch or another
// library; the handle's options become the thread's current ones.
struct Enter {
  int prev_dev = -1; const Tuning* prev_tune; bool ok;
  explicit Enter(doa_cuda_handle* h) : prev_tune(tl_tune) {
    if (cudaGetDevice(&prev_dev) != cudaSuccess) prev_dev = -1;
    ok = cudaSetDevice(h->device) == cudaSuccess;
    tl_tune = &h->tune;
  }
  ~Enter() {
    tl_tune = prev_tune;
    if (prev_dev >= 0 && ok) cudaSetDevice(prev_dev);
  }
};
struct DevSave {   // create / destroy: leave the calling thread's current device as it was
  int prev = -1;
  DevSave() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
  ~DevSave() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ENTER(h)                                                                  \
  Enter enter_guard_(h);                                                          \
  if (!enter_guard_.ok) return fail(h, DOA_CUDA_ECUDA, "cudaSetDevice failed")

#define CK(h, call)                                                                                    \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                   \
      return DOA_CUDA_ECUDA;                                                                           \
    }                                                                                                  \
  } while (0)

static int fail(doa_cuda_handle* h, int code, const std::string& msg) { if (h) h->err = msg; else g_create_err = msg; return code; }

static void free_lane(Lane& l) {
  cudaFree(l.in); cudaFree(l.R); cudaFree(l.G); cudaFree(l.u); cudaFree(l.spec); cudaFree(l.val); cudaFree(l.loc);
  cudaFree(l.bin); cudaFree(l.aoa); cudaFree(l.vecs); cudaFree(l.scratch); cudaFree(l.tc_ws);
  if (l.pin) cudaFreeHost(l.pin);
  if (l.stream) cudaStreamDestroy(l.stream);
  l = Lane();
}

extern "C" void doa_cuda_destroy(doa_cuda_handle* h) {
  if (!h) return;
  h->pool.reset();                          // workers are idle between runs; join them before their devices' handles go
  for (doa_cuda_handle* c : h->children) doa_cuda_destroy(c);
  h->children.clear();
  DevSave dev_save_;
  cudaSetDevice(h->device);
  for (int i = 0; i < 2; ++i) free_lane(h->lane[i]);
  cudaFree(h->d_z); cudaFree(h->d_V); cudaFree(h->d_x); cudaFree(h->d_zpair); cudaFree(h->d_tctab); cudaFree(h->d_gains);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  delete h;
}

static int begin_create(doa_cuda_handle** out, doa_cuda_handle*& h, int kind, int device, int max_frames) {
  if (!out) return fail(nullptr, DOA_CUDA_EINVAL, "null handle pointer");
  *out = nullptr;
  if (max_frames < 1) return fail(nullptr, DOA_CUDA_EINVAL, "max_frames must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(nullptr, DOA_CUDA_ECUDA, "no CUDA device available (libdoa_cuda has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(nullptr, DOA_CUDA_EINVAL, "device index out of range");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, DOA_CUDA_ECUDA, "cudaSetDevice failed");
  h = new doa_cuda_handle();
  h->kind = kind; h->device = device; h->max_frames = max_frames;
  return DOA_CUDA_OK;
}

template <typename Tp>
static bool dalloc(Tp** p, size_t n) { return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(Tp)) == cudaSuccess; }

// 64 elements: the tensor-core HERK's split-tail workspace (counters zeroed once; the kernel leaves them zero), one per lane
// because the chain's two lanes run on two streams.
static bool alloc_tc_ws(const doa_cuda_handle* h, Lane& l) {
  if (h->M != 64) return true;
  const size_t n = covariance_tc_workspace_bytes();
  return cudaMalloc(&l.tc_ws, n) == cudaSuccess && cudaMemset(l.tc_ws, 0, n) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
}

static int finish_create(doa_cuda_handle** out, doa_cuda_handle* h, bool ok) {
  if (!ok) {
    g_create_err = std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    doa_cuda_destroy(h);
    return DOA_CUDA_ENOMEM;
  }
  *out = h;
  return DOA_CUDA_OK;
}

static bool upload_scan_tables(doa_cuda_handle* h, bool need_steering, bool need_x) {
  bool ok = true;
  if (need_steering) {
    build_music_tables(h->d, h->M, h->P, h->h_loc, h->h_theta, h->h_V, h->h_z);
    std::vector<float> zpair;
    build_zpair_table(h->h_z, zpair);
    ok = ok && dalloc(&h->d_z, (size_t)h->P) && dalloc(&h->d_V, (size_t)h->P * h->M) && dalloc(&h->d_zpair, zpair.size());
    if (ok) ok = cudaMemcpy(h->d_zpair, zpair.data(), sizeof(float) * zpair.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) {
      ok = cudaMemcpy(h->d_z, h->h_z.data(), sizeof(float2) * h->P, cudaMemcpyHostToDevice) == cudaSuccess &&
           cudaMemcpy(h->d_V, h->h_V.data(), sizeof(float2) * (size_t)h->P * h->M, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (ok && need_x && scan_tc_covers(h->M, h->P, h->K)) {   // peak-picking plans only: the tensor-core scan's table
      std::vector<float> tc;
      build_scan_tc_table(h->d, h->M, h->P, h->h_theta, tc);
      ok = dalloc(&h->d_tctab, tc.size()) &&
           cudaMemcpy(h->d_tctab, tc.data(), sizeof(float) * tc.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    }
  }
  if (need_x && ok) {
    const int len = h->P;
    build_x_axis(len, h->x_min, h->x_max, h->h_x);
    ok = dalloc(&h->d_x, (size_t)len) &&
         cudaMemcpy(h->d_x, h->h_x.data(), sizeof(float) * len, cudaMemcpyHostToDevice) == cudaSuccess;
  }
  return ok;
}

static ScanTables tables_of(const doa_cuda_handle* h) {
  ScanTables t; t.M = h->M; t.P = h->P; t.z = h->d_z; t.zpair = h->d_zpair; t.V = h->d_V; t.xaxis = h->d_x; t.tctab = reinterpret_cast<const uint8_t*>(h->d_tctab); return t;
}

extern "C" {

int doa_cuda_abi_version(void) { return 1; }
const char* doa_cuda_last_error(const doa_cuda_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }
int doa_cuda_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
int doa_cuda_last_launch_count(const doa_cuda_handle* h) { return h ? h->launches : 0; }
int doa_cuda_set_option(doa_cuda_handle* h, const char* key, int value) {
  static const struct { const char* name; Opt opt; } kNames[] = {
      {"fused", OPT_FUSED}, {"scan_tc", OPT_SCAN_TC}, {"sms_reserve", OPT_SMS_RESERVE}, {"cov_groups", OPT_COV_GROUPS},
      {"cov16_ring", OPT_COV16_RING}, {"herk_tc", OPT_HERK_TC}, {"scan_wide", OPT_SCAN_WIDE}, {"spectrum_smem", OPT_SPECTRUM_SMEM},
      {"root_aberth", OPT_ROOT_ABERTH}, {"jacobi_sweeps", OPT_JACOBI_SWEEPS}, {"tma", OPT_TMA}, {"eig_onesided", OPT_EIG_ONESIDED}, {"herk_split", OPT_HERK_SPLIT}, {"ws_split", OPT_WS_SPLIT}, {"ws_stages", OPT_WS_STAGES},
      {"ws_nbuf", OPT_WS_NBUF}, {"ws4", OPT_WS4}, {"ws_tma", OPT_WS_TMA}, {"ws_fill", OPT_WS_FILL}, {"scan_tc_dbg", OPT_SCAN_TC_DBG}, {"fused16", OPT_FUSED16}};
  if (!h || !key) return DOA_CUDA_EINVAL;
  for (const auto& n : kNames) {
    if (std::strcmp(n.name, key) != 0) continue;
#ifndef DOA_DEV_KNOBS
    if ((int)n.opt >= OPT_FIRST_DEV_ONLY) return fail(h, DOA_CUDA_EINVAL, std::string("option '") + key + "' selects a kernel variant that only a -DDOA_DEV_KNOBS build contains");
#endif
    if (n.opt == OPT_SMS_RESERVE && (value < 0 || value > 64)) return fail(h, DOA_CUDA_EINVAL, "sms_reserve must be in [0, 64]");
    h->tune.v[n.opt] = value;
    for (doa_cuda_handle* c : h->children) c->tune.v[n.opt] = value;
    return DOA_CUDA_OK;
  }
  return fail(h, DOA_CUDA_EINVAL, std::string("unknown option '") + key + "'");
}
int doa_cuda_has_dev_knobs(void) {
#ifdef DOA_DEV_KNOBS
  return 1;
#else
  return 0;
#endif
}

// ---- page-locking a caller's host buffer -----------------------------------------------------------------------------
int doa_cuda_pin_host_buffer(void* p, unsigned long long bytes) {
  if (!p || bytes == 0) return fail(nullptr, DOA_CUDA_EINVAL, "null or empty host range");
  const cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return DOA_CUDA_OK; }
  if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DOA_CUDA_ECUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
  return DOA_CUDA_OK;
}
int doa_cuda_unpin_host_buffer(void* p) {
  if (!p) return fail(nullptr, DOA_CUDA_EINVAL, "null host pointer");
  const cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DOA_CUDA_ECUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); }
  return DOA_CUDA_OK;
}

// ---- channel gains ------------------------------------------------------------------------------------------------
int doa_cuda_set_channel_gains(doa_cuda_handle* h, const float* gains) {
  if (h && h->kind == K_MULTI) {
    for (doa_cuda_handle* c : h->children) {
      const int rc = doa_cuda_set_channel_gains(c, gains);
      if (rc) return fail(h, rc, c->err);
    }
    return DOA_CUDA_OK;
  }
  if (!h || !takes_samples(h->kind)) return DOA_CUDA_EINVAL;
  ENTER(h);
  // runs already queued on the handle's streams may still read the old gains
  for (int i = 0; i < h->nlanes; ++i) if (h->lane[i].stream) CK(h, cudaStreamSynchronize(h->lane[i].stream));
  if (gains == nullptr) { CK(h, cudaDeviceSynchronize()); cudaFree(h->d_gains); h->d_gains = nullptr; return DOA_CUDA_OK; }
  for (int k = 0; k < 2 * h->M; ++k)
    if (!std::isfinite(gains[k])) return fail(h, DOA_CUDA_EINVAL, "channel gains must be finite");
  if (!h->d_gains && cudaMalloc(&h->d_gains, (size_t)h->M * sizeof(float2)) != cudaSuccess) return fail(h, DOA_CUDA_ENOMEM, "cudaMalloc(gains)");
  CK(h, cudaMemcpy(h->d_gains, gains, (size_t)h->M * sizeof(float2), cudaMemcpyHostToDevice));
  return DOA_CUDA_OK;
}

// ---- input sample format ------------------------------------------------------------------------------------------
int doa_cuda_set_input_format(doa_cuda_handle* h, int format, float scale) {
  if (h && h->kind == K_MULTI) {
    for (doa_cuda_handle* c : h->children) {
      const int rc = doa_cuda_set_input_format(c, format, scale);
      if (rc) return fail(h, rc, c->err);
    }
    h->fmt = h->children[0]->fmt;
    return DOA_CUDA_OK;
  }
  if (!h || !takes_samples(h->kind)) return DOA_CUDA_EINVAL;
  if (format != DOA_CUDA_FMT_FC32 && format != DOA_CUDA_FMT_SC16) return fail(h, DOA_CUDA_EINVAL, "unknown input format");
  if (format == DOA_CUDA_FMT_SC16 && !(std::isfinite(scale) && scale > 0.0f))
    return fail(h, DOA_CUDA_EINVAL, "sc16 scale must be finite and > 0");
  ENTER(h);
  for (int i = 0; i < h->nlanes; ++i) if (h->lane[i].stream) CK(h, cudaStreamSynchronize(h->lane[i].stream));
  h->fmt.sc16 = format == DOA_CUDA_FMT_SC16;
  h->fmt.scale = h->fmt.sc16 ? scale : 1.0f;
  return DOA_CUDA_OK;
}

// lib/antenna_correction_impl.cc:54-74, statement by statement: float gain/phase pairs, g = gr_complex(1.0/Gain, 0) * exp(gr_complex(0, -Phase))
int doa_cuda_antenna_gains_from_file(const char* config_filename, int num_ant_ele, float* gains_out) {
  if (!config_filename || !gains_out || num_ant_ele < 1) return fail(nullptr, DOA_CUDA_EINVAL, "bad arguments");
  std::ifstream infile(config_filename);
  if (!infile.good()) return fail(nullptr, DOA_CUDA_EINVAL, "Cannot find configuration file.");
  float GainEst, PhaseEst;
  int i = 0;
  while (infile >> GainEst >> PhaseEst) {
    if (i >= num_ant_ele) return fail(nullptr, DOA_CUDA_EINVAL, "Configuration file has too many inputs.");
    const std::complex<float> g = std::complex<float>((float)(1.0 / GainEst), 0.0f) * std::exp(std::complex<float>(0.0f, -PhaseEst));
    gains_out[2 * i] = g.real(); gains_out[2 * i + 1] = g.imag();
    ++i;
  }
  if (i != num_ant_ele) return fail(nullptr, DOA_CUDA_EINVAL, "Configuration file does not have enough inputs.");
  return DOA_CUDA_OK;
}

// ---- autocorrelate ------------------------------------------------------------------------------------------------
int doa_cuda_autocorrelate_create(doa_cuda_handle** out, int inputs, int snapshot_size, int overlap_size, int avg_method,
                                  int device, int max_frames) {
  if (inputs < 1 || inputs > 64) return fail(nullptr, DOA_CUDA_EINVAL, "inputs must be in [1, 64]");
  if (snapshot_size < 1) return fail(nullptr, DOA_CUDA_EINVAL, "snapshot_size must be > 0");
  if (overlap_size < 0 || overlap_size >= snapshot_size) return fail(nullptr, DOA_CUDA_EINVAL, "need 0 <= overlap_size < snapshot_size");
  if (avg_method != 0 && avg_method != 1) return fail(nullptr, DOA_CUDA_EINVAL, "avg_method must be 0 (forward) or 1 (forward-backward)");
  doa_cuda_handle* h = nullptr;
  DevSave dev_save_;
  int rc = begin_create(out, h, K_AUTOCORR, device, max_frames);
  if (rc) return rc;
  h->M = inputs; h->N = snapshot_size; h->overlap = overlap_size; h->hop = snapshot_size - overlap_size; h->avg = avg_method;
  Lane& l = h->lane[0];
  const size_t Lpad = (((size_t)(max_frames - 1) * h->hop + h->N) + 1) & ~(size_t)1;
  l.in_elems = Lpad * h->M;
  bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess && dalloc(&l.in, l.in_elems) &&
            dalloc(&l.R, (size_t)max_frames * h->M * h->M) && alloc_tc_ws(h, l);
  return finish_create(out, h, ok);
}

int doa_cuda_autocorrelate_forecast(const doa_cuda_handle* h, int noutput_items) {
  if (!h || h->kind != K_AUTOCORR) return DOA_CUDA_EINVAL;
  return h->hop * noutput_items;   // lib/autocorrelate_impl.cc:79
}

int doa_cuda_autocorrelate_run_device(doa_cuda_handle* h, const void* in_dev, long long frame_stride, long long chan_stride,
                                      int nframes, void* out_dev, void* cuda_stream) {
  if (!h || (h->kind != K_AUTOCORR && h->kind != K_CHAIN)) return DOA_CUDA_EINVAL;
  if (nframes < 0) return fail(h, DOA_CUDA_EINVAL, "nframes < 0");
  ENTER(h);
  int n = launch_covariance(in_dev, frame_stride, chan_stride, h->M, h->N, nframes, h->avg, (float2*)out_dev,
                            (cudaStream_t)cuda_stream, h->d_gains, h->fmt, h->lane[0].tc_ws);
  if (n < 0) return fail(h, n, "covariance launch rejected");
  h->launches = n;
  CK(h, cudaGetLastError());
  return DOA_CUDA_OK;
}

// ---- page-locked staging of small host-pointer calls --------------------------------------------------------------------
// A GNU Radio scheduler hands work() pageable buffers and a handful of frames.  Copied straight out of pageable memory every
// channel stream (and every output array) is its own driver-staged, blocking transfer; here the call's samples are first
// gathered into the handle's page-locked buffer in the device layout, so that ONE asynchronous copy moves them, and the
// outputs come back through the same buffer.  Measured (tools/latency.py, B200 box): 81 -> 60 us for a one-frame call at the
// cfg1 shape, even at 1 MB per call, and SLOWER beyond (387 against 326 us at 4 MB: a single host thread's memcpy is no match for
// the driver's own pageable path), hence the 1 MiB limit; larger transfers keep the direct route.  Caller memory that already is
// page-locked (cudaHostAlloc, doa_cuda_pin_host_buffer) is never staged.
static const size_t PIN_STAGE_BYTES = 1u << 20;

static bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}
static char* pin_staging(Lane& l, size_t bytes) {
  if (bytes > PIN_STAGE_BYTES) return nullptr;
  if (l.pin == nullptr) {
    if (cudaHostAlloc((void**)&l.pin, PIN_STAGE_BYTES, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); l.pin = nullptr; return nullptr; }
    l.pin_bytes = PIN_STAGE_BYTES;
  }
  return l.pin;
}

// Copy `inputs` host channel streams into lane.in as [M][Lpad]; returns Lpad.
static int stage_streams(doa_cuda_handle* h, Lane& l, const void* const* in_host, int nframes, size_t* Lpad_out) {
  const size_t L = (size_t)(nframes - 1) * h->hop + h->N;
  const size_t Lpad = (L + 1) & ~(size_t)1;
  if (Lpad * h->M > l.in_elems) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds the staging capacity of this handle (max_frames)");
  const size_t sb = h->sample_bytes();
  char* pin = is_pageable(in_host[0]) ? pin_staging(l, Lpad * h->M * sb) : nullptr;
  if (pin != nullptr) {
    CK(h, cudaStreamSynchronize(l.stream));                 // the previous call's copies out of the staging buffer are done
    for (int k = 0; k < h->M; ++k) std::memcpy(pin + (size_t)k * Lpad * sb, in_host[k], L * sb);
    CK(h, cudaMemcpyAsync(l.in, pin, Lpad * h->M * sb, cudaMemcpyHostToDevice, l.stream));
  } else {
    for (int k = 0; k < h->M; ++k)
      CK(h, cudaMemcpyAsync((char*)l.in + (size_t)k * Lpad * sb, in_host[k], L * sb, cudaMemcpyHostToDevice, l.stream));
  }
  *Lpad_out = Lpad;
  return DOA_CUDA_OK;
}

// Peaks of one call back to the caller's arrays: through the staging buffer (one synchronisation, then plain copies) when the
// destination is pageable and small, directly otherwise.  Synchronises the lane's stream.
static int peaks_to_host(doa_cuda_handle* h, Lane& l, int nframes, void* out_val_host, void* out_loc_host, void* out_bin_host) {
  const size_t nk = (size_t)nframes * h->K, bytes = sizeof(float) * nk;
  char* pin = is_pageable(out_val_host) ? pin_staging(l, 3 * bytes) : nullptr;
  if (pin != nullptr) {
    CK(h, cudaMemcpyAsync(pin, l.val, bytes, cudaMemcpyDeviceToHost, l.stream));
    CK(h, cudaMemcpyAsync(pin + bytes, l.loc, bytes, cudaMemcpyDeviceToHost, l.stream));
    if (out_bin_host) CK(h, cudaMemcpyAsync(pin + 2 * bytes, l.bin, bytes, cudaMemcpyDeviceToHost, l.stream));
    CK(h, cudaStreamSynchronize(l.stream));
    std::memcpy(out_val_host, pin, bytes);
    std::memcpy(out_loc_host, pin + bytes, bytes);
    if (out_bin_host) std::memcpy(out_bin_host, pin + 2 * bytes, bytes);
    return DOA_CUDA_OK;
  }
  CK(h, cudaMemcpyAsync(out_val_host, l.val, bytes, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaMemcpyAsync(out_loc_host, l.loc, bytes, cudaMemcpyDeviceToHost, l.stream));
  if (out_bin_host) CK(h, cudaMemcpyAsync(out_bin_host, l.bin, bytes, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaStreamSynchronize(l.stream));
  return DOA_CUDA_OK;
}

int doa_cuda_autocorrelate_run(doa_cuda_handle* h, const void* const* in_host, int nframes, void* out_host) {
  if (!h || h->kind != K_AUTOCORR) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  Lane& l = h->lane[0];
  size_t Lpad = 0;
  int rc = stage_streams(h, l, in_host, nframes, &Lpad);
  if (rc) return rc;
  rc = doa_cuda_autocorrelate_run_device(h, l.in, h->hop, (long long)Lpad, nframes, l.R, l.stream);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(out_host, l.R, sizeof(float2) * (size_t)nframes * h->M * h->M, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaStreamSynchronize(l.stream));
  return DOA_CUDA_OK;
}

// ---- MUSIC -----------------------------------------------------------------------------------------------------------
static int check_array(float norm_spacing, int num_targets, int num_ant_ele) {
  if (num_ant_ele < 2 || num_ant_ele > 64) return fail(nullptr, DOA_CUDA_EINVAL, "num_ant_ele must be in [2, 64]");
  if (num_targets < 1 || num_targets >= num_ant_ele) return fail(nullptr, DOA_CUDA_EINVAL, "need 1 <= num_targets < num_ant_ele");
  if (!(norm_spacing > 0.0f) || norm_spacing > 0.5f) return fail(nullptr, DOA_CUDA_EINVAL, "need 0 < norm_spacing <= 0.5");
  return DOA_CUDA_OK;
}

int doa_cuda_music_create(doa_cuda_handle** out, float norm_spacing, int num_targets, int num_ant_ele, int pspectrum_len,
                          int device, int max_frames) {
  int rc = check_array(norm_spacing, num_targets, num_ant_ele);
  if (rc) return rc;
  if (pspectrum_len < 2) return fail(nullptr, DOA_CUDA_EINVAL, "pspectrum_len must be >= 2");
  doa_cuda_handle* h = nullptr;
  DevSave dev_save_;
  rc = begin_create(out, h, K_MUSIC, device, max_frames);
  if (rc) return rc;
  h->d = norm_spacing; h->T = num_targets; h->M = num_ant_ele; h->P = pspectrum_len;
  Lane& l = h->lane[0];
  const size_t mm = (size_t)h->M * h->M;
  bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess &&
            dalloc(&l.R, max_frames * mm) && dalloc(&l.u, (size_t)max_frames * h->M) &&
            dalloc(&l.spec, (size_t)max_frames * h->P) && upload_scan_tables(h, true, false);
  return finish_create(out, h, ok);
}

int doa_cuda_music_get_tables(const doa_cuda_handle* h, float* array_loc, float* theta_rad, float* steering) {
  if (!h || (h->kind != K_MUSIC && h->kind != K_CHAIN)) return DOA_CUDA_EINVAL;
  if (array_loc) memcpy(array_loc, h->h_loc.data(), sizeof(float) * h->M);
  if (theta_rad) memcpy(theta_rad, h->h_theta.data(), sizeof(float) * h->P);
  if (steering) memcpy(steering, h->h_V.data(), sizeof(float2) * (size_t)h->P * h->M);
  return DOA_CUDA_OK;
}

int doa_cuda_music_noise_subspace_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* G_dev, void* u_dev,
                                         void* w_dev, void* cuda_stream) {
  if (!h || (h->kind != K_MUSIC && h->kind != K_ROOTMUSIC && h->kind != K_CHAIN)) return DOA_CUDA_EINVAL;
  if (nframes <= 0) return nframes == 0 ? DOA_CUDA_OK : fail(h, DOA_CUDA_EINVAL, "nframes < 0");
  ENTER(h);
  int a = launch_noise_subspace((const float2*)in_dev, h->M, h->T, nframes, (float2*)G_dev, (float2*)u_dev, (float*)w_dev,
                                (cudaStream_t)cuda_stream);
  if (a < 0) return fail(h, a, "eigendecomposition launch rejected");
  h->launches = a;
  CK(h, cudaGetLastError());
  return DOA_CUDA_OK;
}

int doa_cuda_music_run_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* out_dev, void* cuda_stream) {
  if (!h || h->kind != K_MUSIC) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  Lane& l = h->lane[0];
  int a = launch_noise_subspace((const float2*)in_dev, h->M, h->T, nframes, nullptr, l.u, nullptr, st);
  if (a < 0) return fail(h, a, "eigendecomposition launch rejected");
  int b = launch_scan_spectrum(l.u, nullptr, tables_of(h), nframes, (float*)out_dev, st);
  if (b < 0) return fail(h, b, "spectrum launch rejected");
  h->launches = a + b;
  CK(h, cudaGetLastError());
  return DOA_CUDA_OK;
}

int doa_cuda_music_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_host) {
  if (!h || h->kind != K_MUSIC) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  Lane& l = h->lane[0];
  const size_t mm = (size_t)h->M * h->M;
  CK(h, cudaMemcpyAsync(l.R, in_host, sizeof(float2) * nframes * mm, cudaMemcpyHostToDevice, l.stream));
  int rc = doa_cuda_music_run_device(h, l.R, nframes, l.spec, l.stream);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(out_host, l.spec, sizeof(float) * (size_t)nframes * h->P, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaStreamSynchronize(l.stream));
  return DOA_CUDA_OK;
}

// ---- Root-MUSIC ------------------------------------------------------------------------------------------------------
int doa_cuda_rootmusic_create(doa_cuda_handle** out, float norm_spacing, int num_targets, int num_ant_ele, int device,
                              int max_frames) {
  int rc = check_array(norm_spacing, num_targets, num_ant_ele);
  if (rc) return rc;
  doa_cuda_handle* h = nullptr;
  DevSave dev_save_;
  rc = begin_create(out, h, K_ROOTMUSIC, device, max_frames);
  if (rc) return rc;
  h->d = norm_spacing; h->T = num_targets; h->M = num_ant_ele;
  Lane& l = h->lane[0];
  const size_t mm = (size_t)h->M * h->M, n = 2 * (size_t)h->M - 2;
  bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess &&
            dalloc(&l.R, max_frames * mm) && dalloc(&l.u, (size_t)max_frames * h->M) &&
            dalloc(&l.aoa, (size_t)max_frames * h->T) && dalloc(&l.scratch, n * n * (size_t)max_frames);
  return finish_create(out, h, ok);
}

int doa_cuda_rootmusic_run_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* out_dev, void* cuda_stream) {
  if (!h || h->kind != K_ROOTMUSIC) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  Lane& l = h->lane[0];
  int a = launch_noise_subspace((const float2*)in_dev, h->M, h->T, nframes, nullptr, l.u, nullptr, st);
  if (a < 0) return fail(h, a, "eigendecomposition launch rejected");
  int b = launch_rootmusic_scratch(l.u, h->M, h->T, h->d, nframes, l.scratch, h->max_frames, (float*)out_dev, st);
  if (b < 0) return fail(h, b, "root finder launch rejected");
  h->launches = a + b;
  CK(h, cudaGetLastError());
  return DOA_CUDA_OK;
}

int doa_cuda_rootmusic_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_host) {
  if (!h || h->kind != K_ROOTMUSIC) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  Lane& l = h->lane[0];
  const size_t mm = (size_t)h->M * h->M;
  CK(h, cudaMemcpyAsync(l.R, in_host, sizeof(float2) * nframes * mm, cudaMemcpyHostToDevice, l.stream));
  int rc = doa_cuda_rootmusic_run_device(h, l.R, nframes, l.aoa, l.stream);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(out_host, l.aoa, sizeof(float) * (size_t)nframes * h->T, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaStreamSynchronize(l.stream));
  return DOA_CUDA_OK;
}

// ---- calibrate_lin_array ------------------------------------------------------------------------------------------------
int doa_cuda_calibrate_create(doa_cuda_handle** out, float norm_spacing, int num_ant_ele, float pilot_angle, int device,
                              int max_frames) {
  int rc = check_array(norm_spacing, 1, num_ant_ele);
  if (rc) return rc;
  if (!std::isfinite(pilot_angle)) return fail(nullptr, DOA_CUDA_EINVAL, "pilot_angle must be finite");
  doa_cuda_handle* h = nullptr;
  DevSave dev_save_;
  rc = begin_create(out, h, K_CALIB, device, max_frames);
  if (rc) return rc;
  h->d = norm_spacing; h->T = 1; h->M = num_ant_ele;
  // pilot steering vector with the constructor's own arithmetic (lib/calibrate_lin_array_impl.cc:57-74, amv :84-91)
  const double kPi = 3.14159265358979323846;
  const float theta = (float)(kPi * pilot_angle / 180.0);
  const float s = (float)(-1.0 * 2 * kPi * std::cos((double)theta));
  h->h_V.resize(h->M);
  for (int nn = 0; nn < h->M; ++nn) {
    const float loc = (float)(norm_spacing * 0.5 * (h->M - 1 - 2 * nn));
    const float phi = s * loc;
    h->h_V[nn] = make_float2(cosf(phi), sinf(phi));
  }
  Lane& l = h->lane[0];
  const size_t mm = (size_t)h->M * h->M;
  bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess && dalloc(&l.R, max_frames * mm) &&
            dalloc(&l.G, max_frames * mm) && dalloc(&l.u, (size_t)max_frames * h->M) && dalloc(&h->d_V, (size_t)h->M) &&
            cudaMemcpy(h->d_V, h->h_V.data(), sizeof(float2) * h->M, cudaMemcpyHostToDevice) == cudaSuccess;
  return finish_create(out, h, ok);
}

int doa_cuda_calibrate_run_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* out_dev, void* cuda_stream) {
  if (!h || h->kind != K_CALIB) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  Lane& l = h->lane[0];
  int a = launch_noise_subspace((const float2*)in_dev, h->M, 1, nframes, l.G, nullptr, nullptr, st);
  if (a < 0) return fail(h, a, "eigendecomposition launch rejected");
  int b = launch_calibrate_emit((const float2*)in_dev, l.G, h->d_V, h->M, nframes, (float2*)out_dev, st);
  if (b < 0) return fail(h, b, "calibration emit launch rejected");
  h->launches = a + b;
  CK(h, cudaGetLastError());
  return DOA_CUDA_OK;
}

int doa_cuda_calibrate_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_host) {
  if (!h || h->kind != K_CALIB) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  Lane& l = h->lane[0];
  const size_t mm = (size_t)h->M * h->M;
  CK(h, cudaMemcpyAsync(l.R, in_host, sizeof(float2) * nframes * mm, cudaMemcpyHostToDevice, l.stream));
  int rc = doa_cuda_calibrate_run_device(h, l.R, nframes, l.u, l.stream);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(out_host, l.u, sizeof(float2) * (size_t)nframes * h->M, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaStreamSynchronize(l.stream));
  return DOA_CUDA_OK;
}

// ---- find_local_max ----------------------------------------------------------------------------------------------------
int doa_cuda_find_local_max_create(doa_cuda_handle** out, int num_max_vals, int vector_len, float x_min, float x_max,
                                   int device, int max_frames) {
  if (num_max_vals < 1 || num_max_vals > 16) return fail(nullptr, DOA_CUDA_EINVAL, "num_max_vals must be in [1, 16]");
  if (vector_len < 2) return fail(nullptr, DOA_CUDA_EINVAL, "vector_len must be >= 2");
  doa_cuda_handle* h = nullptr;
  DevSave dev_save_;
  int rc = begin_create(out, h, K_FLM, device, max_frames);
  if (rc) return rc;
  h->K = num_max_vals; h->P = vector_len; h->x_min = x_min; h->x_max = x_max;
  Lane& l = h->lane[0];
  bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess &&
            dalloc(&l.vecs, (size_t)max_frames * h->P) && dalloc(&l.val, (size_t)max_frames * h->K) &&
            dalloc(&l.loc, (size_t)max_frames * h->K) && dalloc(&l.bin, (size_t)max_frames * h->K) &&
            upload_scan_tables(h, false, true);
  return finish_create(out, h, ok);
}

int doa_cuda_find_local_max_run_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* out_val_dev,
                                       void* out_loc_dev, void* out_bin_dev, void* cuda_stream) {
  if (!h || h->kind != K_FLM) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0) return fail(h, DOA_CUDA_EINVAL, "nframes < 0");
  ENTER(h);
  int a = launch_find_local_max((const float*)in_dev, h->P, nframes, h->K, h->d_x, (float*)out_val_dev, (float*)out_loc_dev,
                                (int*)out_bin_dev, (cudaStream_t)cuda_stream);
  if (a < 0) return fail(h, a, "find_local_max launch rejected (vector_len too large for shared memory?)");
  h->launches = a;
  CK(h, cudaGetLastError());
  return DOA_CUDA_OK;
}

int doa_cuda_find_local_max_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_val_host,
                                void* out_loc_host, void* out_bin_host) {
  if (!h || h->kind != K_FLM) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  Lane& l = h->lane[0];
  CK(h, cudaMemcpyAsync(l.vecs, in_host, sizeof(float) * (size_t)nframes * h->P, cudaMemcpyHostToDevice, l.stream));
  int rc = doa_cuda_find_local_max_run_device(h, l.vecs, nframes, l.val, l.loc, l.bin, l.stream);
  if (rc) return rc;
  const size_t nk = (size_t)nframes * h->K;
  CK(h, cudaMemcpyAsync(out_val_host, l.val, sizeof(float) * nk, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaMemcpyAsync(out_loc_host, l.loc, sizeof(float) * nk, cudaMemcpyDeviceToHost, l.stream));
  if (out_bin_host) CK(h, cudaMemcpyAsync(out_bin_host, l.bin, sizeof(int) * nk, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaStreamSynchronize(l.stream));
  return DOA_CUDA_OK;
}

// ---- fused chain -----------------------------------------------------------------------------------------------------
static const int CHAIN_HOST_CHUNK_BYTES = 256 << 20;   // H2D granularity of the host path

int doa_cuda_chain_create(doa_cuda_handle** out, int inputs, int snapshot_size, int overlap_size, int avg_method,
                          float norm_spacing, int num_targets, int pspectrum_len, int num_max_vals, float x_min,
                          float x_max, int device, int max_frames) {
  if (inputs < 2 || inputs > 64) return fail(nullptr, DOA_CUDA_EINVAL, "inputs must be in [2, 64]");
  if (snapshot_size < 1) return fail(nullptr, DOA_CUDA_EINVAL, "snapshot_size must be > 0");
  if (overlap_size < 0 || overlap_size >= snapshot_size) return fail(nullptr, DOA_CUDA_EINVAL, "need 0 <= overlap_size < snapshot_size");
  if (avg_method != 0 && avg_method != 1) return fail(nullptr, DOA_CUDA_EINVAL, "avg_method must be 0 or 1");
  int rc = check_array(norm_spacing, num_targets, inputs);
  if (rc) return rc;
  if (pspectrum_len < 2) return fail(nullptr, DOA_CUDA_EINVAL, "pspectrum_len must be >= 2");
  if (num_max_vals < 1 || num_max_vals > 16) return fail(nullptr, DOA_CUDA_EINVAL, "num_max_vals must be in [1, 16]");
  doa_cuda_handle* h = nullptr;
  DevSave dev_save_;
  rc = begin_create(out, h, K_CHAIN, device, max_frames);
  if (rc) return rc;
  h->M = inputs; h->N = snapshot_size; h->overlap = overlap_size; h->hop = snapshot_size - overlap_size; h->avg = avg_method;
  h->d = norm_spacing; h->T = num_targets; h->P = pspectrum_len; h->K = num_max_vals; h->x_min = x_min; h->x_max = x_max;
  h->nlanes = 2;
  const size_t mm = (size_t)h->M * h->M;
  const size_t frame_bytes = sizeof(float2) * (size_t)h->M * h->N;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)max_frames, CHAIN_HOST_CHUNK_BYTES / frame_bytes));
  bool ok = upload_scan_tables(h, true, true);
  for (int i = 0; i < 2 && ok; ++i) {
    Lane& l = h->lane[i];
    // lane 0 carries the intermediates for a full device-resident batch; lane 1 only ever sees host chunks
    const size_t nf = (i == 0) ? (size_t)max_frames : (size_t)chunk;
    l.frames = (int)nf;
    // host staging: a chunk of independent frames, or -- lane 0, doa_cuda_chain_run_streams -- the [M][Lpad] channel streams of
    // up to max_frames frames (hop <= snapshot_size, so that never exceeds max_frames * M * N samples)
    const size_t Lpad_max = (((size_t)(max_frames - 1) * h->hop + h->N) + 1) & ~(size_t)1;
    l.in_elems = std::max((size_t)chunk * h->M * h->N + 2 * (size_t)h->M, i == 0 ? Lpad_max * h->M : (size_t)0);
    ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess && dalloc(&l.in, l.in_elems) &&
         dalloc(&l.R, nf * mm) && dalloc(&l.G, nf * mm) && dalloc(&l.u, nf * h->M) && dalloc(&l.val, nf * h->K) &&
         dalloc(&l.loc, nf * h->K) && dalloc(&l.bin, nf * h->K) && alloc_tc_ws(h, l);
  }
  return finish_create(out, h, ok);
}

static int chain_on_lane(doa_cuda_handle* h, Lane& l, const void* in_dev, long long frame_stride, long long chan_stride,
                         int nframes, float* val, float* loc, int* bin, cudaStream_t st, bool prof) {
  cudaEvent_t* ev = nullptr;
  if (prof) { ev = &h->ev[(size_t)(h->prof_calls % PROF_SETS) * 4]; ++h->prof_calls; }
  if (prof) CK(h, cudaEventRecord(ev[0], st));
  if (dev_option(OPT_FUSED, 1)) {
    // one persistent kernel for the whole chain when the shape allows it; stage events collapse to (0, 0, total)
    if (prof) { CK(h, cudaEventRecord(ev[1], st)); CK(h, cudaEventRecord(ev[2], st)); }
#ifdef DOA_DEV_KNOBS
    if (dev_option(OPT_FUSED, 1) == 2) {
      // experiment: covariance + Jacobi in the persistent kernel (G, u to global memory), the scan on the tensor cores afterwards
      int f2 = launch_chain_fused(in_dev, frame_stride, chan_stride, h->M, h->N, nframes, h->avg, h->T, tables_of(h), h->K, val, loc, bin, st,
                                  h->d_gains, h->fmt, l.G, l.u);
      if (f2 > 0) {
        if (prof) { CK(h, cudaEventRecord(ev[1], st)); CK(h, cudaEventRecord(ev[2], st)); }
        int c2 = launch_scan_peaks_tc(l.u, l.G, tables_of(h), nframes, h->K, val, loc, bin, st);
        if (c2 <= 0) return fail(h, DOA_CUDA_EINVAL, "split form needs the tensor-core scan");
        if (prof) CK(h, cudaEventRecord(ev[3], st));
        h->launches += f2 + c2;
        CK(h, cudaGetLastError());
        return DOA_CUDA_OK;
      }
    }
#endif
    int f = launch_chain_fused(in_dev, frame_stride, chan_stride, h->M, h->N, nframes, h->avg, h->T, tables_of(h), h->K, val,
                               loc, bin, st, h->d_gains, h->fmt);
    if (f < 0) return fail(h, f, "fused chain launch rejected");
    if (f > 0) {
      if (prof) CK(h, cudaEventRecord(ev[3], st));
      h->launches += f;
      CK(h, cudaGetLastError());
      return DOA_CUDA_OK;
    }
  }
  // 16 elements: covariance and eigendecomposition are both bound by the FP32 pipe and by registers (255 per covariance thread,
  // 125 per Jacobi thread): sharing the SMs in one persistent kernel (fused16.cu) leaves each side too few warps and measured
  // 4.9 ms against 1.7 + 2.3 ms for the two stage kernels (65,536 frames).  The experiment stays in the -DDOA_DEV_KNOBS build.
  int a = 0;
#ifdef DOA_DEV_KNOBS
  if (dev_option(OPT_FUSED16, 0))
    a = launch_cov_eig_fused16(in_dev, frame_stride, chan_stride, h->M, h->N, nframes, h->avg, h->T, l.G, l.u, st, h->d_gains, h->fmt);
#endif
  int b = 0;
  if (a > 0) {
    if (prof) { CK(h, cudaEventRecord(ev[1], st)); CK(h, cudaEventRecord(ev[2], st)); }
  } else {
    a = launch_covariance(in_dev, frame_stride, chan_stride, h->M, h->N, nframes, h->avg, l.R, st, h->d_gains, h->fmt, l.tc_ws);
    if (a < 0) return fail(h, a, "covariance launch rejected");
    if (prof) CK(h, cudaEventRecord(ev[1], st));
    b = launch_noise_subspace(l.R, h->M, h->T, nframes, l.G, l.u, nullptr, st);
    if (b < 0) return fail(h, b, "eigendecomposition launch rejected");
    if (prof) CK(h, cudaEventRecord(ev[2], st));
  }
  int c = dev_option(OPT_SCAN_TC, 1) ? launch_scan_peaks_tc(l.u, l.G, tables_of(h), nframes, h->K, val, loc, bin, st) : 0;
  if (c == 0) c = launch_scan_peaks(l.u, l.G, tables_of(h), nframes, h->K, val, loc, bin, st);
  if (c < 0) return fail(h, c, "scan launch rejected (pspectrum_len too large for shared memory?)");
  if (prof) CK(h, cudaEventRecord(ev[3], st));
  h->launches += a + b + c;
  CK(h, cudaGetLastError());
  return DOA_CUDA_OK;
}

int doa_cuda_chain_run_device(doa_cuda_handle* h, const void* in_dev, long long frame_stride, long long chan_stride,
                              int nframes, void* out_val_dev, void* out_loc_dev, void* out_bin_dev, void* cuda_stream) {
  if (!h || h->kind != K_CHAIN) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  h->launches = 0;
  return chain_on_lane(h, h->lane[0], in_dev, frame_stride, chan_stride, nframes, (float*)out_val_dev,
                       (float*)out_loc_dev, (int*)out_bin_dev, (cudaStream_t)cuda_stream, h->profiling);
}

int doa_cuda_chain_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_val_host, void* out_loc_host,
                       void* out_bin_host) {
  if (!h || h->kind != K_CHAIN) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  h->launches = 0;
  const size_t fe = (size_t)h->M * h->N;
  const int chunk = h->lane[1].frames;
  const size_t sb = h->sample_bytes();
  const char* src = (const char*)in_host;
  if (nframes <= chunk && is_pageable(in_host)) {   // a small call out of pageable memory: one staged copy in, one synchronisation out
    Lane& l = h->lane[0];
    char* pin = pin_staging(l, sb * nframes * fe);
    if (pin != nullptr) {
      CK(h, cudaStreamSynchronize(l.stream));
      std::memcpy(pin, src, sb * nframes * fe);
      CK(h, cudaMemcpyAsync(l.in, pin, sb * nframes * fe, cudaMemcpyHostToDevice, l.stream));
      int rc = chain_on_lane(h, l, l.in, (long long)fe, h->N, nframes, l.val, l.loc, l.bin, l.stream, false);
      if (rc) return rc;
      return peaks_to_host(h, l, nframes, out_val_host, out_loc_host, out_bin_host);
    }
  }
  int c = 0;
  for (int f0 = 0; f0 < nframes; f0 += chunk, ++c) {
    Lane& l = h->lane[c & 1];
    const int nf = std::min(chunk, nframes - f0);
    CK(h, cudaMemcpyAsync(l.in, src + (size_t)f0 * fe * sb, sb * nf * fe, cudaMemcpyHostToDevice, l.stream));
    int rc = chain_on_lane(h, l, l.in, (long long)fe, h->N, nf, l.val, l.loc, l.bin, l.stream, false);
    if (rc) return rc;
    const size_t nk = (size_t)nf * h->K, off = (size_t)f0 * h->K;
    CK(h, cudaMemcpyAsync((float*)out_val_host + off, l.val, sizeof(float) * nk, cudaMemcpyDeviceToHost, l.stream));
    CK(h, cudaMemcpyAsync((float*)out_loc_host + off, l.loc, sizeof(float) * nk, cudaMemcpyDeviceToHost, l.stream));
    if (out_bin_host) CK(h, cudaMemcpyAsync((int*)out_bin_host + off, l.bin, sizeof(int) * nk, cudaMemcpyDeviceToHost, l.stream));
  }
  CK(h, cudaStreamSynchronize(h->lane[0].stream));
  CK(h, cudaStreamSynchronize(h->lane[1].stream));
  return DOA_CUDA_OK;
}

int doa_cuda_chain_run_streams(doa_cuda_handle* h, const void* const* in_host, int nframes, void* out_val_host,
                               void* out_loc_host, void* out_bin_host) {
  if (!h || h->kind != K_CHAIN) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  h->launches = 0;
  Lane& l = h->lane[0];
  size_t Lpad = 0;
  int rc = stage_streams(h, l, in_host, nframes, &Lpad);
  if (rc) return rc;
  rc = chain_on_lane(h, l, l.in, h->hop, (long long)Lpad, nframes, l.val, l.loc, l.bin, l.stream, false);
  if (rc) return rc;
  return peaks_to_host(h, l, nframes, out_val_host, out_loc_host, out_bin_host);
}

// ---- autocorrelate -> rootMUSIC_linear_array without the covariance leaving the device (BASELINE configs[1]) ------------
// The reference flowgraph's two blocks (lib/autocorrelate_impl.cc:82-118 -> lib/rootMUSIC_linear_array_impl.cc:90-152) as one
// call: covariance, batched Jacobi + diagonal sums of the noise projector, polynomial roots, root selection; T angles per
// frame come back.  Same kernels as the separate stages, hence the same bits.
int doa_cuda_rootchain_create(doa_cuda_handle** out, int inputs, int snapshot_size, int overlap_size, int avg_method,
                              float norm_spacing, int num_targets, int device, int max_frames) {
  if (inputs < 2 || inputs > 64) return fail(nullptr, DOA_CUDA_EINVAL, "inputs must be in [2, 64]");
  if (snapshot_size < 1) return fail(nullptr, DOA_CUDA_EINVAL, "snapshot_size must be > 0");
  if (overlap_size < 0 || overlap_size >= snapshot_size) return fail(nullptr, DOA_CUDA_EINVAL, "need 0 <= overlap_size < snapshot_size");
  if (avg_method != 0 && avg_method != 1) return fail(nullptr, DOA_CUDA_EINVAL, "avg_method must be 0 or 1");
  int rc = check_array(norm_spacing, num_targets, inputs);
  if (rc) return rc;
  doa_cuda_handle* h = nullptr;
  DevSave dev_save_;
  rc = begin_create(out, h, K_ROOTCHAIN, device, max_frames);
  if (rc) return rc;
  h->M = inputs; h->N = snapshot_size; h->overlap = overlap_size; h->hop = snapshot_size - overlap_size; h->avg = avg_method;
  h->d = norm_spacing; h->T = num_targets;
  Lane& l = h->lane[0];
  const size_t mm = (size_t)h->M * h->M, n = 2 * (size_t)h->M - 2;
  const size_t Lpad = (((size_t)(max_frames - 1) * h->hop + h->N) + 1) & ~(size_t)1;
  l.in_elems = std::max(Lpad * h->M, (size_t)max_frames * h->M * h->N);     // streams or independent frames
  bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess && dalloc(&l.in, l.in_elems) &&
            dalloc(&l.R, max_frames * mm) && dalloc(&l.u, (size_t)max_frames * h->M) &&
            dalloc(&l.aoa, (size_t)max_frames * h->T) && dalloc(&l.scratch, n * n * (size_t)max_frames) && alloc_tc_ws(h, l);
  return finish_create(out, h, ok);
}

int doa_cuda_rootchain_run_device(doa_cuda_handle* h, const void* in_dev, long long frame_stride, long long chan_stride,
                                  int nframes, void* out_aoa_dev, void* cuda_stream) {
  if (!h || h->kind != K_ROOTCHAIN) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  Lane& l = h->lane[0];
  // 4 and 8 elements: covariance + eigensolver in the persistent warp-specialised kernel of the MUSIC chain (fused.cu, split
  // form: the diagonal sums go to global memory, no scan) -- the covariance never leaves the SM and the streaming runs at the
  // fused kernel's rate (cfg2: 2.47 + 0.09 ms as two kernels -> 2.1 ms); same device code, same bits.
  int a = 0, b = 0;
  if (dev_option(OPT_FUSED, 1)) {
    a = launch_chain_fused(in_dev, frame_stride, chan_stride, h->M, h->N, nframes, h->avg, h->T, ScanTables(), 1, nullptr, nullptr, nullptr,
                           st, h->d_gains, h->fmt, nullptr, l.u);
    if (a < 0) return fail(h, a, "fused covariance + eigendecomposition launch rejected");
  }
  if (a == 0) {
    a = launch_covariance(in_dev, frame_stride, chan_stride, h->M, h->N, nframes, h->avg, l.R, st, h->d_gains, h->fmt, l.tc_ws);
    if (a < 0) return fail(h, a, "covariance launch rejected");
    b = launch_noise_subspace(l.R, h->M, h->T, nframes, nullptr, l.u, nullptr, st);
    if (b < 0) return fail(h, b, "eigendecomposition launch rejected");
  }
  int c = launch_rootmusic_scratch(l.u, h->M, h->T, h->d, nframes, l.scratch, h->max_frames, (float*)out_aoa_dev, st);
  if (c < 0) return fail(h, c, "root finder launch rejected");
  h->launches = a + b + c;
  CK(h, cudaGetLastError());
  return DOA_CUDA_OK;
}

int doa_cuda_rootchain_run_streams(doa_cuda_handle* h, const void* const* in_host, int nframes, void* out_aoa_host) {
  if (!h || h->kind != K_ROOTCHAIN) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  Lane& l = h->lane[0];
  size_t Lpad = 0;
  int rc = stage_streams(h, l, in_host, nframes, &Lpad);
  if (rc) return rc;
  rc = doa_cuda_rootchain_run_device(h, l.in, h->hop, (long long)Lpad, nframes, l.aoa, l.stream);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(out_aoa_host, l.aoa, sizeof(float) * (size_t)nframes * h->T, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaStreamSynchronize(l.stream));
  return DOA_CUDA_OK;
}

int doa_cuda_rootchain_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_aoa_host) {
  if (!h || h->kind != K_ROOTCHAIN) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames");
  ENTER(h);
  Lane& l = h->lane[0];
  const size_t fe = (size_t)h->M * h->N;
  CK(h, cudaMemcpyAsync(l.in, in_host, h->sample_bytes() * fe * nframes, cudaMemcpyHostToDevice, l.stream));
  int rc = doa_cuda_rootchain_run_device(h, l.in, (long long)fe, h->N, nframes, l.aoa, l.stream);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(out_aoa_host, l.aoa, sizeof(float) * (size_t)nframes * h->T, cudaMemcpyDeviceToHost, l.stream));
  CK(h, cudaStreamSynchronize(l.stream));
  return DOA_CUDA_OK;
}

// ---- every GPU of the box from ONE process (a GNU Radio flowgraph is one process) --------------------------------------
// Frames are independent (lib/autocorrelate_impl.cc:92, lib/MUSIC_lin_array_impl.cc:121, lib/find_local_max_impl.cc:179), so a
// batch is cut into contiguous blocks of frames, one per listed device; every device runs the chain on its block from the
// caller's host buffer (its own PCIe link, its own streams) and writes its K peaks per frame straight into the caller's
// output arrays at the block's offset.  That write IS the gather: nothing else crosses between devices, so no collective
// is needed in this single-process form (the one-process-per-GPU form gathers with NCCL, gr_doa_b200/sharding.py).
int doa_cuda_multi_create(doa_cuda_handle** out, int inputs, int snapshot_size, int overlap_size, int avg_method,
                          float norm_spacing, int num_targets, int pspectrum_len, int num_max_vals, float x_min, float x_max,
                          const int* devices, int ndevices, int max_frames_per_device) {
  if (!out) return fail(nullptr, DOA_CUDA_EINVAL, "null handle pointer");
  *out = nullptr;
  if (!devices || ndevices < 1 || ndevices > 64) return fail(nullptr, DOA_CUDA_EINVAL, "need 1 <= ndevices <= 64 and a device list");
  doa_cuda_handle* h = nullptr;
  DevSave dev_save_;
  int rc = begin_create(out, h, K_MULTI, devices[0], max_frames_per_device);
  if (rc) return rc;
  h->M = inputs; h->N = snapshot_size; h->K = num_max_vals; h->overlap = overlap_size; h->hop = snapshot_size - overlap_size;
  for (int g = 0; g < ndevices; ++g) {
    doa_cuda_handle* c = nullptr;
    rc = doa_cuda_chain_create(&c, inputs, snapshot_size, overlap_size, avg_method, norm_spacing, num_targets, pspectrum_len,
                               num_max_vals, x_min, x_max, devices[g], max_frames_per_device);
    if (rc) { doa_cuda_destroy(h); return rc; }   // g_create_err holds the child's text
    h->children.push_back(c);
  }
  h->max_frames = max_frames_per_device * ndevices;
  h->pool.reset(new MultiPool(ndevices));
  *out = h;
  return DOA_CUDA_OK;
}

int doa_cuda_multi_device_count(const doa_cuda_handle* h) { return (h && h->kind == K_MULTI) ? (int)h->children.size() : DOA_CUDA_EINVAL; }

// Frames [first, first + count) of device g out of nframes over G devices: contiguous, sizes differ by at most one.
static void multi_block(int nframes, int G, int g, int* first, int* count) {
  const int per = nframes / G, rem = nframes % G;
  *first = g * per + std::min(g, rem);
  *count = per + (g < rem ? 1 : 0);
}

int doa_cuda_multi_block(const doa_cuda_handle* h, int nframes, int index, int* first, int* count) {
  if (!h || h->kind != K_MULTI || nframes < 0 || index < 0 || index >= (int)h->children.size() || !first || !count) return DOA_CUDA_EINVAL;
  multi_block(nframes, (int)h->children.size(), index, first, count);
  return DOA_CUDA_OK;
}

// Run job(g) for every device of the handle (device 0 on the calling thread, the others on the handle's workers) and fold
// the return codes and launch counts.
static int multi_finish(doa_cuda_handle* h, const std::function<int(int)>& job) {
  h->pool->run(job);
  h->launches = 0;
  for (size_t g = 0; g < h->children.size(); ++g) {
    if (h->pool->rcs[g]) return fail(h, h->pool->rcs[g], "device " + std::to_string(h->children[g]->device) + ": " + h->children[g]->err);
    h->launches += h->children[g]->launches;
  }
  return DOA_CUDA_OK;
}

int doa_cuda_multi_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_val_host, void* out_loc_host,
                       void* out_bin_host) {
  if (!h || h->kind != K_MULTI) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames_per_device * ndevices");
  const int G = (int)h->children.size();
  const size_t frame_bytes = (size_t)h->M * h->N * h->children[0]->sample_bytes();
  const std::function<int(int)> work = [&](int g) -> int {
    int first = 0, count = 0;
    multi_block(nframes, G, g, &first, &count);
    if (count == 0) { h->children[g]->launches = 0; return DOA_CUDA_OK; }
    const size_t off = (size_t)first * h->K;
    return doa_cuda_chain_run(h->children[g], (const char*)in_host + (size_t)first * frame_bytes, count,
                              (float*)out_val_host + off, (float*)out_loc_host + off,
                              out_bin_host ? (void*)((int*)out_bin_host + off) : nullptr);
  };
  return multi_finish(h, work);
}

// Streaming form (what a GNU Radio general_work() holds: `inputs` channel pointers, frame i at hop * i): device g gets the
// channel pointers advanced to its first frame; the `overlap` samples its last frame shares with the next device's first are
// simply read by both (the halo of SURVEY section 8(e), free here because every device reads the same host buffer).
int doa_cuda_multi_run_streams(doa_cuda_handle* h, const void* const* in_host, int nframes, void* out_val_host,
                               void* out_loc_host, void* out_bin_host) {
  if (!h || h->kind != K_MULTI) return DOA_CUDA_EINVAL;
  if (nframes == 0) return DOA_CUDA_OK;
  if (nframes < 0 || nframes > h->max_frames) return fail(h, DOA_CUDA_ECAPACITY, "nframes exceeds max_frames_per_device * ndevices");
  const int G = (int)h->children.size();
  const size_t sb = h->children[0]->sample_bytes();
  const std::function<int(int)> work = [&](int g) -> int {
    int first = 0, count = 0;
    multi_block(nframes, G, g, &first, &count);
    if (count == 0) { h->children[g]->launches = 0; return DOA_CUDA_OK; }
    std::vector<const void*> ptrs(h->M);
    for (int k = 0; k < h->M; ++k) ptrs[k] = (const char*)in_host[k] + (size_t)first * h->hop * sb;
    const size_t off = (size_t)first * h->K;
    return doa_cuda_chain_run_streams(h->children[g], ptrs.data(), count, (float*)out_val_host + off, (float*)out_loc_host + off,
                                      out_bin_host ? (void*)((int*)out_bin_host + off) : nullptr);
  };
  return multi_finish(h, work);
}

int doa_cuda_set_profiling(doa_cuda_handle* h, int on) {
  if (!h || h->kind != K_CHAIN) return DOA_CUDA_EINVAL;
  ENTER(h);
  if (on && h->ev.empty()) {
    h->ev.assign((size_t)PROF_SETS * 4, nullptr);
    for (auto& e : h->ev) CK(h, cudaEventCreate(&e));
  }
  h->profiling = on != 0;
  h->prof_calls = 0;
  return DOA_CUDA_OK;
}

// Mean CUDA-event time of each stage over the chain run_device calls recorded since profiling was switched on
// (the last PROF_SETS of them); returns the number of calls averaged, or a negative error.
int doa_cuda_chain_stage_ms(doa_cuda_handle* h, float* cov_ms, float* eig_ms, float* scan_ms) {
  if (!h || h->kind != K_CHAIN || h->ev.empty()) return DOA_CUDA_EINVAL;
  ENTER(h);
  const int n = std::min(h->prof_calls, PROF_SETS);
  double a = 0, b = 0, c = 0;
  for (int i = 0; i < n; ++i) {
    cudaEvent_t* ev = &h->ev[(size_t)i * 4];
    CK(h, cudaEventSynchronize(ev[3]));
    float x = 0, y = 0, z = 0;
    CK(h, cudaEventElapsedTime(&x, ev[0], ev[1]));
    CK(h, cudaEventElapsedTime(&y, ev[1], ev[2]));
    CK(h, cudaEventElapsedTime(&z, ev[2], ev[3]));
    a += x; b += y; c += z;
  }
  if (n > 0) { a /= n; b /= n; c /= n; }
  if (cov_ms) *cov_ms = (float)a;
  if (eig_ms) *eig_ms = (float)b;
  if (scan_ms) *scan_ms = (float)c;
  return n;
}

}  // extern "C"
