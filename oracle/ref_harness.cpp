// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/arma_shim/armadillo).
//
// C entry points around gr-doa's own, unmodified block classes.  oracle/build_ref.py compiles this file together with
//     /root/reference/lib/{autocorrelate,MUSIC_lin_array,rootMUSIC_linear_array,find_local_max,calibrate_lin_array}_impl.cc
// (read where they lie; no reference source is copied into this repository) against the reference's public headers
// (/root/reference/include), the Armadillo stand-in (oracle/arma_shim) and the compile-only GNU Radio stand-in
// (gr_doa_b200/gnuradio/shim), into oracle/_ref/libdoa_ref.so.  Every function below only builds a block through its
// factory -- doa.X::make(...), what SWIG exposes (swig/doa_swig.i:22-36) -- and calls its work() / general_work() the way
// the GNU Radio scheduler would, so the arithmetic that runs is the reference's own statements.
//
// Used by tests/golden/make_ref_golden.py (fixtures for the GPU box, which has no /root/reference), by tests/test_reference_build.py
// (the port oracle/doa_oracle.cpp against this build) and, where present, by bench.py's CPU arm (kind "reference").
#include <armadillo>
#include <dlfcn.h>

#include <cstdio>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <doa/MUSIC_lin_array.h>
#include <doa/autocorrelate.h>
#include <doa/calibrate_lin_array.h>
#include <doa/find_local_max.h>
#include <doa/rootMUSIC_linear_array.h>

namespace arma { namespace shim {
static Lapack g_lapack;
Lapack& lapack() { return g_lapack; }
} }

namespace {
void (*g_set_threads)(int) = nullptr;

// the items a block reads/writes per call, as the scheduler hands them over
template <class B>
int call_work(B& blk, int n, std::vector<const void*> in, std::vector<void*> out) {
  gr_vector_int ninput(in.size(), 0);
  gr_vector_const_void_star iv(in.begin(), in.end());
  gr_vector_void_star ov(out.begin(), out.end());
  return blk->general_work(n, ninput, iv, ov);
}
}  // namespace

extern "C" {

int ref_init(const char* lapack_path, const char* prefix, int use_herk) {
  void* h = dlopen(lapack_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "ref_init: dlopen(%s): %s\n", lapack_path, dlerror()); return -1; }
  const std::string p(prefix ? prefix : "");
  arma::shim::Lapack& L = arma::shim::lapack();
  *(void**)&L.cgemm = dlsym(h, (p + "cgemm_").c_str());
  *(void**)&L.cgemv = dlsym(h, (p + "cgemv_").c_str());
  *(void**)&L.cherk = dlsym(h, (p + "cherk_").c_str());
  *(void**)&L.sgemm = dlsym(h, (p + "sgemm_").c_str());
  *(void**)&L.cheevd = dlsym(h, (p + "cheevd_").c_str());
  *(void**)&L.cgeev = dlsym(h, (p + "cgeev_").c_str());
  *(void**)&g_set_threads = dlsym(h, (p + "openblas_set_num_threads").c_str());
  L.use_herk = use_herk;
  if (!L.cgemm || !L.cgemv || !L.cherk || !L.sgemm || !L.cheevd || !L.cgeev) return -2;
  if (g_set_threads) g_set_threads(1);   // frames are spread over threads by the callers below; BLAS itself stays serial
  return 0;
}
void ref_set_herk(int use_herk) { arma::shim::lapack().use_herk = use_herk; }
int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// autocorrelate: `streams` = M pointers to the channel streams (history included: the first sample of frame 0 first),
// exactly what general_work() receives (lib/autocorrelate_impl.cc:82-118).  Returns the items produced; *consumed = consume_each().
int ref_autocorrelate(int inputs, int snapshot_size, int overlap_size, int avg_method, const void* const* streams, int noutput,
                      void* out, int* forecast_for_noutput, int* history, int* consumed) {
  try {
    gr::doa::autocorrelate::sptr blk = gr::doa::autocorrelate::make(inputs, snapshot_size, overlap_size, avg_method);
    gr_vector_int req(inputs, 0);
    blk->forecast(noutput, req);
    if (forecast_for_noutput) *forecast_for_noutput = req[0];
    if (history) *history = (int)blk->history();
    std::vector<const void*> in(streams, streams + inputs);
    const int r = call_work(blk, noutput, in, {out});
    if (consumed) *consumed = blk->last_consumed();
    return r;
  } catch (const std::exception& e) { fprintf(stderr, "ref_autocorrelate: %s\n", e.what()); return -100; }
}

// MUSIC_lin_array::work over `n` covariance items (lib/MUSIC_lin_array_impl.cc:108-150).
int ref_music(float norm_spacing, int num_targets, int num_ant_ele, int pspectrum_len, const void* R, int n, void* out, int nthreads) {
  int rc = 0;
  const size_t mm = (size_t)num_ant_ele * num_ant_ele;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(min : rc)
  {
    try {
#ifdef _OPENMP
      const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
      const int t = 0, nt = 1;
#endif
      const int lo = (int)((long long)n * t / nt), hi = (int)((long long)n * (t + 1) / nt);
      if (hi > lo) {
        FILE* keep = stdout; (void)keep;
        gr::doa::MUSIC_lin_array::sptr blk = gr::doa::MUSIC_lin_array::make(norm_spacing, num_targets, num_ant_ele, pspectrum_len);
        call_work(blk, hi - lo, {(const char*)R + sizeof(gr_complex) * mm * lo}, {(char*)out + sizeof(float) * (size_t)pspectrum_len * lo});
      }
    } catch (const std::exception& e) { fprintf(stderr, "ref_music: %s\n", e.what()); rc = -100; }
  }
  return rc;
}

// The constructor tables are private members; the steering matrix shows through a one-frame run on R = I with T = M - 1?  No:
// tests compare spectra, and the theta / x-axis grids through find_local_max's second port.  Nothing to export here.

// rootMUSIC_linear_array::work (lib/rootMUSIC_linear_array_impl.cc:90-152).  A frame whose selection throws inside Armadillo
// (no root strictly inside the unit circle: index_min of an empty vector) is reported as NaN angles, frame by frame.
int ref_rootmusic(float norm_spacing, int num_targets, int num_ant_ele, const void* R, int n, void* out, int nthreads, int* max_streams) {
  int rc = 0;
  const size_t mm = (size_t)num_ant_ele * num_ant_ele;
  if (max_streams) {
    gr::doa::rootMUSIC_linear_array::sptr b0 = gr::doa::rootMUSIC_linear_array::make(norm_spacing, num_targets, num_ant_ele);
    *max_streams = b0->output_signature()->max_streams();
  }
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(min : rc)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    const int t = 0, nt = 1;
#endif
    const int lo = (int)((long long)n * t / nt), hi = (int)((long long)n * (t + 1) / nt);
    if (hi > lo) {
      gr::doa::rootMUSIC_linear_array::sptr blk = gr::doa::rootMUSIC_linear_array::make(norm_spacing, num_targets, num_ant_ele);
      for (int i = lo; i < hi; ++i) {
        float* o = (float*)out + (size_t)num_targets * i;
        try {
          call_work(blk, 1, {(const char*)R + sizeof(gr_complex) * mm * i}, {o});
        } catch (const std::exception&) {
          for (int k = 0; k < num_targets; ++k) o[k] = arma::fdatum::nan;
          rc = rc < 1 ? rc : rc;
        }
      }
    }
  }
  return rc;
}

// find_local_max::work (lib/find_local_max_impl.cc:167-194): port 0 = peak heights, port 1 = locations sorted descending.
int ref_find_local_max(int num_max_vals, int vector_len, float x_min, float x_max, const void* in, int n, void* out_val, void* out_loc,
                       int nthreads) {
  int rc = 0;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(min : rc)
  {
    try {
#ifdef _OPENMP
      const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
      const int t = 0, nt = 1;
#endif
      const int lo = (int)((long long)n * t / nt), hi = (int)((long long)n * (t + 1) / nt);
      if (hi > lo) {
        gr::doa::find_local_max::sptr blk = gr::doa::find_local_max::make(num_max_vals, vector_len, x_min, x_max);
        call_work(blk, hi - lo, {(const char*)in + sizeof(float) * (size_t)vector_len * lo},
                  {(char*)out_val + sizeof(float) * (size_t)num_max_vals * lo, (char*)out_loc + sizeof(float) * (size_t)num_max_vals * lo});
      }
    } catch (const std::exception& e) { fprintf(stderr, "ref_find_local_max: %s\n", e.what()); rc = -100; }
  }
  return rc;
}

// calibrate_lin_array::work (lib/calibrate_lin_array_impl.cc:100-134).
int ref_calibrate(float norm_spacing, int num_ant_ele, float pilot_angle, const void* R, int n, void* out) {
  try {
    gr::doa::calibrate_lin_array::sptr blk = gr::doa::calibrate_lin_array::make(norm_spacing, num_ant_ele, pilot_angle);
    return call_work(blk, n, {R}, {out});
  } catch (const std::exception& e) { fprintf(stderr, "ref_calibrate: %s\n", e.what()); return -100; }
}

// The whole flowgraph autocorrelate -> MUSIC_lin_array -> find_local_max on independent frames [B][M][N] (each frame is one
// general_work() call with noutput_items = 1 and no history), one set of block instances per thread: the CPU arm of bench.py.
// spectra (may be null): [B][P] floats.
int ref_chain_frames(int inputs, int snapshot_size, int avg_method, float norm_spacing, int num_targets, int pspectrum_len,
                     int num_max_vals, float x_min, float x_max, const void* frames, int B, void* out_val, void* out_loc, void* spectra,
                     int nthreads) {
  int rc = 0;
  const size_t fe = (size_t)inputs * snapshot_size, mm = (size_t)inputs * inputs;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(min : rc)
  {
    try {
      gr::doa::autocorrelate::sptr ac = gr::doa::autocorrelate::make(inputs, snapshot_size, 0, avg_method);
      gr::doa::MUSIC_lin_array::sptr mu = gr::doa::MUSIC_lin_array::make(norm_spacing, num_targets, inputs, pspectrum_len);
      gr::doa::find_local_max::sptr fl = gr::doa::find_local_max::make(num_max_vals, pspectrum_len, x_min, x_max);
      std::vector<gr_complex> R(mm);
      std::vector<float> spec((size_t)pspectrum_len);
#pragma omp for schedule(static)
      for (int f = 0; f < B; ++f) {
        std::vector<const void*> in(inputs);
        for (int k = 0; k < inputs; ++k) in[k] = (const gr_complex*)frames + fe * f + (size_t)snapshot_size * k;
        call_work(ac, 1, in, {R.data()});
        float* sp = spectra ? (float*)spectra + (size_t)pspectrum_len * f : spec.data();
        call_work(mu, 1, {R.data()}, {sp});
        call_work(fl, 1, {sp}, {(float*)out_val + (size_t)num_max_vals * f, (float*)out_loc + (size_t)num_max_vals * f});
      }
    } catch (const std::exception& e) { fprintf(stderr, "ref_chain_frames: %s\n", e.what()); rc = -100; }
  }
  return rc;
}

// autocorrelate -> rootMUSIC_linear_array on independent frames (BASELINE configs[1]).
int ref_rootchain_frames(int inputs, int snapshot_size, int avg_method, float norm_spacing, int num_targets, const void* frames, int B,
                         void* out_aoa, int nthreads) {
  int rc = 0;
  const size_t fe = (size_t)inputs * snapshot_size, mm = (size_t)inputs * inputs;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(min : rc)
  {
    gr::doa::autocorrelate::sptr ac = gr::doa::autocorrelate::make(inputs, snapshot_size, 0, avg_method);
    gr::doa::rootMUSIC_linear_array::sptr rm = gr::doa::rootMUSIC_linear_array::make(norm_spacing, num_targets, inputs);
    std::vector<gr_complex> R(mm);
#pragma omp for schedule(static)
    for (int f = 0; f < B; ++f) {
      float* o = (float*)out_aoa + (size_t)num_targets * f;
      try {
        std::vector<const void*> in(inputs);
        for (int k = 0; k < inputs; ++k) in[k] = (const gr_complex*)frames + fe * f + (size_t)snapshot_size * k;
        call_work(ac, 1, in, {R.data()});
        call_work(rm, 1, {R.data()}, {o});
      } catch (const std::exception&) {
        for (int k = 0; k < num_targets; ++k) o[k] = arma::fdatum::nan;
      }
    }
  }
  return rc;
}

}  // extern "C"
