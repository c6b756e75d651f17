// doa_internal.h -- shared declarations of libdoa_cuda's translation units (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/doa_cuda.h"

namespace doa {

// ---- launch interfaces (one per kernel family; each returns the number of kernel launches it issued or <0) ----

// Sample format of the covariance input: fc32 (interleaved float re, im; the reference's gr_complex) or sc16 (UHD's
// interleaved int16 I, Q; the value of a sample is int16 * scale).  Strides are in complex samples either way.
struct InputFormat {
  int sc16 = 0;
  float scale = 1.0f;
};

// Stage 1.  Sample (f,k,t) at in[f*frame_stride + k*chan_stride + t]; out[f][r + c*M] column-major.
// gains (may be null): M per-channel complex gains applied in front of the covariance, R' = D R D^H.
int launch_covariance(const void* in, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                      int avg_method, float2* out, cudaStream_t st, const float2* gains = nullptr,
                      InputFormat fmt = InputFormat(), void* tc_ws = nullptr);

// Tensor-core path of stage 1 for M = 64 (herk_tc.cu): 1 if launched, 0 if the shape is not covered.  ws (may be null):
// covariance_tc_workspace_bytes() of zero-initialised device memory, one per stream, for sharing the frames of an incomplete
// last round between SMs (split-K with a fixed-order fold; same bits with or without it).
int launch_covariance_tc(const float2* in, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                         int avg_method, float2* out, cudaStream_t st, const float2* gains = nullptr, void* ws = nullptr);
size_t covariance_tc_workspace_bytes();

// Stage 2a.  Hermitian EVD (batched Jacobi) of R (upper triangle read, like cheevd 'U') and the noise subspace:
//   G[f][r + c*M] = sum_{n < M-T} e_n[r] conj(e_n[c])   (may be null)
//   u[f][l]       = sum_r G[r][r+l], l = 0..M-1          (complex, u[0] real; may be null)
//   w[f][M]       eigenvalues ascending                  (may be null)
int launch_noise_subspace(const float2* R, int M, int T, int nframes, float2* G, float2* u, float* w, cudaStream_t st);
// its CTA-per-matrix form for 17..64 elements (eig_block.cu); launch_noise_subspace dispatches to it
int launch_noise_subspace_block(const float2* R, int M, int T, int nframes, float2* G, float2* u, float* w, cudaStream_t st);

// calibrate_lin_array: from the one-source noise projector G and the pilot steering vector v [M] to the unit-norm
// gain/phase estimate conj(v) o u_S per frame ([nframes][M]).
int launch_calibrate_emit(const float2* R, const float2* G, const float2* v, int M, int nframes, float2* out, cudaStream_t st);

// Tables built on the host by the plan constructor (music_tables.cpp), uploaded once.
struct ScanTables {
  int M = 0, P = 0;
  const float2* z = nullptr;      // [P]   e^{j*psi_i}, psi_i = 2*pi*d*cos(theta_i): the ULA phase step per element
  const float* zpair = nullptr;   // the same z in the scan kernels' per-lane pair layout (scan_device.cuh: ZTab)
  const float2* V = nullptr;      // [P][M] steering table exactly as the reference constructor builds it
  const float* xaxis = nullptr;   // [P]   find_local_max x-axis (float-accumulated)
  const uint8_t* tctab = nullptr; // tensor-core scan: tf32 hi / lo images of the steering-power table (scan_tc.cu), or null
};

// Stage 2b+4 fused: coarse null-spectrum scan (ULA polynomial form) + local-minimum pick + refinement of the
// picked bins with the reference's own v^H G v arithmetic + dB conversion.  Outputs per frame K values/locs/bins.
int launch_scan_peaks(const float2* u, const float2* G, const ScanTables& tb, int nframes, int K, float* out_val,
                      float* out_loc, int* out_bin, cudaStream_t st);

// The same stage on the tensor cores (scan_tc.cu): the scan as a [frames x 2M-1] x [2M-1 x P] contraction (tcgen05, 3xTF32),
// peak picking out of TMEM, the same refinement.  Returns 1 if launched, 0 if the shape is not covered (M <= 16, P a
// multiple of 128, K <= 4, table present).
int launch_scan_peaks_tc(const float2* u, const float2* G, const ScanTables& tb, int nframes, int K, float* out_val,
                         float* out_loc, int* out_bin, cudaStream_t st);
bool scan_tc_covers(int M, int P, int K);
void build_scan_tc_table(float norm_spacing, int M, int P, const std::vector<float>& theta, std::vector<float>& out);

// The whole chain in one persistent kernel (fused.cu).  Returns 1 if launched, 0 if the shape is not covered (the caller
// then runs the three stage kernels), <0 on error.  Bit-identical to the three-kernel path.
int launch_chain_fused(const void* in, long long frame_stride, long long chan_stride, int M, int N, int nframes,
                       int avg_method, int T, const ScanTables& tb, int K, float* out_val, float* out_loc, int* out_bin,
                       cudaStream_t st, const float2* gains = nullptr, InputFormat fmt = InputFormat(),
                       float2* G_out = nullptr, float2* u_out = nullptr);   // G_out / u_out non-null: split form, no scan (dev builds)

// 16-element arrays: covariance + eigendecomposition in one persistent warp-specialised kernel (fused16.cu); G and u as from
// launch_noise_subspace.  Returns 1 if launched, 0 if the shape is not covered.  Bit-identical to the two stage kernels and
// SLOWER than them (measured, tools/fused16_exp.py): an experiment kept in the -DDOA_DEV_KNOBS build only.
int launch_cov_eig_fused16(const void* in, long long frame_stride, long long chan_stride, int M, int N, int nframes, int avg_method,
                           int T, float2* G, float2* u, cudaStream_t st, const float2* gains = nullptr, InputFormat fmt = InputFormat());

// Stage 2b standalone: the full dB pseudo-spectrum [nframes][P].
int launch_scan_spectrum(const float2* u, const float2* G, const ScanTables& tb, int nframes, float* out,
                         cudaStream_t st);

// Stage 4 standalone on arbitrary float vectors.
int launch_find_local_max(const float* in, int len, int nframes, int K, const float* xaxis, float* out_val,
                          float* out_loc, int* out_bin, cudaStream_t st);

// Stage 3: polynomial (from u) -> companion-matrix eigenvalues -> root selection -> angles.
// scratch: (2M-2)^2 * stride double2 (frame-interleaved working matrices), stride >= nframes.
int launch_rootmusic_scratch(const float2* u, int M, int T, float norm_spacing, int nframes, double2* scratch,
                             long long stride, float* out, cudaStream_t st);

// Per-handle options (doa_cuda_set_option, include/doa_cuda.h): kernel-path selection for A/B measurements and the multi-GPU
// SM reserve; the defaults are the shipped path.  The values live in the handle; an ABI entry point makes its handle's set
// the calling thread's current one for the duration of the call (doa_cuda.cu: Enter), the launchers read it with dev_option():
// no global mutable state, no lock, no lookup on the launch path.
enum Opt {
  OPT_FUSED, OPT_SCAN_TC, OPT_SMS_RESERVE, OPT_COV_GROUPS, OPT_COV16_RING, OPT_HERK_TC, OPT_SCAN_WIDE, OPT_SPECTRUM_SMEM,
  OPT_ROOT_ABERTH, OPT_JACOBI_SWEEPS, OPT_TMA, OPT_EIG_ONESIDED, OPT_HERK_SPLIT,
  // kernel variants that only exist in a -DDOA_DEV_KNOBS build (libdoa_cuda_dev.so, used by tools/ and the bit-identity tests)
  OPT_WS_SPLIT, OPT_WS_STAGES, OPT_WS_NBUF, OPT_WS4, OPT_WS_TMA, OPT_WS_FILL, OPT_SCAN_TC_DBG, OPT_FUSED16,
  OPT_COUNT
};
constexpr int OPT_FIRST_DEV_ONLY = OPT_WS_SPLIT;
constexpr int OPT_UNSET = -2147483647 - 1;
struct Tuning {
  int v[OPT_COUNT];
  Tuning() { for (int i = 0; i < OPT_COUNT; ++i) v[i] = OPT_UNSET; }
};
int dev_option(Opt key, int dflt);

// Host-side table construction (reference constructor arithmetic).
void build_music_tables(float norm_spacing, int M, int P, std::vector<float>& array_loc, std::vector<float>& theta,
                        std::vector<float2>& V, std::vector<float2>& z);
void build_x_axis(int len, float x_min, float x_max, std::vector<float>& x);
void build_zpair_table(const std::vector<float2>& z, std::vector<float>& out);   // scan.cu

}  // namespace doa
