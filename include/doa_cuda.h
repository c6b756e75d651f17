/* doa_cuda.h -- C ABI of libdoa_cuda, the B200 (sm_100a) implementation of gr-doa's DoA hot path.
 *
 * One opaque handle per GNU Radio block instance.  Every entry point takes plain pointers and sizes; complex
 * samples are interleaved float pairs (gr_complex).  Every function returns 0 on success or a negative
 * DOA_CUDA_E* code; doa_cuda_last_error(handle) gives the text.  No exception crosses this boundary and no
 * CPU fallback exists behind it: without a usable CUDA device, *_create fails.
 *
 * Each stage S in {autocorrelate, music, rootmusic, find_local_max} has
 *     doa_cuda_S_create      the block constructor   (same parameters as S::make in the reference)
 *     doa_cuda_S_run         the body of work()/general_work(): HOST pointers, synchronous
 *     doa_cuda_S_run_device  the same on DEVICE pointers, asynchronous on the caller's stream
 *     doa_cuda_S_destroy
 * plus doa_cuda_chain_*     = autocorrelate -> MUSIC_lin_array -> find_local_max fused, peaks only;
 *      doa_cuda_rootchain_* = autocorrelate -> rootMUSIC_linear_array, angles only;
 *      doa_cuda_multi_*     = the fused chain over several GPUs from one process;
 *      doa_cuda_calibrate_*, doa_cuda_set_channel_gains, doa_cuda_set_input_format (sc16 samples),
 *      doa_cuda_pin_host_buffer = the steps either side of the path.
 *
 * Reference interfaces replaced (paths relative to the gr-doa tree):
 *     autocorrelate::make / general_work / forecast     include/doa/autocorrelate.h:56, lib/autocorrelate_impl.cc:47-118
 *     MUSIC_lin_array::make / work                      include/doa/MUSIC_lin_array.h:56, lib/MUSIC_lin_array_impl.cc:47-150
 *     rootMUSIC_linear_array::make / work               include/doa/rootMUSIC_linear_array.h:54, lib/rootMUSIC_linear_array_impl.cc:46-152
 *     find_local_max::make / work                       include/doa/find_local_max.h:56, lib/find_local_max_impl.cc:47-194
 */
#ifndef DOA_CUDA_H
#define DOA_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define DOA_CUDA_OK 0
#define DOA_CUDA_EINVAL (-1)    /* bad argument (also the GRC <check>s: overlap < snapshot, inputs > targets, spacing <= 0.5) */
#define DOA_CUDA_ECUDA (-2)     /* a CUDA runtime call failed */
#define DOA_CUDA_ENOMEM (-3)    /* host or device allocation failed */
#define DOA_CUDA_ECAPACITY (-4) /* nframes exceeds the max_frames given at create */

typedef struct doa_cuda_handle doa_cuda_handle; /* opaque; one per block instance, owns a stream + staging */

int doa_cuda_abi_version(void);
/* Text of the last error on this handle (or of the last failed *_create when handle == NULL). */
const char* doa_cuda_last_error(const doa_cuda_handle* h);
int doa_cuda_device_count(void);

/* ---- stage 1: autocorrelate (lib/autocorrelate_impl.cc) --------------------------------------------------
 * R[r,c] = (1/N) sum_t x_r[t] conj(x_c[t]); avg_method 1 adds the reference's forward-backward term
 * 0.5 R + (0.5/N) J conj(R) J (including its extra 1/N).  Output: nframes x (M*M) complex, column-major. */
int doa_cuda_autocorrelate_create(doa_cuda_handle** h, int inputs, int snapshot_size, int overlap_size, int avg_method,
                                  int device, int max_frames);
/* in_host: `inputs` pointers, each to hop*(nframes-1)+snapshot_size complex samples (hop = snapshot-overlap;
 * frame i of channel k starts at in_host[k] + i*hop, exactly as general_work() reads them). */
int doa_cuda_autocorrelate_run(doa_cuda_handle* h, const void* const* in_host, int nframes, void* out_host);
/* Device form.  Sample (frame f, channel k, time t) is at in_dev[f*frame_stride + k*chan_stride + t] (strides in
 * complex samples): streaming = (hop, stream_len); independent frames [B][M][N] = (M*N, N). */
int doa_cuda_autocorrelate_run_device(doa_cuda_handle* h, const void* in_dev, long long frame_stride, long long chan_stride,
                                      int nframes, void* out_dev, void* cuda_stream);
/* forecast(): input items required per port for noutput_items (lib/autocorrelate_impl.cc:74-80). */
int doa_cuda_autocorrelate_forecast(const doa_cuda_handle* h, int noutput_items);

/* ---- stage 2: MUSIC_lin_array (lib/MUSIC_lin_array_impl.cc) ------------------------------------------------
 * in: nframes x (M*M) complex column-major covariance; out: nframes x pspectrum_len float, dB, peak = 0. */
int doa_cuda_music_create(doa_cuda_handle** h, float norm_spacing, int num_targets, int num_ant_ele, int pspectrum_len,
                          int device, int max_frames);
int doa_cuda_music_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_host);
int doa_cuda_music_run_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* out_dev, void* cuda_stream);
/* The eigendecomposition half of the stage on its own (eig_sym + U_N*trans(U_N), lib/MUSIC_lin_array_impl.cc:128-133):
 * G_dev nframes x (M*M) complex noise-subspace projector, u_dev nframes x M complex diagonal sums of G, w_dev nframes x M
 * eigenvalues ascending; any of the three may be NULL.  Valid on music, rootmusic and chain handles. */
int doa_cuda_music_noise_subspace_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* G_dev, void* u_dev,
                                         void* w_dev, void* cuda_stream);
/* Host copies of the constructor tables (for table-parity tests): array_loc[M], theta_rad[P], steering[P][M] complex. */
int doa_cuda_music_get_tables(const doa_cuda_handle* h, float* array_loc, float* theta_rad, float* steering);

/* ---- stage 3: rootMUSIC_linear_array (lib/rootMUSIC_linear_array_impl.cc) ------------------------------------
 * in: nframes x (M*M) complex; out: nframes x num_targets float, degrees ascending (NaN where the reference is
 * undefined: fewer than num_targets roots strictly inside the unit circle). */
int doa_cuda_rootmusic_create(doa_cuda_handle** h, float norm_spacing, int num_targets, int num_ant_ele, int device,
                              int max_frames);
int doa_cuda_rootmusic_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_host);
int doa_cuda_rootmusic_run_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* out_dev, void* cuda_stream);

/* ---- stage 4: find_local_max (lib/find_local_max_impl.cc) ----------------------------------------------------
 * in: nframes x vector_len float; out_val: nframes x K peak heights (descending), out_loc: nframes x K x-axis
 * locations (descending by x, as the reference sorts them), out_bin (optional, may be NULL): nframes x K int32
 * peak bins in out_val order (not a block port; exported for bit-exact comparisons). */
int doa_cuda_find_local_max_create(doa_cuda_handle** h, int num_max_vals, int vector_len, float x_min, float x_max,
                                   int device, int max_frames);
int doa_cuda_find_local_max_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_val_host,
                                void* out_loc_host, void* out_bin_host);
int doa_cuda_find_local_max_run_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* out_val_dev,
                                       void* out_loc_dev, void* out_bin_dev, void* cuda_stream);

/* ---- fused chain: autocorrelate -> MUSIC_lin_array -> find_local_max(K, P, x_min, x_max), peaks only ----------- */
int doa_cuda_chain_create(doa_cuda_handle** h, int inputs, int snapshot_size, int overlap_size, int avg_method,
                          float norm_spacing, int num_targets, int pspectrum_len, int num_max_vals, float x_min,
                          float x_max, int device, int max_frames);
int doa_cuda_chain_run_device(doa_cuda_handle* h, const void* in_dev, long long frame_stride, long long chan_stride,
                              int nframes, void* out_val_dev, void* out_loc_dev, void* out_bin_dev, void* cuda_stream);
/* Host form on independent frames [nframes][inputs][snapshot_size] complex (pageable or pinned): copies are
 * chunked and overlapped with the kernels on the handle's own streams; returns when the outputs are in host memory. */
int doa_cuda_chain_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_val_host, void* out_loc_host,
                       void* out_bin_host);
/* Streaming host form (what an autocorrelate block feeding the chain sees): `inputs` channel pointers. */
int doa_cuda_chain_run_streams(doa_cuda_handle* h, const void* const* in_host, int nframes, void* out_val_host,
                               void* out_loc_host, void* out_bin_host);
/* Number of kernel launches issued by the last run on this handle (for bench accounting). */
int doa_cuda_last_launch_count(const doa_cuda_handle* h);
/* Stage timing of the chain's run_device calls: after doa_cuda_set_profiling(h, 1) every call brackets its three stages
 * (covariance, eigendecomposition, scan+peaks) with CUDA events on the stream the kernels run on;
 * doa_cuda_chain_stage_ms returns the number of calls recorded since then (the last 256 at most) and their mean stage
 * times in ms, or a negative error. */
int doa_cuda_set_profiling(doa_cuda_handle* h, int on);
int doa_cuda_chain_stage_ms(doa_cuda_handle* h, float* cov_ms, float* eig_ms, float* scan_ms);

/* ---- calibrate_lin_array ---------------------------------------------------------------
 * gr::doa::calibrate_lin_array(norm_spacing, num_ant_ele, pilot_angle), lib/calibrate_lin_array_impl.cc:46-134: per input
 * covariance (num_ant_ele^2 complex, column-major) one vector of num_ant_ele complex antenna gain/phase estimates from a
 * pilot at pilot_angle degrees (Soon et al. 1994).  The reference returns the eigenvector LAPACK happens to produce, i.e.
 * the estimate is defined up to a unit-modulus factor; here it has unit norm and the phase reference is the element of
 * largest magnitude (the signal eigenvector is taken with that element real and positive). */
int doa_cuda_calibrate_create(doa_cuda_handle** out, float norm_spacing, int num_ant_ele, float pilot_angle, int device,
                              int max_frames);
int doa_cuda_calibrate_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_host);
int doa_cuda_calibrate_run_device(doa_cuda_handle* h, const void* in_dev, int nframes, void* out_dev, void* cuda_stream);

/* ---- autocorrelate -> rootMUSIC_linear_array in one call (the Root-MUSIC flowgraph, BASELINE configs[1]) ---------------
 * lib/autocorrelate_impl.cc:82-118 followed by lib/rootMUSIC_linear_array_impl.cc:90-152 with the covariance staying on the
 * device: samples in (the three layouts of the chain: device strides, host frames [n][M][N], host channel streams),
 * num_targets ascending angles per frame out.  Same kernels as the two separate stages: identical bits.
 * doa_cuda_set_channel_gains / doa_cuda_set_input_format apply. */
int doa_cuda_rootchain_create(doa_cuda_handle** h, int inputs, int snapshot_size, int overlap_size, int avg_method,
                              float norm_spacing, int num_targets, int device, int max_frames);
int doa_cuda_rootchain_run_device(doa_cuda_handle* h, const void* in_dev, long long frame_stride, long long chan_stride,
                                  int nframes, void* out_aoa_dev, void* cuda_stream);
int doa_cuda_rootchain_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_aoa_host);
int doa_cuda_rootchain_run_streams(doa_cuda_handle* h, const void* const* in_host, int nframes, void* out_aoa_host);

/* ---- the fused chain on several GPUs from one process --------------------------------------
 * A GNU Radio flowgraph is one process; this form lets one block instance use every GPU of the box.  Frames are independent
 * in all four reference blocks (lib/autocorrelate_impl.cc:92, lib/MUSIC_lin_array_impl.cc:121,
 * lib/rootMUSIC_linear_array_impl.cc:105, lib/find_local_max_impl.cc:179), so a batch of independent frames ([nframes][M][N]
 * host memory, as doa_cuda_chain_run) is cut into contiguous blocks, one per entry of `devices` (sizes differ by at most one
 * frame; doa_cuda_multi_block reports them).  Each device copies and processes its block concurrently (the calling thread
 * drives the first device, persistent worker threads owned by the handle the others) and writes its peaks into the caller's output arrays at the block's offset -- that is the whole
 * gather, so per-frame results are bit-identical to doa_cuda_chain_run on one device.  A device may be listed more than once
 * (that many independent stream sets on it).  set_channel_gains / set_input_format apply to every device.
 * The one-process-per-GPU form (torchrun, NCCL gather of device-resident peaks) lives in gr_doa_b200/sharding.py. */
int doa_cuda_multi_create(doa_cuda_handle** h, int inputs, int snapshot_size, int overlap_size, int avg_method,
                          float norm_spacing, int num_targets, int pspectrum_len, int num_max_vals, float x_min, float x_max,
                          const int* devices, int ndevices, int max_frames_per_device);
int doa_cuda_multi_run(doa_cuda_handle* h, const void* in_host, int nframes, void* out_val_host, void* out_loc_host,
                       void* out_bin_host /* may be NULL */);
/* Streaming form, arguments as doa_cuda_chain_run_streams (`inputs` channel pointers, frame i at hop*i): each device reads
 * its block of frames from the same host streams, overlap samples included. */
int doa_cuda_multi_run_streams(doa_cuda_handle* h, const void* const* in_host, int nframes, void* out_val_host,
                               void* out_loc_host, void* out_bin_host /* may be NULL */);
int doa_cuda_multi_device_count(const doa_cuda_handle* h);
int doa_cuda_multi_block(const doa_cuda_handle* h, int nframes, int index, int* first, int* count);

/* ---- channel gains in front of the covariance ---------------------------------------------
 * Replaces the antenna_correction block (lib/antenna_correction_impl.cc:47-99: out_k[i] = g_k * in_k[i]) and
 * python/phase_correct_hier.py:91-102 (g_k = e^{j phi_k}) when they feed autocorrelate: instead of two more passes over the
 * sample stream the gains are folded into the covariance, R' = D R D^H, D = diag(g).  Valid on an autocorrelate, chain or
 * rootchain handle; `gains` = `inputs` complex floats (re, im interleaved, host memory), NULL restores "no gains".  Takes effect for
 * the following runs. */
int doa_cuda_set_channel_gains(doa_cuda_handle* h, const float* gains);
/* The reference constructor's config-file reader (lib/antenna_correction_impl.cc:54-74): one "gain phase" pair per
 * channel, g_k = (1/gain_k) e^{-j phase_k}; fails like the reference (missing file, too many / too few lines).
 * Writes num_ant_ele complex floats to gains_out. */
int doa_cuda_antenna_gains_from_file(const char* config_filename, int num_ant_ele, float* gains_out);

/* ---- sample format of the covariance input -------------------------------------------------
 * The reference's flowgraphs ask UHD for cpu_format "fc32" (python/twinrx_usrp_source.py:57): the host converts the
 * radio's int16 I/Q pairs to gr_complex before autocorrelate reads them.  With DOA_CUDA_FMT_SC16 an autocorrelate, chain,
 * rootchain or multi handle reads the int16 pairs directly (UHD cpu_format "sc16": one little-endian 32-bit word per complex sample,
 * I in the low half), converts exactly inside the covariance kernel, and the value of a sample is int16 * scale (UHD's
 * converter uses 1/32767; 1/32768 is a power of two).  That halves the bytes the chain moves over PCIe and HBM.
 * Every `in` pointer of the run functions (host and device) is then read as sc16; strides stay in complex samples.
 * Result: for a power-of-two scale bit-identical to the fc32 path fed float(int16) * scale; for any other scale the
 * covariance differs from that by at most 2 ulp per entry (the scale enters once, squared, instead of per sample).
 * Takes effect for the following runs; DOA_CUDA_FMT_FC32 (the default) ignores `scale`. */
#define DOA_CUDA_FMT_FC32 0
#define DOA_CUDA_FMT_SC16 1
int doa_cuda_set_input_format(doa_cuda_handle* h, int format, float scale);

/* ---- per-handle options ------------------------------------------------------------------------------------------------
 * Path selection for A/B measurements and the multi-GPU SM reserve.  The defaults are the shipped path; an option is stored in
 * the handle (a multi handle forwards it to its per-device handles) and read at launch time -- there is no process-global
 * state.  Keys (value = int):
 *   "sms_reserve"    0..64, default 0   SMs the persistent chain kernel leaves free, so that a collective's kernel (the NCCL peak
 *                                       gather of a multi-GPU run) can run under the next batch's chain kernel
 *   "fused"          default 1          0: run the chain as its stage kernels (covariance, eigendecomposition, scan + peaks)
 *                                       instead of the persistent warp-specialised kernel (4- and 8-element arrays)
 *   "tma"            default 1          0: per-lane cp.async ring fills in the fused kernel instead of tensor-map TMA boxes
 *   "scan_tc"        default 1          0: Horner scan on the CUDA cores instead of the tensor-core contraction (unfused chain)
 *   "herk_tc"        default 1          0: CUDA-core tiled covariance at 64 elements instead of the tensor-core HERK
 *   "herk_split"     default 1          0: the tensor-core HERK gives every frame to one SM even when the batch does not fill a
 *                                       round of the SMs (1: such frames are shared between SMs, same bits)
 *   "root_aberth"    default 1          0: Root-MUSIC by Hessenberg QR only
 *   "eig_onesided"   default 1          0: two-sided Jacobi eigensolver at 8..64 elements instead of the one-sided Jacobi on the
 *                                       Cholesky factor (results agree to ~1e-6 in the noise projector; NOT bit-identical);
 *                                       2: at 17..64 elements pair columns instead of two-column blocks (comparison)
 *   "scan_wide", "spectrum_smem", "cov16_ring", "cov_groups", "jacobi_sweeps"   kernel variants of single stages
 * A library built with -DDOA_DEV_KNOBS (libdoa_cuda_dev.so: tools/, bit-identity tests) also carries the experimental fused
 * kernel configurations and accepts "ws_split", "ws_stages", "ws_nbuf", "ws4", "ws_tma", "ws_fill", "scan_tc_dbg", "fused16";
 * the product library answers DOA_CUDA_EINVAL to those.  Every variant computes identical results unless its description says
 * otherwise. */
int doa_cuda_set_option(doa_cuda_handle* h, const char* key, int value);
int doa_cuda_has_dev_knobs(void);   /* 1 in a -DDOA_DEV_KNOBS build */

/* ---- threading and streams ---------------------------------------------------------------------------------------------
 * A handle serves ONE call at a time (a GNU Radio block's work() is called serially): *_run_device reuses per-handle scratch
 * buffers on the caller's stream, so calls on the same handle from two streams or two threads must not overlap.  Different
 * handles are independent.  Every entry point makes the handle's device current for the duration of the call and restores the
 * calling thread's previous device before it returns; `cuda_stream` must belong to the handle's device. */

/* ---- page-locking the caller's buffers -------------------------------------------------------------------------------
 * The *_run entry points accept any host memory.  Out of pageable memory (a GNU Radio scheduler's buffers) the host->device
 * copy is staged by the driver and runs at a fraction of the PCIe rate; a flowgraph's circular buffers live as long as the
 * flowgraph, so a maintainer who knows their extent (gr::buffer base and size) can page-lock them once with this wrapper
 * around cudaHostRegister and release them before the buffers are freed.  Registering a range twice is not an error.
 * No handle needed; errors are reported through doa_cuda_last_error(NULL). */
int doa_cuda_pin_host_buffer(void* p, unsigned long long bytes);
int doa_cuda_unpin_host_buffer(void* p);

void doa_cuda_destroy(doa_cuda_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* DOA_CUDA_H */
