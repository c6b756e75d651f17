/* autocorrelate -> MUSIC_lin_array -> find_local_max in one GNU Radio block: the input side (io signature, history,
 * forecast, consume_each) is autocorrelate's (gr-doa lib/autocorrelate_impl.cc:47-118), the output side find_local_max's
 * (lib/find_local_max_impl.cc:47-56: two ports of K floats); everything in between is ONE libdoa_cuda call per work(). */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <gnuradio/io_signature.h>
#include <algorithm>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>
#include "music_chain_impl.h"

namespace gr {
namespace doa {

music_chain::sptr music_chain::make(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing,
                                    int num_targets, int pspectrum_len, int num_max_vals, float x_min, float x_max) {
  return gnuradio::get_initial_sptr(new music_chain_impl(inputs, snapshot_size, overlap_size, avg_method, norm_spacing,
                                                         num_targets, pspectrum_len, num_max_vals, x_min, x_max));
}

music_chain::sptr music_chain::make_sc16(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing,
                                         int num_targets, int pspectrum_len, int num_max_vals, float x_min, float x_max,
                                         float sc16_scale) {
  if (!(sc16_scale > 0.0f)) throw std::invalid_argument("doa.music_chain: sc16_scale must be > 0");
  return gnuradio::get_initial_sptr(new music_chain_impl(inputs, snapshot_size, overlap_size, avg_method, norm_spacing,
                                                         num_targets, pspectrum_len, num_max_vals, x_min, x_max, sc16_scale));
}

music_chain_impl::music_chain_impl(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing,
                                   int num_targets, int pspectrum_len, int num_max_vals, float x_min, float x_max,
                                   float sc16_scale)
    : gr::block("music_chain", gr::io_signature::make(inputs, inputs, sc16_scale > 0.0f ? 2 * sizeof(short) : sizeof(gr_complex)),
                gr::io_signature::make2(2, 2, num_max_vals * sizeof(float), num_max_vals * sizeof(float))),
      d_num_inputs(inputs), d_snapshot_size(snapshot_size), d_overlap_size(overlap_size), d_num_max_vals(num_max_vals),
      d_item_bytes(sc16_scale > 0.0f ? 2 * sizeof(short) : sizeof(gr_complex)), d_cuda(NULL), d_multi(false), d_ptrs(inputs) {
  d_nonoverlap_size = d_snapshot_size - d_overlap_size;
  set_history(d_overlap_size + 1);
  // Scheduler batching: the reference's blocks take whatever noutput_items the scheduler offers (often 1); a GPU call
  // wants a batch.  Never call work() for fewer than DOA_CUDA_MIN_FRAMES frames, and ask for output buffers of four such
  // batches so that the upstream blocks can run ahead while a batch is on the device.
  {
    const int min_frames = std::max(1, doa_env_int("DOA_CUDA_MIN_FRAMES", DOA_CUDA_DEFAULT_MIN_FRAMES));
    set_output_multiple(min_frames);
    set_min_output_buffer(4L * min_frames);
  }
  d_max_frames = doa_env_int("DOA_CUDA_MAX_FRAMES", DOA_CUDA_DEFAULT_MAX_FRAMES);
  const std::vector<int> devs = doa_env_devices();
  d_multi = devs.size() > 1;
  if (d_multi) {   /* one block instance, every listed GPU: frames of a work() call are cut into one block per device */
    const int per_dev = (d_max_frames + (int)devs.size() - 1) / (int)devs.size();
    doa_require_created(doa_cuda_multi_create(&d_cuda, inputs, snapshot_size, overlap_size, avg_method, norm_spacing, num_targets,
                                              pspectrum_len, num_max_vals, x_min, x_max, &devs[0], (int)devs.size(), per_dev),
                        "doa.music_chain");
  } else {
    doa_require_created(doa_cuda_chain_create(&d_cuda, inputs, snapshot_size, overlap_size, avg_method, norm_spacing, num_targets,
                                              pspectrum_len, num_max_vals, x_min, x_max, devs[0], d_max_frames),
                        "doa.music_chain");
  }
  if (sc16_scale > 0.0f && doa_cuda_set_input_format(d_cuda, DOA_CUDA_FMT_SC16, sc16_scale) != DOA_CUDA_OK) {
    const std::string msg = std::string("doa.music_chain: ") + doa_cuda_last_error(d_cuda);
    doa_cuda_destroy(d_cuda);
    throw std::runtime_error(msg);
  }
}

music_chain_impl::~music_chain_impl() { doa_cuda_destroy(d_cuda); }

void music_chain_impl::set_antenna_config(const char* config_filename) {
  if (config_filename == NULL || config_filename[0] == 0) {
    if (doa_cuda_set_channel_gains(d_cuda, NULL) != DOA_CUDA_OK) throw std::runtime_error(doa_cuda_last_error(d_cuda));
    return;
  }
  std::vector<float> g(2 * (size_t)d_num_inputs);
  if (doa_cuda_antenna_gains_from_file(config_filename, d_num_inputs, &g[0]) != DOA_CUDA_OK)
    throw std::invalid_argument(doa_cuda_last_error(NULL));
  if (doa_cuda_set_channel_gains(d_cuda, &g[0]) != DOA_CUDA_OK) throw std::runtime_error(doa_cuda_last_error(d_cuda));
}

void music_chain_impl::forecast(int noutput_items, gr_vector_int& ninput_items_required) {
  for (size_t i = 0; i < ninput_items_required.size(); i++)
    ninput_items_required[i] = d_nonoverlap_size * noutput_items;   // lib/autocorrelate_impl.cc:79
}

int music_chain_impl::general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                                   gr_vector_void_star& output_items) {
  (void)ninput_items;
  float* out1 = (float*)output_items[0];
  float* out2 = (float*)output_items[1];
  for (int done = 0; done < noutput_items; done += d_max_frames) {
    const int n = std::min(d_max_frames, noutput_items - done);
    for (int k = 0; k < d_num_inputs; k++)
      d_ptrs[k] = (const char*)input_items[k] + (size_t)done * d_nonoverlap_size * d_item_bytes;
    const int rc = d_multi ? doa_cuda_multi_run_streams(d_cuda, &d_ptrs[0], n, out1 + (size_t)done * d_num_max_vals,
                                                        out2 + (size_t)done * d_num_max_vals, NULL)
                           : doa_cuda_chain_run_streams(d_cuda, &d_ptrs[0], n, out1 + (size_t)done * d_num_max_vals,
                                                        out2 + (size_t)done * d_num_max_vals, NULL);
    if (rc != DOA_CUDA_OK) {
      std::fprintf(stderr, "doa.music_chain: %s\n", doa_cuda_last_error(d_cuda));
      return -1;  // WORK_DONE
    }
  }
  consume_each(d_nonoverlap_size * noutput_items);
  return noutput_items;
}

}  // namespace doa
}  // namespace gr
