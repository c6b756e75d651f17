/* autocorrelate on the GPU.  Same io signature, history, forecast and consume_each as gr-doa
 * lib/autocorrelate_impl.cc:47-118; the per-frame memcpy + Armadillo product is replaced by ONE libdoa_cuda call for
 * all noutput_items frames (the device kernel reads the overlapping frames in place from the copied stream span). */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <gnuradio/io_signature.h>
#include <algorithm>
#include <cstdio>
#include <stdexcept>
#include <vector>
#include "autocorrelate_impl.h"

namespace gr {
namespace doa {

autocorrelate::sptr autocorrelate::make(int inputs, int snapshot_size, int overlap_size, int avg_method) {
  return gnuradio::get_initial_sptr(new autocorrelate_impl(inputs, snapshot_size, overlap_size, avg_method));
}

autocorrelate_impl::autocorrelate_impl(int inputs, int snapshot_size, int overlap_size, int avg_method)
    : gr::block("autocorrelate", gr::io_signature::make(inputs, inputs, sizeof(gr_complex)),
                gr::io_signature::make(1, 1, sizeof(gr_complex) * inputs * inputs)),
      d_num_inputs(inputs), d_snapshot_size(snapshot_size), d_overlap_size(overlap_size), d_avg_method(avg_method),
      d_cuda(NULL), d_ptrs(inputs) {
  d_nonoverlap_size = d_snapshot_size - d_overlap_size;
  set_history(d_overlap_size + 1);
  d_max_frames = doa_env_int("DOA_CUDA_MAX_FRAMES", DOA_CUDA_DEFAULT_MAX_FRAMES);
  doa_require_created(doa_cuda_autocorrelate_create(&d_cuda, inputs, snapshot_size, overlap_size, avg_method,
                                                    doa_env_int("DOA_CUDA_DEVICE", 0), d_max_frames),
                      "doa.autocorrelate");
}

autocorrelate_impl::~autocorrelate_impl() { doa_cuda_destroy(d_cuda); }

void autocorrelate_impl::set_antenna_config(const char* config_filename) {
  if (config_filename == NULL || config_filename[0] == 0) {
    if (doa_cuda_set_channel_gains(d_cuda, NULL) != DOA_CUDA_OK) throw std::runtime_error(doa_cuda_last_error(d_cuda));
    return;
  }
  std::vector<float> g(2 * (size_t)d_num_inputs);
  if (doa_cuda_antenna_gains_from_file(config_filename, d_num_inputs, &g[0]) != DOA_CUDA_OK)
    throw std::invalid_argument(doa_cuda_last_error(NULL));      // same messages as the reference block
  if (doa_cuda_set_channel_gains(d_cuda, &g[0]) != DOA_CUDA_OK) throw std::runtime_error(doa_cuda_last_error(d_cuda));
}

void autocorrelate_impl::forecast(int noutput_items, gr_vector_int& ninput_items_required) {
  for (size_t i = 0; i < ninput_items_required.size(); i++)
    ninput_items_required[i] = doa_cuda_autocorrelate_forecast(d_cuda, noutput_items);
}

int autocorrelate_impl::general_work(int output_matrices, gr_vector_int& ninput_items,
                                     gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
  (void)ninput_items;
  gr_complex* out = (gr_complex*)output_items[0];
  for (int done = 0; done < output_matrices; done += d_max_frames) {
    const int n = std::min(d_max_frames, output_matrices - done);
    for (int k = 0; k < d_num_inputs; k++)
      d_ptrs[k] = (const gr_complex*)input_items[k] + (size_t)done * d_nonoverlap_size;
    const int rc = doa_cuda_autocorrelate_run(d_cuda, &d_ptrs[0], n, out + (size_t)done * d_num_inputs * d_num_inputs);
    if (rc != DOA_CUDA_OK) {
      std::fprintf(stderr, "doa.autocorrelate: %s\n", doa_cuda_last_error(d_cuda));
      return -1;  // WORK_DONE
    }
  }
  consume_each(d_nonoverlap_size * output_matrices);
  return output_matrices;
}

}  // namespace doa
}  // namespace gr
