/* autocorrelate -> rootMUSIC_linear_array in one GNU Radio block: the input side (io signature, history, forecast,
 * consume_each) is autocorrelate's (gr-doa lib/autocorrelate_impl.cc:47-118), the output is rootMUSIC_linear_array's port 0
 * (lib/rootMUSIC_linear_array_impl.cc:47-49,145: num_targets floats per item); ONE libdoa_cuda call per work(). */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <gnuradio/io_signature.h>
#include <algorithm>
#include <cstdio>
#include <stdexcept>
#include <vector>
#include "rootmusic_chain_impl.h"

namespace gr {
namespace doa {

rootmusic_chain::sptr rootmusic_chain::make(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing,
                                            int num_targets) {
  return gnuradio::get_initial_sptr(new rootmusic_chain_impl(inputs, snapshot_size, overlap_size, avg_method, norm_spacing, num_targets));
}

rootmusic_chain_impl::rootmusic_chain_impl(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing,
                                           int num_targets)
    : gr::block("rootmusic_chain", gr::io_signature::make(inputs, inputs, sizeof(gr_complex)),
                gr::io_signature::make(1, 1, num_targets * sizeof(float))),
      d_num_inputs(inputs), d_snapshot_size(snapshot_size), d_overlap_size(overlap_size), d_num_targets(num_targets),
      d_cuda(NULL), d_ptrs(inputs) {
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
  doa_require_created(doa_cuda_rootchain_create(&d_cuda, inputs, snapshot_size, overlap_size, avg_method, norm_spacing, num_targets,
                                                doa_env_int("DOA_CUDA_DEVICE", 0), d_max_frames),
                      "doa.rootmusic_chain");
}

rootmusic_chain_impl::~rootmusic_chain_impl() { doa_cuda_destroy(d_cuda); }

void rootmusic_chain_impl::set_antenna_config(const char* config_filename) {
  if (config_filename == NULL || config_filename[0] == 0) {
    if (doa_cuda_set_channel_gains(d_cuda, NULL) != DOA_CUDA_OK) throw std::runtime_error(doa_cuda_last_error(d_cuda));
    return;
  }
  std::vector<float> g(2 * (size_t)d_num_inputs);
  if (doa_cuda_antenna_gains_from_file(config_filename, d_num_inputs, &g[0]) != DOA_CUDA_OK)
    throw std::invalid_argument(doa_cuda_last_error(NULL));
  if (doa_cuda_set_channel_gains(d_cuda, &g[0]) != DOA_CUDA_OK) throw std::runtime_error(doa_cuda_last_error(d_cuda));
}

void rootmusic_chain_impl::forecast(int noutput_items, gr_vector_int& ninput_items_required) {
  for (size_t i = 0; i < ninput_items_required.size(); i++)
    ninput_items_required[i] = d_nonoverlap_size * noutput_items;   // lib/autocorrelate_impl.cc:79
}

int rootmusic_chain_impl::general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                                       gr_vector_void_star& output_items) {
  (void)ninput_items;
  float* out = (float*)output_items[0];
  for (int done = 0; done < noutput_items; done += d_max_frames) {
    const int n = std::min(d_max_frames, noutput_items - done);
    for (int k = 0; k < d_num_inputs; k++)
      d_ptrs[k] = (const gr_complex*)input_items[k] + (size_t)done * d_nonoverlap_size;
    if (doa_cuda_rootchain_run_streams(d_cuda, &d_ptrs[0], n, out + (size_t)done * d_num_targets) != DOA_CUDA_OK) {
      std::fprintf(stderr, "doa.rootmusic_chain: %s\n", doa_cuda_last_error(d_cuda));
      return -1;  // WORK_DONE
    }
  }
  consume_each(d_nonoverlap_size * noutput_items);
  return noutput_items;
}

}  // namespace doa
}  // namespace gr
