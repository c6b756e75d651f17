/* rootMUSIC_linear_array on the GPU.  Port layout of gr-doa lib/rootMUSIC_linear_array_impl.cc:46-152 kept: 1..T output
 * ports of T floats, only port 0 is written (:96,145). */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <gnuradio/io_signature.h>
#include <algorithm>
#include <cstdio>
#include "rootMUSIC_linear_array_impl.h"

namespace gr {
namespace doa {

rootMUSIC_linear_array::sptr rootMUSIC_linear_array::make(float norm_spacing, int num_targets, int num_ant_ele) {
  return gnuradio::get_initial_sptr(new rootMUSIC_linear_array_impl(norm_spacing, num_targets, num_ant_ele));
}

rootMUSIC_linear_array_impl::rootMUSIC_linear_array_impl(float norm_spacing, int num_targets, int num_ant_ele)
    : gr::sync_block("rootMUSIC_linear_array", gr::io_signature::make(1, 1, sizeof(gr_complex) * num_ant_ele * num_ant_ele),
                     gr::io_signature::make(1, num_targets, num_targets * sizeof(float))),
      d_norm_spacing(norm_spacing), d_num_targets(num_targets), d_num_ant_ele(num_ant_ele), d_cuda(NULL) {
  d_max_frames = doa_env_int("DOA_CUDA_MAX_FRAMES", DOA_CUDA_DEFAULT_MAX_FRAMES);
  doa_require_created(doa_cuda_rootmusic_create(&d_cuda, norm_spacing, num_targets, num_ant_ele,
                                                doa_env_int("DOA_CUDA_DEVICE", 0), d_max_frames),
                      "doa.rootMUSIC_linear_array");
}

rootMUSIC_linear_array_impl::~rootMUSIC_linear_array_impl() { doa_cuda_destroy(d_cuda); }

int rootMUSIC_linear_array_impl::work(int noutput_items, gr_vector_const_void_star& input_items,
                                      gr_vector_void_star& output_items) {
  const gr_complex* in = (const gr_complex*)input_items[0];
  float* out = (float*)output_items[0];
  const size_t mm = (size_t)d_num_ant_ele * d_num_ant_ele;
  for (int done = 0; done < noutput_items; done += d_max_frames) {
    const int n = std::min(d_max_frames, noutput_items - done);
    if (doa_cuda_rootmusic_run(d_cuda, in + done * mm, n, out + (size_t)done * d_num_targets) != DOA_CUDA_OK) {
      std::fprintf(stderr, "doa.rootMUSIC_linear_array: %s\n", doa_cuda_last_error(d_cuda));
      return -1;
    }
  }
  return noutput_items;
}

}  // namespace doa
}  // namespace gr
