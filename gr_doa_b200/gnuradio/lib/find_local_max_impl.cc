/* find_local_max on the GPU: two output ports of K floats (peak heights; x locations sorted descending), as gr-doa
 * lib/find_local_max_impl.cc:47-194.  The K == 1 / K > 1 function-pointer switch of the reference (:57) lives inside
 * libdoa_cuda. */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <gnuradio/io_signature.h>
#include <algorithm>
#include <cstdio>
#include "find_local_max_impl.h"

namespace gr {
namespace doa {

find_local_max::sptr find_local_max::make(int num_max_vals, int vector_len, float x_min, float x_max) {
  return gnuradio::get_initial_sptr(new find_local_max_impl(num_max_vals, vector_len, x_min, x_max));
}

find_local_max_impl::find_local_max_impl(int num_max_vals, int vector_len, float x_min, float x_max)
    : gr::sync_block("find_local_max", gr::io_signature::make(1, 1, sizeof(float) * vector_len),
                     gr::io_signature::make2(2, 2, num_max_vals * sizeof(float), num_max_vals * sizeof(float))),
      d_num_max_vals(num_max_vals), d_vector_len(vector_len), d_x_min(x_min), d_x_max(x_max), d_cuda(NULL) {
  d_max_frames = doa_env_int("DOA_CUDA_MAX_FRAMES", DOA_CUDA_DEFAULT_MAX_FRAMES);
  doa_require_created(doa_cuda_find_local_max_create(&d_cuda, num_max_vals, vector_len, x_min, x_max,
                                                     doa_env_int("DOA_CUDA_DEVICE", 0), d_max_frames),
                      "doa.find_local_max");
}

find_local_max_impl::~find_local_max_impl() { doa_cuda_destroy(d_cuda); }

int find_local_max_impl::work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
  const float* in = (const float*)input_items[0];
  float* out1 = (float*)output_items[0];
  float* out2 = (float*)output_items[1];
  for (int done = 0; done < noutput_items; done += d_max_frames) {
    const int n = std::min(d_max_frames, noutput_items - done);
    if (doa_cuda_find_local_max_run(d_cuda, in + (size_t)done * d_vector_len, n, out1 + (size_t)done * d_num_max_vals,
                                    out2 + (size_t)done * d_num_max_vals, NULL) != DOA_CUDA_OK) {
      std::fprintf(stderr, "doa.find_local_max: %s\n", doa_cuda_last_error(d_cuda));
      return -1;
    }
  }
  return noutput_items;
}

}  // namespace doa
}  // namespace gr
