/* calibrate_lin_array on the GPU: same io signature as gr-doa lib/calibrate_lin_array_impl.cc:46-52 (one num_ant_ele^2
 * complex vector in, one num_ant_ele complex vector out per item); the two eig_sym calls per item (:118,126) become one
 * batched libdoa_cuda call for all noutput_items covariances. */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <gnuradio/io_signature.h>
#include <algorithm>
#include <cstdio>
#include "calibrate_lin_array_impl.h"

namespace gr {
namespace doa {

calibrate_lin_array::sptr calibrate_lin_array::make(float norm_spacing, int num_ant_ele, float pilot_angle) {
  return gnuradio::get_initial_sptr(new calibrate_lin_array_impl(norm_spacing, num_ant_ele, pilot_angle));
}

calibrate_lin_array_impl::calibrate_lin_array_impl(float norm_spacing, int num_ant_ele, float pilot_angle)
    : gr::sync_block("calibrate_lin_array", gr::io_signature::make(1, 1, num_ant_ele * num_ant_ele * sizeof(gr_complex)),
                     gr::io_signature::make(1, 1, num_ant_ele * sizeof(gr_complex))),
      d_norm_spacing(norm_spacing), d_num_ant_ele(num_ant_ele), d_pilot_angle(pilot_angle), d_cuda(NULL) {
  d_max_frames = doa_env_int("DOA_CUDA_MAX_FRAMES", DOA_CUDA_DEFAULT_MAX_FRAMES);
  doa_require_created(doa_cuda_calibrate_create(&d_cuda, norm_spacing, num_ant_ele, pilot_angle, doa_env_int("DOA_CUDA_DEVICE", 0),
                                                d_max_frames),
                      "doa.calibrate_lin_array");
}

calibrate_lin_array_impl::~calibrate_lin_array_impl() { doa_cuda_destroy(d_cuda); }

int calibrate_lin_array_impl::work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
  const gr_complex* in = (const gr_complex*)input_items[0];
  gr_complex* out = (gr_complex*)output_items[0];
  const size_t mm = (size_t)d_num_ant_ele * d_num_ant_ele;
  for (int done = 0; done < noutput_items; done += d_max_frames) {
    const int n = std::min(d_max_frames, noutput_items - done);
    if (doa_cuda_calibrate_run(d_cuda, in + (size_t)done * mm, n, out + (size_t)done * d_num_ant_ele) != DOA_CUDA_OK) {
      std::fprintf(stderr, "doa.calibrate_lin_array: %s\n", doa_cuda_last_error(d_cuda));
      return -1;  // WORK_DONE
    }
  }
  return noutput_items;
}

}  // namespace doa
}  // namespace gr
