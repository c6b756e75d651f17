/* MUSIC_lin_array on the GPU: io signature and the stdout counter of gr-doa lib/MUSIC_lin_array_impl.cc:47-150 kept;
 * eig_sym + the P-point steering loop + dB normalisation run in libdoa_cuda for all noutput_items matrices at once.
 * The steering/theta tables are built inside doa_cuda_music_create with the reference constructor's arithmetic. */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <gnuradio/io_signature.h>
#include <algorithm>
#include <cstdio>
#include <iostream>
#include "MUSIC_lin_array_impl.h"

namespace gr {
namespace doa {

MUSIC_lin_array::sptr MUSIC_lin_array::make(float norm_spacing, int num_targets, int num_ant_ele, int pspectrum_len) {
  return gnuradio::get_initial_sptr(new MUSIC_lin_array_impl(norm_spacing, num_targets, num_ant_ele, pspectrum_len));
}

MUSIC_lin_array_impl::MUSIC_lin_array_impl(float norm_spacing, int num_targets, int num_ant_ele, int pspectrum_len)
    : gr::sync_block("MUSIC_lin_array", gr::io_signature::make(1, 1, sizeof(gr_complex) * num_ant_ele * num_ant_ele),
                     gr::io_signature::make(1, 1, sizeof(float) * pspectrum_len)),
      d_norm_spacing(norm_spacing), d_num_targets(num_targets), d_num_ant_ele(num_ant_ele), d_pspectrum_len(pspectrum_len),
      d_cuda(NULL), nout_items_total(0) {
  d_max_frames = doa_env_int("DOA_CUDA_MAX_FRAMES", DOA_CUDA_DEFAULT_MAX_FRAMES);
  doa_require_created(doa_cuda_music_create(&d_cuda, norm_spacing, num_targets, num_ant_ele, pspectrum_len,
                                            doa_env_int("DOA_CUDA_DEVICE", 0), d_max_frames),
                      "doa.MUSIC_lin_array");
}

MUSIC_lin_array_impl::~MUSIC_lin_array_impl() {
  std::cout << "Total output items produced: " << nout_items_total << std::endl;
  doa_cuda_destroy(d_cuda);
}

int MUSIC_lin_array_impl::work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
  const gr_complex* in = (const gr_complex*)input_items[0];
  float* out = (float*)output_items[0];
  const size_t mm = (size_t)d_num_ant_ele * d_num_ant_ele;
  for (int done = 0; done < noutput_items; done += d_max_frames) {
    const int n = std::min(d_max_frames, noutput_items - done);
    if (doa_cuda_music_run(d_cuda, in + done * mm, n, out + (size_t)done * d_pspectrum_len) != DOA_CUDA_OK) {
      std::fprintf(stderr, "doa.MUSIC_lin_array: %s\n", doa_cuda_last_error(d_cuda));
      return -1;
    }
  }
  nout_items_total += noutput_items;
  return noutput_items;
}

}  // namespace doa
}  // namespace gr
