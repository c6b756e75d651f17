#ifndef INCLUDED_DOA_MUSIC_LIN_ARRAY_IMPL_H
#define INCLUDED_DOA_MUSIC_LIN_ARRAY_IMPL_H
#include <doa/MUSIC_lin_array.h>
#include "doa_cuda_block_common.h"
namespace gr {
namespace doa {
class MUSIC_lin_array_impl : public MUSIC_lin_array {
 private:
  float d_norm_spacing;
  int d_num_targets, d_num_ant_ele, d_pspectrum_len, d_max_frames;
  doa_cuda_handle* d_cuda;

 public:
  int nout_items_total;   // public counter kept from the reference (lib/MUSIC_lin_array_impl.h:47)
  MUSIC_lin_array_impl(float norm_spacing, int num_targets, int inputs, int pspectrum_len);
  ~MUSIC_lin_array_impl();
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items);
};
}  // namespace doa
}  // namespace gr
#endif
