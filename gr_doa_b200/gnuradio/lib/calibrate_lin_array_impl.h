#ifndef INCLUDED_DOA_CALIBRATE_LIN_ARRAY_IMPL_H
#define INCLUDED_DOA_CALIBRATE_LIN_ARRAY_IMPL_H
#include <doa/calibrate_lin_array.h>
#include "doa_cuda_block_common.h"
namespace gr {
namespace doa {
class calibrate_lin_array_impl : public calibrate_lin_array {
 private:
  const float d_norm_spacing;
  const int d_num_ant_ele;
  const float d_pilot_angle;
  int d_max_frames;
  doa_cuda_handle* d_cuda;

 public:
  calibrate_lin_array_impl(float norm_spacing, int num_ant_ele, float pilot_angle);
  ~calibrate_lin_array_impl();
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items);
};
}  // namespace doa
}  // namespace gr
#endif
