#ifndef INCLUDED_DOA_ROOTMUSIC_LINEAR_ARRAY_IMPL_H
#define INCLUDED_DOA_ROOTMUSIC_LINEAR_ARRAY_IMPL_H
#include <doa/rootMUSIC_linear_array.h>
#include "doa_cuda_block_common.h"
namespace gr {
namespace doa {
class rootMUSIC_linear_array_impl : public rootMUSIC_linear_array {
 private:
  float d_norm_spacing;
  int d_num_targets, d_num_ant_ele, d_max_frames;
  doa_cuda_handle* d_cuda;

 public:
  rootMUSIC_linear_array_impl(float norm_spacing, int num_targets, int num_ant_ele);
  ~rootMUSIC_linear_array_impl();
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items);
};
}  // namespace doa
}  // namespace gr
#endif
