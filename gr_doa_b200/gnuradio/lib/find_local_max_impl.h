#ifndef INCLUDED_DOA_FIND_LOCAL_MAX_IMPL_H
#define INCLUDED_DOA_FIND_LOCAL_MAX_IMPL_H
#include <doa/find_local_max.h>
#include "doa_cuda_block_common.h"
namespace gr {
namespace doa {
class find_local_max_impl : public find_local_max {
 private:
  const int d_num_max_vals, d_vector_len;
  const float d_x_min, d_x_max;
  int d_max_frames;
  doa_cuda_handle* d_cuda;

 public:
  find_local_max_impl(int num_max_vals, int vector_len, float x_min, float x_max);
  ~find_local_max_impl();
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items);
};
}  // namespace doa
}  // namespace gr
#endif
