#ifndef INCLUDED_DOA_ROOTMUSIC_CHAIN_IMPL_H
#define INCLUDED_DOA_ROOTMUSIC_CHAIN_IMPL_H
#include <doa/rootmusic_chain.h>
#include "doa_cuda_block_common.h"
namespace gr {
namespace doa {
class rootmusic_chain_impl : public rootmusic_chain {
 private:
  const int d_num_inputs, d_snapshot_size, d_overlap_size, d_num_targets;
  int d_nonoverlap_size, d_max_frames;
  doa_cuda_handle* d_cuda;
  std::vector<const void*> d_ptrs;

 public:
  rootmusic_chain_impl(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing, int num_targets);
  ~rootmusic_chain_impl();
  void set_antenna_config(const char* config_filename);
  void forecast(int noutput_items, gr_vector_int& ninput_items_required);
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items);
};
}  // namespace doa
}  // namespace gr
#endif
