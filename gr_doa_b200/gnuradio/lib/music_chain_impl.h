#ifndef INCLUDED_DOA_MUSIC_CHAIN_IMPL_H
#define INCLUDED_DOA_MUSIC_CHAIN_IMPL_H
#include <doa/music_chain.h>
#include "doa_cuda_block_common.h"
namespace gr {
namespace doa {
class music_chain_impl : public music_chain {
 private:
  const int d_num_inputs, d_snapshot_size, d_overlap_size, d_num_max_vals;
  int d_nonoverlap_size, d_max_frames;
  const size_t d_item_bytes;   /* sizeof(gr_complex), or 4 for sc16 items */
  doa_cuda_handle* d_cuda;
  bool d_multi;                /* several devices (DOA_CUDA_DEVICES): doa_cuda_multi_* instead of doa_cuda_chain_* */
  std::vector<const void*> d_ptrs;

 public:
  music_chain_impl(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing, int num_targets,
                   int pspectrum_len, int num_max_vals, float x_min, float x_max, float sc16_scale = 0.0f);
  ~music_chain_impl();
  void set_antenna_config(const char* config_filename);
  void forecast(int noutput_items, gr_vector_int& ninput_items_required);
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items);
};
}  // namespace doa
}  // namespace gr
#endif
