/* Minimal gr::sync_block: 1:1 rate, general_work forwards to work() and consumes noutput_items. */
#pragma once
#include <gnuradio/block.h>
namespace gr {
class sync_block : public block {
 public:
  virtual int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) = 0;
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items) {
    (void)ninput_items;
    int r = work(noutput_items, input_items, output_items);
    if (r > 0) consume_each(r);
    return r;
  }
 protected:
  sync_block() {}
  sync_block(const std::string& name, io_signature::sptr in, io_signature::sptr out) : block(name, in, out) {}
};
}  // namespace gr
