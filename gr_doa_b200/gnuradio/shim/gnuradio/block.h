/* Minimal gr::block for compile checks and the fake scheduler: name, signatures, history, forecast, consume_each,
 * general_work.  Semantics follow GNU Radio 3.7's gr::block as far as the four DoA blocks exercise them. */
#pragma once
#include <gnuradio/io_signature.h>
#include <gnuradio/types.h>
#include <string>
namespace gr {
class block {
 public:
  enum { WORK_CALLED_PRODUCE = -2, WORK_DONE = -1 };
  virtual ~block() {}
  const std::string& name() const { return d_name; }
  io_signature::sptr input_signature() const { return d_in; }
  io_signature::sptr output_signature() const { return d_out; }
  unsigned history() const { return d_history; }
  void set_history(unsigned h) { d_history = h; }
  virtual void forecast(int noutput_items, gr_vector_int& ninput_items_required) {
    for (size_t i = 0; i < ninput_items_required.size(); i++) ninput_items_required[i] = noutput_items + history() - 1;
  }
  virtual int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                           gr_vector_void_star& output_items) = 0;
  void set_output_multiple(int m) { d_output_multiple = m; }
  int output_multiple() const { return d_output_multiple; }
  void set_min_output_buffer(long n) { d_min_output_buffer = n; }
  long min_output_buffer() const { return d_min_output_buffer; }
  void consume_each(int n) { d_consumed = n; }
  int last_consumed() const { return d_consumed; }   // shim only: what the scheduler would advance the read pointers by
 protected:
  block() : d_history(1), d_consumed(0), d_output_multiple(1), d_min_output_buffer(-1) {}
  block(const std::string& name, io_signature::sptr in, io_signature::sptr out)
      : d_name(name), d_in(in), d_out(out), d_history(1), d_consumed(0), d_output_multiple(1), d_min_output_buffer(-1) {}
 private:
  std::string d_name; io_signature::sptr d_in, d_out; unsigned d_history; int d_consumed; int d_output_multiple; long d_min_output_buffer;
};
}  // namespace gr
namespace gnuradio {
template <class T> boost::shared_ptr<T> get_initial_sptr(T* p) { return boost::shared_ptr<T>(p); }
}
