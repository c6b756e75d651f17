/* Public API of the peak picker; signature frozen by gr-doa include/doa/find_local_max.h:43-57. */
#ifndef INCLUDED_DOA_FIND_LOCAL_MAX_H
#define INCLUDED_DOA_FIND_LOCAL_MAX_H
#include <doa/api.h>
#include <gnuradio/sync_block.h>
namespace gr {
namespace doa {
/*! vector_len floats in; port 0: the num_max_vals highest local maxima, port 1: their x-axis locations (descending). */
class DOA_API find_local_max : virtual public gr::sync_block {
 public:
  typedef boost::shared_ptr<find_local_max> sptr;
  static sptr make(int num_max_vals, int vector_len, float x_min, float x_max);
};
}  // namespace doa
}  // namespace gr
#endif
