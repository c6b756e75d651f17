/* Public API of the MUSIC pseudo-spectrum block; signature frozen by gr-doa include/doa/MUSIC_lin_array.h:43-57. */
#ifndef INCLUDED_DOA_MUSIC_LIN_ARRAY_H
#define INCLUDED_DOA_MUSIC_LIN_ARRAY_H
#include <doa/api.h>
#include <gnuradio/sync_block.h>
namespace gr {
namespace doa {
/*! num_ant_ele x num_ant_ele covariance in, pspectrum_len-point MUSIC pseudo-spectrum (dB, peak = 0) out. */
class DOA_API MUSIC_lin_array : virtual public gr::sync_block {
 public:
  typedef boost::shared_ptr<MUSIC_lin_array> sptr;
  static sptr make(float norm_spacing, int num_targets, int num_ant_ele, int pspectrum_len);
};
}  // namespace doa
}  // namespace gr
#endif
