/* NOT in gr-doa: autocorrelate -> rootMUSIC_linear_array as ONE block (the Root-MUSIC flowgraphs, apps/run_RootMUSIC_*).
 * Inputs are autocorrelate's (include/doa/autocorrelate.h), the output is rootMUSIC_linear_array's port 0
 * (include/doa/rootMUSIC_linear_array.h: num_targets ascending angles in degrees per item); the covariance never leaves the
 * GPU, so an item costs one PCIe round trip of num_targets floats instead of two with an M*M matrix in between. */
#ifndef INCLUDED_DOA_ROOTMUSIC_CHAIN_H
#define INCLUDED_DOA_ROOTMUSIC_CHAIN_H
#include <doa/api.h>
#include <gnuradio/block.h>
namespace gr {
namespace doa {
class DOA_API rootmusic_chain : virtual public gr::block {
 public:
  typedef boost::shared_ptr<rootmusic_chain> sptr;
  /*! autocorrelate(inputs, snapshot_size, overlap_size, avg_method) followed by
   *  rootMUSIC_linear_array(norm_spacing, num_targets, [num_ant_ele = inputs]). */
  static sptr make(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing, int num_targets);
  /*! Same as autocorrelate::set_antenna_config: fold the Antenna Correction block's config file into the covariance. */
  virtual void set_antenna_config(const char* config_filename) = 0;
};
}  // namespace doa
}  // namespace gr
#endif
