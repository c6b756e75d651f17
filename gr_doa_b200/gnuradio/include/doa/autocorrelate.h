/* Public API of the sample-covariance block; signature frozen by gr-doa include/doa/autocorrelate.h:43-57. */
#ifndef INCLUDED_DOA_AUTOCORRELATE_H
#define INCLUDED_DOA_AUTOCORRELATE_H
#include <doa/api.h>
#include <gnuradio/block.h>
namespace gr {
namespace doa {
/*! Sample covariance of `inputs` complex streams, one inputs x inputs matrix (column-major) per snapshot.
 *  GPU-backed: work is done by libdoa_cuda; there is no CPU path. */
class DOA_API autocorrelate : virtual public gr::block {
 public:
  typedef boost::shared_ptr<autocorrelate> sptr;
  /*! \param inputs number of streams  \param snapshot_size samples per snapshot
   *  \param overlap_size samples shared by consecutive snapshots  \param avg_method 0 forward, 1 forward-backward */
  static sptr make(int inputs, int snapshot_size, int overlap_size, int avg_method);
  /*! Optional, not in gr-doa: fold the antenna_correction block's per-channel gains (its config file, one "gain phase"
   *  pair per line, gr-doa lib/antenna_correction_impl.cc:54-74) into the covariance as R' = D R D^H, so the flowgraph can
   *  drop that block and its pass over the samples.  Throws std::invalid_argument like that block's constructor.
   *  An empty name removes the gains. */
  virtual void set_antenna_config(const char* config_filename) = 0;
};
}  // namespace doa
}  // namespace gr
#endif
