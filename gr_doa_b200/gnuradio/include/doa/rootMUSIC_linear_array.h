/* Public API of the Root-MUSIC block; signature frozen by gr-doa include/doa/rootMUSIC_linear_array.h:42-55. */
#ifndef INCLUDED_DOA_ROOTMUSIC_LINEAR_ARRAY_H
#define INCLUDED_DOA_ROOTMUSIC_LINEAR_ARRAY_H
#include <doa/api.h>
#include <gnuradio/sync_block.h>
namespace gr {
namespace doa {
/*! num_ant_ele x num_ant_ele covariance in, num_targets angles of arrival (degrees, ascending) out. */
class DOA_API rootMUSIC_linear_array : virtual public gr::sync_block {
 public:
  typedef boost::shared_ptr<rootMUSIC_linear_array> sptr;
  static sptr make(float norm_spacing, int num_targets, int num_ant_ele);
};
}  // namespace doa
}  // namespace gr
#endif
