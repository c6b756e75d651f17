/* Public API of the array-calibration block; signature frozen by gr-doa include/doa/calibrate_lin_array.h:42-55. */
#ifndef INCLUDED_DOA_CALIBRATE_LIN_ARRAY_H
#define INCLUDED_DOA_CALIBRATE_LIN_ARRAY_H
#include <doa/api.h>
#include <gnuradio/sync_block.h>
namespace gr {
namespace doa {
/*! Antenna gain/phase estimates (num_ant_ele complex values per item) from the covariance of a pilot at a known angle.
 *  GPU-backed: work is done by libdoa_cuda; there is no CPU path.  Like the reference's eigenvector output the estimate
 *  is defined up to a unit-modulus factor. */
class DOA_API calibrate_lin_array : virtual public gr::sync_block {
 public:
  typedef boost::shared_ptr<calibrate_lin_array> sptr;
  /*! \param norm_spacing element spacing / wavelength  \param num_ant_ele number of antenna elements
   *  \param pilot_angle known angle of the pilot transmitter in degrees */
  static sptr make(float norm_spacing, int num_ant_ele, float pilot_angle);
};
}  // namespace doa
}  // namespace gr
#endif
