/* NOT in gr-doa: autocorrelate -> MUSIC_lin_array -> find_local_max as ONE block.
 * Inputs are autocorrelate's (include/doa/autocorrelate.h), outputs are find_local_max's (include/doa/find_local_max.h);
 * the covariance and the pseudo-spectrum never leave the GPU, so a frame costs one PCIe round trip instead of three and the
 * host ring buffers between the three blocks disappear. */
#ifndef INCLUDED_DOA_MUSIC_CHAIN_H
#define INCLUDED_DOA_MUSIC_CHAIN_H
#include <doa/api.h>
#include <gnuradio/block.h>
namespace gr {
namespace doa {
class DOA_API music_chain : virtual public gr::block {
 public:
  typedef boost::shared_ptr<music_chain> sptr;
  /*! The union of the three blocks' parameters, in their order:
   *  autocorrelate(inputs, snapshot_size, overlap_size, avg_method),
   *  MUSIC_lin_array(norm_spacing, num_targets, [num_ant_ele = inputs], pspectrum_len),
   *  find_local_max(num_max_vals, [vector_len = pspectrum_len], x_min, x_max). */
  static sptr make(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing, int num_targets,
                   int pspectrum_len, int num_max_vals, float x_min, float x_max);
  /*! The same block fed UHD cpu_format "sc16" items (std::complex<short>, 4 bytes: what the radio delivers before the host
   *  converts it to gr_complex for python/twinrx_usrp_source.py:57's "fc32"): half the bytes per sample into the GPU.
   *  A sample's value is int16 * sc16_scale (UHD's converter: 1/32767; a power of two such as 1/32768 makes the block
   *  bit-identical to make() fed the converted samples). */
  static sptr make_sc16(int inputs, int snapshot_size, int overlap_size, int avg_method, float norm_spacing, int num_targets,
                        int pspectrum_len, int num_max_vals, float x_min, float x_max, float sc16_scale);
  /*! Same as autocorrelate::set_antenna_config: fold the Antenna Correction block's config file into the covariance. */
  virtual void set_antenna_config(const char* config_filename) = 0;
};
}  // namespace doa
}  // namespace gr
#endif
