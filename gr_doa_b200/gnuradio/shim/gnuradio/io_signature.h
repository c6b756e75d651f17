/* Minimal gr::io_signature: port-count range and per-port item sizes (make / make2 as the four blocks use them). */
#pragma once
#include <boost/shared_ptr.hpp>
#include <vector>
namespace gr {
class io_signature {
 public:
  typedef boost::shared_ptr<io_signature> sptr;
  static sptr make(int min_streams, int max_streams, int sizeof_stream_item) {
    return sptr(new io_signature(min_streams, max_streams, std::vector<int>(1, sizeof_stream_item)));
  }
  static sptr make2(int min_streams, int max_streams, int s1, int s2) {
    std::vector<int> v; v.push_back(s1); v.push_back(s2);
    return sptr(new io_signature(min_streams, max_streams, v));
  }
  int min_streams() const { return d_min; }
  int max_streams() const { return d_max; }
  int sizeof_stream_item(int i) const { return d_sizes[i < (int)d_sizes.size() ? i : d_sizes.size() - 1]; }
 private:
  io_signature(int mn, int mx, const std::vector<int>& s) : d_min(mn), d_max(mx), d_sizes(s) {}
  int d_min, d_max; std::vector<int> d_sizes;
};
}  // namespace gr
