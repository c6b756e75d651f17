/* Shared by the four *_impl.cc: handle ownership and the environment knobs the frozen constructor signatures cannot carry.
 *   DOA_CUDA_DEVICE      CUDA device index (default 0)
 *   DOA_CUDA_MAX_FRAMES  largest noutput_items processed per libdoa_cuda call (default 256; larger calls are chunked).  Device
 *                        buffers are sized for it at construction: at 4 x 2048 samples per frame 256 frames are 16 MB per
 *                        block, and a scheduler call carries a few dozen frames at most (64 KB items in its default buffers)
 *   DOA_CUDA_MIN_FRAMES  scheduler hint for the fused chain blocks (default 8): set_output_multiple() keeps the scheduler from
 *                        calling work() for fewer frames, set_min_output_buffer() sizes the output buffers for 4x as many
 *   DOA_CUDA_DEVICES     comma-separated device list for doa.music_chain: the frames of every work() call are spread over
 *                        these GPUs (doa_cuda_multi_*); unset = the single DOA_CUDA_DEVICE */
#ifndef INCLUDED_DOA_CUDA_BLOCK_COMMON_H
#define INCLUDED_DOA_CUDA_BLOCK_COMMON_H
#include <doa_cuda.h>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#define DOA_CUDA_DEFAULT_MAX_FRAMES 256
#define DOA_CUDA_DEFAULT_MIN_FRAMES 8
namespace gr {
namespace doa {
inline int doa_env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}
/* DOA_CUDA_DEVICES="0,1,2,3": the devices a multi-GPU capable block (doa.music_chain) spreads its frames over; unset or a
 * single entry: one device (DOA_CUDA_DEVICE, default 0). */
inline std::vector<int> doa_env_devices() {
  std::vector<int> out;
  const char* v = std::getenv("DOA_CUDA_DEVICES");
  if (v) {
    for (const char* p = v; *p;) {
      char* end = NULL;
      const long d = std::strtol(p, &end, 10);
      if (end == p) break;
      out.push_back((int)d);
      p = (*end == ',') ? end + 1 : end;
    }
  }
  if (out.empty()) out.push_back(doa_env_int("DOA_CUDA_DEVICE", 0));
  return out;
}
/* Constructor failures become exceptions (the reference's blocks throw std::invalid_argument the same way,
 * lib/antenna_correction_impl.cc:58-73); work() never throws: it logs and returns WORK_DONE (-1). */
inline void doa_require_created(int rc, const char* what) {
  if (rc != DOA_CUDA_OK) throw std::runtime_error(std::string(what) + ": " + doa_cuda_last_error(NULL));
}
}  // namespace doa
}  // namespace gr
#endif
