/* Shared by the four *_impl.cc: handle ownership and the environment knobs the frozen constructor signatures cannot carry.
 *   DOA_CUDA_DEVICE      CUDA device index (default 0)
 *   DOA_CUDA_MAX_FRAMES  largest noutput_items processed per libdoa_cuda call (default 8192; larger calls are chunked) */
#ifndef INCLUDED_DOA_CUDA_BLOCK_COMMON_H
#define INCLUDED_DOA_CUDA_BLOCK_COMMON_H
#include <doa_cuda.h>
#include <cstdlib>
#include <stdexcept>
#include <string>
namespace gr {
namespace doa {
inline int doa_env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}
/* Constructor failures become exceptions (the reference's blocks throw std::invalid_argument the same way,
 * lib/antenna_correction_impl.cc:58-73); work() never throws: it logs and returns WORK_DONE (-1). */
inline void doa_require_created(int rc, const char* what) {
  if (rc != DOA_CUDA_OK) throw std::runtime_error(std::string(what) + ": " + doa_cuda_last_error(NULL));
}
}  // namespace doa
}  // namespace gr
#endif
