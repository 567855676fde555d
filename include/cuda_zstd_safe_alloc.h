// cuda_zstd_safe_alloc.h -- headroom-checked allocation helpers.
//
// Drop-in for the reference's include/cuda_zstd_safe_alloc.h:50-306, which its batch-path tests and benchmarks
// include (tests/test_nvcomp_batch.cu:6, tests/test_c_api.cpp:6, tests/cuda_error_checking.h:12).  Same names and
// contracts: an allocation is refused (no allocator call) when it would leave less than VRAM_SAFETY_BUFFER_BYTES of
// free device memory / RAM_SAFETY_BUFFER_BYTES of free host memory; null out-pointer -> cudaErrorInvalidValue / false;
// size 0 succeeds with *ptr = nullptr.  Header-only; nothing in the library itself uses it (managers never allocate).
#ifndef CUDA_ZSTD_SAFE_ALLOC_H
#define CUDA_ZSTD_SAFE_ALLOC_H

#include <cuda_runtime.h>
#include <cstddef>
#include <cstdlib>
#ifdef __linux__
#include <sys/sysinfo.h>
#endif

#include "cuda_zstd_types.h"

namespace cuda_zstd {
namespace safe_alloc_detail {

// free bytes beyond the reserve, 0 when the query fails or the reserve is already eaten
inline size_t device_headroom() {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
  return free_b > VRAM_SAFETY_BUFFER_BYTES ? free_b - VRAM_SAFETY_BUFFER_BYTES : 0;
}
// host side: `known` is false where the platform gives no figure (then nothing is refused)
inline size_t host_headroom(bool *known) {
  *known = false;
#ifdef __linux__
  struct sysinfo si;
  if (sysinfo(&si) == 0) {
    *known = true;
    const size_t free_b = (size_t)si.freeram * (size_t)si.mem_unit;
    return free_b > RAM_SAFETY_BUFFER_BYTES ? free_b - RAM_SAFETY_BUFFER_BYTES : 0;
  }
#endif
  return 0;
}
template <class Alloc> inline cudaError_t device_alloc(void **ptr, size_t size, Alloc alloc) {
  if (!ptr) return cudaErrorInvalidValue;
  *ptr = nullptr;
  if (size == 0) return cudaSuccess;
  size_t free_b = 0, total_b = 0;
  const cudaError_t q = cudaMemGetInfo(&free_b, &total_b);
  if (q != cudaSuccess) return q;
  if (free_b < size + VRAM_SAFETY_BUFFER_BYTES) return cudaErrorMemoryAllocation;
  return alloc(ptr, size);
}
inline bool host_fits(size_t size) {
  bool known = false;
  const size_t room = host_headroom(&known);
  return !known || size <= room;
}

} // namespace safe_alloc_detail

inline cudaError_t safe_cuda_malloc(void **ptr, size_t size) {
  return safe_alloc_detail::device_alloc(ptr, size, [](void **p, size_t n) { return cudaMalloc(p, n); });
}
template <typename T> inline cudaError_t safe_cuda_malloc(T **ptr, size_t size) { return safe_cuda_malloc(reinterpret_cast<void **>(ptr), size); }

inline cudaError_t safe_cuda_malloc_async(void **ptr, size_t size, cudaStream_t stream) {
  return safe_alloc_detail::device_alloc(ptr, size, [stream](void **p, size_t n) { return cudaMallocAsync(p, n, stream); });
}
template <typename T> inline cudaError_t safe_cuda_malloc_async(T **ptr, size_t size, cudaStream_t stream) {
  return safe_cuda_malloc_async(reinterpret_cast<void **>(ptr), size, stream);
}

inline cudaError_t safe_cuda_malloc_host(void **ptr, size_t size) {
  if (!ptr) return cudaErrorInvalidValue;
  *ptr = nullptr;
  if (size == 0) return cudaSuccess;
  if (!safe_alloc_detail::host_fits(size)) return cudaErrorMemoryAllocation;
  return cudaMallocHost(ptr, size);
}
template <typename T> inline cudaError_t safe_cuda_malloc_host(T **ptr, size_t size) { return safe_cuda_malloc_host(reinterpret_cast<void **>(ptr), size); }

inline bool safe_host_malloc(void **ptr, size_t size) {
  if (!ptr) return false;
  *ptr = nullptr;
  if (size == 0) return true;
  if (!safe_alloc_detail::host_fits(size)) return false;
  *ptr = std::malloc(size);
  return *ptr != nullptr;
}
template <typename T> inline bool safe_host_malloc(T **ptr, size_t size) { return safe_host_malloc(reinterpret_cast<void **>(ptr), size); }

inline size_t get_usable_vram() { return safe_alloc_detail::device_headroom(); }
inline size_t get_usable_host_ram() {
  bool known = false;
  const size_t room = safe_alloc_detail::host_headroom(&known);
  return known ? room : 0;
}

} // namespace cuda_zstd
#endif // CUDA_ZSTD_SAFE_ALLOC_H
