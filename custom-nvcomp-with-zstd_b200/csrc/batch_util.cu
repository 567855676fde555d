// batch_util.cu -- device-side exclusive scan of per-chunk sizes and frame packing.
// Replaces the reference's thrust::exclusive_scan wrapper (src/cuda_zstd_utils.cu:50-90) for the one
// place the batch path needs it: turning per-chunk compressed sizes into packed offsets
// ("device-side offset gather").
#include "zstd_device_api.h"
#include "zstd_common.cuh"
#include <cuda_runtime.h>

namespace b200zstd {

constexpr int SCAN_THREADS = 1024;

// One CTA walks the array in tiles of 1024 with a carried prefix: warp shuffles for the intra-warp
// scan, one SMEM pass across the 32 warp totals.  N <= a few hundred thousand sizes, so a single CTA
// (one pass over <= 1 MiB) is launch-latency-, not bandwidth-, bound.
__global__ void __launch_bounds__(SCAN_THREADS) scan_sizes_kernel(const size_t *__restrict__ sizes, size_t n, uint64_t base,
                                                                  uint64_t *offsets, const uint64_t *base_ptr) {
  __shared__ uint64_t warp_tot[32];
  __shared__ uint64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = base_ptr ? *base_ptr : base;        // (base_ptr may be offsets[0]: read before anything is written)
  __syncthreads();
  for (size_t t0 = 0; t0 < n; t0 += SCAN_THREADS) {
    const size_t i = t0 + tid;
    const uint64_t v = (i < n) ? (uint64_t)sizes[i] : 0;
    uint64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint64_t w = warp_tot[lane], s = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_tot[lane] = s - w;                 // exclusive prefix of warp totals
      if (lane == 31) warp_tot[31] = s - w;   // (kept for clarity)
    }
    __syncthreads();
    const uint64_t carry = carry_s;
    const uint64_t excl = carry + warp_tot[warp] + (x - v);
    if (i < n) offsets[i] = excl;
    __syncthreads();
    if (tid == SCAN_THREADS - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) offsets[n] = carry_s;
}

// One CTA per frame; 16-byte vector copies when source and destination agree on alignment.
__global__ void __launch_bounds__(256) pack_kernel(const void *const *__restrict__ ptrs, const size_t *__restrict__ sizes,
                                                   const uint64_t *__restrict__ offsets, size_t n, uint8_t *__restrict__ packed) {
  for (size_t c = blockIdx.x; c < n; c += gridDim.x) {
    const uint8_t *src = (const uint8_t *)ptrs[c];
    uint8_t *dst = packed + offsets[c];
    size_t len = sizes[c];
    size_t head = 0;
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) == 0) {
      head = (16 - ((uintptr_t)dst & 15)) & 15;
      if (head > len) head = len;
      const size_t body = (len - head) >> 4;
      const uint4 *s4 = (const uint4 *)(src + head);
      uint4 *d4 = (uint4 *)(dst + head);
      for (size_t k = threadIdx.x; k < body; k += blockDim.x) d4[k] = s4[k];
      for (size_t k = threadIdx.x; k < head; k += blockDim.x) dst[k] = src[k];
      for (size_t k = head + (body << 4) + threadIdx.x; k < len; k += blockDim.x) dst[k] = src[k];
    } else {
      for (size_t k = threadIdx.x; k < len; k += blockDim.x) dst[k] = src[k];
    }
  }
}

// Content checksum of a frame assembled from independently encoded blocks: XXH64 is one sequential chain of 32-byte
// stripes, so one warp walks the whole buffer (16-byte loads, four accumulator lanes) and drops the low 32 bits behind
// the last block.  `where` = device word holding the offset at which the 4 bytes go.
__global__ void __launch_bounds__(32) frame_checksum_kernel(const uint8_t *__restrict__ src, size_t n, uint8_t *__restrict__ dst,
                                                            const uint64_t *__restrict__ where) {
  const uint64_t h = xxh64_warp(src, (uint32_t)n, threadIdx.x);
  if (threadIdx.x == 0) {
    uint8_t *p = dst + *where;
    p[0] = (uint8_t)h; p[1] = (uint8_t)(h >> 8); p[2] = (uint8_t)(h >> 16); p[3] = (uint8_t)(h >> 24);
  }
}
cudaError_t launch_frame_checksum(const void *d_src, size_t n, void *d_dst, const uint64_t *d_where, cudaStream_t stream) {
  frame_checksum_kernel<<<1, 32, 0, stream>>>((const uint8_t *)d_src, n, (uint8_t *)d_dst, d_where);
  return cudaGetLastError();
}

// Outcome of one block-parallel frame (compress_big) folded into 16 bytes so that a caller who does not want to
// synchronise can fetch it with a single small copy: {frame bytes, first non-zero block status, 0}.
__global__ void __launch_bounds__(256) big_result_kernel(const uint32_t *__restrict__ statuses, size_t n, const uint64_t *__restrict__ total,
                                                         uint64_t add, uint64_t *__restrict__ result) {
  __shared__ unsigned int first_bad;
  if (threadIdx.x == 0) first_bad = 0xFFFFFFFFu;
  __syncthreads();
  for (size_t i = threadIdx.x; i < n; i += blockDim.x)
    if (statuses[i] != 0) atomicMin(&first_bad, (unsigned int)i);
  __syncthreads();
  if (threadIdx.x == 0) {
    result[0] = *total + add;
    result[1] = first_bad == 0xFFFFFFFFu ? 0 : (uint64_t)statuses[first_bad];
  }
}
cudaError_t launch_big_result(const uint32_t *d_statuses, size_t n, const uint64_t *d_total, uint64_t add, uint64_t *d_result,
                              cudaStream_t stream) {
  big_result_kernel<<<1, 256, 0, stream>>>(d_statuses, n, d_total, add, d_result);
  return cudaGetLastError();
}

cudaError_t launch_scan_sizes(const size_t *d_sizes, size_t n, uint64_t base, uint64_t *d_offsets, cudaStream_t stream) {
  scan_sizes_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_sizes, n, base, d_offsets, nullptr);
  return cudaGetLastError();
}
// the same with the base taken from device memory: chains the scans of consecutive ranges (d_base = the previous range's
// closing offset, which is this range's d_offsets[0])
cudaError_t launch_scan_sizes_from(const size_t *d_sizes, size_t n, const uint64_t *d_base, uint64_t *d_offsets, cudaStream_t stream) {
  scan_sizes_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_sizes, n, 0, d_offsets, d_base);
  return cudaGetLastError();
}
cudaError_t launch_pack(const void *const *d_ptrs, const size_t *d_sizes, const uint64_t *d_offsets, size_t n, void *d_packed,
                        cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  int grid = (int)(n < 148 * 8 ? n : 148 * 8);
  pack_kernel<<<grid, 256, 0, stream>>>(d_ptrs, d_sizes, d_offsets, n, (uint8_t *)d_packed);
  return cudaGetLastError();
}

} // namespace b200zstd
