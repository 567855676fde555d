/* cuda_zstd_batch_c.h -- the C-ABI drop-in boundary of the B200-native batched Zstandard codec.
 *
 * Plain C: pointers, sizes and ints only; no C++ or torch types.  Every entry point names the
 * reference interface it stands in for (file:line under /root/reference).  The reference's own C
 * ABI is single-buffer only (include/cuda_zstd_manager.h:433-479, include/cuda_zstd_nvcomp.h:
 * 272-336); the cuda_zstd_batch_* calls are the additive batch form BASELINE.json's north_star asks
 * for, taking exactly the argument lists of NvcompV5BatchManager (include/cuda_zstd_nvcomp.h:
 * 93-134) so that a binding written against either maps 1:1.
 *
 * Conventions (same as the reference):
 *  - return value: 0 = success, else a cuda_zstd::Status value passed through the reference's lossy
 *    map {0,2,3,4,6,7,10,12 -> same; everything else -> 1} (src/cuda_zstd_nvcomp.cpp:75-96);
 *  - `*_sizes` arrays are in = capacity, out = bytes written;
 *  - data buffers and the workspace are DEVICE memory owned by the caller; pointer/size arrays may
 *    live in host OR device memory (src/cuda_zstd_nvcomp.cpp:319-437 accepts both);
 *  - calls enqueue on `stream` and, unless the name says `_nosync`, return after the results are
 *    visible to the host (the reference synchronises too: src/cuda_zstd_nvcomp.cpp:458,610);
 *  - no allocation happens inside a call; no CPU codec is ever used; a missing GPU is an error.
 */
#ifndef CUDA_ZSTD_BATCH_C_H
#define CUDA_ZSTD_BATCH_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st *cudaStream_t;
#endif

typedef struct cuda_zstd_batch cuda_zstd_batch_t;

/* NvcompV5BatchManager(const NvcompV5Options&) (include/cuda_zstd_nvcomp.h:95, options :41-51) /
 * create_batch_manager(level) (include/cuda_zstd_manager.h:361).  level 1..22; checksum != 0 writes
 * and verifies the XXH64 content checksum (ChecksumPolicy::COMPUTE_AND_VERIFY).  NULL on failure. */
cuda_zstd_batch_t *cuda_zstd_batch_create(int level, int enable_checksum);
void cuda_zstd_batch_destroy(cuda_zstd_batch_t *mgr);

/* NvcompV5BatchManager::get_max_compressed_chunk_size (src/cuda_zstd_nvcomp.cpp:289-298) ->
 * ZstdBatchManager::get_max_compressed_size -> estimate_compressed_size
 * (src/cuda_zstd_types.cpp:831-853): n + n/255 + 3*ceil(n/128K) + 512.  0 on a null handle. */
size_t cuda_zstd_batch_get_max_compressed_size(cuda_zstd_batch_t *mgr, size_t uncompressed_chunk_size);

/* NvcompV5BatchManager::get_compress_temp_size / get_decompress_temp_size
 * (src/cuda_zstd_nvcomp.cpp:207-287) -> ZstdBatchManager::get_batch_*_temp_size
 * (src/cuda_zstd_manager.cu:5661-5712).  `sizes` is a HOST array.  0 on a null handle.
 * decompress_temp <= compress_temp for the same batch, so one workspace serves both
 * (tests/test_c_api.cpp:62-64 in the reference relies on that). */
size_t cuda_zstd_batch_get_compress_temp_size(cuda_zstd_batch_t *mgr, const size_t *chunk_sizes, size_t num_chunks);
size_t cuda_zstd_batch_get_decompress_temp_size(cuda_zstd_batch_t *mgr, const size_t *compressed_sizes, size_t num_chunks);

/* NvcompV5BatchManager::compress_async (include/cuda_zstd_nvcomp.h:112-121,
 * src/cuda_zstd_nvcomp.cpp:300-484) -> ZstdBatchManager::compress_batch
 * (src/cuda_zstd_manager.cu:5715-5797).  Every output is one complete Zstandard frame that stock
 * libzstd decodes.  Empty batch -> 0; null array -> 2; workspace too small -> 7. */
int cuda_zstd_batch_compress(cuda_zstd_batch_t *mgr, const void *const *uncompressed_ptrs, const size_t *uncompressed_sizes,
                             size_t num_chunks, void *const *compressed_ptrs, size_t *compressed_sizes, void *d_temp,
                             size_t temp_bytes, cudaStream_t stream);

/* NvcompV5BatchManager::decompress_async (include/cuda_zstd_nvcomp.h:124-133,
 * src/cuda_zstd_nvcomp.cpp:486-644) -> ZstdBatchManager::decompress_batch
 * (src/cuda_zstd_manager.cu:5799-5859). */
int cuda_zstd_batch_decompress(cuda_zstd_batch_t *mgr, const void *const *compressed_ptrs, const size_t *compressed_sizes,
                               size_t num_chunks, void *const *uncompressed_ptrs, size_t *uncompressed_sizes, void *d_temp,
                               size_t temp_bytes, cudaStream_t stream);

/* Fully device-resident, no host synchronisation: all five arrays are DEVICE memory, d_statuses
 * (uint32 per chunk, cuda_zstd::Status values, may be NULL) is written by the kernel.  This is the
 * "true no-sync" form of ZstdBatchManager::decompress_async_no_sync
 * (include/cuda_zstd_manager.h:254-259; the reference's version still synchronises,
 * src/cuda_zstd_manager.cu:5979) extended to batches, and what bench.py times for `value`. */
int cuda_zstd_batch_compress_nosync(cuda_zstd_batch_t *mgr, const void *const *d_uncompressed_ptrs,
                                    const size_t *d_uncompressed_sizes, size_t num_chunks, void *const *d_compressed_ptrs,
                                    size_t *d_compressed_sizes, uint32_t *d_statuses, void *d_temp, size_t temp_bytes,
                                    cudaStream_t stream);
int cuda_zstd_batch_decompress_nosync(cuda_zstd_batch_t *mgr, const void *const *d_compressed_ptrs,
                                      const size_t *d_compressed_sizes, size_t num_chunks, void *const *d_uncompressed_ptrs,
                                      size_t *d_uncompressed_sizes, uint32_t *d_statuses, void *d_temp, size_t temp_bytes,
                                      cudaStream_t stream);

/* Host-resident batches: the payloads live in HOST memory (pinned for copy / kernel overlap; pageable works).  The
 * library stages them through d_temp in waves -- H2D of wave k+1 and D2H of wave k-1 run on two internal streams beside
 * the codec's kernels for wave k on `stream` -- and returns when the results are on the host.  This is the batch form of
 * what the reference does for host data with one synchronous single-buffer call per item: HybridEngine::compress_batch /
 * decompress_batch (src/cuda_zstd_hybrid.cu:926-951 -> :402-458) and NvcompV5BatchManager callers that copy around
 * compress_async (src/cuda_zstd_nvcomp.cpp:300-484).  Items that sit back to back in host memory move with one copy.
 *   decompress_host: frame i at h_compressed_ptrs[i] -> bytes at h_uncompressed_ptrs[i]; sizes in = capacity, out = bytes.
 *   compress_host_packed: chunk i at h_uncompressed_ptrs[i] -> frames packed back to back in h_packed,
 *     h_offsets[0..n] = exclusive scan of the frame sizes (device-side scan + gather, then two copies home).
 * h_statuses (n x uint32, may be NULL) receives the per-item status.  Return values as above; 7 = d_temp / h_packed too small. */
size_t cuda_zstd_batch_get_host_decompress_temp_size(cuda_zstd_batch_t *mgr, const size_t *compressed_sizes,
                                                     const size_t *uncompressed_capacities, size_t num_chunks);
size_t cuda_zstd_batch_get_host_compress_temp_size(cuda_zstd_batch_t *mgr, const size_t *chunk_sizes, size_t num_chunks);
int cuda_zstd_batch_decompress_host(cuda_zstd_batch_t *mgr, const void *const *h_compressed_ptrs, const size_t *compressed_sizes,
                                    size_t num_chunks, void *const *h_uncompressed_ptrs, size_t *uncompressed_sizes,
                                    uint32_t *h_statuses, void *d_temp, size_t temp_bytes, cudaStream_t stream);
int cuda_zstd_batch_compress_host_packed(cuda_zstd_batch_t *mgr, const void *const *h_uncompressed_ptrs, const size_t *chunk_sizes,
                                         size_t num_chunks, void *h_packed, size_t packed_capacity, uint64_t *h_offsets,
                                         uint32_t *h_statuses, void *d_temp, size_t temp_bytes, cudaStream_t stream);

/* Multi-GPU inside ONE process (the reference lists multi-GPU as future work, README.md:1648).
 * The batch is sharded by chunk index; shard g describes its own chunks with device-resident tables on its own GPU and is
 * driven by its own host thread (cudaSetDevice + the no-sync batch call on the shard's stream).  No payload crosses GPUs.
 * The one exchange: every shard's per-chunk output sizes are copied device-to-device (cudaMemcpyPeerAsync: NVLink when
 * peer access is on) into each shard's d_all_sizes, at the shard's base index, and scanned there on the device into
 * d_all_offsets -- the global packed offsets -- so every GPU ends up knowing where every frame goes.  Both are optional
 * (NULL: no exchange).  Returns when all shards are done: 0, or 1 if any chunk failed (see d_statuses), or the launch error. */
typedef struct cuda_zstd_shard {
  int device;                      /* CUDA device ordinal of this shard */
  cuda_zstd_batch_t *mgr;          /* a manager of this shard (one per shard) */
  const void *const *d_in_ptrs;    /* this shard's chunks: device tables on `device` */
  const size_t *d_in_sizes;
  void *const *d_out_ptrs;
  size_t *d_out_sizes;             /* in = capacity, out = bytes */
  uint32_t *d_statuses;            /* per chunk, required */
  size_t num_chunks;
  void *d_temp;
  size_t temp_bytes;
  cudaStream_t stream;             /* a stream of `device` */
  size_t *d_all_sizes;             /* optional: sum of all shards' num_chunks entries, on `device` */
  uint64_t *d_all_offsets; /* optional: that + 1 entries, on `device` */ } cuda_zstd_shard_t;
int cuda_zstd_batch_compress_sharded(cuda_zstd_shard_t *shards, int num_shards);
int cuda_zstd_batch_decompress_sharded(cuda_zstd_shard_t *shards, int num_shards);

/* Device-side exclusive scan of per-chunk compressed sizes -> packed offsets, plus the grand total
 * in d_offsets[num_chunks]; `base` is this GPU's starting offset in a multi-GPU job (SURVEY.md
 * section 8e).  Replaces the reference's thrust::exclusive_scan wrapper
 * (src/cuda_zstd_utils.cu:50-90).  d_offsets has num_chunks + 1 entries. */
int cuda_zstd_batch_scan_sizes(const size_t *d_sizes, size_t num_chunks, uint64_t base, uint64_t *d_offsets, cudaStream_t stream);

/* Gather frames written at `compressed_ptrs[i]` into one packed buffer at d_offsets[i]. */
int cuda_zstd_batch_pack(const void *const *d_compressed_ptrs, const size_t *d_sizes, const uint64_t *d_offsets,
                         size_t num_chunks, void *d_packed, cudaStream_t stream);

/* launch statistics of the most recent call on this manager: number of kernel launches it made */
int cuda_zstd_batch_last_launch_count(cuda_zstd_batch_t *mgr);

/* cuda_zstd_get_error_string (include/cuda_zstd_manager.h:476). */
const char *cuda_zstd_batch_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif /* CUDA_ZSTD_BATCH_C_H */
