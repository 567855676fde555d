// pipeline_manager.hpp -- host-staged streaming compression over the batch codec.
//
// Drop-in for the reference's src/pipeline_manager.hpp:12-66 (class PipelinedBatchManager, struct
// RingBufferSlot): same public calls and the same data members in the same order, because the
// reference's callers (tests/test_pipeline_integration.cu:51, benchmarks/benchmark_pipeline.cu:94)
// include the reference header by relative path and construct the object on their own stack -- the
// layout is therefore part of the boundary here, unlike the pimpl managers.
//
// What differs is behind it.  The reference calls the blocking single-buffer compress() per batch
// from the host loop (src/pipeline_manager.cu), so its three streams never hold more than one
// stage.  This build enqueues H2D -> compress_async_no_sync -> D2H of the 16-byte result on three
// streams chained by events (every slot compresses on its own stream, so small batches, which fill
// only a fraction of the GPU, overlap each other as well), fills the next pinned slot while the
// GPU works, and waits for a batch's size only when it is there or when the ring would stall,
// to issue a D2H of exactly the bytes produced.  Every batch becomes
// one Zstandard frame (blocks of 128 KB encoded side by side), so the output is a concatenation of
// frames that `zstd -d` / ZSTD_decompress-in-a-loop reads back.
#ifndef CUDA_ZSTD_PIPELINE_MANAGER_HPP
#define CUDA_ZSTD_PIPELINE_MANAGER_HPP

#include "cuda_zstd_manager.h"

#ifdef __cplusplus
#include <functional>
#include <memory>
#include <vector>

namespace cuda_zstd {

// One stage buffer set of the ring (reference src/pipeline_manager.hpp:12-33; field order is ABI).
struct RingBufferSlot {
  void *d_input = nullptr;          // device: batch as uploaded
  void *d_output = nullptr;         // device: frame as produced
  void *d_workspace = nullptr;      // device: codec scratch of this slot
  void *h_input = nullptr;          // pinned host: filled by the input callback
  void *h_output = nullptr;         // pinned host: handed to the output callback
  size_t input_capacity = 0;
  size_t output_capacity = 0;
  size_t workspace_capacity = 0;
  size_t current_input_size = 0;
  size_t current_output_size = 0;
  cudaEvent_t event_uploaded = nullptr;     // H2D of the batch finished
  cudaEvent_t event_compressed = nullptr;   // frame and its 16-byte result are ready
  cudaEvent_t event_downloaded = nullptr;   // D2H finished: the slot may be refilled
};

class PipelinedBatchManager {
public:
  // batch_size_bytes: bytes asked from the input callback per batch (one frame each); num_slots >= 2.
  explicit PipelinedBatchManager(const CompressionConfig &config, size_t batch_size_bytes = 64 * 1024 * 1024, int num_slots = 3);
  ~PipelinedBatchManager();
  PipelinedBatchManager(const PipelinedBatchManager &) = delete;
  PipelinedBatchManager &operator=(const PipelinedBatchManager &) = delete;

  // input_callback(buffer, max_len, &len): write up to max_len bytes, set len; return false when
  // this was the last batch (a batch with len == 0 is skipped).  output_callback(frame, size) is
  // called once per batch, in input order, from the calling thread.
  Status compress_stream_pipeline(std::function<bool(void *h_input, size_t max_len, size_t *out_len)> input_callback,
                                  std::function<void(const void *h_output, size_t size)> output_callback);

private:
  std::unique_ptr<ZstdManager> manager_;
  CompressionConfig config_;
  size_t batch_size_;
  int num_slots_;
  std::vector<RingBufferSlot> ring_buffer_;
  std::vector<cudaStream_t> streams_;       // [0] upload, [1] compress (slot 0), [2] download, [3..] compress of slots 1..

  Status init_resources();
  void cleanup_resources();
};

} // namespace cuda_zstd

extern "C" {
#endif

// ---- C ABI over the class (additive: the reference exposes the pipeline to C++ only) ----
typedef struct cuda_zstd_pipeline cuda_zstd_pipeline_t;
// return 0 when this was the last batch, non-zero when more follows; *out_len = bytes written (<= max_len)
typedef int (*cuda_zstd_pipeline_input_fn)(void *user, void *h_input, size_t max_len, size_t *out_len);
typedef void (*cuda_zstd_pipeline_output_fn)(void *user, const void *h_output, size_t size);
cuda_zstd_pipeline_t *cuda_zstd_pipeline_create(int level, int enable_checksum, size_t batch_size_bytes, int num_slots);
void cuda_zstd_pipeline_destroy(cuda_zstd_pipeline_t *p);
int cuda_zstd_pipeline_compress(cuda_zstd_pipeline_t *p, cuda_zstd_pipeline_input_fn in_fn, cuda_zstd_pipeline_output_fn out_fn, void *user);

#ifdef __cplusplus
}
#endif
#endif
