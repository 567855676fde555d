// cuda_zstd_hybrid.h -- HybridEngine surface of the reference (include/cuda_zstd_hybrid.h:88-263, C API :287-343),
// GPU-only in this build.
//
// The reference's engine picks between host libzstd and its kernels per call (src/cuda_zstd_hybrid.cu:779-905).  This
// library has no host codec and links no libzstd, so the engine here is a convenience front end of
// ZstdBatchManager for callers that hold HOST buffers -- the reference's Python binding (python/src/binding.cpp:419-539)
// is the one on the hot path's next ring:
//   * every HybridMode, FORCE_CPU included, runs the CUDA path; the mode and thresholds are stored for get_config();
//   * query_routing() / HybridResult::backend_used answer GPU_KERNELS (single buffer) or GPU_BATCH, never a CPU backend;
//   * HOST buffers are staged through an engine-owned device arena (H2D -> kernels -> D2H on the caller's stream);
//     DEVICE / MANAGED buffers are used in place.
// Names, argument order and Status values follow the reference so that its callers compile unchanged.
#ifndef CUDA_ZSTD_HYBRID_H
#define CUDA_ZSTD_HYBRID_H

#include "cuda_zstd_types.h"
#include <cuda_runtime.h>

#ifdef __cplusplus
#include <memory>
#include <vector>

namespace cuda_zstd {

class HybridEngine {
public:
  HybridEngine();
  explicit HybridEngine(const HybridConfig &config);
  ~HybridEngine();
  HybridEngine(const HybridEngine &) = delete;
  HybridEngine &operator=(const HybridEngine &) = delete;
  HybridEngine(HybridEngine &&) noexcept;
  HybridEngine &operator=(HybridEngine &&) noexcept;

  Status configure(const HybridConfig &config);          // level outside 1..22 -> ERROR_INVALID_PARAMETER
  HybridConfig get_config() const;
  Status set_compression_level(int level);

  // *output_size: in = capacity, out = bytes written.  UNKNOWN locations are resolved with cudaPointerGetAttributes.
  // Null pointers or a zero input_size -> ERROR_INVALID_PARAMETER (src/cuda_zstd_hybrid.cu:783, :843).
  Status compress(const void *input, size_t input_size, void *output, size_t *output_size,
                  DataLocation input_loc = DataLocation::HOST, DataLocation output_loc = DataLocation::HOST,
                  HybridResult *result = nullptr, cudaStream_t stream = 0);
  Status decompress(const void *input, size_t input_size, void *output, size_t *output_size,
                    DataLocation input_loc = DataLocation::HOST, DataLocation output_loc = DataLocation::HOST,
                    HybridResult *result = nullptr, cudaStream_t stream = 0);

  // The reference loops over single-buffer calls (src/cuda_zstd_hybrid.cu:926-951); here the items go through ONE batch
  // launch.  Returns SUCCESS iff every item succeeded, else ERROR_COMPRESSION / ERROR_DECOMPRESSION like the reference,
  // with the per-item verdicts in results[i].status.
  Status compress_batch(const void *const *inputs, const size_t *input_sizes, void **outputs, size_t *output_sizes, size_t count,
                        DataLocation input_loc = DataLocation::HOST, DataLocation output_loc = DataLocation::HOST,
                        BatchRoutingResult *results = nullptr, cudaStream_t stream = 0);
  Status decompress_batch(const void *const *inputs, const size_t *input_sizes, void **outputs, size_t *output_sizes, size_t count,
                          DataLocation input_loc = DataLocation::HOST, DataLocation output_loc = DataLocation::HOST,
                          BatchRoutingResult *results = nullptr, cudaStream_t stream = 0);

  size_t get_max_compressed_size(size_t input_size) const;
  ExecutionBackend query_routing(size_t data_size, DataLocation input_loc, DataLocation output_loc, bool is_compression) const;
  CompressionStats get_stats() const;
  void reset_stats();
  static DataLocation detect_location(const void *ptr);
  // MB/s of the last profiled calls on `backend` (0 for the CPU backends, which never run here)
  double get_observed_throughput(ExecutionBackend backend, bool is_compression) const;
  void reset_profiling();

private:
  class Impl;
  std::unique_ptr<Impl> pimpl_;
};

Status hybrid_compress(const void *input, size_t input_size, void *output, size_t *output_size,
                       DataLocation input_loc = DataLocation::HOST, DataLocation output_loc = DataLocation::HOST,
                       int compression_level = 3, HybridResult *result = nullptr, cudaStream_t stream = 0);
Status hybrid_decompress(const void *input, size_t input_size, void *output, size_t *output_size,
                         DataLocation input_loc = DataLocation::HOST, DataLocation output_loc = DataLocation::HOST,
                         HybridResult *result = nullptr, cudaStream_t stream = 0);
std::unique_ptr<HybridEngine> create_hybrid_engine(const HybridConfig &config = HybridConfig{});
std::unique_ptr<HybridEngine> create_hybrid_engine(int compression_level);

} // namespace cuda_zstd
#endif // __cplusplus

// ---- C API (reference include/cuda_zstd_hybrid.h:287-343); return values are Status codes as int ----
typedef struct cuda_zstd_hybrid_engine_t cuda_zstd_hybrid_engine_t;
typedef struct {
  unsigned int mode;
  size_t cpu_size_threshold;
  size_t gpu_device_threshold;
  int compression_level;
  int enable_profiling;
  unsigned int cpu_thread_count;
} cuda_zstd_hybrid_config_t;
typedef struct {
  unsigned int backend_used;
  unsigned int input_location;
  unsigned int output_location;
  double total_time_ms;
  double transfer_time_ms;
  double compute_time_ms;
  double throughput_mbps;
  size_t input_bytes;
  size_t output_bytes;
  float compression_ratio;
} cuda_zstd_hybrid_result_t;

#ifdef __cplusplus
extern "C" {
#endif
cuda_zstd_hybrid_engine_t *cuda_zstd_hybrid_create(const cuda_zstd_hybrid_config_t *config);
cuda_zstd_hybrid_engine_t *cuda_zstd_hybrid_create_default(void);
void cuda_zstd_hybrid_destroy(cuda_zstd_hybrid_engine_t *engine);
int cuda_zstd_hybrid_compress(cuda_zstd_hybrid_engine_t *engine, const void *input, size_t input_size, void *output,
                              size_t *output_size, unsigned int input_loc, unsigned int output_loc,
                              cuda_zstd_hybrid_result_t *result, cudaStream_t stream);
int cuda_zstd_hybrid_decompress(cuda_zstd_hybrid_engine_t *engine, const void *input, size_t input_size, void *output,
                                size_t *output_size, unsigned int input_loc, unsigned int output_loc,
                                cuda_zstd_hybrid_result_t *result, cudaStream_t stream);
size_t cuda_zstd_hybrid_max_compressed_size(cuda_zstd_hybrid_engine_t *engine, size_t input_size);
unsigned int cuda_zstd_hybrid_query_routing(cuda_zstd_hybrid_engine_t *engine, size_t data_size, unsigned int input_loc,
                                            unsigned int output_loc, int is_compression);
#ifdef __cplusplus
}
#endif
#endif // CUDA_ZSTD_HYBRID_H
