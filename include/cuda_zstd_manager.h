// cuda_zstd_manager.h -- manager classes of the batch-codec boundary (B200-native build).
//
// Same names, signatures and error behaviour as the public part of the reference header of the
// same name (ZstdManager include/cuda_zstd_manager.h:45-100, ZstdBatchManager :113-278, factories
// :358-363, single-shot helpers :369-386, utilities :392-420, C API :433-479).  Both classes are
// pimpl in the reference too, so object layout is private to each build.  ZstdStreamingManager and
// the dictionary trainer are out of scope.
#ifndef CUDA_ZSTD_MANAGER_H
#define CUDA_ZSTD_MANAGER_H

#include "cuda_zstd_types.h"

#ifdef __cplusplus
#include <memory>
#include <vector>

namespace cuda_zstd {

class ZstdManager {
public:
  virtual ~ZstdManager() = default;
  virtual Status configure(const CompressionConfig &config) = 0;
  virtual CompressionConfig get_config() const = 0;
  virtual size_t get_compress_temp_size(size_t uncompressed_size) const = 0;
  virtual size_t get_decompress_temp_size(size_t compressed_size) const = 0;
  virtual size_t get_max_compressed_size(size_t uncompressed_size) const = 0;
  // *compressed_size / *uncompressed_size: in = capacity, out = bytes written.  Buffers and the
  // workspace are device memory.  Returns once the result is visible to the host.
  virtual Status compress(const void *uncompressed_data, size_t uncompressed_size, void *compressed_data,
                          size_t *compressed_size, void *temp_workspace, size_t temp_size, const void *dict_buffer,
                          size_t dict_size, cudaStream_t stream, void *streaming_context = nullptr) = 0;
  virtual Status decompress(const void *compressed_data, size_t compressed_size, void *uncompressed_data,
                            size_t *uncompressed_size, void *temp_workspace, size_t temp_size,
                            cudaStream_t stream = 0) = 0;
  virtual Status set_dictionary(const dictionary::Dictionary &dict) = 0;
  virtual Status get_dictionary(dictionary::Dictionary &dict) const = 0;
  virtual Status clear_dictionary() = 0;
  virtual const CompressionStats &get_stats() const = 0;
  virtual Status set_compression_level(int level) = 0;
  virtual int get_compression_level() const = 0;
  virtual void reset_stats() = 0;

  enum class ExecutionPath { CPU, GPU_BATCH, GPU_CHUNK };
  // Kept for source compatibility.  This build never takes the CPU path, whatever this returns.
  static ExecutionPath select_execution_path(size_t size, int cpu_threshold = 1024 * 1024);
  virtual Status preallocate_tables(cudaStream_t = 0) { return Status::SUCCESS; }
  virtual Status free_tables(cudaStream_t = 0) { return Status::SUCCESS; }
};

class ZstdBatchManager : public ZstdManager {
public:
  ZstdBatchManager();
  explicit ZstdBatchManager(const CompressionConfig &config);
  ~ZstdBatchManager() override;

  Status configure(const CompressionConfig &config) override;
  CompressionConfig get_config() const override;
  size_t get_compress_temp_size(size_t uncompressed_size) const override;
  size_t get_decompress_temp_size(size_t compressed_size) const override;
  size_t get_max_compressed_size(size_t uncompressed_size) const override;
  Status compress(const void *uncompressed_data, size_t uncompressed_size, void *compressed_data, size_t *compressed_size,
                  void *temp_workspace, size_t temp_size, const void *dict_buffer, size_t dict_size, cudaStream_t stream = 0,
                  void *streaming_context = nullptr) override;
  Status decompress(const void *compressed_data, size_t compressed_size, void *uncompressed_data, size_t *uncompressed_size,
                    void *temp_workspace, size_t temp_size, cudaStream_t stream = 0) override;
  Status set_dictionary(const dictionary::Dictionary &dict) override;
  Status get_dictionary(dictionary::Dictionary &dict) const override;
  Status clear_dictionary() override;
  const CompressionStats &get_stats() const override;
  Status set_compression_level(int level) override;
  int get_compression_level() const override;
  void reset_stats() override;

  // Batch calls: one kernel pipeline over all items on `stream`; per-item Status is written into
  // items[i].status and sizes into items[i].output_size (the reference does the same through a
  // const_cast, src/cuda_zstd_manager.cu:5770-5795).  SUCCESS iff every item succeeded.
  Status compress_batch(const std::vector<BatchItem> &items, void *temp_workspace, size_t temp_size, cudaStream_t stream = 0);
  Status decompress_batch(const std::vector<BatchItem> &items, void *temp_workspace, size_t temp_size, cudaStream_t stream = 0);
  size_t get_batch_compress_temp_size(const std::vector<size_t> &uncompressed_sizes) const;
  size_t get_batch_decompress_temp_size(const std::vector<size_t> &compressed_sizes) const;

  // Inference API (reference include/cuda_zstd_manager.h:196-270).
  Status decompress_to_preallocated(const void *compressed_data, size_t compressed_size, void *preallocated_output,
                                    size_t output_capacity, size_t *actual_output_size, void *temp_workspace, size_t temp_size,
                                    cudaStream_t stream = 0);
  Status decompress_batch_preallocated(std::vector<BatchItem> &items, void *temp_workspace, size_t temp_size,
                                       cudaStream_t stream = 0);
  // d_actual_size is a DEVICE pointer; nothing is synchronised (unlike the reference, which still
  // syncs at src/cuda_zstd_manager.cu:5979).  The caller synchronises `stream`.
  Status decompress_async_no_sync(const void *compressed_data, size_t compressed_size, void *preallocated_output,
                                  size_t output_capacity, size_t *d_actual_size, void *temp_workspace, size_t temp_size,
                                  cudaStream_t stream);
  // Additive (no reference counterpart): compress without any host synchronisation.  Device buffers only; result16 gets
  // {bytes written, Status} as two 64-bit words in stream order (pinned host or device memory).  Used by
  // PipelinedBatchManager to keep H2D, compress and D2H of neighbouring batches in flight together.
  Status compress_async_no_sync(const void *uncompressed_data, size_t uncompressed_size, void *compressed_data,
                                size_t compressed_capacity, unsigned long long *result16, void *temp_workspace, size_t temp_size,
                                cudaStream_t stream);
  size_t get_inference_workspace_size(size_t max_compressed_size, size_t max_output_size) const;
  Status allocate_inference_workspace(size_t max_compressed_size, size_t max_output_size, void **workspace_ptr,
                                      size_t *workspace_size);
  Status free_inference_workspace(void *workspace_ptr);

  class Impl;
  Impl *impl() { return pimpl_.get(); }   // used by the C ABI in this build only

private:
  std::unique_ptr<Impl> pimpl_;
};

std::unique_ptr<ZstdManager> create_manager(int compression_level = 3);
std::unique_ptr<ZstdManager> create_manager(const CompressionConfig &config);
std::unique_ptr<ZstdBatchManager> create_batch_manager(int compression_level = 3);

Status compress_simple(const void *uncompressed_data, size_t uncompressed_size, void *compressed_data,
                       size_t *compressed_size, int compression_level = 3, cudaStream_t stream = 0);
Status decompress_simple(const void *compressed_data, size_t compressed_size, void *uncompressed_data,
                         size_t *uncompressed_size, cudaStream_t stream = 0);

// Frame-header helpers; the pointer may be host or device memory (reference src/cuda_zstd_types.cpp:1058-1170).
// get_decompressed_size: a frame that declares no content size answers SUCCESS with *decompressed_size = 0.
Status get_decompressed_size(const void *compressed_data, size_t compressed_size, size_t *decompressed_size);
// magic + header check only (the reference does not decode either); check_checksum is accepted and unused
Status validate_compressed_data(const void *compressed_data, size_t compressed_size, bool check_checksum = true);
size_t estimate_compressed_size(size_t uncompressed_size, int compression_level);
Status validate_config(const CompressionConfig &config);
void apply_level_parameters(CompressionConfig &config);
u32 get_optimal_block_size(u32 input_size, u32 compression_level);
constexpr const char *get_format_name() { return "cuda_zstd"; }
constexpr u32 get_format_version() { return 0x00010000; }
bool is_nvcomp_zstd_format(const void *compressed_data, size_t compressed_size);
Status extract_metadata(const void *compressed_data, size_t compressed_size, NvcompMetadata &metadata);

} // namespace cuda_zstd
#endif // __cplusplus

// ---- single-buffer C API (reference include/cuda_zstd_manager.h:433-479, src/cuda_zstd_c_api.cpp) ----
#ifdef __cplusplus
extern "C" {
#endif
typedef struct cuda_zstd_manager_t cuda_zstd_manager_t;
typedef struct cuda_zstd_dict_t cuda_zstd_dict_t;
cuda_zstd_manager_t *cuda_zstd_create_manager(int compression_level);
void cuda_zstd_destroy_manager(cuda_zstd_manager_t *manager);
int cuda_zstd_compress(cuda_zstd_manager_t *manager, const void *src, size_t src_size, void *dst, size_t *dst_size,
                       void *workspace, size_t workspace_size, cudaStream_t stream);
int cuda_zstd_decompress(cuda_zstd_manager_t *manager, const void *src, size_t src_size, void *dst, size_t *dst_size,
                         void *workspace, size_t workspace_size, cudaStream_t stream);
size_t cuda_zstd_get_compress_workspace_size(cuda_zstd_manager_t *manager, size_t src_size);
size_t cuda_zstd_get_decompress_workspace_size(cuda_zstd_manager_t *manager, size_t compressed_size);
// Dictionary calls keep their symbols; dictionaries are out of scope: train returns NULL and
// set returns ERROR_NOT_IMPLEMENTED (24 -> mapped to 1).
cuda_zstd_dict_t *cuda_zstd_train_dictionary(const void **samples, const size_t *sample_sizes, size_t num_samples,
                                             size_t dict_size);
void cuda_zstd_destroy_dictionary(cuda_zstd_dict_t *dict);
int cuda_zstd_set_dictionary(cuda_zstd_manager_t *manager, cuda_zstd_dict_t *dict);
const char *cuda_zstd_get_error_string(int error_code);
int cuda_zstd_is_error(int code);
#ifdef __cplusplus
}
#endif
#endif // CUDA_ZSTD_MANAGER_H
