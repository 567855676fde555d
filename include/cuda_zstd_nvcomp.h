// cuda_zstd_nvcomp.h -- nvCOMP-v5-style batch facade and its C ABI (B200-native build).
//
// Same public surface as the reference header of the same name for the batch path:
// NvcompV5Options (include/cuda_zstd_nvcomp.h:41-51), NvcompV5BatchManager (:93-137), option and
// status conversions (:54-58, :227-233), metadata helpers (:144-196), nvcomp_zstd_*_v5 (:272-336).
// The benchmark_level helpers (:243-266) are host-side loops around compress/decompress.
#ifndef CUDA_ZSTD_NVCOMP_H
#define CUDA_ZSTD_NVCOMP_H

#include "cuda_zstd_manager.h"

#ifdef __cplusplus
#include <memory>
#include <vector>

namespace cuda_zstd {
namespace nvcomp_v5 {

bool is_nvcomp_v5_zstd_format(const void *compressed_data, size_t compressed_size);   // device or host pointer
constexpr u32 get_nvcomp_v5_format_version() { return 0x00050000; }
bool is_compatible_with_nvcomp_v5(u32 format_version);

struct NvcompV5Options {
  int level;            // 1..22
  int algorithm;        // reserved
  u32 chunk_size;       // bytes per chunk the caller intends to use
  bool enable_checksum; // XXH64 content checksum in every frame
  NvcompV5Options() : level(3), algorithm(0), chunk_size(64 * 1024), enable_checksum(false) {}
};
NvcompV5Options to_nvcomp_v5_opts(const CompressionConfig &config);
CompressionConfig from_nvcomp_v5_opts(const NvcompV5Options &opts);
std::unique_ptr<ZstdManager> create_nvcomp_v5_manager(const NvcompV5Options &opts);

// Pointer-array batch API.  The four arrays may be host or device memory; data buffers and the
// workspace are device memory; sizes arrays are in = capacity, out = bytes written.
class NvcompV5BatchManager {
public:
  explicit NvcompV5BatchManager(const NvcompV5Options &opts);
  ~NvcompV5BatchManager();
  size_t get_compress_temp_size(const size_t *chunk_sizes, size_t num_chunks, cudaStream_t stream = 0) const;
  size_t get_decompress_temp_size(const size_t *compressed_sizes, size_t num_chunks, cudaStream_t stream = 0) const;
  size_t get_max_compressed_chunk_size(size_t uncompressed_chunk_size) const;
  Status compress_async(const void *const *d_uncompressed_ptrs, const size_t *uncompressed_sizes, size_t num_chunks,
                        void *const *d_compressed_ptrs, size_t *compressed_sizes, void *d_temp_storage,
                        size_t temp_storage_bytes, cudaStream_t stream = 0);
  Status decompress_async(const void *const *d_compressed_ptrs, const size_t *compressed_sizes, size_t num_chunks,
                          void *const *d_uncompressed_ptrs, size_t *uncompressed_sizes, void *d_temp_storage,
                          size_t temp_storage_bytes, cudaStream_t stream = 0);
  const CompressionStats &get_stats() const;
  ZstdBatchManager &batch_manager();     // the manager this facade drives (this build only)

private:
  class Impl;
  std::unique_ptr<Impl> pimpl_;
};

struct NvcompV5Metadata {
  u32 format_version;
  u32 library_version;
  int compression_level;
  u64 uncompressed_size;
  u64 compressed_size;
  u32 num_chunks;
  u32 chunk_size;
  u32 dictionary_id;
  ChecksumPolicy checksum_policy;
  bool has_dictionary;
  u64 checksum;
  NvcompV5Metadata()
      : format_version(get_nvcomp_v5_format_version()), library_version(0x00010000), compression_level(3),
        uncompressed_size(0), compressed_size(0), num_chunks(0), chunk_size(0), dictionary_id(0),
        checksum_policy(ChecksumPolicy::NO_COMPUTE_NO_VERIFY), has_dictionary(false), checksum(0) {}
};
Status get_metadata_async(const void *d_compressed_data, size_t compressed_size, NvcompV5Metadata *h_metadata,
                          cudaStream_t stream = 0);
Status get_metadata(const void *d_compressed_data, size_t compressed_size, NvcompV5Metadata &metadata);
bool validate_metadata(const NvcompV5Metadata &metadata);
Status get_decompressed_size_async(const void *d_compressed_data, size_t compressed_size, size_t *h_decompressed_size,
                                   cudaStream_t stream = 0);
Status get_num_chunks(const void *d_compressed_data, size_t compressed_size, size_t *num_chunks);
Status get_chunk_sizes(const void *d_compressed_data, size_t compressed_size, size_t *chunk_sizes, size_t max_chunks);

int status_to_nvcomp_error(Status status);
Status nvcomp_error_to_status(int nvcomp_error);
const char *get_nvcomp_v5_error_string(int error_code);

struct NvcompV5BenchmarkResult {
  int level;
  double compress_time_ms;
  double decompress_time_ms;
  double compress_throughput_mbps;
  double decompress_throughput_mbps;
  float compression_ratio;
  size_t compressed_size;
};
NvcompV5BenchmarkResult benchmark_level(const void *d_input, size_t input_size, int level, int iterations = 100,
                                        cudaStream_t stream = 0);
std::vector<NvcompV5BenchmarkResult> benchmark_all_levels(const void *d_input, size_t input_size, int iterations = 100,
                                                          cudaStream_t stream = 0);

} // namespace nvcomp_v5
} // namespace cuda_zstd
#endif // __cplusplus

#ifdef __cplusplus
extern "C" {
#endif
typedef void *nvcompZstdManagerHandle;
nvcompZstdManagerHandle nvcomp_zstd_create_manager_v5(int compression_level);
void nvcomp_zstd_destroy_manager_v5(nvcompZstdManagerHandle handle);
int nvcomp_zstd_compress_async_v5(nvcompZstdManagerHandle handle, const void *d_uncompressed, size_t uncompressed_size,
                                  void *d_compressed, size_t *compressed_size, void *d_temp, size_t temp_size,
                                  cudaStream_t stream);
int nvcomp_zstd_decompress_async_v5(nvcompZstdManagerHandle handle, const void *d_compressed, size_t compressed_size,
                                    void *d_uncompressed, size_t *uncompressed_size, void *d_temp, size_t temp_size,
                                    cudaStream_t stream);
size_t nvcomp_zstd_get_compress_temp_size_v5(nvcompZstdManagerHandle handle, size_t uncompressed_size);
size_t nvcomp_zstd_get_decompress_temp_size_v5(nvcompZstdManagerHandle handle, size_t compressed_size);
#ifdef __cplusplus
int nvcomp_zstd_get_metadata_v5(const void *d_compressed_data, size_t compressed_size,
                                cuda_zstd::nvcomp_v5::NvcompV5Metadata *h_metadata, cudaStream_t stream);
#endif
#ifdef __cplusplus
}
#endif
#endif // CUDA_ZSTD_NVCOMP_H
