// cuda_zstd_types.h -- plain data types of the batch-codec boundary (B200-native build).
//
// Drop-in for the PUBLIC part of the reference header of the same name: every enum value,
// struct field and default below matches what a caller of the batch path can observe
// (reference include/cuda_zstd_types.h: Status :92-128, ErrorContext :132-156, Strategy :162-171,
// ChecksumPolicy :186-190, CompressionConfig :196-232, CompressionStats :238-262, BatchItem
// :268-274, NvcompMetadata :290-298, constants :304-310).  Hybrid/streaming/memory-pool types of
// the reference are out of scope and are not declared.
#ifndef CUDA_ZSTD_TYPES_H
#define CUDA_ZSTD_TYPES_H

#include <cuda_runtime_api.h>

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace cuda_zstd {

using u8 = std::uint8_t;
using u16 = std::uint16_t;
using u32 = std::uint32_t;
using u64 = std::uint64_t;
using i32 = std::int32_t;
using i64 = std::int64_t;
using byte_t = unsigned char;

// Numeric values are ABI: the C entry points return them as int (see status_to_nvcomp_error).
enum class Status : u32 {
  SUCCESS = 0,
  ERROR_GENERIC = 1,
  ERROR_INVALID_PARAMETER = 2,
  ERROR_OUT_OF_MEMORY = 3,
  ERROR_CUDA_ERROR = 4,
  ERROR_INVALID_MAGIC = 5,
  ERROR_CORRUPT_DATA = 6,
  ERROR_CORRUPTED_DATA = 6,
  ERROR_BUFFER_TOO_SMALL = 7,
  ERROR_UNSUPPORTED_VERSION = 8,
  ERROR_DICTIONARY_MISMATCH = 9,
  ERROR_CHECKSUM_FAILED = 10,
  ERROR_IO = 11,
  ERROR_COMPRESSION = 12,
  ERROR_COMPRESSION_FAILED = 12,
  ERROR_DECOMPRESSION = 13,
  ERROR_DECOMPRESSION_FAILED = 13,
  ERROR_WORKSPACE_INVALID = 14,
  ERROR_STREAM_ERROR = 15,
  ERROR_ALLOCATION_FAILED = 16,
  ERROR_HASH_TABLE_FULL = 17,
  ERROR_SEQUENCE_ERROR = 18,
  ERROR_NOT_INITIALIZED = 19,
  ERROR_ALREADY_INITIALIZED = 20,
  ERROR_INVALID_STATE = 21,
  ERROR_TIMEOUT = 22,
  ERROR_CANCELLED = 23,
  ERROR_NOT_IMPLEMENTED = 24,
  ERROR_INTERNAL = 25,
  ERROR_UNKNOWN = 26,
  ERROR_DICTIONARY_FAILED = 27,
  ERROR_UNSUPPORTED_FORMAT = 28
};

struct ErrorContext {
  Status status = Status::SUCCESS;
  const char *file = nullptr;
  int line = 0;
  const char *function = nullptr;
  const char *message = nullptr;
  cudaError_t cuda_error = cudaSuccess;
  ErrorContext() = default;
  ErrorContext(Status s, const char *f, int l, const char *fn, const char *msg = nullptr)
      : status(s), file(f), line(l), function(fn), message(msg) {}
};
typedef void (*ErrorCallback)(const ErrorContext &ctx);

const char *status_to_string(Status status);
const char *get_detailed_error_message(const ErrorContext &ctx);
void set_error_callback(ErrorCallback callback);
void log_error(const ErrorContext &ctx);
ErrorContext get_last_error();
void clear_last_error();

enum class Strategy : u32 { FAST = 0, DFAST = 1, GREEDY = 2, LAZY = 3, LAZY2 = 4, BTLAZY2 = 5, BTOPT = 6, BTULTRA = 7 };
enum class CompressionMode : u32 { LEVEL_BASED = 0, STRATEGY_BASED = 1 };
enum class ChecksumPolicy : u32 { NO_COMPUTE_NO_VERIFY = 0, COMPUTE_NO_VERIFY = 1, COMPUTE_AND_VERIFY = 2 };

// Same fields, order and defaults as the reference struct.  In this build the batch path reads
// `level` (1..22), `checksum` and `block_size`; window/hash/chain/search logs are accepted and kept
// for get_config() round-trips but the device parse derives its own table sizes from the level and
// the chunk size (DESIGN.md "level mapping").  `cpu_threshold` is accepted and IGNORED: there is no
// CPU route inside the batch path.
struct CompressionConfig {
  CompressionMode compression_mode = CompressionMode::LEVEL_BASED;
  int level = 3;
  bool use_exact_level = true;
  Strategy strategy = Strategy::GREEDY;
  u32 window_log = 20;
  u32 hash_log = 17;
  u32 chain_log = 17;
  u32 search_log = 8;
  u32 min_match = 3;
  u32 target_length = 0;
  u32 block_size = 128 * 1024;
  bool enable_ldm = false;
  u32 ldm_hash_log = 20;
  ChecksumPolicy checksum = ChecksumPolicy::NO_COMPUTE_NO_VERIFY;
  u32 cpu_threshold = 1024 * 1024;

  static CompressionConfig from_level(int level);
  static CompressionConfig optimal(size_t input_size);
  static int strategy_to_default_level(Strategy s);
  static Strategy level_to_strategy(int level);
  Status validate() const;
  static CompressionConfig get_default();
};

struct CompressionStats {
  uint64_t input_bytes = 0;
  uint64_t output_bytes = 0;
  uint64_t num_blocks = 0;
  uint64_t num_sequences = 0;
  uint64_t num_literals = 0;
  uint64_t matches_found = 0;
  uint64_t bytes_compressed = 0;
  uint64_t bytes_produced = 0;
  uint64_t bytes_decompressed = 0;
  uint64_t blocks_processed = 0;
  double compression_time_ms = 0.0;
  double decompression_time_ms = 0.0;
  float get_ratio() const { return output_bytes ? static_cast<float>(input_bytes) / output_bytes : 0.0f; }
  double get_compression_throughput_gbps() const {
    return compression_time_ms > 0 ? (input_bytes / 1e9) / (compression_time_ms / 1000.0) : 0.0;
  }
};

// One chunk of a batch.  output_size is capacity on entry and bytes written on return.
struct BatchItem {
  void *input_ptr = nullptr;
  void *output_ptr = nullptr;
  size_t input_size = 0;
  size_t output_size = 0;
  Status status = Status::SUCCESS;
};

struct DictionaryContent {
  const unsigned char *d_buffer = nullptr;
  size_t size = 0;
  u32 dict_id = 0;
};

namespace dictionary {
// Dictionaries are out of scope for the batch path (the reference passes nullptr,0 there:
// src/cuda_zstd_manager.cu:5765); the type exists so the virtual interface keeps its shape.
struct Dictionary {
  const unsigned char *raw_content = nullptr;
  size_t raw_size = 0;
  u32 dict_id = 0;
};
} // namespace dictionary

struct NvcompMetadata {
  u32 format_version = 0;
  u32 compression_level = 0;
  u64 uncompressed_size = 0;
  u32 num_chunks = 0;
  u32 chunk_size = 0;
  u32 dictionary_id = 0;
  ChecksumPolicy checksum_policy = ChecksumPolicy::NO_COMPUTE_NO_VERIFY;
};

// ---- types of the HybridEngine surface (reference include/cuda_zstd_types.h:323-437) ----------------------------------
// The reference's engine routes between host libzstd and its kernels.  This build has no host codec: every mode runs the
// CUDA path, the mode / thresholds are kept for get_config() round trips, and results always name a GPU backend
// (include/cuda_zstd_hybrid.h).  Enumerator values and field names are the reference's, since callers (its Python
// binding, python/src/binding.cpp:689-751) spell them out.
enum class HybridMode : u32 { AUTO = 0, PREFER_CPU = 1, PREFER_GPU = 2, FORCE_CPU = 3, FORCE_GPU = 4, ADAPTIVE = 5 };
enum class DataLocation : u32 { HOST = 0, DEVICE = 1, MANAGED = 2, UNKNOWN = 3 };
enum class ExecutionBackend : u32 { CPU_LIBZSTD = 0, GPU_KERNELS = 1, CPU_PARALLEL = 2, GPU_BATCH = 3 };

struct HybridConfig {
  HybridMode mode = HybridMode::AUTO;
  size_t cpu_size_threshold = 1024 * 1024;      // kept, never consulted
  size_t gpu_device_threshold = 64 * 1024;      // kept, never consulted
  bool enable_profiling = false;
  int compression_level = 3;
  u32 cpu_thread_count = 0;
  bool use_pinned_memory = true;
  bool overlap_transfers = true;
};

struct HybridResult {
  ExecutionBackend backend_used = ExecutionBackend::GPU_KERNELS;
  DataLocation input_location = DataLocation::HOST;
  DataLocation output_location = DataLocation::HOST;
  double total_time_ms = 0.0;
  double transfer_time_ms = 0.0;
  double compute_time_ms = 0.0;
  double throughput_mbps = 0.0;
  size_t input_bytes = 0;
  size_t output_bytes = 0;
  float compression_ratio = 1.0f;
  const char *routing_reason = nullptr;
};

struct BatchRoutingResult {
  size_t item_index = 0;
  ExecutionBackend backend_used = ExecutionBackend::GPU_BATCH;
  Status status = Status::SUCCESS;
  size_t input_bytes = 0;
  size_t output_bytes = 0;
  double compute_time_ms = 0.0;
};

constexpr u32 ZSTD_MAGIC = 0xFD2FB528;
constexpr u32 MIN_COMPRESSION_LEVEL = 1;
constexpr u32 MAX_COMPRESSION_LEVEL = 22;
constexpr u32 DEFAULT_COMPRESSION_LEVEL = 3;
constexpr u32 MIN_WINDOW_LOG = 10;
constexpr u32 MAX_WINDOW_LOG = 31;
constexpr u32 DEFAULT_BLOCK_SIZE = 128 * 1024;
// headroom the allocation helpers (cuda_zstd_safe_alloc.h) leave free; reference include/cuda_zstd_types.h:447-450
constexpr size_t VRAM_SAFETY_BUFFER_BYTES = 256ULL * 1024 * 1024;
constexpr size_t RAM_SAFETY_BUFFER_BYTES = 512ULL * 1024 * 1024;

// small helpers callers of the batch path use (reference include/cuda_zstd_types.h:505-511)
inline float get_compression_ratio(size_t uncompressed, size_t compressed) { return compressed ? (float)uncompressed / (float)compressed : 0.0f; }
inline bool is_valid_compression_level(int level) { return level >= (int)MIN_COMPRESSION_LEVEL && level <= (int)MAX_COMPRESSION_LEVEL; }

} // namespace cuda_zstd
#endif // CUDA_ZSTD_TYPES_H
