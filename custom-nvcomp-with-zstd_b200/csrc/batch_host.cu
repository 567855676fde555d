// batch_host.cu -- host side of the drop-in boundary: ZstdBatchManager, NvcompV5BatchManager and the
// extern "C" entry points, all thin marshalling over the two batch kernels.
//
// Mirrors the reference's interface for this path (same names, argument meaning, error behaviour):
//   ZstdBatchManager            src/cuda_zstd_manager.cu:5540-6037
//   NvcompV5BatchManager        src/cuda_zstd_nvcomp.cpp:192-644
//   nvcomp_zstd_*_v5            src/cuda_zstd_nvcomp.cpp:766-840
//   cuda_zstd_*                 src/cuda_zstd_c_api.cpp:18-211
//   Status/config helpers       src/cuda_zstd_types.cpp:36-261, 831-950
// What is different by design: no per-item host loop, no CPU route (cpu_threshold is ignored), no
// allocation and no hidden streams inside a call -- one persistent kernel per batch on the caller's
// stream, with pointer/size tables staged in the caller's workspace.
#include "../../include/cuda_zstd_batch_c.h"
#include "../../include/cuda_zstd_nvcomp.h"
#include "zstd_device_api.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>
#include <new>
#include <vector>

namespace cuda_zstd {

// ============================================================================================
// Status strings and the global last-error slot (reference src/cuda_zstd_types.cpp:25-141)
// ============================================================================================
const char *status_to_string(Status s) {
  switch (static_cast<u32>(s)) {
  case 0: return "SUCCESS";
  case 1: return "ERROR_GENERIC";
  case 2: return "ERROR_INVALID_PARAMETER";
  case 3: return "ERROR_OUT_OF_MEMORY";
  case 4: return "ERROR_CUDA_ERROR";
  case 5: return "ERROR_INVALID_MAGIC";
  case 6: return "ERROR_CORRUPT_DATA";
  case 7: return "ERROR_BUFFER_TOO_SMALL";
  case 8: return "ERROR_UNSUPPORTED_VERSION";
  case 9: return "ERROR_DICTIONARY_MISMATCH";
  case 10: return "ERROR_CHECKSUM_FAILED";
  case 11: return "ERROR_IO";
  case 12: return "ERROR_COMPRESSION";
  case 13: return "ERROR_DECOMPRESSION";
  case 14: return "ERROR_WORKSPACE_INVALID";
  case 15: return "ERROR_STREAM_ERROR";
  case 16: return "ERROR_ALLOCATION_FAILED";
  case 17: return "ERROR_HASH_TABLE_FULL";
  case 18: return "ERROR_SEQUENCE_ERROR";
  case 19: return "ERROR_NOT_INITIALIZED";
  case 20: return "ERROR_ALREADY_INITIALIZED";
  case 21: return "ERROR_INVALID_STATE";
  case 22: return "ERROR_TIMEOUT";
  case 23: return "ERROR_CANCELLED";
  case 24: return "ERROR_NOT_IMPLEMENTED";
  case 25: return "ERROR_INTERNAL";
  case 26: return "ERROR_UNKNOWN";
  case 27: return "ERROR_DICTIONARY_FAILED";
  case 28: return "ERROR_UNSUPPORTED_FORMAT";
  default: return "UNKNOWN_STATUS";
  }
}

namespace {
std::mutex g_err_mu;
ErrorContext g_last_err;
ErrorCallback g_err_cb = nullptr;

thread_local bool g_quiet_errors = false;      // set around speculative attempts whose failure is not the call's result
Status fail(Status s, const char *fn, const char *msg, cudaError_t ce = cudaSuccess) {
  if (g_quiet_errors) return s;
  ErrorContext c(s, __FILE__, 0, fn, msg);
  c.cuda_error = ce;
  log_error(c);
  return s;
}
Status cuda_fail(cudaError_t e, const char *fn) { return fail(Status::ERROR_CUDA_ERROR, fn, cudaGetErrorString(e), e); }
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
} // namespace

const char *get_detailed_error_message(const ErrorContext &ctx) { return ctx.message ? ctx.message : status_to_string(ctx.status); }
void set_error_callback(ErrorCallback cb) { std::lock_guard<std::mutex> l(g_err_mu); g_err_cb = cb; }
void log_error(const ErrorContext &ctx) {
  ErrorCallback cb;
  { std::lock_guard<std::mutex> l(g_err_mu); g_last_err = ctx; cb = g_err_cb; }
  if (cb) cb(ctx);
}
ErrorContext get_last_error() { std::lock_guard<std::mutex> l(g_err_mu); return g_last_err; }
void clear_last_error() { std::lock_guard<std::mutex> l(g_err_mu); g_last_err = ErrorContext(); }

// ============================================================================================
// CompressionConfig (reference src/cuda_zstd_types.cpp:147-261, 860-950): same level bands
// ============================================================================================
Strategy CompressionConfig::level_to_strategy(int level) {
  static const int upper[] = {1, 3, 6, 12, 15, 18, 20};
  for (int i = 0; i < 7; ++i) if (level <= upper[i]) return static_cast<Strategy>(i);
  return Strategy::BTULTRA;
}
int CompressionConfig::strategy_to_default_level(Strategy s) {
  static const int lv[] = {1, 3, 5, 9, 13, 16, 19, 22};
  u32 i = static_cast<u32>(s);
  return i < 8 ? lv[i] : 3;
}
CompressionConfig CompressionConfig::from_level(int level) {
  CompressionConfig c;
  c.compression_mode = CompressionMode::LEVEL_BASED;
  c.level = level;
  c.use_exact_level = true;
  c.strategy = level_to_strategy(level);
  // table-size fields are informational in this build (the kernels size their own SMEM tables);
  // they keep the reference's band values so get_config() round-trips look the same
  struct Band { int upto; u32 wlog, hlog, clog; };
  static const Band bands[] = {{1, 18, 15, 15}, {3, 19, 17, 17}, {6, 20, 17, 17}, {9, 22, 18, 18}, {12, 23, 19, 19},
                               {14, 23, 19, 19}, {15, 23, 20, 19}, {22, 23, 20, 20}};
  for (const Band &b : bands) if (level <= b.upto) { c.window_log = b.wlog; c.hash_log = b.hlog; c.chain_log = b.clog; break; }
  static const u32 depth[] = {1, 1, 1, 1, 2, 4, 8, 8, 16, 32, 64, 128, 256, 256, 256, 512, 512, 512, 999, 999, 999, 999, 999};
  c.search_log = depth[std::min(std::max(level, 0), 22)];
  return c;
}
CompressionConfig CompressionConfig::get_default() { return from_level(static_cast<int>(DEFAULT_COMPRESSION_LEVEL)); }
CompressionConfig CompressionConfig::optimal(size_t) { return from_level(3); }
Status CompressionConfig::validate() const {
  if (level < static_cast<int>(MIN_COMPRESSION_LEVEL) || level > static_cast<int>(MAX_COMPRESSION_LEVEL)) return Status::ERROR_INVALID_PARAMETER;
  if (window_log < MIN_WINDOW_LOG || window_log > MAX_WINDOW_LOG) return Status::ERROR_INVALID_PARAMETER;
  if (block_size < 1024) return Status::ERROR_INVALID_PARAMETER;
  return Status::SUCCESS;
}
Status validate_config(const CompressionConfig &c) { return c.validate(); }
void apply_level_parameters(CompressionConfig &c) {
  CompressionConfig t = CompressionConfig::from_level(c.level);
  c.strategy = t.strategy; c.window_log = t.window_log; c.hash_log = t.hash_log; c.chain_log = t.chain_log; c.search_log = t.search_log;
}
u32 get_optimal_block_size(u32 input_size, u32) { return std::min<u32>(std::max<u32>(input_size, 1024u), 128u * 1024u); }

size_t estimate_compressed_size(size_t n, int) {
  size_t blocks = (n + (128 * 1024 - 1)) / (128 * 1024);
  if (blocks == 0) blocks = 1;
  return n + n / 255 + blocks * 3 + 512;
}

ZstdManager::ExecutionPath ZstdManager::select_execution_path(size_t size, int cpu_threshold) {
  return (cpu_threshold > 0 && size < static_cast<size_t>(cpu_threshold)) ? ExecutionPath::CPU : ExecutionPath::GPU_BATCH;
}

// ============================================================================================
// frame-header peek on a HOST copy of the first bytes (reference parse_frame_header,
// src/cuda_zstd_manager.cu:4108-4225, but with the single-segment 1-byte FCS handled per RFC)
// ============================================================================================
namespace {
struct HeaderInfo { bool ok = false; bool has_size = false; u64 content_size = 0; bool checksum = false; u32 dict_id = 0; u32 header_bytes = 0; };
HeaderInfo peek_header(const unsigned char *p, size_t n) {
  HeaderInfo h;
  if (n < 5) return h;
  u32 magic = p[0] | (p[1] << 8) | (p[2] << 16) | ((u32)p[3] << 24);
  if (magic != ZSTD_MAGIC) return h;
  const u32 fhd = p[4];
  const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
  const int did = did_flag == 3 ? 4 : did_flag, fcs = fcs_flag == 0 ? single : (1 << fcs_flag);
  size_t pos = 5 + (single ? 0 : 1);
  if (pos + did + fcs > n) return h;
  for (int k = 0; k < did; ++k) h.dict_id |= (u32)p[pos + k] << (8 * k);
  pos += did;
  if (fcs) {
    for (int k = 0; k < fcs; ++k) h.content_size |= (u64)p[pos + k] << (8 * k);
    if (fcs == 2) h.content_size += 256;
    h.has_size = true;
    pos += fcs;
  }
  h.checksum = (fhd >> 2) & 1;
  h.header_bytes = (u32)pos;
  h.ok = true;
  return h;
}
} // namespace

namespace {
// strict form: a frame without a content size is an error (what the metadata helpers below want)
Status frame_content_size(const void *data, size_t size, size_t *out) {
  if (!data || !out) return Status::ERROR_INVALID_PARAMETER;
  unsigned char head[18];
  size_t n = std::min<size_t>(size, sizeof head);
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, data) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
    if (cudaMemcpy(head, data, n, cudaMemcpyDeviceToHost) != cudaSuccess) return Status::ERROR_CUDA_ERROR;
  } else { (void)cudaGetLastError(); std::memcpy(head, data, n); }
  HeaderInfo h = peek_header(head, n);
  if (!h.ok) return Status::ERROR_INVALID_MAGIC;
  if (!h.has_size) return Status::ERROR_CORRUPT_DATA;
  *out = (size_t)h.content_size;
  return Status::SUCCESS;
}
} // namespace
// Public form, host or device pointer (reference src/cuda_zstd_types.cpp:1058-1106): a frame that does not declare its size
// answers SUCCESS with 0 -- the reference's Python binding then falls back to an estimate (python/src/binding.cpp:247-250).
Status get_decompressed_size(const void *data, size_t size, size_t *out) {
  if (!data || !out || size < 4) return Status::ERROR_INVALID_PARAMETER;
  *out = 0;
  const Status s = frame_content_size(data, size, out);
  if (s == Status::ERROR_CORRUPT_DATA) { *out = 0; return Status::SUCCESS; }
  return s;
}
// Header-level check only, like the reference (src/cuda_zstd_types.cpp:1108-1170: magic, then "the header parses");
// check_checksum is accepted and, as there, does not decode the frame.
Status validate_compressed_data(const void *data, size_t size, bool) {
  if (!data) return Status::ERROR_INVALID_PARAMETER;
  if (size < 4) return Status::ERROR_CORRUPT_DATA;
  size_t s = 0;
  const Status st = frame_content_size(data, size, &s);
  return st == Status::ERROR_CORRUPT_DATA ? Status::SUCCESS : st;
}
bool is_nvcomp_zstd_format(const void *data, size_t size) {
  size_t s = 0;
  Status st = frame_content_size(data, size, &s);
  return st == Status::SUCCESS || st == Status::ERROR_CORRUPT_DATA;
}
Status extract_metadata(const void *data, size_t size, NvcompMetadata &m) {
  size_t s = 0;
  Status st = frame_content_size(data, size, &s);
  if (st != Status::SUCCESS) return st;
  m = NvcompMetadata();
  m.format_version = get_format_version();
  m.uncompressed_size = s;
  m.num_chunks = 1;
  m.chunk_size = (u32)std::min<size_t>(s, 0xFFFFFFFFu);
  return Status::SUCCESS;
}

// ============================================================================================
// ZstdBatchManager
// ============================================================================================
// Workspace layout (caller-owned device memory), identical for both directions:
//   [0, 256)                         work-queue counter and padding
//   [256, 256 + align256(40 n))      staged tables: in_ptrs | in_sizes | out_ptrs | out_sizes | statuses
//   [.., .. + grid * per_cta)        per-CTA scratch (literal buffer for decode; parse scratch for encode)
// grid = min(n, SMs * resident CTAs) -- scratch scales with the GPU, not with the batch.
class ZstdBatchManager::Impl {
public:
  CompressionConfig cfg;
  CompressionStats stats;
  int sm_count = 0;
  int dec_ctas_per_sm = 1;
  int last_launches = 0;
  bool bare_mode = false;      // run(): the items are the block units of one frame (decompress_big)
  std::mutex mu;   // one manager per thread is the contract; the lock keeps misuse safe (reference: api_mutex)
  b200zstd::FastOverlap overlap{};     // helper stream + events, created on first decode (no device memory)
  bool overlap_ready = false;
  ~Impl() {
    if (overlap_ready) {
      for (auto &e : overlap.ev) cudaEventDestroy(e);
      cudaEventDestroy(overlap.done);
      cudaEventDestroy(overlap.prep);
      cudaStreamDestroy(overlap.side);
    }
  }
  const b200zstd::FastOverlap *get_overlap() {
    if (!overlap_ready) {
      if (cudaStreamCreateWithFlags(&overlap.side, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
      bool ok = cudaEventCreateWithFlags(&overlap.done, cudaEventDisableTiming) == cudaSuccess;
      ok = ok && cudaEventCreateWithFlags(&overlap.prep, cudaEventDisableTiming) == cudaSuccess;
      for (auto &e : overlap.ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
      if (!ok) { (void)cudaGetLastError(); return nullptr; }
      overlap_ready = true;
    }
    return &overlap;
  }

  int cached_dev = -1;
  std::vector<unsigned char> dict_bytes;      // raw content of the dictionary set on this manager (kept, not referenced by the encoder)
  u32 dict_id = 0;
  // the SM count and the decoder's residency belong to the CURRENT device: re-read when a call arrives on another one
  void refresh_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = -1; }
    if (dev == cached_dev && sm_count > 0) return;
    sm_count = 0;
    if (dev >= 0) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) { (void)cudaGetLastError(); sm_count = 148; }     // size queries still work without a device
    dec_ctas_per_sm = b200zstd::decode_ctas_per_sm();
    (void)cudaGetLastError();
    cached_dev = dev;
  }
  Impl() { refresh_device(); }
  b200zstd::EncodeParams enc_params() const {
    return b200zstd::encode_params_for_level(cfg.level, cfg.checksum != ChecksumPolicy::NO_COMPUTE_NO_VERIFY);
  }
  int dec_grid(size_t n) const { return (int)std::min<size_t>(n, (size_t)sm_count * dec_ctas_per_sm); }
  int enc_grid(size_t n) const {
    int per = b200zstd::encode_ctas_per_sm(enc_params());
    (void)cudaGetLastError();
    return (int)std::min<size_t>(n, (size_t)sm_count * per);
  }
  static size_t table_bytes(size_t n) { return align_up(n * 40, 256); }
  // ---- encoder placement: levels 1-4 run the decoupled pipeline (zstd_encode_esd.cu), the chain levels and items of
  // more than one block the general kernel (zstd_encode.cu) ----
  bool use_esd() const { return b200zstd::esd_level(std::max(cfg.level, 1)); }
  // slot geometry of the levels 1-4 pipeline: 64 KB blocks when every item is known to fit, else 128 KB
  static uint32_t esd_block_max(size_t max_item) { return (max_item == 0 || max_item > 65536) ? 131072u : 65536u; }
  size_t lists_bytes(size_t n) const { return align_up(n * 8 + (use_esd() ? b200zstd::esd_counter_words(n, 131072u) * 4 : 0), 256); }
  size_t enc_scratch_bytes(size_t n, uint32_t bm) const {
    const size_t general = b200zstd::encode_cta_scratch_bytes(enc_params());
    if (use_esd()) return std::max(b200zstd::esd_scratch_bytes(n, bm, sm_count), general);
    return (size_t)enc_grid(n) * general;
  }
  size_t enc_need(size_t n, uint32_t bm) const { return b200zstd::WS_HEADER_BYTES + table_bytes(n) + lists_bytes(n) + enc_scratch_bytes(n, bm); }
  // what the size QUERIES report: the need of whichever level wants the most, so that a workspace sized with one manager
  // serves a manager at any other level (the reference's tests size one workspace at "the highest level tested",
  // tests/test_c_api_edge_cases.cu:268-272); run() itself only asks for what its own level needs
  size_t enc_scratch_any_level(size_t n, uint32_t bm) const {
    const bool ck = cfg.checksum != ChecksumPolicy::NO_COMPUTE_NO_VERIFY;
    size_t scratch = b200zstd::esd_scratch_bytes(n, bm, sm_count);
    for (int level : {1, 3, 5, 7, 9, 12, 19}) {
      const b200zstd::EncodeParams prm = b200zstd::encode_params_for_level(level, ck);
      const size_t general = b200zstd::encode_cta_scratch_bytes(prm);
      int per = b200zstd::encode_ctas_per_sm(prm);
      (void)cudaGetLastError();
      const size_t grid = std::min<size_t>(n, (size_t)sm_count * (size_t)std::max(per, 1));
      scratch = std::max(scratch, b200zstd::esd_level(level) ? general : grid * general);
    }
    return scratch;
  }
  size_t lists_bytes_any_level(size_t n) const { return align_up(n * 8 + b200zstd::esd_counter_words(n, 131072u) * 4, 256); }
  size_t enc_need_any_level(size_t n, uint32_t bm) const {
    return b200zstd::WS_HEADER_BYTES + table_bytes(n) + lists_bytes_any_level(n) + enc_scratch_any_level(n, bm);
  }
  // a.counter / a.scratch are set here; w = workspace base, lists = 2 n words followed by the pipeline's counters
  cudaError_t enqueue_encode(b200zstd::EncodeArgs &a, unsigned char *w, unsigned char *lists, unsigned char *scratch, size_t scratch_bytes,
                             uint32_t bm, size_t min_item, size_t max_item, cudaStream_t stream, int *launches) {
    a.scratch = scratch;
    if (use_esd()) {
      b200zstd::EsdLaunch L{};
      L.lists = reinterpret_cast<u32 *>(lists);
      L.counters = L.lists + 2 * (size_t)a.n;
      L.counter_words = b200zstd::esd_counter_words(a.n, 131072u);
      L.scratch = scratch; L.scratch_bytes = scratch_bytes;
      L.sm_count = sm_count; L.block_max = bm; L.min_item_bytes = min_item; L.max_item_bytes = max_item;
      return b200zstd::launch_encode_esd(a, L, stream, launches);
    }
    a.counter = reinterpret_cast<u32 *>(w);
    *launches = 1;
    return b200zstd::launch_encode_batch(a, enc_grid(a.n), stream);
  }
  // decode workspace: header | staged tables | slow list | general-kernel literal scratch | fast-path slots
  static size_t wave_of(size_t n) { return std::min<size_t>(n, b200zstd::FAST_WAVE); }
  static size_t slow_list_bytes(size_t n) { return align_up(wave_of(n) * 8, 256); }      // slow list / KA order, then KC's order
  size_t general_scratch_bytes(size_t n) const { return align_up((size_t)dec_grid(n) * b200zstd::LIT_SCRATCH_BYTES, 256); }
  size_t dec_fixed(size_t n) const {
    return b200zstd::WS_HEADER_BYTES + table_bytes(n) + slow_list_bytes(n) + general_scratch_bytes(n) +
           wave_of(n) * b200zstd::FAST_SLOT_BYTES;
  }
  // pool for one wave (literals and 16-byte sequence records share it), sized from the compressed bytes of the
  // largest wave: Huffman literals regenerate to at most ~2x their coded size on compressible data and a sequence
  // costs >= ~2.2 coded bytes (match-heavy P=0.90 frames: 4,400 sequences in 9.5 KB), so 8 x compressed covers both
  // extremes.  Too small a pool is not an error: the overflow chunks take the general kernel.
  static size_t pool_share(size_t compressed) { return std::min<size_t>(8 * compressed, 192 * 1024) + 3072; }
  size_t dec_temp(size_t n, const size_t *sizes = nullptr) const {
    if (n == 0) return 0;
    size_t worst = 0;
    if (sizes) {
      for (size_t w0 = 0; w0 < n; w0 += b200zstd::FAST_WAVE) {
        size_t s = 0;
        for (size_t i = w0; i < std::min<size_t>(n, w0 + b200zstd::FAST_WAVE); ++i) s += pool_share(sizes[i]);
        worst = std::max(worst, s);
      }
    } else worst = wave_of(n) * pool_share(16 * 1024);      // no sizes given: assume 16 KB frames
    return dec_fixed(n) + align_up(worst, 256);
  }
  size_t enc_temp(size_t n, const size_t *sizes = nullptr) const {
    if (n == 0) return 0;
    size_t max_item = 0;
    if (sizes) for (size_t i = 0; i < n; ++i) max_item = std::max(max_item, sizes[i]);
    size_t e = enc_need_any_level(n, esd_block_max(max_item));
    // a compress workspace can always be reused for decompress (reference tests/test_c_api.cpp:62-64):
    // the decoder needs dec_fixed(); pool space beyond that only decides how many chunks take the fast path
    return std::max(e, dec_fixed(n) + wave_of(n) * (size_t)(64 * 1024));
  }

  // ---- one large buffer as ONE frame of independently encoded 128 KB blocks ----
  // Every block is a work item of the batch encoder (block mode: header + payload only, repeat offsets unknown at the
  // block start), a device scan places the blocks behind the frame header and the pack kernel moves them.  Workspace:
  // [header | 5 block tables + offsets | encoder scratch for min(B, resident) CTAs | B staging slots].
  static constexpr size_t BIG_BLOCK = 128 * 1024;
  // Buffers up to 4 MB are cut into 64 KB blocks instead (window descriptor 64 KB, so that any decoder knows the block
  // size): twice the blocks in flight and the shared-memory block geometry halve the latency of one call, for 0.4 - 2 % of
  // size on real data (matches end at block borders).  Above that 128 KB blocks fill the GPU anyway.
  static constexpr size_t BIG_SMALL_BLOCK = 64 * 1024, BIG_SMALL_LIMIT = 4u << 20;
  static size_t big_block_for(size_t n) { return n <= BIG_SMALL_LIMIT ? BIG_SMALL_BLOCK : BIG_BLOCK; }
  static size_t big_blocks(size_t n) { const size_t b = big_block_for(n); return (n + b - 1) / b; }
  static size_t big_slot(size_t blk) { return align_up(blk + 3 + 64, 256); }
  static size_t big_tables(size_t B) { return align_up(B * 44 + (B + 1) * 8, 256); }
  size_t big_temp(size_t n) const {
    const size_t B = big_blocks(n), blk = big_block_for(n);
    return b200zstd::WS_HEADER_BYTES + big_tables(B) + lists_bytes(B) + enc_scratch_bytes(B, (uint32_t)blk) + B * big_slot(blk);
  }
  size_t big_temp_any_level(size_t n) const {          // for the size queries (see enc_need_any_level)
    const size_t B = big_blocks(n), blk = big_block_for(n);
    return b200zstd::WS_HEADER_BYTES + big_tables(B) + lists_bytes_any_level(B) + enc_scratch_any_level(B, (uint32_t)blk) + B * big_slot(blk);
  }
  // Enqueue only: nothing is synchronised.  The outcome {frame bytes, first failing block status} lands in the 16-byte
  // device mailbox at ws + 64 and, when h_result is given (pinned host memory), is copied there on the same stream.
  Status compress_big_enqueue(const void *d_src, size_t n, void *d_dst, size_t cap, void *ws, size_t ws_bytes, cudaStream_t stream,
                              u64 *h_result) {
    const char *fn = "compress";
    const size_t B = big_blocks(n), blk = big_block_for(n);
    const bool ck = cfg.checksum != ChecksumPolicy::NO_COMPUTE_NO_VERIFY;
    if (n > 0xFFFF0000ull) return fail(Status::ERROR_UNSUPPORTED_VERSION, fn, "single buffers of 4 GiB and more are not supported");
    if (ws_bytes < big_temp(n)) return fail(Status::ERROR_BUFFER_TOO_SMALL, fn, "workspace too small");
    // frame header: windowed (one block: no match leaves its block) with a 4-byte content size (RFC 8878 3.1.1.1)
    const unsigned wlog = blk == BIG_BLOCK ? 17u : 16u;
    unsigned char hdr[10] = {0x28, 0xB5, 0x2F, 0xFD, (unsigned char)(0x80 | (ck ? 0x04 : 0)), (unsigned char)((wlog - 10) << 3),
                             (unsigned char)n, (unsigned char)(n >> 8), (unsigned char)(n >> 16), (unsigned char)(n >> 24)};
    if (cap < sizeof hdr + n + 3 * B + (ck ? 4 : 0)) return fail(Status::ERROR_BUFFER_TOO_SMALL, fn, "output capacity below the worst case");
    std::lock_guard<std::mutex> lock(mu);
    unsigned char *w = static_cast<unsigned char *>(ws);
    u32 *counter = reinterpret_cast<u32 *>(w);
    u64 *d_result = reinterpret_cast<u64 *>(w + 64);
    unsigned char *tab = w + b200zstd::WS_HEADER_BYTES;
    unsigned char *lists = tab + big_tables(B);
    unsigned char *scratch = lists + lists_bytes(B);
    unsigned char *slots = scratch + enc_scratch_bytes(B, (uint32_t)blk);
    std::vector<u64> host(4 * B);
    for (size_t i = 0; i < B; ++i) {
      host[i] = (u64)(uintptr_t)(static_cast<const unsigned char *>(d_src) + i * blk);
      host[B + i] = std::min(blk, n - i * blk);
      host[2 * B + i] = (u64)(uintptr_t)(slots + i * big_slot(blk));
      host[3 * B + i] = big_slot(blk);
    }
    cudaError_t e;
    // (pageable -> device async copies are staged by the runtime before they return, so the locals may go out of scope)
    if ((e = cudaMemcpyAsync(tab, host.data(), 4 * B * 8, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, fn);
    if ((e = cudaMemcpyAsync(d_dst, hdr, sizeof hdr, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, fn);
    u64 *d_offsets = reinterpret_cast<u64 *>(tab + 4 * B * 8);
    u32 *d_status = reinterpret_cast<u32 *>(tab + 4 * B * 8 + (B + 1) * 8);
    b200zstd::EncodeArgs a{};
    a.in_ptrs = reinterpret_cast<const void *const *>(tab); a.in_sizes = reinterpret_cast<const size_t *>(tab + B * 8);
    a.out_ptrs = reinterpret_cast<void *const *>(tab + 2 * B * 8); a.out_sizes = reinterpret_cast<size_t *>(tab + 3 * B * 8);
    a.statuses = d_status; a.counter = counter; a.n = (uint32_t)B; a.block_mode = 1; a.prm = enc_params();
    a.prm.checksum = 0;
    int enc_launches = 0;
    if ((e = enqueue_encode(a, w, lists, scratch, enc_scratch_bytes(B, (uint32_t)blk), (uint32_t)blk, n - (B - 1) * blk, std::min(n, blk), stream,
                            &enc_launches)) != cudaSuccess) return cuda_fail(e, fn);
    if ((e = b200zstd::launch_scan_sizes(a.out_sizes, B, sizeof hdr, reinterpret_cast<uint64_t *>(d_offsets), stream)) != cudaSuccess) return cuda_fail(e, fn);
    if ((e = b200zstd::launch_pack(a.out_ptrs, a.out_sizes, reinterpret_cast<const uint64_t *>(d_offsets), B, d_dst, stream)) != cudaSuccess) return cuda_fail(e, fn);
    last_launches = 3 + enc_launches;
    if (ck) {
      if ((e = b200zstd::launch_frame_checksum(d_src, n, d_dst, reinterpret_cast<const uint64_t *>(d_offsets + B), stream)) != cudaSuccess) return cuda_fail(e, fn);
      last_launches++;
    }
    if ((e = b200zstd::launch_big_result(d_status, B, reinterpret_cast<const uint64_t *>(d_offsets + B), ck ? 4 : 0, d_result, stream)) != cudaSuccess) return cuda_fail(e, fn);
    if (h_result && (e = cudaMemcpyAsync(h_result, d_result, 16, cudaMemcpyDefault, stream)) != cudaSuccess) return cuda_fail(e, fn);
    return Status::SUCCESS;
  }
  Status compress_big(const void *d_src, size_t n, void *d_dst, size_t *dst_size, void *ws, size_t ws_bytes, cudaStream_t stream) {
    const char *fn = "compress";
    u64 res[2] = {0, 0};
    Status s = compress_big_enqueue(d_src, n, d_dst, *dst_size, ws, ws_bytes, stream, res);
    if (s != Status::SUCCESS) return s;
    cudaError_t e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return cuda_fail(e, fn);
    if (res[1] != 0) return fail(static_cast<Status>((u32)res[1]), fn, "a block failed to encode");
    *dst_size = (size_t)res[0];
    stats.input_bytes += n; stats.output_bytes += *dst_size; stats.bytes_compressed += n; stats.bytes_produced += *dst_size;
    return Status::SUCCESS;
  }

  // ---- one multi-block frame decoded block-parallel (decode half) ----
  // launch_split_frame cuts the frame into block units (speculating that every block but the last regenerates 128 KB),
  // the units go through the batch fast path as bare blocks, the whole-content checksum is verified by one warp.
  // Returns ERROR_NOT_IMPLEMENTED when the frame is not of that kind or the workspace is too small for it: the caller
  // then decodes serially (general kernel), which is also what settles any error a unit reports.
  size_t big_dec_tail(size_t B) const { return align_up(B * 32 + sizeof(b200zstd::SplitInfo) + 64, 256); }
  size_t big_dec_temp(size_t n, size_t B) const { return dec_fixed(B) + align_up(8 * n + 3072 * B, 256) + big_dec_tail(B); }
  Status decompress_big(const void *d_src, size_t n, void *d_dst, size_t cap, size_t *out, void *ws, size_t ws_bytes, cudaStream_t stream) {
    const char *fn = "decompress";
    const size_t Bcap = (cap + BIG_SMALL_BLOCK - 1) / BIG_SMALL_BLOCK;     // (a frame of 128 KB blocks uses half of them)
    if (Bcap < 2 || Bcap > 0xFFFFFFu || ws_bytes < big_dec_temp(n, Bcap)) return Status::ERROR_NOT_IMPLEMENTED;
    std::lock_guard<std::mutex> lock(mu);
    const size_t tail = big_dec_tail(Bcap), body = (ws_bytes - tail) & ~(size_t)255;
    unsigned char *t = static_cast<unsigned char *>(ws) + body;
    const void **u_in = reinterpret_cast<const void **>(t);
    size_t *u_in_sz = reinterpret_cast<size_t *>(t + Bcap * 8);
    void **u_out = reinterpret_cast<void **>(t + Bcap * 16);
    size_t *u_out_sz = reinterpret_cast<size_t *>(t + Bcap * 24);
    b200zstd::SplitInfo *d_info = reinterpret_cast<b200zstd::SplitInfo *>(t + Bcap * 32);
    u32 *d_flag = reinterpret_cast<u32 *>(t + Bcap * 32 + sizeof(b200zstd::SplitInfo));
    cudaError_t e;
    if ((e = b200zstd::launch_split_frame(d_src, n, d_dst, cap, (uint32_t)Bcap, u_in, u_in_sz, u_out, u_out_sz, d_info, stream)) != cudaSuccess)
      return cuda_fail(e, fn);
    b200zstd::SplitInfo info{};
    if ((e = cudaMemcpyAsync(&info, d_info, sizeof info, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return cuda_fail(e, fn);
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return cuda_fail(e, fn);
    if (!info.ok || info.units < 2) return Status::ERROR_NOT_IMPLEMENTED;
    std::vector<u32> st;
    bare_mode = true;
    g_quiet_errors = true;        // a unit that is not self-contained is not an error of the call: the serial decode decides
    Status s = run(false, u_in, u_in_sz, info.units, u_out, u_out_sz, nullptr, true, ws, body, stream, true, &st, nullptr);
    g_quiet_errors = false;
    bare_mode = false;
    const int launches = last_launches + 1;
    if (s != Status::SUCCESS) return Status::ERROR_NOT_IMPLEMENTED;                  // some unit was not self-contained: serial decode decides
    if (info.has_checksum && cfg.checksum == ChecksumPolicy::COMPUTE_AND_VERIFY) {
      u32 bad = 0;
      if ((e = b200zstd::launch_verify_checksum(d_dst, (size_t)info.content_size, static_cast<const unsigned char *>(d_src) + info.checksum_off,
                                                d_flag, stream)) != cudaSuccess) return cuda_fail(e, fn);
      if ((e = cudaMemcpyAsync(&bad, d_flag, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return cuda_fail(e, fn);
      if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return cuda_fail(e, fn);
      if (bad) return fail(Status::ERROR_CHECKSUM_FAILED, fn, "content checksum mismatch");
      last_launches = launches + 1;
    } else last_launches = launches;
    *out = (size_t)info.content_size;
    stats.input_bytes += n; stats.output_bytes += *out; stats.bytes_decompressed += *out;
    return Status::SUCCESS;
  }

  // Direction-agnostic batch driver.  tables_on_device: the five tables already live in device
  // memory (no staging, and with sync == false no host synchronisation at all).
  Status run(bool compress, const void *const *in_ptrs, const size_t *in_sizes, size_t n, void *const *out_ptrs,
             size_t *out_sizes, u32 *statuses, bool tables_on_device, void *ws, size_t ws_bytes, cudaStream_t stream,
             bool sync, std::vector<u32> *host_status, std::vector<size_t> *host_out_sizes) {
    const char *fn = compress ? "compress_batch" : "decompress_batch";
    last_launches = 0;
    if (n == 0) return Status::SUCCESS;
    refresh_device();
    if (!in_ptrs || !in_sizes || !out_ptrs || !out_sizes) return fail(Status::ERROR_INVALID_PARAMETER, fn, "null pointer table");
    if (n > 0xFFFFFFF0ull) return fail(Status::ERROR_INVALID_PARAMETER, fn, "too many chunks");
    size_t min_item = 0, max_item = 0;              // 0 / 0 = unknown (device-resident size table)
    if (compress && !tables_on_device) {
      min_item = ~(size_t)0;
      for (size_t i = 0; i < n; ++i) { min_item = std::min(min_item, in_sizes[i]); max_item = std::max(max_item, in_sizes[i]); }
      if (max_item == 0) max_item = 1;              // all-empty batch: every item is an error, any kernel reports it
    }
    // levels 1-4: 128 KB slots unless every item is known to fit 64 KB -- or, with a device-resident size table, unless
    // the workspace was sized for 64 KB chunks (anything larger then takes the general kernel)
    uint32_t enc_bm = esd_block_max(max_item);
    if (compress && tables_on_device && ws_bytes < enc_need(n, 131072u)) enc_bm = 65536u;
    const size_t need = compress ? enc_need(n, enc_bm) : dec_fixed(n);     // decode pools use whatever lies beyond the fixed part
    if (!ws) return fail(Status::ERROR_INVALID_PARAMETER, fn, "null workspace");
    if (ws_bytes < need) return fail(Status::ERROR_BUFFER_TOO_SMALL, fn, "workspace too small");
    unsigned char *w = static_cast<unsigned char *>(ws);
    u32 *counter = reinterpret_cast<u32 *>(w);
    unsigned char *tab = w + b200zstd::WS_HEADER_BYTES;
    unsigned char *scratch = tab + table_bytes(n);
    const void *const *d_in = in_ptrs;
    const size_t *d_in_sz = in_sizes;
    void *const *d_out = out_ptrs;
    size_t *d_out_sz = out_sizes;
    u32 *d_status = statuses;
    cudaError_t e;
    if (!tables_on_device) {
      // stage the four host tables into the workspace (4 async copies, stream-ordered)
      const size_t b = n * 8;
      if ((e = cudaMemcpyAsync(tab, in_ptrs, b, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, fn);
      if ((e = cudaMemcpyAsync(tab + b, in_sizes, b, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, fn);
      if ((e = cudaMemcpyAsync(tab + 2 * b, out_ptrs, b, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, fn);
      if ((e = cudaMemcpyAsync(tab + 3 * b, out_sizes, b, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, fn);
      d_in = reinterpret_cast<const void *const *>(tab);
      d_in_sz = reinterpret_cast<const size_t *>(tab + b);
      d_out = reinterpret_cast<void *const *>(tab + 2 * b);
      d_out_sz = reinterpret_cast<size_t *>(tab + 3 * b);
    }
    if (!d_status) d_status = reinterpret_cast<u32 *>(tab + 4 * n * 8);
    if (compress) {
      b200zstd::EncodeArgs a{};
      a.in_ptrs = d_in; a.in_sizes = d_in_sz; a.out_ptrs = d_out; a.out_sizes = d_out_sz; a.statuses = d_status;
      a.counter = counter; a.n = (uint32_t)n; a.prm = enc_params();
      e = enqueue_encode(a, w, scratch, scratch + lists_bytes(n), enc_scratch_bytes(n, enc_bm), enc_bm, min_item, max_item, stream, &last_launches);
    } else {
      // fast path in waves of FAST_WAVE chunks (bounds the scratch); each wave is 4 launches:
      // prep, entropy, execute, and the general kernel for whatever the fast path declined
      unsigned char *slow_list = scratch;
      unsigned char *gen_scratch = slow_list + slow_list_bytes(n);
      unsigned char *slots = gen_scratch + general_scratch_bytes(n);
      unsigned char *pools = slots + wave_of(n) * b200zstd::FAST_SLOT_BYTES;
      const size_t pool_bytes = (ws_bytes - (size_t)(pools - w)) & ~(size_t)255;
      e = cudaSuccess;
      for (size_t w0 = 0; w0 < n && e == cudaSuccess; w0 += b200zstd::FAST_WAVE) {
        const size_t m = std::min<size_t>(b200zstd::FAST_WAVE, n - w0);
        b200zstd::FastDecodeArgs f{};
        f.base.in_ptrs = d_in + w0; f.base.in_sizes = d_in_sz + w0; f.base.out_ptrs = d_out + w0; f.base.out_sizes = d_out_sz + w0;
        f.base.statuses = d_status + w0; f.base.counter = counter; f.base.lit_scratch = gen_scratch; f.base.n = (uint32_t)m;
        f.base.verify_checksum = cfg.checksum == ChecksumPolicy::COMPUTE_AND_VERIFY;   // reference gates on the manager's policy (manager.cu:3654)
        f.slots = slots;
        f.lit_pool = f.seq_pool = pools; f.lit_pool_bytes = f.seq_pool_bytes = pool_bytes;
        f.pool_heads = reinterpret_cast<unsigned long long *>(w + 128);
        f.slow_list = reinterpret_cast<u32 *>(slow_list);
        f.slow_count = counter + 16;
        f.group_counters = counter + 48;
        f.lit_buckets = counter + 60;               // 65 words
        f.seq_buckets = counter + 128;              // FAST_ORDER_SUBS x 64 words (the header is 2 KB)
        f.kc_order = reinterpret_cast<u32 *>(slow_list) + wave_of(n);
        f.sm_count = sm_count;
        f.general_grid = dec_grid(m);
        f.bare_blocks = bare_mode ? 1u : 0u;
        f.unit_base = (uint32_t)w0;
        int k = 0;
        e = b200zstd::launch_decode_fast(f, stream, get_overlap(), &k);
        last_launches += k;
      }
    }
    if (e != cudaSuccess) return cuda_fail(e, fn);
    if (!sync) return Status::SUCCESS;
    // results back to the host: sizes (when the caller's table is host memory) and statuses
    std::vector<u32> st_local;
    std::vector<u32> &st = host_status ? *host_status : st_local;
    st.resize(n);
    if ((e = cudaMemcpyAsync(st.data(), d_status, n * sizeof(u32), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return cuda_fail(e, fn);
    if (!tables_on_device) {
      if ((e = cudaMemcpyAsync(out_sizes, d_out_sz, n * 8, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return cuda_fail(e, fn);
    } else if (host_out_sizes) {
      host_out_sizes->resize(n);
      if ((e = cudaMemcpyAsync(host_out_sizes->data(), d_out_sz, n * 8, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return cuda_fail(e, fn);
    }
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return cuda_fail(e, fn);
    Status overall = Status::SUCCESS;
    for (size_t i = 0; i < n; ++i)
      if (st[i] != 0) { overall = Status::ERROR_GENERIC; break; }
    if (overall != Status::SUCCESS) {
      Status first = Status::ERROR_GENERIC;
      for (size_t i = 0; i < n; ++i) if (st[i] != 0) { first = static_cast<Status>(st[i]); break; }
      fail(first, fn, "one or more chunks failed; see per-item status");
    }
    return overall;
  }
};

ZstdBatchManager::ZstdBatchManager() : pimpl_(new Impl) { pimpl_->cfg = CompressionConfig::get_default(); }
ZstdBatchManager::ZstdBatchManager(const CompressionConfig &c) : pimpl_(new Impl) {
  pimpl_->cfg = CompressionConfig::get_default();
  configure(c);
}
ZstdBatchManager::~ZstdBatchManager() = default;

Status ZstdBatchManager::configure(const CompressionConfig &c) {
  Status s = c.validate();
  if (s != Status::SUCCESS) return s;
  pimpl_->cfg = c;
  return Status::SUCCESS;
}
CompressionConfig ZstdBatchManager::get_config() const { return pimpl_->cfg; }
// single-buffer sizes include the tail that stages pageable host buffers (input; for compress also the worst-case output)
size_t ZstdBatchManager::get_compress_temp_size(size_t n) const {
  if (n == 0) return 0;
  const size_t core = n > Impl::BIG_BLOCK ? std::max(pimpl_->big_temp_any_level(n), pimpl_->enc_temp(1, &n)) : pimpl_->enc_temp(1, &n);
  return core + align_up(n, 256) + align_up(estimate_compressed_size(n, pimpl_->cfg.level), 256) + 256;
}
size_t ZstdBatchManager::get_decompress_temp_size(size_t n) const {
  if (n == 0) return 0;
  // the decompressed size is not known here: provision the block-parallel path for a ratio of 16 (at most 4096 blocks);
  // a frame that needs more decodes serially
  const size_t est = std::min<size_t>(4096, (16 * n + Impl::BIG_SMALL_BLOCK - 1) / Impl::BIG_SMALL_BLOCK);
  const size_t core = est >= 2 ? std::max(pimpl_->dec_temp(1, &n), pimpl_->big_dec_temp(n, est)) : pimpl_->dec_temp(1, &n);
  // staging of pageable host buffers: the input, and the output for ratios up to 16 (a larger pageable output needs a
  // larger workspace or a pinned / device destination)
  return core + align_up(n, 256) + align_up(std::min<size_t>(16 * n, (size_t)1 << 28), 256) + 256;
}
size_t ZstdBatchManager::get_max_compressed_size(size_t n) const { return estimate_compressed_size(n, pimpl_->cfg.level); }
size_t ZstdBatchManager::get_batch_compress_temp_size(const std::vector<size_t> &v) const { return pimpl_->enc_temp(v.size(), v.data()); }
size_t ZstdBatchManager::get_batch_decompress_temp_size(const std::vector<size_t> &v) const { return pimpl_->dec_temp(v.size(), v.data()); }
// Dictionaries (reference: set_dictionary src/cuda_zstd_manager.cu:3712-3760).  At the sizes of this path the reference
// compresses through its host route, plain ZSTD_compress without the dictionary (manager.cu:1604-1668), so a manager with a
// dictionary set writes ordinary frames there: no Dictionary_ID, decodable with or without the dictionary.  Same here: the
// dictionary is validated, kept (a copy of the raw content) and reported back by get_dictionary; the match stage does not
// reference it.  Frames that DO carry a Dictionary_ID are refused by the decoder (ERROR_DICTIONARY_MISMATCH).
Status ZstdBatchManager::set_dictionary(const dictionary::Dictionary &d) {
  if (!d.raw_content || d.raw_size == 0) return fail(Status::ERROR_INVALID_PARAMETER, "set_dictionary", "empty dictionary");
  if (d.raw_size > ((size_t)1 << 27)) return fail(Status::ERROR_INVALID_PARAMETER, "set_dictionary", "dictionary too large");
  std::lock_guard<std::mutex> lock(pimpl_->mu);
  cudaPointerAttributes at{};
  pimpl_->dict_bytes.resize(d.raw_size);
  const bool on_device = cudaPointerGetAttributes(&at, d.raw_content) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  if (!on_device) (void)cudaGetLastError();
  if (on_device) { if (cudaMemcpy(pimpl_->dict_bytes.data(), d.raw_content, d.raw_size, cudaMemcpyDeviceToHost) != cudaSuccess) return Status::ERROR_CUDA_ERROR; }
  else std::memcpy(pimpl_->dict_bytes.data(), d.raw_content, d.raw_size);
  pimpl_->dict_id = d.dict_id;
  return Status::SUCCESS;
}
Status ZstdBatchManager::get_dictionary(dictionary::Dictionary &d) const {
  if (pimpl_->dict_bytes.empty()) return Status::ERROR_INVALID_PARAMETER;
  d.raw_content = pimpl_->dict_bytes.data(); d.raw_size = pimpl_->dict_bytes.size(); d.dict_id = pimpl_->dict_id;
  return Status::SUCCESS;
}
Status ZstdBatchManager::clear_dictionary() { pimpl_->dict_bytes.clear(); pimpl_->dict_id = 0; return Status::SUCCESS; }
const CompressionStats &ZstdBatchManager::get_stats() const { return pimpl_->stats; }
Status ZstdBatchManager::set_compression_level(int level) {
  if (level < 1 || level > 22) return Status::ERROR_INVALID_PARAMETER;      // level left unchanged (manager.cu:1517-1522)
  ChecksumPolicy ck = pimpl_->cfg.checksum;
  u32 bs = pimpl_->cfg.block_size;
  pimpl_->cfg = CompressionConfig::from_level(level);
  pimpl_->cfg.checksum = ck;
  pimpl_->cfg.block_size = bs;
  return Status::SUCCESS;
}
int ZstdBatchManager::get_compression_level() const { return pimpl_->cfg.level; }
void ZstdBatchManager::reset_stats() { pimpl_->stats = CompressionStats(); }

namespace {
Status run_items(ZstdBatchManager::Impl &I, bool compress, const std::vector<BatchItem> &items, void *ws, size_t ws_bytes,
                 cudaStream_t stream) {
  const size_t n = items.size();
  if (n == 0) return Status::SUCCESS;
  std::lock_guard<std::mutex> lock(I.mu);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<const void *> in(n);
  std::vector<void *> out(n);
  std::vector<size_t> in_sz(n), out_sz(n);
  std::vector<u32> st;
  // null pointers and zero-size inputs are reported per item by the kernels as ERROR_INVALID_PARAMETER
  for (size_t i = 0; i < n; ++i) {
    in[i] = items[i].input_ptr; out[i] = items[i].output_ptr; in_sz[i] = items[i].input_size; out_sz[i] = items[i].output_size;
  }
  Status s = I.run(compress, in.data(), in_sz.data(), n, out.data(), out_sz.data(), nullptr, false, ws, ws_bytes, stream, true, &st, nullptr);
  if (st.size() == n) {
    BatchItem *mut = const_cast<BatchItem *>(items.data());      // the reference writes results the same way (manager.cu:5770-5795)
    u64 in_total = 0, out_total = 0;
    for (size_t i = 0; i < n; ++i) {
      mut[i].status = static_cast<Status>(st[i]);
      mut[i].output_size = st[i] == 0 ? out_sz[i] : 0;
      if (st[i] == 0) { in_total += in_sz[i]; out_total += out_sz[i]; }
    }
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (compress) {
      I.stats.input_bytes += in_total; I.stats.output_bytes += out_total; I.stats.bytes_compressed += in_total;
      I.stats.bytes_produced += out_total; I.stats.compression_time_ms += ms;
    } else {
      I.stats.bytes_decompressed += out_total; I.stats.decompression_time_ms += ms;
    }
    I.stats.blocks_processed += n; I.stats.num_blocks += n;
  }
  return s;
}
} // namespace

Status ZstdBatchManager::compress_batch(const std::vector<BatchItem> &items, void *ws, size_t ws_bytes, cudaStream_t stream) {
  return run_items(*pimpl_, true, items, ws, ws_bytes, stream);
}
Status ZstdBatchManager::decompress_batch(const std::vector<BatchItem> &items, void *ws, size_t ws_bytes, cudaStream_t stream) {
  return run_items(*pimpl_, false, items, ws, ws_bytes, stream);
}
Status ZstdBatchManager::decompress_batch_preallocated(std::vector<BatchItem> &items, void *ws, size_t ws_bytes, cudaStream_t stream) {
  return run_items(*pimpl_, false, items, ws, ws_bytes, stream);
}

// Single-buffer calls take device memory, pinned host memory (read/written in place over the bus) or, like the
// reference's CPU route (manager.cu:1604-1668; tests/test_two_phase_unit.cu:56 passes a std::vector), plain pageable
// host memory.  Pageable buffers are staged through the tail of the caller's workspace: the temp-size queries include
// room for the input and, for compress, the worst-case output.  Nothing is allocated here.
namespace {
enum class Mem { DEVICE, PINNED, PAGEABLE };
Mem mem_of(const void *p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return Mem::PAGEABLE; }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return Mem::DEVICE;
  return at.type == cudaMemoryTypeHost ? Mem::PINNED : Mem::PAGEABLE;
}
size_t stage_room(size_t bytes) { return align_up(bytes, 256); }
}  // namespace

static Status single_buffer(ZstdBatchManager::Impl &I, bool compress, const void *src, size_t n, void *dst, size_t *dst_size, void *ws,
                            size_t ws_bytes, cudaStream_t stream) {
  const char *fn = compress ? "compress" : "decompress";
  const bool stage_in = mem_of(src) == Mem::PAGEABLE, stage_out = mem_of(dst) == Mem::PAGEABLE;
  // a compressor never needs more than the worst-case size: capacities beyond it would only inflate the staging room
  const size_t cap = compress ? std::min(*dst_size, estimate_compressed_size(n, I.cfg.level)) : *dst_size;
  const size_t tail = (stage_in ? stage_room(n) : 0) + (stage_out ? stage_room(cap) : 0);
  if (tail > ws_bytes) return fail(Status::ERROR_BUFFER_TOO_SMALL, fn, "workspace has no room to stage pageable host buffers");
  const size_t body = (ws_bytes - tail) & ~(size_t)255;
  unsigned char *stage = static_cast<unsigned char *>(ws) + body;
  const void *d_src = src;
  void *d_dst = dst;
  cudaError_t e;
  if (stage_in) {
    if ((e = cudaMemcpyAsync(stage, src, n, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, fn);
    d_src = stage;
  }
  if (stage_out) d_dst = stage + (stage_in ? stage_room(n) : 0);
  std::vector<BatchItem> it(1);
  it[0].input_ptr = const_cast<void *>(d_src); it[0].input_size = n; it[0].output_ptr = d_dst; it[0].output_size = cap;
  Status s;
  if (compress && n > ZstdBatchManager::Impl::BIG_BLOCK) {
    // more than one block: the blocks are encoded side by side and assembled into one frame
    size_t out = cap;
    s = I.compress_big(d_src, n, d_dst, &out, ws, body, stream);
    if (s != Status::SUCCESS) return s;
    it[0].output_size = out;
  } else {
    s = Status::ERROR_NOT_IMPLEMENTED;
    if (!compress && cap > ZstdBatchManager::Impl::BIG_BLOCK) {
      // possibly a multi-block frame: try its blocks side by side; anything but success or a checksum verdict falls through
      size_t out = 0;
      s = I.decompress_big(d_src, n, d_dst, cap, &out, ws, body, stream);
      if (s == Status::SUCCESS) it[0].output_size = out;
      else if (s == Status::ERROR_CHECKSUM_FAILED) return s;
    }
    if (s != Status::SUCCESS) {
      s = run_items(I, compress, it, ws, body, stream);
      if (s != Status::SUCCESS) return it[0].status != Status::SUCCESS ? it[0].status : s;
    }
  }
  if (stage_out) {
    if ((e = cudaMemcpyAsync(dst, d_dst, it[0].output_size, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return cuda_fail(e, fn);
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return cuda_fail(e, fn);
  }
  *dst_size = it[0].output_size;
  return s;
}

Status ZstdBatchManager::compress(const void *src, size_t n, void *dst, size_t *dst_size, void *ws, size_t ws_bytes, const void *dict,
                                  size_t dict_size, cudaStream_t stream, void *) {
  if (!src || !dst || !dst_size || !ws) return fail(Status::ERROR_INVALID_PARAMETER, "compress", "null argument");   // manager.cu:1549-1552
  if (n == 0) return fail(Status::ERROR_INVALID_PARAMETER, "compress", "zero-size input");                            // manager.cu:1554-1558
  (void)dict; (void)dict_size;        // a per-call dictionary is accepted and not referenced (see set_dictionary)
  return single_buffer(*pimpl_, true, src, n, dst, dst_size, ws, ws_bytes, stream);
}
Status ZstdBatchManager::decompress(const void *src, size_t n, void *dst, size_t *dst_size, void *ws, size_t ws_bytes, cudaStream_t stream) {
  if (!src || !dst || !dst_size || !ws) return fail(Status::ERROR_INVALID_PARAMETER, "decompress", "null argument");
  if (n < 4) return fail(Status::ERROR_INVALID_PARAMETER, "decompress", "input shorter than a magic number");         // manager.cu:3202-3206
  if (*dst_size == 0) return fail(Status::ERROR_BUFFER_TOO_SMALL, "decompress", "zero output capacity");
  return single_buffer(*pimpl_, false, src, n, dst, dst_size, ws, ws_bytes, stream);
}
Status ZstdBatchManager::decompress_to_preallocated(const void *src, size_t n, void *out, size_t cap, size_t *actual, void *ws,
                                                    size_t ws_bytes, cudaStream_t stream) {
  if (!src || !out || !actual) return fail(Status::ERROR_INVALID_PARAMETER, "decompress_to_preallocated", "null argument");
  if (cap == 0) return fail(Status::ERROR_BUFFER_TOO_SMALL, "decompress_to_preallocated", "zero output capacity");
  size_t sz = cap;
  Status s = decompress(src, n, out, &sz, ws, ws_bytes, stream);
  if (s == Status::SUCCESS) *actual = sz;
  return s;
}
Status ZstdBatchManager::decompress_async_no_sync(const void *src, size_t n, void *out, size_t cap, size_t *d_actual, void *ws,
                                                  size_t ws_bytes, cudaStream_t stream) {
  if (!src || !out || !d_actual || !ws) return fail(Status::ERROR_INVALID_PARAMETER, "decompress_async_no_sync", "null argument");
  if (cap == 0) return fail(Status::ERROR_BUFFER_TOO_SMALL, "decompress_async_no_sync", "zero output capacity");
  Impl &I = *pimpl_;
  if (ws_bytes < I.dec_temp(1)) return fail(Status::ERROR_BUFFER_TOO_SMALL, "decompress_async_no_sync", "workspace too small");
  // tables for a batch of one are built in the workspace by stream-ordered copies of by-value words
  unsigned char *tab = static_cast<unsigned char *>(ws) + b200zstd::WS_HEADER_BYTES;
  const void *hin = src; size_t hn = n; void *hout = out; size_t hcap = cap;
  cudaError_t e;
  // pageable -> device async copies are staged by the runtime before returning, so the stack words are safe
  if ((e = cudaMemcpyAsync(tab, &hin, 8, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, "decompress_async_no_sync");
  if ((e = cudaMemcpyAsync(tab + 8, &hn, 8, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, "decompress_async_no_sync");
  if ((e = cudaMemcpyAsync(tab + 16, &hout, 8, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, "decompress_async_no_sync");
  if ((e = cudaMemcpyAsync(d_actual, &hcap, 8, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, "decompress_async_no_sync");
  b200zstd::DecodeArgs a{};
  a.in_ptrs = reinterpret_cast<const void *const *>(tab); a.in_sizes = reinterpret_cast<const size_t *>(tab + 8);
  a.out_ptrs = reinterpret_cast<void *const *>(tab + 16); a.out_sizes = d_actual; a.statuses = reinterpret_cast<u32 *>(tab + 32);
  a.counter = reinterpret_cast<u32 *>(ws); a.lit_scratch = tab + Impl::table_bytes(1); a.n = 1;
  a.verify_checksum = I.cfg.checksum == ChecksumPolicy::COMPUTE_AND_VERIFY;
  e = b200zstd::launch_decode_batch(a, 1, stream);
  return e == cudaSuccess ? Status::SUCCESS : cuda_fail(e, "decompress_async_no_sync");
}
// Compress counterpart of decompress_async_no_sync (additive; the reference has none): device buffers only, nothing is
// synchronised; result16 receives {frame bytes, status} as two 64-bit words once `stream` reaches that point (pinned
// host memory or device memory).  Inputs of at most one block go through the batch encoder as a batch of one.
Status ZstdBatchManager::compress_async_no_sync(const void *src, size_t n, void *dst, size_t cap, unsigned long long *result16, void *ws,
                                                size_t ws_bytes, cudaStream_t stream) {
  const char *fn = "compress_async_no_sync";
  if (!src || !dst || !result16 || !ws) return fail(Status::ERROR_INVALID_PARAMETER, fn, "null argument");
  if (n == 0) return fail(Status::ERROR_INVALID_PARAMETER, fn, "zero-size input");
  Impl &I = *pimpl_;
  if (n > Impl::BIG_BLOCK) return I.compress_big_enqueue(src, n, dst, cap, ws, ws_bytes, stream, reinterpret_cast<u64 *>(result16));
  const uint32_t bm = Impl::esd_block_max(n);
  if (ws_bytes < I.enc_need(1, bm)) return fail(Status::ERROR_BUFFER_TOO_SMALL, fn, "workspace too small");
  std::lock_guard<std::mutex> lock(I.mu);
  unsigned char *w = static_cast<unsigned char *>(ws);
  unsigned char *tab = w + b200zstd::WS_HEADER_BYTES;
  u64 *d_result = reinterpret_cast<u64 *>(w + 64);
  const u64 words[4] = {(u64)(uintptr_t)src, (u64)n, (u64)(uintptr_t)dst, (u64)cap};
  cudaError_t e;
  if ((e = cudaMemcpyAsync(tab, words, 32, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return cuda_fail(e, fn);
  b200zstd::EncodeArgs a{};
  a.in_ptrs = reinterpret_cast<const void *const *>(tab); a.in_sizes = reinterpret_cast<const size_t *>(tab + 8);
  a.out_ptrs = reinterpret_cast<void *const *>(tab + 16); a.out_sizes = reinterpret_cast<size_t *>(tab + 24);
  a.statuses = reinterpret_cast<u32 *>(tab + 32); a.counter = reinterpret_cast<u32 *>(w);
  a.n = 1; a.prm = I.enc_params();
  int enc_launches = 0;
  unsigned char *lists = tab + Impl::table_bytes(1);
  if ((e = I.enqueue_encode(a, w, lists, lists + I.lists_bytes(1), I.enc_scratch_bytes(1, bm), bm, n, n, stream, &enc_launches)) != cudaSuccess) return cuda_fail(e, fn);
  if ((e = b200zstd::launch_big_result(a.statuses, 1, reinterpret_cast<const uint64_t *>(tab + 24), 0, d_result, stream)) != cudaSuccess) return cuda_fail(e, fn);
  if ((e = cudaMemcpyAsync(result16, d_result, 16, cudaMemcpyDefault, stream)) != cudaSuccess) return cuda_fail(e, fn);
  I.last_launches = 1 + enc_launches;
  return Status::SUCCESS;
}
// both sizes are known here, so the block-parallel path for multi-block frames is provisioned exactly
size_t ZstdBatchManager::get_inference_workspace_size(size_t mc, size_t mo) const {
  const size_t B = (mo + Impl::BIG_SMALL_BLOCK - 1) / Impl::BIG_SMALL_BLOCK;
  const size_t core = B >= 2 ? std::max(pimpl_->dec_temp(1, &mc), pimpl_->big_dec_temp(mc, B)) : pimpl_->dec_temp(1, &mc);
  return core + align_up(mc, 256) + 256;
}
Status ZstdBatchManager::allocate_inference_workspace(size_t mc, size_t mo, void **p, size_t *sz) {
  if (!p || !sz) return Status::ERROR_INVALID_PARAMETER;
  *sz = get_inference_workspace_size(mc, mo);
  cudaError_t e = cudaMalloc(p, *sz);
  return e == cudaSuccess ? Status::SUCCESS : fail(Status::ERROR_OUT_OF_MEMORY, "allocate_inference_workspace", cudaGetErrorString(e), e);
}
Status ZstdBatchManager::free_inference_workspace(void *p) {
  if (!p) return Status::ERROR_INVALID_PARAMETER;
  return cudaFree(p) == cudaSuccess ? Status::SUCCESS : Status::ERROR_CUDA_ERROR;
}

std::unique_ptr<ZstdManager> create_manager(int level) {
  auto m = std::make_unique<ZstdBatchManager>();
  m->set_compression_level(level);
  return m;
}
std::unique_ptr<ZstdManager> create_manager(const CompressionConfig &c) { return std::make_unique<ZstdBatchManager>(c); }
std::unique_ptr<ZstdBatchManager> create_batch_manager(int level) {
  auto m = std::make_unique<ZstdBatchManager>();
  m->set_compression_level(level);
  return m;
}

namespace {
Status simple(bool compress, const void *src, size_t n, void *dst, size_t *dst_size, int level, cudaStream_t stream) {
  ZstdBatchManager m;
  if (compress) m.set_compression_level(level);
  size_t ws_bytes = compress ? m.get_compress_temp_size(n) : m.get_decompress_temp_size(n);
  void *ws = nullptr;
  if (cudaMalloc(&ws, ws_bytes) != cudaSuccess) return Status::ERROR_OUT_OF_MEMORY;
  Status s = compress ? m.compress(src, n, dst, dst_size, ws, ws_bytes, nullptr, 0, stream) : m.decompress(src, n, dst, dst_size, ws, ws_bytes, stream);
  cudaFree(ws);
  return s;
}
} // namespace
Status compress_simple(const void *src, size_t n, void *dst, size_t *dst_size, int level, cudaStream_t stream) {
  return simple(true, src, n, dst, dst_size, level, stream);
}
Status decompress_simple(const void *src, size_t n, void *dst, size_t *dst_size, cudaStream_t stream) {
  return simple(false, src, n, dst, dst_size, 3, stream);
}

// ============================================================================================
// nvCOMP-v5 facade
// ============================================================================================
namespace nvcomp_v5 {

bool is_compatible_with_nvcomp_v5(u32 v) { return v == get_nvcomp_v5_format_version(); }   // reference nvcomp.cpp:153-155
// a Zstandard or skippable frame magic at the front; the data may be device or host memory (reference nvcomp.cpp:129-151)
bool is_nvcomp_v5_zstd_format(const void *data, size_t size) {
  if (!data || size < 4) return false;
  u32 magic = 0;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, data) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
    if (cudaMemcpy(&magic, data, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  } else { (void)cudaGetLastError(); std::memcpy(&magic, data, 4); }
  return magic == ZSTD_MAGIC || (magic & 0xFFFFFFF0u) == 0x184D2A50u;
}
NvcompV5Options to_nvcomp_v5_opts(const CompressionConfig &c) {
  NvcompV5Options o;
  o.level = c.level; o.chunk_size = c.block_size; o.enable_checksum = c.checksum != ChecksumPolicy::NO_COMPUTE_NO_VERIFY;
  return o;
}
CompressionConfig from_nvcomp_v5_opts(const NvcompV5Options &o) {
  CompressionConfig c = CompressionConfig::from_level(o.level);
  c.block_size = o.chunk_size;
  c.checksum = o.enable_checksum ? ChecksumPolicy::COMPUTE_AND_VERIFY : ChecksumPolicy::NO_COMPUTE_NO_VERIFY;
  return c;
}
std::unique_ptr<ZstdManager> create_nvcomp_v5_manager(const NvcompV5Options &o) { return create_manager(from_nvcomp_v5_opts(o)); }

int status_to_nvcomp_error(Status s) {
  switch (static_cast<u32>(s)) {
  case 0: case 2: case 3: case 4: case 6: case 7: case 10: case 12: return static_cast<int>(s);
  default: return 1;
  }
}
Status nvcomp_error_to_status(int e) {
  switch (e) {
  case 0: case 2: case 3: case 4: case 6: case 7: case 10: case 12: return static_cast<Status>(e);
  default: return Status::ERROR_GENERIC;
  }
}
const char *get_nvcomp_v5_error_string(int e) { return status_to_string(nvcomp_error_to_status(e)); }

class NvcompV5BatchManager::Impl {
public:
  ZstdBatchManager mgr;
  explicit Impl(const NvcompV5Options &o) : mgr(sanitize(o)) {}
  static CompressionConfig sanitize(const NvcompV5Options &o) {
    NvcompV5Options t = o;
    if (t.level < 1 || t.level > 22) t.level = 3;
    if (t.chunk_size < 1024) t.chunk_size = 64 * 1024;
    return from_nvcomp_v5_opts(t);
  }
};

namespace {
bool on_device(const void *p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
// Tables may be host or device memory, in any mix (reference src/cuda_zstd_nvcomp.cpp:319-437).
// All-device -> used in place; otherwise everything is brought to host vectors and staged once.
Status run_tables(ZstdBatchManager &m, bool compress, const void *const *in_ptrs, const size_t *in_sizes, size_t n,
                  void *const *out_ptrs, size_t *out_sizes, void *ws, size_t ws_bytes, cudaStream_t stream) {
  if (n == 0) return Status::SUCCESS;
  if (!in_ptrs || !in_sizes || !out_ptrs || !out_sizes) return fail(Status::ERROR_INVALID_PARAMETER, "batch", "null pointer table");
  ZstdBatchManager::Impl &I = *m.impl();
  std::lock_guard<std::mutex> lock(I.mu);
  const bool d0 = on_device(in_ptrs), d1 = on_device(in_sizes), d2 = on_device(out_ptrs), d3 = on_device(out_sizes);
  std::vector<u32> st;
  std::vector<size_t> hs;
  Status s;
  if (d0 && d1 && d2 && d3) {
    s = I.run(compress, in_ptrs, in_sizes, n, out_ptrs, out_sizes, nullptr, true, ws, ws_bytes, stream, true, &st, &hs);
  } else {
    std::vector<const void *> in(n);
    std::vector<void *> out(n);
    std::vector<size_t> isz(n), osz(n);
    cudaError_t e = cudaSuccess;
    if (d0 || d1 || d2 || d3) if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return cuda_fail(e, "batch");
    auto fetch = [&](void *dst, const void *src, bool dev) {
      if (dev) { cudaError_t r = cudaMemcpy(dst, src, n * 8, cudaMemcpyDeviceToHost); if (r != cudaSuccess) e = r; }
      else std::memcpy(dst, src, n * 8);
    };
    fetch(in.data(), in_ptrs, d0); fetch(isz.data(), in_sizes, d1); fetch(out.data(), out_ptrs, d2); fetch(osz.data(), out_sizes, d3);
    if (e != cudaSuccess) return cuda_fail(e, "batch");
    s = I.run(compress, in.data(), isz.data(), n, out.data(), osz.data(), nullptr, false, ws, ws_bytes, stream, true, &st, nullptr);
    if (st.size() == n) {
      for (size_t i = 0; i < n; ++i) if (st[i] != 0) osz[i] = 0;
      if (d3) { if ((e = cudaMemcpy(out_sizes, osz.data(), n * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "batch"); }
      else std::memcpy(out_sizes, osz.data(), n * 8);
      hs = osz;
    }
    if (st.size() == n) {
      u64 it = 0, ot = 0;
      for (size_t i = 0; i < n; ++i) if (st[i] == 0) { it += isz[i]; ot += osz[i]; }
      if (compress) { I.stats.input_bytes += it; I.stats.output_bytes += ot; I.stats.bytes_compressed += it; I.stats.bytes_produced += ot; }
      else I.stats.bytes_decompressed += ot;
      I.stats.blocks_processed += n;
    }
  }
  return s;
}
} // namespace

NvcompV5BatchManager::NvcompV5BatchManager(const NvcompV5Options &o) : pimpl_(new Impl(o)) {}
NvcompV5BatchManager::~NvcompV5BatchManager() = default;
ZstdBatchManager &NvcompV5BatchManager::batch_manager() { return pimpl_->mgr; }
size_t NvcompV5BatchManager::get_compress_temp_size(const size_t *s, size_t n, cudaStream_t) const { return pimpl_->mgr.impl()->enc_temp(n, s); }
size_t NvcompV5BatchManager::get_decompress_temp_size(const size_t *s, size_t n, cudaStream_t) const { return pimpl_->mgr.impl()->dec_temp(n, s); }
size_t NvcompV5BatchManager::get_max_compressed_chunk_size(size_t n) const { return pimpl_->mgr.get_max_compressed_size(n); }
Status NvcompV5BatchManager::compress_async(const void *const *in, const size_t *in_sz, size_t n, void *const *out, size_t *out_sz,
                                            void *tmp, size_t tmp_bytes, cudaStream_t stream) {
  return run_tables(pimpl_->mgr, true, in, in_sz, n, out, out_sz, tmp, tmp_bytes, stream);
}
Status NvcompV5BatchManager::decompress_async(const void *const *in, const size_t *in_sz, size_t n, void *const *out, size_t *out_sz,
                                              void *tmp, size_t tmp_bytes, cudaStream_t stream) {
  return run_tables(pimpl_->mgr, false, in, in_sz, n, out, out_sz, tmp, tmp_bytes, stream);
}
const CompressionStats &NvcompV5BatchManager::get_stats() const { return pimpl_->mgr.get_stats(); }

Status get_metadata_async(const void *d, size_t n, NvcompV5Metadata *m, cudaStream_t stream) {
  if (!d || !m) return Status::ERROR_INVALID_PARAMETER;
  if (n < 5) return Status::ERROR_INVALID_PARAMETER;
  unsigned char head[18];
  size_t k = std::min<size_t>(n, sizeof head);
  if (on_device(d)) {
    if (cudaMemcpyAsync(head, d, k, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return Status::ERROR_CUDA_ERROR;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return Status::ERROR_CUDA_ERROR;
  } else std::memcpy(head, d, k);
  HeaderInfo h = peek_header(head, k);
  if (!h.ok) return Status::ERROR_INVALID_MAGIC;
  *m = NvcompV5Metadata();
  m->uncompressed_size = h.has_size ? h.content_size : 0;
  m->compressed_size = n;
  m->num_chunks = 1;
  m->chunk_size = (u32)std::min<u64>(m->uncompressed_size, 0xFFFFFFFFu);
  m->dictionary_id = h.dict_id;
  m->has_dictionary = h.dict_id != 0;
  m->checksum_policy = h.checksum ? ChecksumPolicy::COMPUTE_AND_VERIFY : ChecksumPolicy::NO_COMPUTE_NO_VERIFY;
  return Status::SUCCESS;
}
Status get_metadata(const void *d, size_t n, NvcompV5Metadata &m) { return get_metadata_async(d, n, &m, 0); }
bool validate_metadata(const NvcompV5Metadata &m) { return is_compatible_with_nvcomp_v5(m.format_version) && m.num_chunks >= 1; }
Status get_decompressed_size_async(const void *d, size_t n, size_t *out, cudaStream_t stream) {
  NvcompV5Metadata m;
  if (!out) return Status::ERROR_INVALID_PARAMETER;
  Status s = get_metadata_async(d, n, &m, stream);
  if (s == Status::SUCCESS) *out = (size_t)m.uncompressed_size;
  return s;
}
Status get_num_chunks(const void *d, size_t n, size_t *out) {
  if (!d || !out || n < 5) return Status::ERROR_INVALID_PARAMETER;
  *out = 1;
  return Status::SUCCESS;
}
Status get_chunk_sizes(const void *d, size_t n, size_t *sizes, size_t max_chunks) {
  if (!sizes || max_chunks < 1) return Status::ERROR_INVALID_PARAMETER;
  return get_decompressed_size_async(d, n, &sizes[0], 0);
}

NvcompV5BenchmarkResult benchmark_level(const void *d_input, size_t n, int level, int iterations, cudaStream_t stream) {
  NvcompV5BenchmarkResult r{};
  r.level = level;
  if (!d_input || n == 0 || iterations < 1) return r;
  ZstdBatchManager m;
  m.set_compression_level(level);
  const size_t cap = m.get_max_compressed_size(n), ws_bytes = m.get_compress_temp_size(n);
  void *ws = nullptr, *comp = nullptr, *back = nullptr;
  if (cudaMalloc(&ws, ws_bytes) != cudaSuccess || cudaMalloc(&comp, cap) != cudaSuccess || cudaMalloc(&back, n) != cudaSuccess) {
    cudaFree(ws); cudaFree(comp); cudaFree(back);
    return r;
  }
  size_t csz = cap, dsz = n;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iterations; ++i) { csz = cap; if (m.compress(d_input, n, comp, &csz, ws, ws_bytes, nullptr, 0, stream) != Status::SUCCESS) csz = 0; }
  auto t1 = std::chrono::steady_clock::now();
  for (int i = 0; i < iterations && csz; ++i) { dsz = n; m.decompress(comp, csz, back, &dsz, ws, ws_bytes, stream); }
  auto t2 = std::chrono::steady_clock::now();
  r.compress_time_ms = std::chrono::duration<double, std::milli>(t1 - t0).count() / iterations;
  r.decompress_time_ms = std::chrono::duration<double, std::milli>(t2 - t1).count() / iterations;
  r.compress_throughput_mbps = r.compress_time_ms > 0 ? (n / 1e6) / (r.compress_time_ms / 1e3) : 0;
  r.decompress_throughput_mbps = r.decompress_time_ms > 0 ? (n / 1e6) / (r.decompress_time_ms / 1e3) : 0;
  r.compressed_size = csz;
  r.compression_ratio = csz ? (float)n / (float)csz : 0.f;
  cudaFree(ws); cudaFree(comp); cudaFree(back);
  return r;
}
std::vector<NvcompV5BenchmarkResult> benchmark_all_levels(const void *d_input, size_t n, int iterations, cudaStream_t stream) {
  std::vector<NvcompV5BenchmarkResult> v;
  for (int l = 1; l <= 22; ++l) v.push_back(benchmark_level(d_input, n, l, iterations, stream));
  return v;
}

} // namespace nvcomp_v5
} // namespace cuda_zstd

// ================================================================================================
// extern "C"
// ================================================================================================
using cuda_zstd::Status;
using cuda_zstd::ZstdBatchManager;
using cuda_zstd::nvcomp_v5::status_to_nvcomp_error;
using cuda_zstd::u32;
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct cuda_zstd_dict_t { std::vector<unsigned char> bytes; uint32_t id = 0; };

// streams and events of the host-resident batch calls (cuda_zstd_batch_*_host*), created on first use
struct HostPipe {
  static constexpr int MAX_WAVES = 8;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_entry = nullptr, ev_in[MAX_WAVES] = {}, ev_run[MAX_WAVES] = {}, ev_done = nullptr;
  unsigned long long *h_totals = nullptr;      // pinned: the packed size after every wave (compress_host_packed)
  bool ready = false;
  bool init() {
    if (ready) return true;
    bool ok = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ev_entry, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < MAX_WAVES && ok; i++)
      ok = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ev_run[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaHostAlloc(reinterpret_cast<void **>(&h_totals), MAX_WAVES * sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
    if (!ok) { (void)cudaGetLastError(); return false; }
    return ready = true;
  }
  ~HostPipe() {
    if (!ready) return;
    for (int i = 0; i < MAX_WAVES; i++) { cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_run[i]); }
    cudaEventDestroy(ev_entry); cudaEventDestroy(ev_done);
    cudaStreamDestroy(s_in); cudaStreamDestroy(s_out);
    cudaFreeHost(h_totals);
  }
};
struct cuda_zstd_batch { cuda_zstd::nvcomp_v5::NvcompV5BatchManager *m; HostPipe pipe; };

namespace {
// Host buffers of a batch usually sit back to back in a few large allocations: a RUN is a maximal stretch of items whose
// host addresses ascend with gaps of at most max_gap bytes, staged by ONE copy; the device image keeps every item's
// address modulo 256.  Items that do not line up simply form runs of their own.  Inputs may bridge small gaps (the gap
// bytes are only read); outputs must be exactly adjacent, because a copy home writes every byte of its run.
struct HostRun { size_t first, count; const unsigned char *h_begin; size_t bytes; size_t d_off; };
size_t plan_runs(const void *const *ptrs, const size_t *sizes, size_t lo, size_t hi, size_t d_off, std::vector<HostRun> &runs,
                 std::vector<size_t> &item_off, size_t max_gap) {
  size_t i = lo;
  while (i < hi) {
    const unsigned char *b = static_cast<const unsigned char *>(ptrs[i]);
    const unsigned char *e = b + sizes[i];
    size_t j = i + 1;
    while (j < hi) {
      const unsigned char *q = static_cast<const unsigned char *>(ptrs[j]);
      if (q < e || (size_t)(q - e) > max_gap) break;
      e = q + sizes[j];
      j++;
    }
    d_off = align_up(d_off, 256) + ((uintptr_t)b & 255);
    runs.push_back(HostRun{i, j - i, b, (size_t)(e - b), d_off});
    for (size_t k = i; k < j; k++) item_off[k] = d_off + (size_t)(static_cast<const unsigned char *>(ptrs[k]) - b);
    d_off += (size_t)(e - b);
    i = j;
  }
  return align_up(d_off, 256);
}
// chunk ranges of the pipeline waves.  Decompress: 1 : 2 : 4 : 9 sixteenths, so that the first results leave early (the
// device-to-host copy of the output is the long pole).  Compress: the host-to-device copy of the input is, and the encoder
// runs about as fast as the link, so the call ends one wave's compress time after the last input byte has arrived: equal
// waves of 3,072 chunks or more (measured on 16,384 x 64 KiB at level 3, GB/s host to host: 1:2:4:9 31.1; equal waves: 2 33.3,
// 3 34.5, 4 38.7, 5 39.1, 6 38.4, 8 35.9 -- small waves leave the encoder's kernels under-filled; CUDA_ZSTD_HOST_WAVES overrides).
int plan_waves(size_t n, size_t edges[HostPipe::MAX_WAVES + 1], bool compress = false) {
  if (n < 1024) { edges[0] = 0; edges[1] = n; return 1; }
  if (compress) {
    static const int env = getenv("CUDA_ZSTD_HOST_WAVES") ? atoi(getenv("CUDA_ZSTD_HOST_WAVES")) : 0;
    int k = env > 0 ? env : (int)std::min<size_t>(HostPipe::MAX_WAVES, n / 3072);
    k = std::max(1, std::min(k, HostPipe::MAX_WAVES));
    for (int w = 0; w <= k; w++) edges[w] = n * (size_t)w / (size_t)k;
    return k;
  }
  edges[0] = 0; edges[1] = n / 16; edges[2] = 3 * n / 16; edges[3] = 7 * n / 16; edges[4] = n;
  return 4;
}
size_t staged_bytes_bound(const size_t *sizes, size_t n) {          // worst case of plan_runs: every item a run of its own
  size_t t = 0;
  for (size_t i = 0; i < n; i++) t += align_up(sizes[i], 256) + 512;
  return t + 1024;
}
} // namespace

extern "C" {

// ---- nvcomp_zstd_*_v5 (reference src/cuda_zstd_nvcomp.cpp:766-840) ----
nvcompZstdManagerHandle nvcomp_zstd_create_manager_v5(int level) {
  try {
    auto m = cuda_zstd::create_batch_manager((level < 1 || level > 22) ? 3 : level);
    return static_cast<cuda_zstd::ZstdManager *>(m.release());
  } catch (...) { return nullptr; }
}
void nvcomp_zstd_destroy_manager_v5(nvcompZstdManagerHandle h) { delete static_cast<cuda_zstd::ZstdManager *>(h); }
int nvcomp_zstd_compress_async_v5(nvcompZstdManagerHandle h, const void *src, size_t n, void *dst, size_t *dst_size, void *tmp,
                                  size_t tmp_bytes, cudaStream_t stream) {
  if (!h) return status_to_nvcomp_error(Status::ERROR_INVALID_PARAMETER);
  try { return status_to_nvcomp_error(static_cast<cuda_zstd::ZstdManager *>(h)->compress(src, n, dst, dst_size, tmp, tmp_bytes, nullptr, 0, stream)); }
  catch (...) { return 1; }
}
int nvcomp_zstd_decompress_async_v5(nvcompZstdManagerHandle h, const void *src, size_t n, void *dst, size_t *dst_size, void *tmp,
                                    size_t tmp_bytes, cudaStream_t stream) {
  if (!h) return status_to_nvcomp_error(Status::ERROR_INVALID_PARAMETER);
  try { return status_to_nvcomp_error(static_cast<cuda_zstd::ZstdManager *>(h)->decompress(src, n, dst, dst_size, tmp, tmp_bytes, stream)); }
  catch (...) { return 1; }
}
size_t nvcomp_zstd_get_compress_temp_size_v5(nvcompZstdManagerHandle h, size_t n) {
  return h ? static_cast<cuda_zstd::ZstdManager *>(h)->get_compress_temp_size(n) : 0;
}
size_t nvcomp_zstd_get_decompress_temp_size_v5(nvcompZstdManagerHandle h, size_t n) {
  return h ? static_cast<cuda_zstd::ZstdManager *>(h)->get_decompress_temp_size(n) : 0;
}
int nvcomp_zstd_get_metadata_v5(const void *d, size_t n, cuda_zstd::nvcomp_v5::NvcompV5Metadata *m, cudaStream_t stream) {
  return status_to_nvcomp_error(cuda_zstd::nvcomp_v5::get_metadata_async(d, n, m, stream));
}

// ---- cuda_zstd_* single-buffer C API (reference src/cuda_zstd_c_api.cpp:18-211) ----
cuda_zstd_manager_t *cuda_zstd_create_manager(int level) {
  return reinterpret_cast<cuda_zstd_manager_t *>(nvcomp_zstd_create_manager_v5(level));
}
void cuda_zstd_destroy_manager(cuda_zstd_manager_t *m) { nvcomp_zstd_destroy_manager_v5(m); }
int cuda_zstd_compress(cuda_zstd_manager_t *m, const void *src, size_t n, void *dst, size_t *dst_size, void *ws, size_t ws_bytes,
                       cudaStream_t stream) {
  if (!m) return static_cast<int>(Status::ERROR_INVALID_PARAMETER);
  try { return static_cast<int>(reinterpret_cast<cuda_zstd::ZstdManager *>(m)->compress(src, n, dst, dst_size, ws, ws_bytes, nullptr, 0, stream)); }
  catch (...) { return 1; }
}
int cuda_zstd_decompress(cuda_zstd_manager_t *m, const void *src, size_t n, void *dst, size_t *dst_size, void *ws, size_t ws_bytes,
                         cudaStream_t stream) {
  if (!m) return static_cast<int>(Status::ERROR_INVALID_PARAMETER);
  try { return static_cast<int>(reinterpret_cast<cuda_zstd::ZstdManager *>(m)->decompress(src, n, dst, dst_size, ws, ws_bytes, stream)); }
  catch (...) { return 1; }
}
size_t cuda_zstd_get_compress_workspace_size(cuda_zstd_manager_t *m, size_t n) { return nvcomp_zstd_get_compress_temp_size_v5(m, n); }
size_t cuda_zstd_get_decompress_workspace_size(cuda_zstd_manager_t *m, size_t n) { return nvcomp_zstd_get_decompress_temp_size_v5(m, n); }
// cuda_zstd_train_dictionary (src/cuda_zstd_c_api.cpp:128-177; the reference runs COVER): a RAW-CONTENT dictionary made of
// the last dict_size bytes of the samples laid end to end (the format gives the END of a raw dictionary the cheapest
// offsets), with an id derived from the content.  Host memory only.
cuda_zstd_dict_t *cuda_zstd_train_dictionary(const void **samples, const size_t *sizes, size_t n, size_t dict_size) {
  if (!samples || !sizes || n == 0 || dict_size == 0) return nullptr;
  try {
    size_t total = 0;
    for (size_t i = 0; i < n; i++) { if (!samples[i] && sizes[i]) return nullptr; total += sizes[i]; }
    if (total == 0) return nullptr;
    auto *d = new cuda_zstd_dict_t;
    const size_t keep = std::min(total, dict_size);
    d->bytes.resize(keep);
    size_t skip = total - keep, w = 0;
    for (size_t i = 0; i < n; i++) {
      const unsigned char *p = static_cast<const unsigned char *>(samples[i]);
      size_t len = sizes[i];
      if (skip >= len) { skip -= len; continue; }
      std::memcpy(d->bytes.data() + w, p + skip, len - skip);
      w += len - skip; skip = 0;
    }
    uint32_t h = 2166136261u;                                  // FNV-1a of the content: ids below 32768 are reserved by the format
    for (unsigned char c : d->bytes) h = (h ^ c) * 16777619u;
    d->id = h | 0x8000u;
    return d;
  } catch (...) { return nullptr; }
}
void cuda_zstd_destroy_dictionary(cuda_zstd_dict_t *d) { delete d; }
int cuda_zstd_set_dictionary(cuda_zstd_manager_t *m, cuda_zstd_dict_t *d) {
  if (!m || !d) return static_cast<int>(Status::ERROR_INVALID_PARAMETER);
  try {
    cuda_zstd::dictionary::Dictionary x;
    x.raw_content = d->bytes.data(); x.raw_size = d->bytes.size(); x.dict_id = d->id;
    return static_cast<int>(reinterpret_cast<cuda_zstd::ZstdManager *>(m)->set_dictionary(x));
  } catch (...) { return 1; }
}
const char *cuda_zstd_get_error_string(int code) { return cuda_zstd::status_to_string(static_cast<Status>(code)); }
int cuda_zstd_is_error(int code) { return code != 0; }

// ---- cuda_zstd_batch_* (additive batch ABI, include/cuda_zstd_batch_c.h) ----
cuda_zstd_batch_t *cuda_zstd_batch_create(int level, int enable_checksum) {
  try {
    cuda_zstd::nvcomp_v5::NvcompV5Options o;
    o.level = level; o.enable_checksum = enable_checksum != 0; o.chunk_size = 128 * 1024;
    auto *b = new cuda_zstd_batch;
    b->m = new cuda_zstd::nvcomp_v5::NvcompV5BatchManager(o);
    return b;
  } catch (...) { return nullptr; }
}
void cuda_zstd_batch_destroy(cuda_zstd_batch_t *b) { if (b) { delete b->m; delete b; } }
size_t cuda_zstd_batch_get_max_compressed_size(cuda_zstd_batch_t *b, size_t n) { return b ? b->m->get_max_compressed_chunk_size(n) : 0; }
size_t cuda_zstd_batch_get_compress_temp_size(cuda_zstd_batch_t *b, const size_t *sizes, size_t n) { return b ? b->m->get_compress_temp_size(sizes, n) : 0; }
size_t cuda_zstd_batch_get_decompress_temp_size(cuda_zstd_batch_t *b, const size_t *sizes, size_t n) { return b ? b->m->get_decompress_temp_size(sizes, n) : 0; }
int cuda_zstd_batch_compress(cuda_zstd_batch_t *b, const void *const *in, const size_t *in_sz, size_t n, void *const *out, size_t *out_sz,
                             void *tmp, size_t tmp_bytes, cudaStream_t stream) {
  if (!b) return 2;
  try { return status_to_nvcomp_error(b->m->compress_async(in, in_sz, n, out, out_sz, tmp, tmp_bytes, stream)); } catch (...) { return 1; }
}
int cuda_zstd_batch_decompress(cuda_zstd_batch_t *b, const void *const *in, const size_t *in_sz, size_t n, void *const *out, size_t *out_sz,
                               void *tmp, size_t tmp_bytes, cudaStream_t stream) {
  if (!b) return 2;
  try { return status_to_nvcomp_error(b->m->decompress_async(in, in_sz, n, out, out_sz, tmp, tmp_bytes, stream)); } catch (...) { return 1; }
}
int cuda_zstd_batch_compress_nosync(cuda_zstd_batch_t *b, const void *const *in, const size_t *in_sz, size_t n, void *const *out,
                                    size_t *out_sz, uint32_t *st, void *tmp, size_t tmp_bytes, cudaStream_t stream) {
  if (!b) return 2;
  try {
    auto *I = b->m->batch_manager().impl();
    std::lock_guard<std::mutex> lock(I->mu);
    return status_to_nvcomp_error(I->run(true, in, in_sz, n, out, out_sz, st, true, tmp, tmp_bytes, stream, false, nullptr, nullptr));
  } catch (...) { return 1; }
}
int cuda_zstd_batch_decompress_nosync(cuda_zstd_batch_t *b, const void *const *in, const size_t *in_sz, size_t n, void *const *out,
                                      size_t *out_sz, uint32_t *st, void *tmp, size_t tmp_bytes, cudaStream_t stream) {
  if (!b) return 2;
  try {
    auto *I = b->m->batch_manager().impl();
    std::lock_guard<std::mutex> lock(I->mu);
    return status_to_nvcomp_error(I->run(false, in, in_sz, n, out, out_sz, st, true, tmp, tmp_bytes, stream, false, nullptr, nullptr));
  } catch (...) { return 1; }
}
// ---- host-resident batches (include/cuda_zstd_batch_c.h): payloads in HOST memory, staged in waves over two copy
// streams around the caller's stream; the codec calls are the device-table no-sync form ----
size_t cuda_zstd_batch_get_host_decompress_temp_size(cuda_zstd_batch_t *b, const size_t *comp_sizes, const size_t *caps, size_t n) {
  if (!b || !comp_sizes || !caps || n == 0) return 0;
  return staged_bytes_bound(comp_sizes, n) + staged_bytes_bound(caps, n) + align_up(n * 40, 256) + 256 +
         b->m->get_decompress_temp_size(comp_sizes, n);
}
size_t cuda_zstd_batch_get_host_compress_temp_size(cuda_zstd_batch_t *b, const size_t *sizes, size_t n) {
  if (!b || !sizes || n == 0) return 0;
  size_t frames = 0;
  for (size_t i = 0; i < n; i++) frames += align_up(b->m->get_max_compressed_chunk_size(sizes[i]), 16);
  return staged_bytes_bound(sizes, n) + 2 * align_up(frames, 256) + align_up(n * 40 + (n + 1) * 8, 256) + 256 +
         b->m->get_compress_temp_size(sizes, n);
}
int cuda_zstd_batch_decompress_host(cuda_zstd_batch_t *b, const void *const *h_in, const size_t *in_sz, size_t n, void *const *h_out,
                                    size_t *out_sz, uint32_t *h_status, void *tmp, size_t tmp_bytes, cudaStream_t stream) {
  if (!b) return 2;
  if (n == 0) return 0;
  if (!h_in || !in_sz || !h_out || !out_sz || !tmp) return 2;
  try {
    auto *I = b->m->batch_manager().impl();
    std::lock_guard<std::mutex> lock(I->mu);
    if (!b->pipe.init()) return 4;
    HostPipe &P = b->pipe;
    size_t edges[HostPipe::MAX_WAVES + 1];
    const int nw = plan_waves(n, edges);
    std::vector<HostRun> in_runs[HostPipe::MAX_WAVES], out_runs[HostPipe::MAX_WAVES];
    std::vector<size_t> in_off(n), out_off(n);
    size_t off = 0;
    for (int w = 0; w < nw; w++) off = plan_runs(h_in, in_sz, edges[w], edges[w + 1], off, in_runs[w], in_off, 255);
    const size_t out_base = off;
    for (int w = 0; w < nw; w++) off = plan_runs(h_out, out_sz, edges[w], edges[w + 1], off, out_runs[w], out_off, 0);
    const size_t tab_base = off, tab_bytes = align_up(n * 36, 256);
    const size_t ws_base = tab_base + tab_bytes;
    const size_t codec_need = b->m->get_decompress_temp_size(in_sz, n);
    if (tmp_bytes < ws_base + codec_need) return 7;
    unsigned char *d = static_cast<unsigned char *>(tmp);
    (void)out_base;
    // device-side tables: in_ptrs | in_sizes | out_ptrs | out_sizes | statuses
    std::vector<unsigned long long> tab(4 * n);
    for (size_t i = 0; i < n; i++) {
      tab[i] = (unsigned long long)(uintptr_t)(d + in_off[i]); tab[n + i] = in_sz[i];
      tab[2 * n + i] = (unsigned long long)(uintptr_t)(d + out_off[i]); tab[3 * n + i] = out_sz[i];
    }
    unsigned char *d_tab = d + tab_base;
    u32 *d_status = reinterpret_cast<u32 *>(d_tab + 4 * n * 8);
    cudaError_t e;
    if ((e = cudaEventRecord(P.ev_entry, stream)) != cudaSuccess) return 4;
    cudaStreamWaitEvent(P.s_in, P.ev_entry, 0);
    cudaStreamWaitEvent(P.s_out, P.ev_entry, 0);
    if ((e = cudaMemcpyAsync(d_tab, tab.data(), 4 * n * 8, cudaMemcpyHostToDevice, P.s_in)) != cudaSuccess) return 4;
    int launches = 0;
    Status overall = Status::SUCCESS;
    for (int w = 0; w < nw; w++) {
      const size_t lo = edges[w], m = edges[w + 1] - edges[w];
      if (m == 0) continue;
      for (const HostRun &r : in_runs[w])
        if ((e = cudaMemcpyAsync(d + r.d_off, r.h_begin, r.bytes, cudaMemcpyHostToDevice, P.s_in)) != cudaSuccess) return 4;
      cudaEventRecord(P.ev_in[w], P.s_in);
      cudaStreamWaitEvent(stream, P.ev_in[w], 0);
      Status s = I->run(false, reinterpret_cast<const void *const *>(d_tab) + lo, reinterpret_cast<const size_t *>(d_tab + n * 8) + lo, m,
                        reinterpret_cast<void *const *>(d_tab + 2 * n * 8) + lo, reinterpret_cast<size_t *>(d_tab + 3 * n * 8) + lo, d_status + lo,
                        true, d + ws_base, tmp_bytes - ws_base, stream, false, nullptr, nullptr);
      if (s != Status::SUCCESS) overall = s;
      launches += I->last_launches;
      cudaEventRecord(P.ev_run[w], stream);
      cudaStreamWaitEvent(P.s_out, P.ev_run[w], 0);
      for (const HostRun &r : out_runs[w])
        if ((e = cudaMemcpyAsync(const_cast<unsigned char *>(r.h_begin), d + r.d_off, r.bytes, cudaMemcpyDeviceToHost, P.s_out)) != cudaSuccess) return 4;
    }
    std::vector<u32> st(n);
    cudaMemcpyAsync(out_sz, d_tab + 3 * n * 8, n * 8, cudaMemcpyDeviceToHost, P.s_out);
    cudaMemcpyAsync(st.data(), d_status, n * 4, cudaMemcpyDeviceToHost, P.s_out);
    cudaEventRecord(P.ev_done, P.s_out);
    cudaStreamWaitEvent(stream, P.ev_done, 0);
    if ((e = cudaStreamSynchronize(P.s_out)) != cudaSuccess) return 4;
    I->last_launches = launches;
    if (overall != Status::SUCCESS) return status_to_nvcomp_error(overall);
    int rc = 0;
    for (size_t i = 0; i < n; i++) {
      if (h_status) h_status[i] = st[i];
      if (st[i] != 0 && rc == 0) rc = 1;
    }
    return rc;
  } catch (...) { return 1; }
}
int cuda_zstd_batch_compress_host_packed(cuda_zstd_batch_t *b, const void *const *h_in, const size_t *in_sz, size_t n, void *h_packed,
                                         size_t packed_cap, uint64_t *h_offsets, uint32_t *h_status, void *tmp, size_t tmp_bytes,
                                         cudaStream_t stream) {
  if (!b) return 2;
  if (n == 0) return 0;
  if (!h_in || !in_sz || !h_packed || !h_offsets || !tmp) return 2;
  try {
    auto *I = b->m->batch_manager().impl();
    std::lock_guard<std::mutex> lock(I->mu);
    if (!b->pipe.init()) return 4;
    HostPipe &P = b->pipe;
    size_t edges[HostPipe::MAX_WAVES + 1];
    const int nw = plan_waves(n, edges, true);
    std::vector<HostRun> in_runs[HostPipe::MAX_WAVES];
    std::vector<size_t> in_off(n), caps(n), frame_off(n);
    size_t off = 0;
    for (int w = 0; w < nw; w++) off = plan_runs(h_in, in_sz, edges[w], edges[w + 1], off, in_runs[w], in_off, 255);
    size_t frames = 0;
    for (size_t i = 0; i < n; i++) { caps[i] = b->m->get_max_compressed_chunk_size(in_sz[i]); frame_off[i] = frames; frames += align_up(caps[i], 16); }
    const size_t frames_base = off, packed_base = frames_base + align_up(frames, 256);
    const size_t tab_base = packed_base + align_up(frames, 256), tab_bytes = align_up(n * 36 + (n + 1) * 8, 256);
    const size_t ws_base = tab_base + tab_bytes;
    if (tmp_bytes < ws_base + b->m->get_compress_temp_size(in_sz, n)) return 7;
    unsigned char *d = static_cast<unsigned char *>(tmp);
    std::vector<unsigned long long> tab(4 * n);
    for (size_t i = 0; i < n; i++) {
      tab[i] = (unsigned long long)(uintptr_t)(d + in_off[i]); tab[n + i] = in_sz[i];
      tab[2 * n + i] = (unsigned long long)(uintptr_t)(d + frames_base + frame_off[i]); tab[3 * n + i] = caps[i];
    }
    unsigned char *d_tab = d + tab_base;
    u32 *d_status = reinterpret_cast<u32 *>(d_tab + 4 * n * 8);
    uint64_t *d_offsets = reinterpret_cast<uint64_t *>(d_tab + align_up(n * 36, 8));
    cudaError_t e;
    if ((e = cudaEventRecord(P.ev_entry, stream)) != cudaSuccess) return 4;
    cudaStreamWaitEvent(P.s_in, P.ev_entry, 0);
    cudaStreamWaitEvent(P.s_out, P.ev_entry, 0);
    if ((e = cudaMemcpyAsync(d_tab, tab.data(), 4 * n * 8, cudaMemcpyHostToDevice, P.s_in)) != cudaSuccess) return 4;
    int launches = 0;
    Status overall = Status::SUCCESS;
    // all input copies are queued at once (every wave has its own staging range): the link never waits for the host
    for (int w = 0; w < nw; w++) {
      for (const HostRun &r : in_runs[w])
        if ((e = cudaMemcpyAsync(d + r.d_off, r.h_begin, r.bytes, cudaMemcpyHostToDevice, P.s_in)) != cudaSuccess) return 4;
      cudaEventRecord(P.ev_in[w], P.s_in);
    }
    // Per wave: compress, scan of the wave's frame sizes chained to the previous wave's closing offset, pack, and the closing
    // offset to pinned host memory.  The host follows one wave behind: as soon as wave w - 1 has its closing offset it
    // queues that wave's packed bytes on the copy-out stream, beside the compress of wave w.
    unsigned long long done_bytes = 0;
    int rc_cap = 0;
    auto copy_home = [&](int w) -> int {
      if (cudaEventSynchronize(P.ev_run[w]) != cudaSuccess) return 4;
      const unsigned long long upto = P.h_totals[w];
      if (upto > packed_cap) { rc_cap = 7; return 0; }
      if (upto > done_bytes && rc_cap == 0 &&
          cudaMemcpyAsync(static_cast<unsigned char *>(h_packed) + done_bytes, d + packed_base + done_bytes, (size_t)(upto - done_bytes),
                          cudaMemcpyDeviceToHost, P.s_out) != cudaSuccess) return 4;
      done_bytes = upto;
      return 0;
    };
    int prev = -1;
    for (int w = 0; w < nw; w++) {
      const size_t lo = edges[w], m = edges[w + 1] - edges[w];
      if (m == 0) continue;
      cudaStreamWaitEvent(stream, P.ev_in[w], 0);
      Status s = I->run(true, reinterpret_cast<const void *const *>(d_tab) + lo, reinterpret_cast<const size_t *>(d_tab + n * 8) + lo, m,
                        reinterpret_cast<void *const *>(d_tab + 2 * n * 8) + lo, reinterpret_cast<size_t *>(d_tab + 3 * n * 8) + lo, d_status + lo,
                        true, d + ws_base, tmp_bytes - ws_base, stream, false, nullptr, nullptr);
      if (s != Status::SUCCESS) overall = s;
      launches += I->last_launches;
      const size_t *wsizes = reinterpret_cast<const size_t *>(d_tab + 3 * n * 8) + lo;
      if ((lo == 0 ? b200zstd::launch_scan_sizes(wsizes, m, 0, d_offsets, stream)
                   : b200zstd::launch_scan_sizes_from(wsizes, m, d_offsets + lo, d_offsets + lo, stream)) != cudaSuccess) return 4;
      if (b200zstd::launch_pack(reinterpret_cast<const void *const *>(d_tab + 2 * n * 8) + lo, wsizes, d_offsets + lo, m, d + packed_base, stream) !=
          cudaSuccess) return 4;
      launches += 2;
      if (cudaMemcpyAsync(P.h_totals + w, d_offsets + lo + m, 8, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return 4;
      cudaEventRecord(P.ev_run[w], stream);
      if (prev >= 0) { const int r2 = copy_home(prev); if (r2) return r2; }
      prev = w;
    }
    std::vector<u32> st(n);
    cudaMemcpyAsync(h_offsets, d_offsets, (n + 1) * 8, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(st.data(), d_status, n * 4, cudaMemcpyDeviceToHost, stream);
    if (prev >= 0) { const int r2 = copy_home(prev); if (r2) return r2; }
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return 4;
    if ((e = cudaStreamSynchronize(P.s_out)) != cudaSuccess) return 4;
    I->last_launches = launches;
    if (overall != Status::SUCCESS) return status_to_nvcomp_error(overall);
    if (rc_cap) return rc_cap;
    int rc = 0;
    for (size_t i = 0; i < n; i++) {
      if (h_status) h_status[i] = st[i];
      if (st[i] != 0 && rc == 0) rc = 1;
    }
    return rc;
  } catch (...) { return 1; }
}
// ---- shards of one batch on several GPUs of this process ----
static int run_sharded(bool compress, cuda_zstd_shard_t *sh, int G) {
  if (!sh || G <= 0) return 2;
  for (int g = 0; g < G; g++)
    if (!sh[g].mgr || !sh[g].d_in_ptrs || !sh[g].d_in_sizes || !sh[g].d_out_ptrs || !sh[g].d_out_sizes || !sh[g].d_statuses) return 2;
  int home = 0;
  cudaGetDevice(&home);
  std::vector<int> rc(G, 0);
  std::vector<cudaEvent_t> ev(G, nullptr);
  std::vector<std::thread> th;
  // one host thread per shard: the launches of different GPUs are issued side by side
  for (int g = 0; g < G; g++)
    th.emplace_back([&, g]() {
      cuda_zstd_shard_t &s = sh[g];
      if (cudaSetDevice(s.device) != cudaSuccess) { rc[g] = 4; return; }
      if (cudaEventCreateWithFlags(&ev[g], cudaEventDisableTiming) != cudaSuccess) { rc[g] = 4; return; }
      if (s.num_chunks)
        rc[g] = compress ? cuda_zstd_batch_compress_nosync(s.mgr, s.d_in_ptrs, s.d_in_sizes, s.num_chunks, s.d_out_ptrs, s.d_out_sizes, s.d_statuses,
                                                           s.d_temp, s.temp_bytes, s.stream)
                         : cuda_zstd_batch_decompress_nosync(s.mgr, s.d_in_ptrs, s.d_in_sizes, s.num_chunks, s.d_out_ptrs, s.d_out_sizes, s.d_statuses,
                                                             s.d_temp, s.temp_bytes, s.stream);
      cudaEventRecord(ev[g], s.stream);
    });
  for (auto &t : th) t.join();
  int overall = 0;
  for (int g = 0; g < G; g++) if (rc[g] != 0 && overall == 0) overall = rc[g];
  size_t total = 0;
  std::vector<size_t> base(G);
  for (int g = 0; g < G; g++) { base[g] = total; total += sh[g].num_chunks; }
  // the exchange: every shard's sizes to every shard's table, then the scan on each device
  if (overall == 0)
    for (int h = 0; h < G; h++) {
      if (!sh[h].d_all_sizes) continue;
      cudaSetDevice(sh[h].device);
      for (int g = 0; g < G; g++) {
        if (!sh[g].num_chunks) continue;
        cudaStreamWaitEvent(sh[h].stream, ev[g], 0);
        if (cudaMemcpyPeerAsync(sh[h].d_all_sizes + base[g], sh[h].device, sh[g].d_out_sizes, sh[g].device, sh[g].num_chunks * sizeof(size_t),
                                sh[h].stream) != cudaSuccess) overall = 4;
      }
      if (sh[h].d_all_offsets && total &&
          b200zstd::launch_scan_sizes(sh[h].d_all_sizes, total, 0, sh[h].d_all_offsets, sh[h].stream) != cudaSuccess) overall = 4;
    }
  // completion and the per-chunk verdicts
  for (int g = 0; g < G; g++) {
    cudaSetDevice(sh[g].device);
    std::vector<u32> st(sh[g].num_chunks);
    if (sh[g].num_chunks && cudaMemcpyAsync(st.data(), sh[g].d_statuses, st.size() * 4, cudaMemcpyDeviceToHost, sh[g].stream) != cudaSuccess && overall == 0) overall = 4;
    if (cudaStreamSynchronize(sh[g].stream) != cudaSuccess && overall == 0) overall = 4;
    if (overall == 0) for (u32 v : st) if (v != 0) { overall = 1; break; }
  }
  for (int g = 0; g < G; g++) if (ev[g]) { cudaSetDevice(sh[g].device); cudaEventDestroy(ev[g]); }
  cudaSetDevice(home);
  return overall;
}
int cuda_zstd_batch_compress_sharded(cuda_zstd_shard_t *shards, int num_shards) { try { return run_sharded(true, shards, num_shards); } catch (...) { return 1; } }
int cuda_zstd_batch_decompress_sharded(cuda_zstd_shard_t *shards, int num_shards) { try { return run_sharded(false, shards, num_shards); } catch (...) { return 1; } }

int cuda_zstd_batch_scan_sizes(const size_t *d_sizes, size_t n, uint64_t base, uint64_t *d_offsets, cudaStream_t stream) {
  if (!d_sizes || !d_offsets) return 2;
  return b200zstd::launch_scan_sizes(d_sizes, n, base, d_offsets, stream) == cudaSuccess ? 0 : 4;
}
int cuda_zstd_batch_pack(const void *const *d_ptrs, const size_t *d_sizes, const uint64_t *d_offsets, size_t n, void *d_packed,
                         cudaStream_t stream) {
  if (!d_ptrs || !d_sizes || !d_offsets || !d_packed) return 2;
  return b200zstd::launch_pack(d_ptrs, d_sizes, d_offsets, n, d_packed, stream) == cudaSuccess ? 0 : 4;
}
int cuda_zstd_batch_last_launch_count(cuda_zstd_batch_t *b) { return b ? b->m->batch_manager().impl()->last_launches : 0; }
const char *cuda_zstd_batch_error_string(int code) { return cuda_zstd::status_to_string(static_cast<Status>(code)); }

} // extern "C"
