// zstd_device_api.h -- host<->kernel launch interface of the sm_100a batch codec (internal).
#pragma once
#include <cuda_runtime_api.h>
#include <stddef.h>
#include <stdint.h>

#include "zstd_encode_params.h"

namespace b200zstd {

constexpr size_t LIT_SCRATCH_BYTES = 128 * 1024 + 256;   // per in-flight chunk: one block's literals
constexpr size_t WS_HEADER_BYTES = 2048;                   // work counters etc. at the start of the workspace

// All pointers are DEVICE memory.  out_sizes: in = capacity, out = bytes produced (0 on failure).
struct DecodeArgs {
  const void *const *in_ptrs;
  const size_t *in_sizes;
  void *const *out_ptrs;
  size_t *out_sizes;
  uint32_t *statuses;       // may be null
  uint32_t *counter;        // work-queue head, zeroed by the launcher
  uint8_t *lit_scratch;     // grid * LIT_SCRATCH_BYTES
  uint32_t n;
  int verify_checksum;
  const uint32_t *list;        // optional: chunk indices to process (general kernel behind the fast path)
  const uint32_t *list_count;  // number of valid entries in list (device memory)
};
cudaError_t launch_decode_batch(const DecodeArgs &args, int grid, cudaStream_t stream);
cudaError_t launch_decode_batch_nomemset(const DecodeArgs &args, int grid, cudaStream_t stream);   // counter already zero
int decode_ctas_per_sm();

// ---- batch fast path: three kernels over the whole batch (zstd_decode_fast.cu) --------------------
constexpr uint32_t FAST_WAVE = 16384;                         // chunks per fast-path wave (bounds the scratch)
constexpr uint32_t FAST_ORDER_SUBS = 4;                       // KC's work order is built for up to this many sub-waves per wave
constexpr uint32_t FAST_SEQ_CAP = 65536;                      // sequences per chunk the fast path accepts
constexpr size_t FAST_TABLE_BYTES = 64 + 4096 + 3840;         // sequence-stream info, Huffman table, packed LL/ML/OF tables (KP -> KA, KB)
constexpr size_t FAST_DESC_BYTES = 192;
constexpr size_t FAST_SLOT_BYTES = FAST_DESC_BYTES + FAST_TABLE_BYTES;     // fixed per-chunk slot
// Literals and sequence records come from ONE bump-allocated pool sized from the COMPRESSED sizes the caller passes
// to the temp-size query (lit_pool == seq_pool, one head); a chunk that does not fit takes the general kernel instead.
struct FastDecodeArgs {
  DecodeArgs base;          // tables, sizes, statuses; base.counter / lit_scratch serve the general kernel
  uint8_t *slots;           // n * FAST_SLOT_BYTES
  uint8_t *lit_pool;        // 16-byte aligned
  uint8_t *seq_pool;
  uint64_t lit_pool_bytes, seq_pool_bytes;   // both = the pool's size
  unsigned long long *pool_heads;   // [0] the pool's head; zeroed by the launcher
  uint32_t *slow_list;      // n entries
  uint32_t *slow_count;     // 1 entry, zeroed by the launcher
  uint32_t *group_counters; // [0] literal kernel, [1] sequence kernel work queues; zeroed by the launcher
  uint32_t *lit_buckets;    // 64 counters + [64] their sum: Huffman chunks by literal count (KP -> order kernel -> KA); zeroed by the launcher
  uint32_t *seq_buckets;    // FAST_ORDER_SUBS x 64 counters: chunks of a KB sub-wave by sequence count (KP -> order kernel -> KC's queue)
  uint32_t *kc_order;       // n entries: the chunks of every sub-wave, most sequences first; null = chunk order
  uint32_t sub_chunks;      // chunks per KB sub-wave (set by the launcher)
  int general_grid;
  int sm_count;
  uint32_t lo, hi, sub;     // set by the launcher: chunk sub-range and work-queue index of a KB / KC launch
  // ... and, when a sub-wave's sequences are decoded in several KB launches so that KC can start on the first parts of every
  // chunk while KB decodes the later ones: this KB launch decodes the sequences of parts [part_lo, part_hi) (the chain's state
  // waits in the slot in between) and the KC launch behind it executes them; kb_queue / kc_queue = the launches' work counters
  uint32_t part_lo, part_hi, kb_queue, kc_queue;
  // bare-block mode (a multi-block frame cut into units by launch_split_frame): every item starts at a block header,
  // must regenerate exactly its capacity, and -- except global unit 0 -- starts with an unknown repeat-offset history
  uint32_t bare_blocks, unit_base;
};
// Result of cutting one frame into block units (device memory, 64 bytes)
struct SplitInfo {
  uint32_t ok;              // 1: the tables hold `units` self-describing block units
  uint32_t units;
  uint64_t content_size;
  uint32_t has_checksum, checksum_off;
  uint32_t pad[10];
};
// one warp walks the block headers of the frame at d_src and fills the four unit tables (max_units entries each)
cudaError_t launch_split_frame(const void *d_src, size_t n, void *d_dst, size_t cap, uint32_t max_units, const void **d_in_ptrs,
                               size_t *d_in_sizes, void **d_out_ptrs, size_t *d_out_sizes, SplitInfo *d_info, cudaStream_t stream);
// compares the low 32 bits of XXH64(d_data[0..n)) with the 4 bytes at d_expect; *d_flag = 1 on mismatch
cudaError_t launch_verify_checksum(const void *d_data, size_t n, const void *d_expect, uint32_t *d_flag, cudaStream_t stream);
// optional helper stream + events (owned by the manager) that let KC overlap the next KB sub-wave
struct FastOverlap {
  static constexpr int MAX_SUB = 8;
  cudaStream_t side;
  cudaEvent_t ev[MAX_SUB];
  cudaEvent_t done;
  cudaEvent_t prep;         // KP + order kernel finished: KA may start on the side stream while KB runs on the caller's
};
// returns the number of kernels launched through *launches
cudaError_t launch_decode_fast(const FastDecodeArgs &args, cudaStream_t stream, const FastOverlap *overlap, int *launches);

struct EncodeArgs {
  const void *const *in_ptrs;
  const size_t *in_sizes;
  void *const *out_ptrs;
  size_t *out_sizes;        // in = capacity, out = frame bytes (0 on failure)
  uint32_t *statuses;       // may be null
  uint32_t *counter;
  uint8_t *scratch;         // grid * encode_cta_scratch_bytes()
  uint32_t n;
  uint32_t block_mode;      // 1: items are the consecutive <= 128 KB blocks of one frame (no frame header / checksum, see kernel)
  EncodeParams prm;
  const uint32_t *list;        // optional: item indices to process instead of 0 .. n-1 (device memory) ...
  const uint32_t *list_count;  // ... and how many of them
};
size_t encode_cta_scratch_bytes(const EncodeParams &prm);
cudaError_t launch_encode_batch(const EncodeArgs &args, int grid, cudaStream_t stream);
cudaError_t launch_encode_batch_nomemset(const EncodeArgs &args, int grid, cudaStream_t stream);   // counter already zero

// ---- levels 1-4: the decoupled pipeline (zstd_encode_esd.cu) ----
struct EsdLaunch {
  uint32_t *counters;       // esd_counter_words() words, zeroed by the launcher: work heads and list counts per wave
  size_t counter_words;
  uint32_t *lists;          // 2 * n words: items handed to the 128 KB-block kernel | to the general kernel
  uint8_t *scratch;         // esd_scratch_bytes(): the wave's sequence lists, then the finish kernel's per-warp buffers
  size_t scratch_bytes;
  int sm_count;
  uint32_t block_max;       // 65536: every item above 64 KB takes the general kernel; 131072: blocks up to 128 KB are parsed here
  size_t min_item_bytes, max_item_bytes;    // 0 / 0: unknown (size table lives on the device)
};
size_t esd_scratch_bytes(size_t n_items, uint32_t block_max, int sm_count);
size_t esd_counter_words(size_t n_items, uint32_t block_max);
cudaError_t launch_encode_esd(const EncodeArgs &args, const EsdLaunch &L, cudaStream_t stream, int *launches);
int encode_ctas_per_sm(const EncodeParams &prm);

// low 32 bits of XXH64(src[0..n)) written at dst + *d_where (n < 4 GiB)
cudaError_t launch_frame_checksum(const void *d_src, size_t n, void *d_dst, const uint64_t *d_where, cudaStream_t stream);
// d_result[0] = *d_total + add, d_result[1] = first non-zero entry of d_statuses[0..n) (0 if none)
cudaError_t launch_big_result(const uint32_t *d_statuses, size_t n, const uint64_t *d_total, uint64_t add, uint64_t *d_result,
                              cudaStream_t stream);
// exclusive scan of sizes (+ base) -> offsets[0..n], offsets[n] = base + total; single CTA.
cudaError_t launch_scan_sizes(const size_t *d_sizes, size_t n, uint64_t base, uint64_t *d_offsets, cudaStream_t stream);
cudaError_t launch_scan_sizes_from(const size_t *d_sizes, size_t n, const uint64_t *d_base, uint64_t *d_offsets, cudaStream_t stream);
// gather frames into a packed buffer
cudaError_t launch_pack(const void *const *d_ptrs, const size_t *d_sizes, const uint64_t *d_offsets, size_t n, void *d_packed,
                        cudaStream_t stream);

} // namespace b200zstd
