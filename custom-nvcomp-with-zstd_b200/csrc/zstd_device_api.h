// zstd_device_api.h -- host<->kernel launch interface of the sm_100a batch codec (internal).
#pragma once
#include <cuda_runtime_api.h>
#include <stddef.h>
#include <stdint.h>

#include "zstd_encode_params.h"

namespace b200zstd {

constexpr size_t LIT_SCRATCH_BYTES = 128 * 1024 + 256;   // per in-flight chunk: one block's literals
constexpr size_t WS_HEADER_BYTES = 256;                   // work counters etc. at the start of the workspace

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
};
cudaError_t launch_decode_batch(const DecodeArgs &args, int grid, cudaStream_t stream);
int decode_ctas_per_sm();

struct EncodeArgs {
  const void *const *in_ptrs;
  const size_t *in_sizes;
  void *const *out_ptrs;
  size_t *out_sizes;        // in = capacity, out = frame bytes (0 on failure)
  uint32_t *statuses;       // may be null
  uint32_t *counter;
  uint8_t *scratch;         // grid * encode_cta_scratch_bytes()
  uint32_t n;
  EncodeParams prm;
};
size_t encode_cta_scratch_bytes(const EncodeParams &prm);
cudaError_t launch_encode_batch(const EncodeArgs &args, int grid, cudaStream_t stream);
int encode_ctas_per_sm(const EncodeParams &prm);

// exclusive scan of sizes (+ base) -> offsets[0..n], offsets[n] = base + total; single CTA.
cudaError_t launch_scan_sizes(const size_t *d_sizes, size_t n, uint64_t base, uint64_t *d_offsets, cudaStream_t stream);
// gather frames into a packed buffer
cudaError_t launch_pack(const void *const *d_ptrs, const size_t *d_sizes, const uint64_t *d_offsets, size_t n, void *d_packed,
                        cudaStream_t stream);

} // namespace b200zstd
