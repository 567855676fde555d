// zstd_encode_params.h -- level -> parse parameters and the hash functions of the match finder.
// Shared (host + device) so that the kernel, the host-side model and the workspace sizing agree.
//
// Level mapping.  The reference maps level -> {strategy, window/hash/chain log, search depth}
// (src/cuda_zstd_types.cpp:147-210, 860-950) but never branches on the strategy; the
// compressed-size target is libzstd's output at the same level, whose parameters for <= 128 KB inputs
// are fast(L1) / dfast(L3) / greedy(L5) / lazy2(L9).  This build follows those
// strategy bands with tables sized for shared memory:
//   L1-2   FAST    one 5-byte hash table, greedy
//   L3-4   DFAST   8-byte "long" table + 5-byte "short" table, one-step lazy
//   L5     GREEDY  4-byte hash + chain, 8 candidates
//   L6     LAZY    8 candidates, one-step lazy
//   L7-10  LAZY2   8/16/32/64 candidates, two-step lazy   (L11+: 128 candidates; no optimal parser)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZP_HD __host__ __device__ __forceinline__
#else
#define ZP_HD inline
#endif

namespace b200zstd {

constexpr uint32_t BLOCK_BYTES = 128u * 1024u;
constexpr uint32_t MAX_SEQ_PER_BLOCK = BLOCK_BYTES / 4 + 64;     // every sequence covers >= 4 input bytes

struct EncodeParams {
  int level;          // 1..22 as given by the caller
  int strategy;       // 0 fast, 1 dfast, 2 greedy, 3 lazy, 4 lazy2
  int hash_log;       // short/primary table: 1 << hash_log uint16 entries
  int hash_bytes;     // bytes hashed for the primary table (4, 5 or 6)
  int long_log;       // 8-byte-hash table (dfast); 0 = absent
  int chain_depth;    // 0 = no chain; else candidates walked per position
  int min_match;      // shortest non-repeat match emitted
  int lazy;           // lanes after the first match that may replace it (0, 1, 2)
  int insert_all;     // insert every position covered by a match (else only two)
  uint32_t lane_cap;  // per-lane match-length evaluation cap (chain levels)
  uint32_t rep_bonus; // a repeat-offset match wins when its length + bonus exceeds the best other
  int checksum;       // 1 = append XXH64 content checksum
};

ZP_HD EncodeParams encode_params_for_level(int level, int checksum) {
  EncodeParams p{};
  if (level < 1) level = 1;
  if (level > 22) level = 22;
  p.level = level;
  p.checksum = checksum ? 1 : 0;
  p.lane_cap = 8;
  p.rep_bonus = 1;
  if (level <= 2) {
    p.strategy = 0; p.hash_log = 13; p.hash_bytes = 5; p.long_log = 0; p.chain_depth = 0; p.min_match = 5; p.lazy = 0; p.insert_all = 1;
  } else if (level <= 4) {
    p.strategy = 1; p.hash_log = 11; p.hash_bytes = 5; p.long_log = 13; p.chain_depth = 0; p.min_match = 5; p.lazy = 1; p.insert_all = 1;
  } else {
    // chain levels: libzstd's <=128 KB rows are greedy(5) lazy(6) lazy2(7..10) with 2^3..2^6 attempts
    static const int depth[] = {8, 8, 8, 16, 32, 64};
    p.hash_log = 13; p.hash_bytes = 4; p.long_log = 0; p.min_match = 4; p.insert_all = 1;
    p.chain_depth = level <= 10 ? depth[level - 5] : 128;
    p.lazy = level == 5 ? 0 : level == 6 ? 1 : 2;
    p.strategy = level == 5 ? 2 : level == 6 ? 3 : 4;
    p.lane_cap = level <= 6 ? 64 : 128;
  }
  return p;
}

// ---- levels 1-4, blocks <= 128 KB: the match / select / finish pipeline of zstd_encode_esd.cu ------------------------
// One CTA per block finds, for EVERY position, its best table candidate (block and hash tables in shared memory, all
// positions in parallel, window by window); one warp per block then walks the result (lanes walk sub-segments
// speculatively and are stitched together exactly), and a third kernel codes the block.  libzstd's <= 128 KB rows for
// these levels are fast (L1-2) and dfast (L3-4); sizes land within +-2 % of them or below (tools/model_ratio.cpp,
// tests/test_gpu_encode.py).  The parse arithmetic lives in zstd_encode_lz.cuh.
#ifndef LZ_MAX_LEVEL
#define LZ_MAX_LEVEL 9
#endif
struct EsdParams {
  int dfast;          // 0: one table (FAST), 1: 5-byte table + 8-byte "long" table (DFAST)
  int hash_log;       // primary table: 1 << hash_log uint32 entries
  int hash_bytes;     // bytes hashed for the primary table
  int long_log;       // 8-byte-hash table; 0 = absent
  int lazy;           // 1: the position after the first candidate may replace it when its match is longer
  int rows;           // levels 5+: the 8-byte table holds rows of LZ_ROW_WAYS tagged entries, matches are measured up to
                      // LZ_QCAP bytes and the walk is lazy by gain with depth `lazy` (0, 1, 2); 0: the levels 1-4 form
};
ZP_HD bool esd_level(int level) { return level <= LZ_MAX_LEVEL; }
// big: the 128 KB block geometry (the block itself takes twice the shared memory, the tables half)
ZP_HD EsdParams esd_params_for_level(int level, int big) {
  EsdParams e{};
  if (level <= 2) { e.dfast = 0; e.hash_log = big ? 14 : 15; e.hash_bytes = level <= 1 ? 6 : 5; e.long_log = 0; e.lazy = level >= 2 ? 1 : 0; }
  else if (level <= 4) { e.dfast = 1; e.hash_log = big ? 13 : 14; e.hash_bytes = 5; e.long_log = big ? 13 : 14; e.lazy = 1; }
  else {
    // rows: 2^11 rows x 16 ways x 4 B = 128 KB for the 8-byte hash, 2^12 single entries for the 4-byte hash
    e.dfast = 1; e.rows = 1; e.hash_log = 12; e.hash_bytes = 4; e.long_log = 11; e.lazy = level == 5 ? 0 : level == 6 ? 1 : 2;
  }
  return e;
}

// hashes of the little-endian 8 bytes at a position, built from 32-bit multiplies only (IMAD on the GPU)
ZP_HD uint32_t hash_short(uint64_t v, int bytes, int log) {
  const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  if (bytes == 4) return (lo * 2654435761u) >> (32 - log);
  if (bytes == 5) return ((lo * 2654435761u) ^ ((hi & 0xFFu) * 2246822519u)) >> (32 - log);
  return ((lo * 2654435761u) ^ ((hi & 0xFFFFu) * 2246822519u)) >> (32 - log);
}
ZP_HD uint32_t hash_long(uint64_t v, int log) {
  const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  return ((lo * 2654435761u) + (hi * 2246822519u) * 3266489917u) >> (32 - log);
}

} // namespace b200zstd
