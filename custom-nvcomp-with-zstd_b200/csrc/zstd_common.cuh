// zstd_common.cuh -- shared device-side definitions for the sm_100a Zstandard batch codec.
//
// Format constants follow RFC 8878; the reference keeps the same tables in
// include/cuda_zstd_internal.h:235-449 (LL/ML/OF code tables) and src/cuda_zstd_fse.cu:2507-2528
// (predefined distributions).  Nothing here is copied from those files: the values are the
// format's, the layout (packed 8-byte decode entries, constant-memory tables) is this build's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200zstd {

// cuda_zstd::Status values used by the kernels (include/cuda_zstd_types.h)
enum : uint32_t {
  ST_OK = 0,
  ST_GENERIC = 1,
  ST_INVALID_PARAMETER = 2,
  ST_INVALID_MAGIC = 5,
  ST_CORRUPT = 6,
  ST_BUFFER_TOO_SMALL = 7,
  ST_DICT_MISMATCH = 9,
  ST_CHECKSUM = 10,
  ST_COMPRESSION = 12,
  ST_UNSUPPORTED = 28
};

constexpr uint32_t ZSTD_FRAME_MAGIC = 0xFD2FB528u;
constexpr uint32_t ZSTD_SKIP_MAGIC = 0x184D2A50u;
constexpr uint32_t BLOCK_MAX = 128u * 1024u;
constexpr int LL_MAX_SYM = 35, ML_MAX_SYM = 52, OF_MAX_SYM = 31;
constexpr int LL_MAX_LOG = 9, ML_MAX_LOG = 9, OF_MAX_LOG = 8;
constexpr int LL_DEF_LOG = 6, ML_DEF_LOG = 6, OF_DEF_LOG = 5;
constexpr int HUF_MAX_LOG = 11;

static __constant__ uint32_t c_ll_base[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40,
                                       48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
static __constant__ uint8_t c_ll_bits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3,
                                      4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static __constant__ uint32_t c_ml_base[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28,
                                       29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051,
                                       4099, 8195, 16387, 32771, 65539};
static __constant__ uint8_t c_ml_bits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                      0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static __constant__ int16_t c_ll_def[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static __constant__ int16_t c_ml_def[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                     1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static __constant__ int16_t c_of_def[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

__device__ __forceinline__ int highbit32(uint32_t v) { return 31 - __clz(v); }   // v != 0
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ uint32_t ld_le16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ uint32_t ld_le24(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
__device__ __forceinline__ uint32_t ld_le32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// ---------------------------------------------------------------------------------------------
// Backward bit reader (RFC 8878 section 4.1).  The stream occupies [src, src+n); the last byte holds
// the sentinel bit.  The reader only ever issues ALIGNED 32-bit loads (shared or global), walking
// down from the word that holds the last byte; words below `floor` read as zero, so a corrupt
// stream cannot walk out of the buffer.  `left` counts unread payload bits and goes negative on
// over-read (checked by the caller).
// ---------------------------------------------------------------------------------------------
struct BackBits {
  const uint32_t *wp;     // address of `nextw` (the next word to enter the window)
  const uint32_t *floor;  // lowest word that may be fetched
  uint64_t win;           // unread bits, MSB-aligned
  uint32_t nextw;         // prefetched: loaded one refill ahead so its latency is off the critical path
  int avail;              // valid bits in win
  int left;               // payload bits not yet consumed

  // returns false when the stream is malformed (empty or zero last byte)
  __device__ __forceinline__ bool init(const uint8_t *src, uint32_t n) {
    if (n == 0) return false;
    const uint8_t *lastp = src + n - 1;
    uint32_t last = *lastp;
    if (last == 0) return false;
    uintptr_t a = (uintptr_t)lastp;
    const uint32_t *wa = (const uint32_t *)(a & ~(uintptr_t)3);
    uint32_t vb = (uint32_t)(a & 3) + 1;               // valid bytes in the top word
    uint32_t w = *wa << (8 * (4 - vb));
    int skip = 8 - highbit32(last);                    // zero padding + sentinel
    floor = (const uint32_t *)((uintptr_t)src & ~(uintptr_t)3);
    wp = wa - 1;
    nextw = (wp >= floor) ? *wp : 0u;
    win = ((uint64_t)w << 32) << skip;
    avail = (int)(8 * vb) - skip;
    left = (int)(8 * n) - skip;
    return true;
  }
  __device__ __forceinline__ void refill() {
    if (avail <= 32) {
      win |= (uint64_t)nextw << (32 - avail);
      avail += 32;
      wp--;
      nextw = (wp >= floor) ? *wp : 0u;
    }
  }
  __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)((win >> 1) >> (63 - n)); }   // n in [0,32]
  __device__ __forceinline__ void skip(int n) { win <<= n; avail -= n; left -= n; }
  __device__ __forceinline__ uint32_t read(int n) { uint32_t v = peek(n); skip(n); return v; }
};

// Forward little-endian bit reader for FSE table descriptions.  Refills four bytes at a time from the aligned words under
// the read position (the byte-by-byte loop this replaces was a seventh of the prepare kernel's instructions); bytes past
// the end of the buffer read as zero, and no word that lies wholly outside [p, p + n) is touched.
struct FwdBits {
  const uint8_t *p;
  uint32_t n, ip;
  uint64_t acc;
  int have;
  __device__ __forceinline__ void init(const uint8_t *src, uint32_t len) { p = src; n = len; ip = 0; acc = 0; have = 0; }
  __device__ __forceinline__ void need(int k) {          // k <= 32
    if (have < k) {
      uint32_t w = 0;
      if (ip < n) {
        const uintptr_t a = (uintptr_t)(p + ip);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        const uint32_t mis = (uint32_t)(a & 3), valid = min(n - ip, 4u);
        const uint32_t lo = wp[0], hi = (mis + valid > 4) ? wp[1] : 0u;
        w = __funnelshift_r(lo, hi, mis * 8);
        if (valid < 4) w &= (1u << (8 * valid)) - 1u;
      }
      acc |= (uint64_t)w << have; have += 32; ip += 4;
    }
  }
  __device__ __forceinline__ uint32_t peek(int k) { need(k); return (uint32_t)(acc & ((1ull << k) - 1)); }
  __device__ __forceinline__ void drop(int k) { acc >>= k; have -= k; }
  __device__ __forceinline__ uint32_t bytes_used() const { return (ip * 8 - (uint32_t)have + 7) >> 3; }
};

// Reads a normalised-count header (FSE_readNCount semantics; RFC 8878 4.1.1).  Single thread.
// On success returns bytes consumed (>0), fills norm[0..*max_sym] and *log.
__device__ inline int read_ncount(const uint8_t *src, uint32_t n, int16_t *norm, int max_sym_allowed, int max_log, int *max_sym,
                                  int *log) {
  FwdBits b;
  b.init(src, n);
  if (n < 1) return -1;
  int al = (int)b.peek(4) + 5;
  b.drop(4);
  if (al > max_log) return -1;
  int remaining = (1 << al) + 1, threshold = 1 << al, nb = al + 1, sym = 0;
  for (int i = 0; i <= max_sym_allowed; i++) norm[i] = 0;
  while (remaining > 1 && sym <= max_sym_allowed) {
    int mx = (2 * threshold - 1) - remaining, count;
    uint32_t v = b.peek(nb);
    if ((int)(v & (uint32_t)(threshold - 1)) < mx) { count = (int)(v & (uint32_t)(threshold - 1)); b.drop(nb - 1); }
    else { count = (int)(v & (uint32_t)(2 * threshold - 1)); if (count >= threshold) count -= mx; b.drop(nb); }
    count--;
    remaining -= count < 0 ? -count : count;
    norm[sym++] = (int16_t)count;
    if (count == 0) {
      for (;;) {
        uint32_t rep = b.peek(2);
        b.drop(2);
        sym += (int)rep;
        if (rep != 3 || sym > max_sym_allowed + 1) break;
      }
    }
    while (remaining < threshold && threshold > 1) { nb--; threshold >>= 1; }
  }
  if (remaining != 1 || sym > max_sym_allowed + 1) return -1;
  uint32_t used = b.bytes_used();
  if (used > n) return -1;
  *max_sym = sym - 1;
  *log = al;
  return (int)used;
}

// XXH64 primes and rounds (reference: src/cuda_zstd_xxhash.cu:72-138 restates the same public algorithm)
constexpr uint64_t XXP1 = 0x9E3779B185EBCA87ull, XXP2 = 0xC2B2AE3D27D4EB4Full, XXP3 = 0x165667B19E3779F9ull,
                   XXP4 = 0x85EBCA77C2B2AE63ull, XXP5 = 0x27D4EB2F165667C5ull;
__device__ __forceinline__ uint64_t xx_rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xx_round(uint64_t acc, uint64_t in) { return xx_rotl(acc + in * XXP2, 31) * XXP1; }
__device__ __forceinline__ uint64_t xx_merge(uint64_t h, uint64_t v) { return (h ^ xx_round(0, v)) * XXP1 + XXP4; }

// XXH64(seed 0) of [p, p+len) computed by lanes 0..3 of one warp (one accumulator per lane, the
// algorithm's own 4-way parallelism); 16-byte vector loads when p is 16-byte aligned.  All 32 lanes
// must call; the result is returned in every lane.
__device__ inline uint64_t xxh64_warp(const uint8_t *p, uint32_t len, int lane) {
  uint64_t h;
  uint32_t done = 0;
  if (len >= 32) {
    uint64_t v = (lane == 0) ? XXP1 + XXP2 : (lane == 1) ? XXP2 : (lane == 2) ? 0ull : 0ull - XXP1;
    uint32_t stripes = len >> 5;
    if (lane < 4) {
      bool al16 = (((uintptr_t)p) & 15) == 0;
      if (al16) {
        // lanes 0/1 share one 16-byte load, lanes 2/3 the next: each lane keeps its own half
        const uint4 *q = (const uint4 *)p + (lane >> 1);
        for (uint32_t s = 0; s < stripes; s++) {
          uint4 x = q[2 * s];          // plain load: the hashed bytes may have been written by this kernel (decoder output)
          uint64_t in = (lane & 1) ? ((uint64_t)x.w << 32 | x.z) : ((uint64_t)x.y << 32 | x.x);
          v = xx_round(v, in);
        }
      } else {
        const uint8_t *q = p + 8 * lane;
        for (uint32_t s = 0; s < stripes; s++) {
          uint64_t in = (uint64_t)ld_le32(q) | ((uint64_t)ld_le32(q + 4) << 32);
          v = xx_round(v, in);
          q += 32;
        }
      }
    }
    uint64_t v1 = __shfl_sync(0xffffffffu, v, 0), v2 = __shfl_sync(0xffffffffu, v, 1), v3 = __shfl_sync(0xffffffffu, v, 2),
             v4 = __shfl_sync(0xffffffffu, v, 3);
    h = xx_rotl(v1, 1) + xx_rotl(v2, 7) + xx_rotl(v3, 12) + xx_rotl(v4, 18);
    h = xx_merge(h, v1); h = xx_merge(h, v2); h = xx_merge(h, v3); h = xx_merge(h, v4);
    done = stripes << 5;
  } else {
    h = XXP5;
  }
  h += (uint64_t)len;
  const uint8_t *q = p + done;
  uint32_t rem = len - done;
  while (rem >= 8) {
    uint64_t in = (uint64_t)ld_le32(q) | ((uint64_t)ld_le32(q + 4) << 32);
    h ^= xx_round(0, in);
    h = xx_rotl(h, 27) * XXP1 + XXP4;
    q += 8; rem -= 8;
  }
  if (rem >= 4) { h ^= (uint64_t)ld_le32(q) * XXP1; h = xx_rotl(h, 23) * XXP2 + XXP3; q += 4; rem -= 4; }
  while (rem) { h ^= (uint64_t)(*q) * XXP5; h = xx_rotl(h, 11) * XXP1; q++; rem--; }
  h ^= h >> 33; h *= XXP2; h ^= h >> 29; h *= XXP3; h ^= h >> 32;
  return h;
}

} // namespace b200zstd
