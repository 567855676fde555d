// zstd_encode.cu -- batched Zstandard compressor for sm_100a: one persistent warp-CTA per in-flight chunk.
//
// Replaces, for the batch path, ZstdBatchManager::compress_batch -> DefaultZstdManager::compress
// (src/cuda_zstd_manager.cu:5715-5797, 1536-3112 in the reference): find_matches_kernel +
// greedy_parse_kernel + build_sequences_gpu_kernel (src/lz77_parallel.cu:26-268), compress_literals
// (manager.cu:4406-4484, raw only), compress_sequences (:4864-4974, predefined FSE only),
// write_frame_header / write_block (:3998-4106, 4227-4286), compute_xxhash64
// (src/cuda_zstd_xxhash.cu:238-249).  Differences by design (north_star): the hash tables live in
// shared memory, the parse is deterministic, repeat offsets are used, literals are Huffman coded
// and sequence tables are FSE-compressed when that is cheaper, and one launch handles the batch.
//
// Parse ("window" parse).  The warp looks at 32 consecutive positions at a time: every lane hashes
// the bytes at its own position, fetches its candidate(s) (long table, short table or hash chain,
// repeat offset) and measures them; a ballot picks the first lane that holds a match, optionally
// displaced by one of the next lanes (lazy evaluation); the chosen match is extended cooperatively
// (32 x 8 bytes per step), emitted, and the tables are updated for all positions it covers with
// sequential semantics (match_any resolves equal hashes inside a stripe: the highest position
// wins, chain links point to the previous equal-hash lane).  tests/model/enc_model.cpp restates the
// same procedure lane by lane on the host; the two must agree byte for byte.
#include "zstd_common.cuh"
#include "zstd_device_api.h"
#include "zstd_encode_core.cuh"
#include "zstd_encode_entropy.cuh"

namespace b200zstd {

using namespace enc;

constexpr int ENC_THREADS = 32;
constexpr size_t ENC_SMEM_TABLE_MAX = 4096;             // a hash table larger than this costs more in occupancy than shared memory saves in latency

struct EncScratch {          // layout of one CTA's slice of the global workspace
  static constexpr size_t lits_off = 0;
  static constexpr size_t ll_off = BLOCK_BYTES + 64;
  static constexpr size_t ml_off = ll_off + (size_t)MAX_SEQ_PER_BLOCK * 4;
  static constexpr size_t of_off = ml_off + (size_t)MAX_SEQ_PER_BLOCK * 4;
  static constexpr size_t chain_off = of_off + (size_t)MAX_SEQ_PER_BLOCK * 4;
  static constexpr size_t bytes_nochain = chain_off + (size_t)64 * 1024;     // room for the long-hash table (<= 2^15 u16) of DFAST
  static constexpr size_t bytes_chain = chain_off + (size_t)BLOCK_BYTES * 2 + (size_t)64 * 1024;   // chain + room for the primary table
  static constexpr size_t chain_tab_off = chain_off + (size_t)BLOCK_BYTES * 2;
};

size_t encode_cta_scratch_bytes(const EncodeParams &p) {
  size_t b = p.chain_depth > 0 ? EncScratch::bytes_chain : EncScratch::bytes_nochain;
  return (b + 255) & ~(size_t)255;
}
static size_t encode_smem_bytes(const EncodeParams &p) {
  // only the short table lives in shared memory; the long table of DFAST sits in the CTA's L2-resident scratch
  // (10 -> 19 resident warps per SM; the parse is latency-bound, so occupancy buys more than the slower lookup costs)
  size_t tabs = ((size_t)2 << p.hash_log);
  if (tabs > ENC_SMEM_TABLE_MAX) tabs = 0;               // big primary tables also live in the CTA's L2-resident scratch
  size_t ent = sizeof(EntropyWs);
  return (tabs > ent ? tabs : ent) + 16;
}

// ---- unaligned 8-byte read through three aligned 32-bit loads and two funnel shifts, clamped to the chunk's last
// word (bytes past the end are never used by the parser).  (Two 8-byte loads + 64-bit shifts measured 10 % slower.)
struct Src {
  const uint32_t *w;      // 4-byte aligned base (<= chunk start)
  uint32_t delta;         // chunk start - aligned base (0..3)
  uint32_t last_word;     // index of the last word that holds chunk bytes
  __device__ __forceinline__ uint64_t ld64(uint32_t pos) const {      // pos relative to chunk start
    const uint32_t off = pos + delta, a = off >> 2, sh = (off & 3) * 8;
    const uint32_t w0 = __ldg(w + min(a, last_word)), w1 = __ldg(w + min(a + 1, last_word)), w2 = __ldg(w + min(a + 2, last_word));
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    return ((uint64_t)hi << 32) | lo;
  }
};
__device__ __forceinline__ uint32_t common8(uint64_t a, uint64_t b) {
  const uint32_t xl = (uint32_t)a ^ (uint32_t)b, xh = (uint32_t)(a >> 32) ^ (uint32_t)(b >> 32);
  if (xl) return (uint32_t)(__ffs((int)xl) - 1) >> 3;
  if (xh) return 4u + ((uint32_t)(__ffs((int)xh) - 1) >> 3);
  return 8u;
}

struct ParseCtx {
  Src src;
  const uint8_t *chunk;
  uint32_t blk_off, bn, ilimit;
  uint16_t *tab1, *tab2, *chain;
  bool t1_global;         // tab1 lives in global scratch: L2-only accesses
  EncodeParams P;
};

// forward length of the match (s, s - off) inside the block, given `have` (<= 8) verified bytes;
// the whole warp compares 32 x 8 bytes per round.  Uniform arguments, uniform result.
__device__ uint32_t extend_warp(const ParseCtx &C, uint32_t s, uint32_t off, uint32_t have, int lane) {
  if (have < 8) return have;
  uint32_t len = 8;
  for (;;) {
    const uint32_t p = s + len + 8u * (uint32_t)lane;
    uint32_t c = 0;
    if (p < C.bn) {
      c = common8(C.src.ld64(C.blk_off + p), C.src.ld64(C.blk_off + p - off));
      const uint32_t room = C.bn - p;
      if (c > room) c = room;
    }
    const uint32_t stop = __ballot_sync(0xffffffffu, c < 8);
    if (stop) {
      const int j = __ffs(stop) - 1;
      return len + 8u * (uint32_t)j + __shfl_sync(0xffffffffu, c, j);
    }
    len += 256;
  }
}

// two matches measured together: lanes 0-15 walk (s1, off1), lanes 16-31 walk (s2, off2), 16 x 8 bytes per round each
__device__ void extend_pair(const ParseCtx &C, uint32_t s1, uint32_t off1, uint32_t have1, uint32_t s2, uint32_t off2, uint32_t have2,
                            int lane, uint32_t *len1, uint32_t *len2) {
  const bool second = lane >= 16;
  const uint32_t s = second ? s2 : s1, off = second ? off2 : off1, l16 = (uint32_t)(lane & 15);
  uint32_t la = have1, lb = have2;
  bool run_a = have1 >= 8, run_b = have2 >= 8;
  uint32_t base_a = 8, base_b = 8;
  while (run_a || run_b) {
    const uint32_t base = second ? base_b : base_a;
    const bool act = second ? run_b : run_a;
    const uint32_t p = s + base + 8u * l16;
    uint32_t c = 0;
    if (act && p < C.bn) {
      c = common8(C.src.ld64(C.blk_off + p), C.src.ld64(C.blk_off + p - off));
      const uint32_t room = C.bn - p;
      if (c > room) c = room;
    }
    const uint32_t stop = __ballot_sync(0xffffffffu, c < 8);
    if (run_a) {
      const uint32_t sa = stop & 0xFFFFu;
      if (sa) { const int j = __ffs(sa) - 1; la = base_a + 8u * (uint32_t)j + __shfl_sync(0xffffffffu, c, j); run_a = false; }
      else base_a += 128;
    }
    if (run_b) {
      const uint32_t sb = stop >> 16;
      if (sb) { const int j = __ffs(sb) - 1; lb = base_b + 8u * (uint32_t)j + __shfl_sync(0xffffffffu, c, 16 + j); run_b = false; }
      else base_b += 128;
    }
  }
  *len1 = la; *len2 = lb;
}

// per-lane forward length capped at `cap` (chain levels)
__device__ __forceinline__ uint32_t extend_lane(const ParseCtx &C, uint32_t pos, uint32_t off, uint32_t cap) {
  uint32_t len = 8;
  while (len < cap && pos + len < C.bn) {
    uint32_t c = common8(C.src.ld64(C.blk_off + pos + len), C.src.ld64(C.blk_off + pos + len - off));
    const uint32_t room = C.bn - pos - len;
    if (c > room) c = room;
    len += c;
    if (c < 8) break;
  }
  return len < cap ? len : cap;
}

// table update for one stripe of up to 32 consecutive positions [p0, p0+cnt): sequential semantics.
// h1 / h2 are the lane's hashes of position p0+lane (only read where `act`).
template <int S> __device__ __forceinline__ void insert_hashed(const ParseCtx &C, uint32_t p0, bool act, uint32_t h1_in, uint32_t h2_in, int lane) {
  const uint32_t pos = p0 + (uint32_t)lane;
  // inactive lanes get unique keys above any real hash so they never group with active lanes
  const uint32_t h1 = act ? h1_in : 0x80000000u + (uint32_t)lane;
  const uint32_t g1 = __match_any_sync(0xffffffffu, h1);
  if (act) {
    const uint32_t lower = g1 & lanemask_lt();
    if (S == 2) {
      const uint32_t prev = lower ? (p0 + (uint32_t)(31 - __clz(lower))) : (uint32_t)(S != 1 ? __ldcg(C.tab1 + h1) : C.tab1[h1]);
      C.chain[pos] = (uint16_t)((pos - prev) & 0xFFFF);
    }
  }
  __syncwarp();
  if (act && (g1 >> lane) == 1u) { if (S != 1) __stcg(C.tab1 + h1, (uint16_t)pos); else C.tab1[h1] = (uint16_t)pos; }
  if (S == 1) {
    const uint32_t h2 = act ? h2_in : 0x80000000u + (uint32_t)lane;
    const uint32_t g2 = __match_any_sync(0xffffffffu, h2);
    if (act && (g2 >> lane) == 1u) __stcg(C.tab2 + h2, (uint16_t)pos);
  }
  __syncwarp();
}
template <int S> __device__ __forceinline__ void insert_stripe(const ParseCtx &C, uint32_t p0, uint32_t cnt, int lane) {
  const uint32_t pos = p0 + (uint32_t)lane;
  const bool act = (uint32_t)lane < cnt && pos < C.ilimit;
  uint32_t h1 = 0, h2 = 0;
  if (act) {
    const uint64_t v = C.src.ld64(C.blk_off + pos);
    h1 = hash_short(v, S == 2 ? 4 : 5, C.P.hash_log);
    if (S == 1) h2 = hash_long(v, C.P.long_log);
  }
  insert_hashed<S>(C, p0, act, h1, h2, lane);
}

// S: strategy class fixed at compile time -- 0 FAST (one table, L2-resident), 1 DFAST (short table in shared memory + long
// table), 2 chain levels (L2-resident table + hash chain) -- so that the parse carries neither the branches nor the
// registers of the other two
// LZ: lanes behind the first match that may replace it (lazy depth 0..2).  Hashed bytes and the shortest match follow the class.
template <int S, int LZ>
__global__ void __launch_bounds__(ENC_THREADS, 24) zstd_encode_batch_kernel(EncodeArgs A, size_t cta_scratch) {
  constexpr int HASH_BYTES = S == 2 ? 4 : 5, MIN_MATCH = S == 2 ? 4 : 5;
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t s_chunk;
  const int lane = threadIdx.x;
  const EncodeParams P = A.prm;
  uint8_t *const scratch = A.scratch + (size_t)blockIdx.x * cta_scratch;
  uint8_t *const lits = scratch + EncScratch::lits_off;
  uint32_t *const s_ll = (uint32_t *)(scratch + EncScratch::ll_off);
  uint32_t *const s_ml = (uint32_t *)(scratch + EncScratch::ml_off);
  uint32_t *const s_of = (uint32_t *)(scratch + EncScratch::of_off);
  EntropyWs &W = *reinterpret_cast<EntropyWs *>(smem);

  for (;;) {
    if (lane == 0) {
      const uint32_t idx = atomicAdd(A.counter, 1u);
      s_chunk = A.list ? (idx < *A.list_count ? A.list[idx] : 0xFFFFFFFFu) : idx;
    }
    __syncwarp();
    const uint32_t chunk_id = s_chunk;
    __syncwarp();
    if (chunk_id >= A.n) break;

    const uint8_t *const chunk = (const uint8_t *)A.in_ptrs[chunk_id];
    const size_t n = A.in_sizes[chunk_id];
    uint8_t *const dst = (uint8_t *)A.out_ptrs[chunk_id];
    const size_t cap = A.out_sizes[chunk_id];
    uint32_t status = ST_OK;
    size_t op = 0;
    if (!chunk || !dst) status = ST_INVALID_PARAMETER;
    else if (n == 0) status = ST_INVALID_PARAMETER;                       // reference: manager.cu:1554-1558
    else if (n > 0xFFFF0000ull) status = ST_UNSUPPORTED;
    const size_t nblocks = (n + BLOCK_BYTES - 1) / BLOCK_BYTES;
    // block mode: the items are the consecutive blocks of ONE frame (a large single buffer cut into independent
    // 128 KB blocks, one warp each); an item emits block header + payload only, the caller assembles the frame
    const bool blocks_only = A.block_mode != 0;
    if (status == ST_OK && blocks_only && n > BLOCK_BYTES) status = ST_INVALID_PARAMETER;
    if (status == ST_OK && cap < (blocks_only ? 0 : (size_t)frame_header_size(n) + (P.checksum ? 4 : 0)) + n + 3 * nblocks) status = ST_BUFFER_TOO_SMALL;

    if (status == ST_OK) {
      ParseCtx C;
      C.P = P;
      C.chunk = chunk;
      C.src.w = (const uint32_t *)((uintptr_t)chunk & ~(uintptr_t)3);
      C.src.delta = (uint32_t)((uintptr_t)chunk & 3);
      C.src.last_word = (uint32_t)((n - 1 + C.src.delta) >> 2);
      C.t1_global = S != 1;
      // scratch homes: [chain_off, +64 KB) long table (DFAST) or primary table (FAST); chain levels keep the chain there
      // and the primary table right behind it
      C.tab1 = S == 1 ? (uint16_t *)smem
                            : (uint16_t *)(scratch + (P.chain_depth > 0 ? EncScratch::chain_tab_off : EncScratch::chain_off + (P.long_log ? 32 * 1024 : 0)));
      C.tab2 = S == 1 ? (uint16_t *)(scratch + EncScratch::chain_off) : nullptr;
      C.chain = S == 2 ? (uint16_t *)(scratch + EncScratch::chain_off) : nullptr;

      if (!blocks_only) {
        if (lane == 0) op = write_frame_header(dst, n, P.checksum != 0);
        op = __shfl_sync(0xffffffffu, (unsigned long long)op, 0);
      }
      // a block encoded on its own does not know the repeat offsets the decoder will hold when it gets there: 0 =
      // unknown, never matched against and never equal to a real offset, so such a block only uses the history it
      // builds itself (the decoder's own history shifts the same way, RFC 8878 3.1.1.5)
      uint32_t rep[3] = {1, 4, 8};
      if (blocks_only && chunk_id != 0) { rep[0] = 0; rep[1] = 0; rep[2] = 0; }
      size_t ip_chunk = 0;
      while (ip_chunk < n) {
        const uint32_t bn = (uint32_t)min((size_t)BLOCK_BYTES, n - ip_chunk);
        const bool last = blocks_only ? chunk_id + 1 == A.n : ip_chunk + bn == n;
        const uint32_t blk_off = (uint32_t)ip_chunk;
        // ---- RLE block? ----
        {
          const uint8_t b0 = chunk[blk_off];
          bool same = true;
          for (uint32_t k = lane; k < bn && same; k += 32) same = chunk[blk_off + k] == b0;
          if (__all_sync(0xffffffffu, same) && bn > 1) {
            if (lane == 0) { write_block_header(dst + op, last, 1, bn); dst[op + 3] = b0; }
            op += 4;
            ip_chunk += bn;
            continue;
          }
        }
        // ---- parse ----
        C.blk_off = blk_off; C.bn = bn; C.ilimit = bn > 8 ? bn - 8 : 0;
        {
          if (S == 1) {
            const uint32_t words = ((uint32_t)2 << P.hash_log) >> 2;
            uint32_t *z = (uint32_t *)smem;
            for (uint32_t k = lane; k < words; k += 32) z[k] = 0;
          } else {
            uint4 *z1 = (uint4 *)C.tab1;
            const uint32_t vecs = ((uint32_t)2 << P.hash_log) >> 4;
            for (uint32_t k = lane; k < vecs; k += 32) __stcg(z1 + k, make_uint4(0, 0, 0, 0));
          }
          if (S == 1) {
            uint4 *z2 = (uint4 *)C.tab2;
            const uint32_t vecs = ((uint32_t)2 << P.long_log) >> 4;
            for (uint32_t k = lane; k < vecs; k += 32) __stcg(z2 + k, make_uint4(0, 0, 0, 0));
          }
          // "ghost" candidates of never-written buckets are positions 0 and 65536: their chain links
          // must read as end-of-chain until those positions are really inserted
          if (S == 2 && lane == 0) { C.chain[0] = 0; if (bn > 65536) C.chain[65536] = 0; }
        }
        __syncwarp();
        const uint32_t rep_save0 = rep[0], rep_save1 = rep[1], rep_save2 = rep[2];
        uint32_t ip = 0, anchor = 0, nseq = 0, nlit = 0;
        while (ip < C.ilimit && nseq < MAX_SEQ_PER_BLOCK) {
          const uint32_t pos = ip + (uint32_t)lane;
          uint32_t best = 0, bo = 0, wh1 = 0, wh2 = 0;        // wh*: this position's hashes, reused by the table update
          bool has = false;
          if (pos < C.ilimit) {
            const uint64_t v = C.src.ld64(blk_off + pos);
            wh1 = hash_short(v, HASH_BYTES, P.hash_log);
            if (S == 1) wh2 = hash_long(v, P.long_log);
            if (S != 2) {
              // table candidates and the repeat offset: positions first, then all loads in flight together
              int64_t c2 = -1, c1;
              if (S == 1) {
                c2 = (int64_t)((pos & ~0xFFFFu) | __ldcg(C.tab2 + wh2));
                if (c2 >= (int64_t)pos) c2 -= 0x10000;
              }
              c1 = (int64_t)((pos & ~0xFFFFu) | (S != 1 ? __ldcg(C.tab1 + wh1) : C.tab1[wh1]));
              if (c1 >= (int64_t)pos) c1 -= 0x10000;
              const bool vr = rep[0] != 0 && blk_off + pos >= rep[0];
              const uint64_t x2 = c2 >= 0 ? C.src.ld64(blk_off + (uint32_t)c2) : ~v;
              const uint64_t x1 = c1 >= 0 ? C.src.ld64(blk_off + (uint32_t)c1) : ~v;
              const uint64_t xr = vr ? C.src.ld64(blk_off + pos - rep[0]) : ~v;
              if (c2 >= 0 && common8(v, x2) == 8) { best = 8; bo = pos - (uint32_t)c2; }
              if (c1 >= 0) {
                const uint32_t l = common8(v, x1);
                if (l >= (uint32_t)MIN_MATCH && l > best) { best = l; bo = pos - (uint32_t)c1; }
              }
              if (vr) {
                const uint32_t l = common8(v, xr);
                if (l >= 4 && l + P.rep_bonus > best) { best = l; bo = rep[0]; }
              }
            } else {
              int64_t c = (int64_t)((pos & ~0xFFFFu) | (S != 1 ? __ldcg(C.tab1 + wh1) : C.tab1[wh1]));
              if (c >= (int64_t)pos) c -= 0x10000;
              int depth = P.chain_depth;
              while (depth-- > 0 && c >= 0) {
                uint32_t l = common8(v, C.src.ld64(blk_off + (uint32_t)c));
                if (l == 8) l = extend_lane(C, pos, pos - (uint32_t)c, P.lane_cap);
                if (l >= (uint32_t)MIN_MATCH && l > best) { best = l; bo = pos - (uint32_t)c; }
                const uint32_t d = C.chain[(uint32_t)c];
                if (d == 0) break;
                c -= d;
              }
              if (rep[0] != 0 && blk_off + pos >= rep[0]) {
                uint32_t l = common8(v, C.src.ld64(blk_off + pos - rep[0]));
                if (l == 8) l = extend_lane(C, pos, rep[0], P.lane_cap);
                if (l >= 4 && l + P.rep_bonus > best) { best = l; bo = rep[0]; }
              }
            }
            has = best >= (uint32_t)MIN_MATCH || (bo == rep[0] && best >= 4);
          }
          // the input is read once, front to back: keep the line four windows ahead on its way from HBM
          if (lane == 0 && ip + 640 < bn) asm volatile("prefetch.global.L2 [%0];" ::"l"(chunk + blk_off + ip + 512));
          const uint32_t mask = __ballot_sync(0xffffffffu, has);
          if (mask == 0) {
            insert_hashed<S>(C, ip, pos < C.ilimit, wh1, wh2, lane);
            ip += 32;
            continue;
          }
          const int f = __ffs(mask) - 1;
          uint32_t s = ip + (uint32_t)f;
          uint32_t off = __shfl_sync(0xffffffffu, bo, f);
          const uint32_t bl_f = __shfl_sync(0xffffffffu, best, f);
          uint32_t len;
          uint32_t len_g1 = 0;
          const bool g1 = LZ >= 1 && f + 1 < 32 && ((mask >> (f + 1)) & 1);
          if (g1) {
            // first match and its lazy rival measured in the same memory round trips: 16 lanes each
            const uint32_t bl_g = __shfl_sync(0xffffffffu, best, f + 1), off_g = __shfl_sync(0xffffffffu, bo, f + 1);
            extend_pair(C, s, off, bl_f < 8 ? bl_f : 8, s + 1, off_g, bl_g < 8 ? bl_g : 8, lane, &len, &len_g1);
          } else len = extend_warp(C, s, off, bl_f < 8 ? bl_f : 8, lane);
          #pragma unroll
          for (int step = 1; step <= LZ; step++) {
            const int g = f + step;
            if (g >= 32 || !((mask >> g) & 1)) continue;
            const uint32_t bl_g = __shfl_sync(0xffffffffu, best, g), off2 = __shfl_sync(0xffffffffu, bo, g);
            if (bl_g < 8 && bl_g <= len) continue;
            const uint32_t s2 = ip + (uint32_t)g;
            const uint32_t len2 = (step == 1) ? len_g1 : extend_warp(C, s2, off2, bl_g < 8 ? bl_g : 8, lane);
            const int gain1 = (int)len * 4 - hb32(off + 1) + 3 * step + (off == rep[0] ? hb32(off + 1) : 0);
            const int gain2 = (int)len2 * 4 - hb32(off2 + 1) + (off2 == rep[0] ? hb32(off2 + 1) : 0);
            if (gain2 > gain1) { s = s2; off = off2; len = len2; }
          }
          // backward extension into the pending literals
          for (;;) {
            const uint32_t k = (uint32_t)lane;
            bool m = false;
            if (k < s - anchor && (uint64_t)k + off < (uint64_t)blk_off + s)
              m = chunk[blk_off + s - 1 - k] == chunk[blk_off + s - 1 - k - off];
            const uint32_t bal = __ballot_sync(0xffffffffu, m);
            const uint32_t ext = (bal == 0xffffffffu) ? 32u : (uint32_t)(__ffs(~bal) - 1);
            s -= ext; len += ext;
            if (ext < 32) break;
          }
          // emit ...
          const uint32_t llen = s - anchor;
          for (uint32_t k = lane; k < llen; k += 32) lits[nlit + k] = chunk[blk_off + anchor + k];
          const uint32_t code = offset_to_code(off, llen, rep);
          if (lane == 0) { s_ll[nseq] = llen; s_ml[nseq] = len; s_of[nseq] = code; }
          nlit += llen; nseq++;
          // ... and the tables: every position up to the end of the match, in order.  The window's own positions (before
          // the match and under it) already have their hashes: one update for all of them -- the same table state as
          // separate updates, since the highest position of a hash wins either way -- and only what lies beyond the
          // window is loaded and hashed again
          {
            const uint32_t end = s + len;
            if (true) {                                   // every level inserts all covered positions (P.insert_all)
              insert_hashed<S>(C, ip, (uint32_t)lane < min(end - ip, 32u) && pos < C.ilimit, wh1, wh2, lane);
              for (uint32_t p = ip + 32; p < end; p += 32) insert_stripe<S>(C, p, min(end - p, 32u), lane);
            } else {
              const uint32_t from = s > ip ? s : ip;
              if (s > ip) insert_hashed<S>(C, ip, (uint32_t)lane < min(s - ip, 32u) && pos < C.ilimit, wh1, wh2, lane);
              insert_stripe<S>(C, from, 1, lane); if (end >= 2) insert_stripe<S>(C, end - 2, 1, lane);
            }
          }
          ip = anchor = s + len;
        }
        {
          const uint32_t rest = bn - anchor;
          for (uint32_t k = lane; k < rest; k += 32) lits[nlit + k] = chunk[blk_off + anchor + k];
          nlit += rest;
        }
        __syncwarp();
        __threadfence_block();
        // ---- entropy stage (the hash tables are dead: the same shared memory now holds EntropyWs) ----
        uint32_t payload = 0;
        if (nseq < MAX_SEQ_PER_BLOCK) payload = entropy_stage_warp(W, lits, nlit, s_ll, s_ml, s_of, nseq, dst + op + 3, bn - 1, lane);
        if (payload == 0 || payload >= bn) {
          rep[0] = rep_save0; rep[1] = rep_save1; rep[2] = rep_save2;
          if (lane == 0) write_block_header(dst + op, last, 0, bn);
          for (uint32_t k = lane; k < bn; k += 32) dst[op + 3 + k] = chunk[blk_off + k];
          op += 3 + (size_t)bn;
        } else {
          if (lane == 0) write_block_header(dst + op, last, 2, payload);
          op += 3 + (size_t)payload;
        }
        __syncwarp();
        ip_chunk += bn;
      }
      if (P.checksum && !blocks_only) {
        // hash in <= 1 GiB pieces is not needed: chunks are far below 4 GiB (checked above)
        const uint64_t h = xxh64_warp(chunk, (uint32_t)n, lane);
        if (lane == 0) { dst[op] = (uint8_t)h; dst[op + 1] = (uint8_t)(h >> 8); dst[op + 2] = (uint8_t)(h >> 16); dst[op + 3] = (uint8_t)(h >> 24); }
        op += 4;
      }
    }
    __syncwarp();
    if (lane == 0) {
      A.out_sizes[chunk_id] = status == ST_OK ? op : 0;
      if (A.statuses) A.statuses[chunk_id] = status;
    }
  }
}

// strategy class of a parameter set; the shared-memory / L2 placement of the primary table is tied to it
static int strategy_class(const EncodeParams &p) { return p.long_log ? 1 : p.chain_depth > 0 ? 2 : 0; }
static bool class_consistent(const EncodeParams &p) {
  const bool t1_in_smem = ((size_t)2 << p.hash_log) <= ENC_SMEM_TABLE_MAX;
  const int s = strategy_class(p);
  return t1_in_smem == (s == 1) && p.hash_bytes == (s == 2 ? 4 : 5) && p.min_match == (s == 2 ? 4 : 5) && p.insert_all == 1 &&
         p.lazy == (s == 0 ? 0 : s == 1 ? 1 : p.lazy) && p.lazy >= 0 && p.lazy <= 2;
}
template <int S, int LZ> static cudaError_t launch_class(const EncodeArgs &args, int grid, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(zstd_encode_batch_kernel<S, LZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  zstd_encode_batch_kernel<S, LZ><<<grid, ENC_THREADS, smem, stream>>>(args, encode_cta_scratch_bytes(args.prm));
  return cudaGetLastError();
}
// the five instantiations: FAST, DFAST, and the chain levels with lazy depth 0 / 1 / 2
static int kernel_variant(const EncodeParams &p) { const int s = strategy_class(p); return s < 2 ? s : 2 + p.lazy; }
cudaError_t launch_encode_batch(const EncodeArgs &args, int grid, cudaStream_t stream) {
  if (args.n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(args.counter, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  return launch_encode_batch_nomemset(args, grid, stream);
}
cudaError_t launch_encode_batch_nomemset(const EncodeArgs &args, int grid, cudaStream_t stream) {
  if (args.n == 0) return cudaSuccess;
  if (!class_consistent(args.prm)) return cudaErrorInvalidValue;
  const size_t smem = encode_smem_bytes(args.prm);
  switch (kernel_variant(args.prm)) {
    case 0: return launch_class<0, 0>(args, grid, smem, stream);
    case 1: return launch_class<1, 1>(args, grid, smem, stream);
    case 2: return launch_class<2, 0>(args, grid, smem, stream);
    case 3: return launch_class<2, 1>(args, grid, smem, stream);
    default: return launch_class<2, 2>(args, grid, smem, stream);
  }
}

static int g_enc_ctas_cap = 0;                      // 0 = whatever fits; tools/tune_enc.py sweeps it
int encode_ctas_per_sm(const EncodeParams &prm) {
  int n = 0;
  const size_t smem = encode_smem_bytes(prm);
  cudaError_t e;
  switch (kernel_variant(prm)) {
    case 0: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, zstd_encode_batch_kernel<0, 0>, ENC_THREADS, smem); break;
    case 1: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, zstd_encode_batch_kernel<1, 1>, ENC_THREADS, smem); break;
    case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, zstd_encode_batch_kernel<2, 0>, ENC_THREADS, smem); break;
    case 3: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, zstd_encode_batch_kernel<2, 1>, ENC_THREADS, smem); break;
    default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, zstd_encode_batch_kernel<2, 2>, ENC_THREADS, smem); break;
  }
  if (e != cudaSuccess || n < 1) n = 8;
  if (g_enc_ctas_cap > 0 && n > g_enc_ctas_cap) n = g_enc_ctas_cap;
  return n;
}

} // namespace b200zstd
extern "C" void cuda_zstd_b200_tune_enc_ctas(int v) { b200zstd::g_enc_ctas_cap = v; }
