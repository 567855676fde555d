// zstd_decode_fast.cu -- batch fast path of the Zstandard decoder for sm_100a.
//
// The general decoder (zstd_decode.cu) keeps one chunk per CTA and is bound by the single lane that walks the FSE
// sequence bitstream.  The serial parts of Zstandard -- one Huffman stream, one interleaved FSE sequence stream -- cannot
// be split, but a batch has tens of thousands of them.  The fast path therefore runs every serial stream of the WHOLE
// BATCH as its own thread with its table in SHARED memory, and everything else warp- or piece-parallel:
//
//   KP prepare    warp per chunk: frame / block / section headers, pool allocation, the Huffman decode table and the three
//                 packed FSE decode tables, built in shared-memory staging and written to the chunk's slot (global, L2)
//   KA literals   28 chunks per CTA, two CTAs per SM: the Huffman tables arrive by cp.async.bulk (TMA) copies on an
//                 mbarrier, then one THREAD per Huffman stream decodes against them
//   KB sequences  56 chunks per CTA (the shared memory of an SM): the FSE tables arrive the same way, then one LANE per
//                 chunk walks its interleaved sequence bitstream and streams out 16-byte records
//   (a first version kept the tables in global memory: every lookup missed L2 -- 7.8 GB of DRAM reads per GiB decoded --
//   hence shared memory; building the tables inside KA / KB left their SMs idle for a third of each pass, hence KP)
//   KC execute    one warp per chunk, one LANE per sequence: record range checks, then literal runs and every match whose
//                 source lies before the current group of 32 sequences are copied as destination-aligned 16-byte pieces;
//                 the matches that depend on the group itself are replayed in order by the whole warp.  Runs on a side
//                 stream beside KB of the next sub-wave.
//
// It handles the batch case: a chunk that is exactly one frame with one block (what libzstd and this library's compressor
// emit for <= 128 KB chunks) -- or, in bare-block mode, one block of a multi-block frame that launch_split_frame cut into
// units.  Anything else (several blocks or frames, skippable frames, more than FAST_SEQ_CAP sequences) is appended to a
// device-side list and decoded by the general kernel in the same call.  Same reference functions replaced as zstd_decode.cu.
#include <cstdio>
#include <cstdlib>
#include <atomic>

#include "zstd_common.cuh"
#include "zstd_decode_tables.cuh"
#include "zstd_device_api.h"

namespace b200zstd {

// KA takes the Huffman chunks in descending order of literal count (buckets of 2 KB, counting sort: KP ranks, the order kernel
// places), so that the streams a group decodes side by side are of one length and the long ones start first.
constexpr uint32_t KA_BUCKETS = 64, KA_BUCKET_SHIFT = 11, KC_BUCKET_SHIFT = 7;      // (KC: buckets of 128 sequences, the same sort per sub-wave)
struct __align__(16) FastDesc {
  uint32_t state;        // 0 fast path continues, 1 finished in prep, 2 routed to the general kernel
  uint32_t status;
  uint32_t content;      // frame content size, 0xFFFFFFFF when the header has none
  uint32_t cap;          // output capacity (clamped to 32 bits)
  uint32_t lit_type;     // 0 raw, 1 rle, 2 huffman
  uint32_t lit_size;
  uint32_t lit_src;      // raw / rle literals: offset of the bytes inside the frame
  uint32_t n_streams;    // huffman streams: 0, 1 or 4
  uint32_t huf_log;
  uint32_t seg;          // literals decoded by each of the first three streams
  uint32_t st_off[4];
  uint32_t st_len[4];
  uint32_t nseq;
  uint32_t bits_off, bits_len;
  uint32_t ll_log, of_log, ml_log;
  uint32_t ck_off;       // offset of the 4-byte content checksum, 0 = none
  uint32_t tab_off;      // offset of the first FSE table description
  uint32_t blk_end;      // offset one past the block
  uint32_t modes;        // symbol compression modes byte
  uint32_t seq_status, out_end, lit_end;      // written by the sequence thread
  uint32_t lit_status[4];                     // written by the Huffman threads
  uint32_t lit_slot, seq_slot;                // pool offsets in 16-byte units
  uint32_t ka_key;                            // Huffman chunks: literal-count bucket << 16 | rank inside the bucket (KA's work order)
  uint32_t kc_key;                            // every chunk: sequence-count bucket << 16 | rank inside the bucket of its sub-wave (KC's work order)
};
// The tail of the descriptor area holds KC's per-chunk hand-over words: [0] parts of the chunk executed so far, [1] the status
// the finished parts arrived at (zeroed by KP; see zstd_fast_exec_kernel).
constexpr uint32_t FAST_PROG_OFF = 160;
static_assert(sizeof(FastDesc) <= FAST_PROG_OFF && FAST_PROG_OFF + 8 <= FAST_DESC_BYTES, "descriptor slot too small");

struct ChunkSlot {
  uint8_t *base, *lit_pool, *seq_pool;
  __device__ __forceinline__ FastDesc *desc() const { return reinterpret_cast<FastDesc *>(base); }
  __device__ __forceinline__ uint8_t *seq_info() const { return base + FAST_DESC_BYTES; }          // SeqInfo, 64 B reserved
  __device__ __forceinline__ uint8_t *huf_tab() const { return base + FAST_DESC_BYTES + 64; }      // 2048 x uint16 Huffman decode table
  __device__ __forceinline__ uint8_t *seq_tabs() const { return base + FAST_DESC_BYTES + 64 + 4096; }   // 2560 B uint16 plane, 1280 B uint8 plane
  __device__ __forceinline__ uint8_t *lits() const { return lit_pool + (size_t)desc()->lit_slot * 16; }
  __device__ __forceinline__ uint4 *seqs() const { return reinterpret_cast<uint4 *>(seq_pool + (size_t)desc()->seq_slot * 16); }
};
__device__ __forceinline__ ChunkSlot slot_of(const FastDecodeArgs &F, uint32_t chunk) {
  return ChunkSlot{F.slots + (size_t)chunk * FAST_SLOT_BYTES, F.lit_pool, F.seq_pool};
}
__device__ __forceinline__ uint32_t seg_padded(uint32_t seg) { return (seg + 15u) & ~15u; }

// =================================================================================================
// KP: per-chunk preparation, one warp per chunk -- frame/block/section headers, the Huffman decode table and the three
// packed FSE decode tables, all written to the chunk's slot.  Every step here is a short serial chain per chunk, so the
// kernel runs with every warp slot of the GPU instead of inside the decode kernels, whose shared memory is full of tables.
// =================================================================================================
constexpr int KP_WARPS = 8;
struct __align__(16) HufScratch {                              // per-warp scratch of huf_read_table_warp
  uint16_t *huf;                                               // -> the table being built (shared-memory staging)
  uint8_t weights[256];
  uint32_t huf_ft[64];
  int16_t huf_norm[16];
  uint32_t rank_cnt[16];
  uint32_t rank_start[16];
  int huf_log, huf_valid;
};

// Raw and RLE blocks are finished by the prepare kernel's warp.  The payload of a raw block sits at an arbitrary byte
// offset of its frame (10 for a single-block frame), the destination is usually 16-byte aligned: aligned 32-bit loads
// realigned with funnel shifts feed 16-byte stores, two vectors per lane in flight.  `src_room` = readable bytes from src
// on (the rest of the frame); the words in front of src hold the block header, so reading them is safe.
template <int U>       // U > 1: that many 16-byte vectors per lane in flight
__device__ __forceinline__ void warp_copy_block(uint8_t *dst, const uint8_t *src, uint32_t n, uint32_t src_room, int lane) {
  uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
  if (head > n) head = n;
  for (uint32_t k = lane; k < head; k += 32) dst[k] = src[k];
  // vector j covers src bytes [head + 16 j, head + 16 j + 16) and needs the aligned words up to 3 bytes past them
  uint32_t vecs = (n - head) >> 4;
  while (vecs && head + 16 * vecs + 3 > src_room) vecs--;
  const uint8_t *s0 = src + head;
  const uint32_t sh = (uint32_t)((uintptr_t)s0 & 3) * 8;
  const uint32_t *w = reinterpret_cast<const uint32_t *>((uintptr_t)s0 & ~(uintptr_t)3);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
  uint32_t j = (uint32_t)lane;
  for (; U > 1 && j + 32 * (U - 1) < vecs; j += 32 * U) {
    uint32_t a[U][5];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int q = 0; q < 5; q++) a[u][q] = __ldg(w + 4 * (j + 32 * u) + q);
#pragma unroll
    for (int u = 0; u < U; u++)
      d4[j + 32 * u] = make_uint4(__funnelshift_r(a[u][0], a[u][1], sh), __funnelshift_r(a[u][1], a[u][2], sh),
                                  __funnelshift_r(a[u][2], a[u][3], sh), __funnelshift_r(a[u][3], a[u][4], sh));
  }
  for (; j < vecs; j += 32) {
    uint32_t a[5];
#pragma unroll
    for (int q = 0; q < 5; q++) a[q] = __ldg(w + 4 * j + q);
    d4[j] = make_uint4(__funnelshift_r(a[0], a[1], sh), __funnelshift_r(a[1], a[2], sh), __funnelshift_r(a[2], a[3], sh),
                       __funnelshift_r(a[3], a[4], sh));
  }
  for (uint32_t k = head + 16 * vecs + lane; k < n; k += 32) dst[k] = src[k];
}
__device__ __forceinline__ void warp_fill_block(uint8_t *dst, uint8_t v, uint32_t n, int lane) {
  uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
  if (head > n) head = n;
  for (uint32_t k = lane; k < head; k += 32) dst[k] = v;
  const uint32_t vecs = (n - head) >> 4, x = 0x01010101u * v;
  uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
  for (uint32_t j = lane; j < vecs; j += 32) d4[j] = make_uint4(x, x, x, x);
  for (uint32_t k = head + 16 * vecs + lane; k < n; k += 32) dst[k] = v;
}

__device__ __forceinline__ void prep_chunk(const FastDecodeArgs &F, uint32_t chunk, uint8_t *stage, HufScratch &S, int lane) {
    const DecodeArgs &A = F.base;
    if (lane == 0) S.huf = reinterpret_cast<uint16_t *>(stage);
    __syncwarp();
    const uint8_t *const src = (const uint8_t *)A.in_ptrs[chunk];
    const size_t src_size = A.in_sizes[chunk];
    uint8_t *const dst = (uint8_t *)A.out_ptrs[chunk];
    const size_t dst_cap = A.out_sizes[chunk];
    ChunkSlot slot = slot_of(F, chunk);
    FastDesc D{};
    uint32_t status = ST_OK;
    bool route = false, finished = false;
    uint32_t produced = 0;
    __syncwarp();
    do {
      if (src == nullptr || (dst == nullptr && dst_cap != 0) || src_size < 4) { status = ST_INVALID_PARAMETER; break; }
      uint32_t n, h;
      uint64_t fcs;
      int has_ck;
      if (F.bare_blocks) {
        // a unit cut out of a multi-block frame: no frame header, and it must regenerate exactly its capacity
        if (src_size > 0x7FFFFFFFull || dst_cap > BLOCK_MAX) { status = ST_CORRUPT; break; }
        n = (uint32_t)src_size; h = 0; fcs = dst_cap; has_ck = 0;
      } else {
      const uint32_t magic = ld_le32(src);
      if ((magic & 0xFFFFFFF0u) == ZSTD_SKIP_MAGIC) { route = true; break; }
      if (magic != ZSTD_FRAME_MAGIC) { status = ST_INVALID_MAGIC; break; }
      if (src_size > 0x7FFFFFFFull) { route = true; break; }
      n = (uint32_t)src_size;
      h = 4;
      if (h >= n) { status = ST_CORRUPT; break; }
      const uint32_t fhd = src[h++];
      const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
      has_ck = (fhd >> 2) & 1;
      if (fhd & 0x08) { status = ST_UNSUPPORTED; break; }
      const uint32_t did_size = did_flag == 3 ? 4 : did_flag, fcs_size = fcs_flag == 0 ? single : (1u << fcs_flag);
      if (h + (single ? 0 : 1) + did_size + fcs_size > n) { status = ST_CORRUPT; break; }
      if (!single) { if ((src[h++] >> 3) > 21) { status = ST_UNSUPPORTED; break; } }
      uint32_t dict_id = 0;
      for (uint32_t k = 0; k < did_size; k++) dict_id |= (uint32_t)src[h + k] << (8 * k);
      h += did_size;
      fcs = ~0ull;
      if (fcs_size) {
        fcs = 0;
        for (uint32_t k = 0; k < fcs_size; k++) fcs |= (uint64_t)src[h + k] << (8 * k);
        if (fcs_size == 2) fcs += 256;
        h += fcs_size;
      }
      if (dict_id != 0) { status = ST_DICT_MISMATCH; break; }
      if (fcs != ~0ull && fcs > dst_cap) { status = ST_BUFFER_TOO_SMALL; break; }
      }
      const uint32_t cap = dst_cap > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)dst_cap;
      // ---- the single block ----
      if (n - h < 3) { status = ST_CORRUPT; break; }
      const uint32_t bh = ld_le24(src + h);
      h += 3;
      const int last = F.bare_blocks ? 1 : (int)(bh & 1), btype = (bh >> 1) & 3;
      const uint32_t bsize = bh >> 3;
      if (btype == 3 || bsize > BLOCK_MAX) { status = ST_CORRUPT; break; }
      const uint32_t body = (btype == 1) ? 1u : bsize;
      if (body > n - h) { status = ST_CORRUPT; break; }
      if (!last) { route = true; break; }
      const uint32_t after = h + body;
      if (has_ck && n - after < 4) { status = ST_CORRUPT; break; }
      if (after + (has_ck ? 4u : 0u) != n) { route = true; break; }          // more frames follow
      const uint32_t ck_off = has_ck ? after : 0;
      if (btype == 0 || btype == 1) {
        if (bsize > cap) { status = ST_BUFFER_TOO_SMALL; break; }
        if (fcs != ~0ull && fcs != bsize) { status = ST_CORRUPT; break; }
        if (btype == 0) warp_copy_block<2>(dst, src + h, bsize, n - h, lane);
        else warp_fill_block(dst, src[h], bsize, lane);
        __syncwarp();
        if (has_ck && A.verify_checksum) {
          const uint64_t hs = xxh64_warp(dst, bsize, lane);
          if ((uint32_t)hs != ld_le32(src + ck_off)) { status = ST_CHECKSUM; break; }
        }
        produced = bsize;
        finished = true;
        break;
      }
      if (bsize < 2) { status = ST_CORRUPT; break; }
      const uint8_t *const bp = src + h;
      // -- literals section header --
      const uint32_t b0 = bp[0];
      const int ltype = b0 & 3, sf = (b0 >> 2) & 3;
      uint32_t lit_size, lit_comp = 0, lhs, nstreams = 1;
      if (ltype < 2) {
        lhs = (sf == 1) ? 2 : (sf == 3) ? 3 : 1;
        if (bsize < lhs) { status = ST_CORRUPT; break; }
        lit_size = (lhs == 1) ? (b0 >> 3) : (lhs == 2) ? ((b0 >> 4) | ((uint32_t)bp[1] << 4))
                                                       : ((b0 >> 4) | ((uint32_t)bp[1] << 4) | ((uint32_t)bp[2] << 12));
        lit_comp = (ltype == 0) ? lit_size : 1;
      } else {
        lhs = (sf < 2) ? 3 : (sf == 2) ? 4 : 5;
        nstreams = (sf == 0) ? 1 : 4;
        if (bsize < lhs) { status = ST_CORRUPT; break; }
        if (sf < 2) { lit_size = (b0 >> 4) | (((uint32_t)bp[1] & 0x3F) << 4); lit_comp = ((uint32_t)bp[1] >> 6) | ((uint32_t)bp[2] << 2); }
        else if (sf == 2) { lit_size = (b0 >> 4) | ((uint32_t)bp[1] << 4) | (((uint32_t)bp[2] & 3) << 12); lit_comp = ((uint32_t)bp[2] >> 2) | ((uint32_t)bp[3] << 6); }
        else { lit_size = (b0 >> 4) | ((uint32_t)bp[1] << 4) | (((uint32_t)bp[2] & 0x3F) << 12); lit_comp = ((uint32_t)bp[2] >> 6) | ((uint32_t)bp[3] << 2) | ((uint32_t)bp[4] << 10); }
      }
      if (lit_size > BLOCK_MAX || lhs + lit_comp > bsize) { status = ST_CORRUPT; break; }
      // -- sequences section header --
      uint32_t sp = lhs + lit_comp;
      if (sp >= bsize) { status = ST_CORRUPT; break; }
      uint32_t nseq;
      {
        const uint32_t s0 = bp[sp];
        if (s0 < 128) { nseq = s0; sp += 1; }
        else if (s0 < 255) { if (sp + 2 > bsize) { status = ST_CORRUPT; break; } nseq = ((s0 - 128) << 8) + bp[sp + 1]; sp += 2; }
        else { if (sp + 3 > bsize) { status = ST_CORRUPT; break; } nseq = (uint32_t)bp[sp + 1] + ((uint32_t)bp[sp + 2] << 8) + 0x7F00; sp += 3; }
      }
      uint32_t modes = 0;
      if (nseq) {
        if (sp >= bsize) { status = ST_CORRUPT; break; }
        modes = bp[sp++];
        if (modes & 3) { status = ST_CORRUPT; break; }
      } else if (sp != bsize) { status = ST_CORRUPT; break; }
      if (nseq > FAST_SEQ_CAP) { route = true; break; }
      D.content = fcs == ~0ull ? 0xFFFFFFFFu : (uint32_t)fcs;
      D.cap = cap;
      D.lit_type = ltype >= 2 ? 2u : (uint32_t)ltype;
      D.lit_size = lit_size;
      D.lit_src = h + lhs;
      D.nseq = nseq;
      D.ck_off = ck_off;
      // -- Huffman table + stream layout --
      if (ltype == 3) { status = ST_CORRUPT; break; }                         // treeless needs a previous block
      if (ltype == 2) {
        const uint8_t *hp = bp + lhs;
        uint32_t rem = lit_comp;
        const int used = huf_read_table_warp(S, hp, rem, lane);
        if (used < 0) { status = ST_CORRUPT; break; }
        hp += used; rem -= (uint32_t)used;
        const uint32_t hoff = h + lhs + (uint32_t)used;
        D.huf_log = (uint32_t)S.huf_log;
        D.n_streams = nstreams;
        if (nstreams == 1) { D.seg = lit_size; D.st_off[0] = hoff; D.st_len[0] = rem; }
        else {
          const uint32_t seg = (lit_size + 3) >> 2;
          if (rem < 6 || seg * 3 > lit_size) { status = ST_CORRUPT; break; }
          const uint32_t s1 = ld_le16(hp), s2 = ld_le16(hp + 2), s3 = ld_le16(hp + 4);
          if (6 + s1 + s2 + s3 > rem) { status = ST_CORRUPT; break; }
          D.seg = seg;
          D.st_off[0] = hoff + 6; D.st_len[0] = s1;
          D.st_off[1] = hoff + 6 + s1; D.st_len[1] = s2;
          D.st_off[2] = hoff + 6 + s1 + s2; D.st_len[2] = s3;
          D.st_off[3] = hoff + 6 + s1 + s2 + s3; D.st_len[3] = rem - 6 - s1 - s2 - s3;
        }
      }
      // -- FSE tables are built by KB; remember where their descriptions start --
      D.tab_off = h + sp; D.blk_end = h + bsize; D.modes = modes;
    } while (false);
    __syncwarp();
    if (lane == 0 && !route && !finished && status == ST_OK) {
      // bump-allocate this chunk's literal and sequence storage; pool exhaustion -> general kernel
      const unsigned long long lit_need = D.lit_type == 2 ? (unsigned long long)4 * seg_padded(D.seg ? D.seg : 1) : 0ull;
      const unsigned long long seq_need = (unsigned long long)(D.nseq + 1) * 16;
      // (one pool, one head: a literal-only block may take all of its share for literals, a match-heavy one for records)
      const unsigned long long lo = atomicAdd(&F.pool_heads[0], lit_need + seq_need);
      if (lo + lit_need + seq_need > F.lit_pool_bytes) route = true;
      D.lit_slot = (uint32_t)(lo >> 4); D.seq_slot = (uint32_t)((lo + lit_need) >> 4);
      if (!route && D.lit_type == 2) {
        const uint32_t b = min(D.lit_size >> KA_BUCKET_SHIFT, KA_BUCKETS - 1u);
        D.ka_key = (b << 16) | atomicAdd(F.lit_buckets + b, 1u);
      }
    }
    route = __shfl_sync(0xffffffffu, route ? 1 : 0, 0) != 0;
    if (lane == 0) {
      if (route) {
        D.state = 2;
        F.slow_list[atomicAdd(F.slow_count, 1u)] = chunk;
      } else if (status != ST_OK || finished) {
        D.state = 1;
        A.out_sizes[chunk] = status == ST_OK ? produced : 0;
        if (A.statuses) A.statuses[chunk] = status;
      } else D.state = 0;
      D.status = status;
      if (F.kc_order) {
        // KC takes the chunks of a sub-wave heaviest first (a chunk's parts are a serial chain: the long chains must start early)
        const uint32_t sb = D.state == 0 ? min(D.nseq >> KC_BUCKET_SHIFT, KA_BUCKETS - 1u) : 0u;
        D.kc_key = (sb << 16) | atomicAdd(F.seq_buckets + (chunk / F.sub_chunks) * KA_BUCKETS + sb, 1u);
      }
      *slot.desc() = D;
    }
    __syncwarp();
    // the Huffman table leaves the staging buffer before the FSE build reuses it
    if (slot.desc()->state == 0 && D.lit_type == 2) {
      const uint4 *const from = reinterpret_cast<const uint4 *>(stage);
      uint4 *const to = reinterpret_cast<uint4 *>(slot.huf_tab());
      const uint32_t n16 = (2u << D.huf_log) / 16;                             // 2^log entries x 2 bytes
      for (uint32_t k = lane; k < n16; k += 32) to[k] = from[k];
    }
    __syncwarp();
}

// =================================================================================================
// KA: literals -- 28 chunks' Huffman tables pulled into shared memory by bulk async copies (two CTAs per SM), then one
// thread per Huffman stream
// =================================================================================================
constexpr int KA_GROUP = 24;                                   // chunks per CTA pass: 24 x (4 KB table + 4 stream rings), two CTAs per SM
constexpr int KA_THREADS = 4 * KA_GROUP;
constexpr int KA_RING_WORDS = 32;                              // per stream: the 128 bytes of bitstream around its read position
constexpr size_t KA_SMEM = (size_t)KA_GROUP * 4096 + (size_t)KA_THREADS * KA_RING_WORDS * 4;

__device__ __forceinline__ uint32_t top_bits(uint32_t hi, int skip, int n) { return ((hi << skip) >> 1) >> (31 - n); }   // n in [0,31]
__device__ __forceinline__ uint32_t peek32(const uint32_t *W, int t) {             // bits [t-32, t) of the stream, t may be anything
  const int k = max(t >> 5, 0);
  return __funnelshift_r(W[k - 1], W[k], (uint32_t)t & 31u);
}

__device__ __forceinline__ void fast_decode_huffman(const uint8_t *src, const FastDesc *D, uint8_t *lits, const uint16_t *tab,
                                                    uint32_t k, uint32_t *status_out, uint32_t lead, uint32_t *ring);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

// counting sort, second half: bucket sizes -> start positions (largest bucket first); every Huffman chunk writes its index
// to its place.  The order occupies the END of the slow-list array, last entry first (a chunk is in one list or the other).
__global__ void __launch_bounds__(256) zstd_fast_order_kernel(FastDecodeArgs F) {
  __shared__ uint32_t start[KA_BUCKETS];
  if (threadIdx.x < 32) {
    const uint32_t lane = threadIdx.x;
    const uint32_t hi = F.lit_buckets[KA_BUCKETS - 1 - lane], lo = F.lit_buckets[KA_BUCKETS - 33 - lane];     // descending bucket order
    uint32_t a = hi, b = lo;
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, a, d), y = __shfl_up_sync(0xffffffffu, b, d);
      if (lane >= (uint32_t)d) { a += x; b += y; }
    }
    const uint32_t first_half = __shfl_sync(0xffffffffu, a, 31);
    start[KA_BUCKETS - 1 - lane] = a - hi;
    start[KA_BUCKETS - 33 - lane] = first_half + b - lo;
    if (blockIdx.x == 0 && lane == 31) F.lit_buckets[KA_BUCKETS] = first_half + b;
  }
  __syncthreads();
  const uint32_t chunk = blockIdx.x * 256 + threadIdx.x;
  if (chunk >= F.base.n) return;
  const FastDesc *const D = slot_of(F, chunk).desc();
  if (F.kc_order) {
    const uint32_t key = D->kc_key, b = key >> 16, sub = chunk / F.sub_chunks;
    uint32_t first = 0;
    for (uint32_t k = KA_BUCKETS - 1; k > b; k--) first += F.seq_buckets[sub * KA_BUCKETS + k];
    F.kc_order[sub * F.sub_chunks + first + (key & 0xFFFFu)] = chunk;
  }
  if (D->state != 0 || D->lit_type != 2) return;
  const uint32_t key = D->ka_key;
  F.slow_list[F.base.n - 1 - (start[key >> 16] + (key & 0xFFFFu))] = chunk;
}

__global__ void __launch_bounds__(KA_THREADS) zstd_fast_lit_kernel(FastDecodeArgs F) {
  extern __shared__ __align__(128) uint8_t ka_smem[];
  __shared__ uint32_t s_group;
  __shared__ __align__(8) uint64_t s_bar;
  uint16_t *const tables = reinterpret_cast<uint16_t *>(ka_smem);
  const DecodeArgs &A = F.base;
  if (threadIdx.x == 0) mbar_init(&s_bar, KA_GROUP);           // one arrival per chunk of the group
  uint32_t phase = 0;
  for (;;) {
    if (threadIdx.x == 0) s_group = atomicAdd(F.group_counters + 0, 1u);
    __syncthreads();                                            // also: every thread is done with the previous group's tables
    const uint32_t g0 = s_group * KA_GROUP, n_huf = F.lit_buckets[KA_BUCKETS];
    if (g0 >= n_huf) break;
    const uint32_t c = threadIdx.x >> 2, k = threadIdx.x & 3;
    const bool have = g0 + c < n_huf;
    const uint32_t chunk = F.slow_list[A.n - 1 - (have ? g0 + c : g0)];
    bool work = false;
    ChunkSlot slot = slot_of(F, chunk);
    FastDesc *const D = slot.desc();
    if (c < KA_GROUP) {
      const bool live = have && D->state == 0 && D->lit_type == 2;
      work = live && k < D->n_streams;
      if (k == 0) {                                             // the chunk's first thread pulls its table
        if (live) {
          const uint32_t bytes = 2u << D->huf_log;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive_tx(&s_bar, bytes);
          bulk_g2s(tables + (size_t)c * 2048, slot.huf_tab(), bytes, &s_bar);
        } else mbar_arrive(&s_bar);
      }
    }
    mbar_wait(&s_bar, phase);
    phase ^= 1;
    if (work)
      fast_decode_huffman((const uint8_t *)A.in_ptrs[chunk], D, slot.lits(), tables + (size_t)c * 2048, k, &D->lit_status[k], F.bare_blocks ? 6u : 0u,
                          reinterpret_cast<uint32_t *>(ka_smem + (size_t)KA_GROUP * 4096) + (size_t)threadIdx.x * KA_RING_WORDS);
  }
}

// =================================================================================================
// Huffman stream decode (one thread), table in shared memory
// =================================================================================================
// `lead` = bytes known to be readable in front of src (0 for a frame, 6 for a block unit inside a frame)
// The bitstream reaches the thread through its own 128-byte ring in shared memory, filled by 16-byte `cp.async.cg` copies
// (LDGSTS, past the L1) that run eight granules ahead of the read position.  Direct loads -- every thread of the SM walking
// its own stream -- need one 128-byte L1 line per stream, and 2 x 112 of them beside 224 KB of tables do not fit what is
// left of the L1: ncu on literal-only chunks showed 30 % L1 hits, 13 x the stream bytes fetched from the L2 and the warps
// waiting on those loads 65 % of the time (3.8 ms per 16,384 chunks; 2.2 ms of the 6.1 of a config-5 wave).
__device__ __forceinline__ void cp_async16(uint32_t *smem_dst, const uint32_t *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void fast_decode_huffman(const uint8_t *src, const FastDesc *D, uint8_t *lits, const uint16_t *tab,
                                                    uint32_t k, uint32_t *status_out, uint32_t lead, uint32_t *ring) {
  const uint32_t seg = D->seg, lit_size = D->lit_size;
  const uint32_t count = (D->n_streams == 1) ? lit_size : (k < 3 ? seg : lit_size - 3 * seg);
  uint8_t *dst = lits + (size_t)k * seg_padded(seg);                          // 16-byte aligned segment
  const int sh = 64 - (int)D->huf_log;
  const uint8_t *const p = src + D->st_off[k];
  const uint32_t nb = D->st_len[k];
  // position-based reader (see the sequence loop in KB): the words under bit `t` are taken each round, so there is no
  // refill branch.  W[-1], W[-2] are read under the first stream bits: >= 12 header bytes always precede a Huffman stream.
  bool ok = nb != 0 && D->st_off[k] + lead >= 12 && p[nb - 1] != 0;
  if (ok) {
    const uint32_t *const W = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    const int d = (int)((uintptr_t)p & 3), low = 8 * d;
    int t = 8 * (d + (int)nb - 1) + highbit32(p[nb - 1]);
    uint32_t i = 0;
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
    // Ring numbering: word W[k] is word j = k + bias of an array G that starts on a 16-byte boundary one granule below
    // W's own (never read below W[-2]); granule g = words 4g .. 4g + 3 lives in ring slot g & 7.
    const int bias = (int)(((uintptr_t)W >> 2) & 3) + 4;
    const uint32_t *const G = W - bias;
    const int jmin = bias - 2, glo = jmin >> 2;
    int j = max(t >> 5, 0) + bias;
    // the top granule by word loads (a 16-byte copy could read up to 15 bytes past the stream), seven more by async copies
    for (int w = max((j >> 2) * 4, jmin); w <= j; w++) ring[w & (KA_RING_WORDS - 1)] = G[w];
    int gnext = (j >> 2) - 1;
#pragma unroll
    for (int r = 0; r < 7; r++)
      if (gnext >= glo) { cp_async16(ring + (gnext & 7) * 4, G + gnext * 4); gnext--; }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // Five words ride in registers: q0..q2 = W[k], W[k-1], W[k-2] feed this round, q3 and q4 were taken one round
    // earlier.  A round consumes at most 44 bits, so the window moves down by 0, 1 or 2 words (three selects) and the ring
    // by at most one granule: one copy request per round keeps it full, and the two granules under the window are always
    // the two oldest of the eight, so all but the six youngest requests must have landed.
#define KA_RING(jj) ring[max((jj), jmin) & (KA_RING_WORDS - 1)]
    uint32_t q0 = KA_RING(j), q1 = KA_RING(j - 1), q2 = KA_RING(j - 2), q3 = KA_RING(j - 3), q4 = KA_RING(j - 4);
    for (; i + 4 <= count; i += 4) {                                          // 4 symbols (<= 44 bits) -> one aligned 32-bit store
      const uint32_t s = (uint32_t)t & 31u;
      uint64_t win = ((uint64_t)__funnelshift_r(q1, q0, s) << 32) | __funnelshift_r(q2, q1, s);
      const uint32_t e0 = tab[win >> sh]; win <<= (e0 >> 8);
      const uint32_t e1 = tab[win >> sh]; win <<= (e1 >> 8);
      const uint32_t e2 = tab[win >> sh]; win <<= (e2 >> 8);
      const uint32_t e3 = tab[win >> sh];
      t -= (int)((e0 >> 8) + (e1 >> 8) + (e2 >> 8) + (e3 >> 8));
      d32[i >> 2] = (e0 & 0xFF) | ((e1 & 0xFF) << 8) | ((e2 & 0xFF) << 16) | (e3 << 24);
      const int jn = max(t >> 5, 0) + bias, dk = j - jn;
      const uint32_t n0 = dk == 0 ? q0 : dk == 1 ? q1 : q2;
      const uint32_t n1 = dk == 0 ? q1 : dk == 1 ? q2 : q3;
      const uint32_t n2 = dk == 0 ? q2 : dk == 1 ? q3 : q4;
      q0 = n0; q1 = n1; q2 = n2; j = jn;
      if ((j >> 2) - gnext < 8 && gnext >= glo) {
        cp_async16(ring + (gnext & 7) * 4, G + gnext * 4);
        asm volatile("cp.async.commit_group;" ::: "memory");
        gnext--;
      }
      asm volatile("cp.async.wait_group 6;" ::: "memory");
      q3 = KA_RING(j - 3); q4 = KA_RING(j - 4);
    }
#undef KA_RING
    asm volatile("cp.async.wait_group 0;" ::: "memory");                      // (the ring is reused by the next group's stream)
    for (; i < count; i++) {
      const uint32_t e = tab[((uint64_t)peek32(W, t) << 32) >> sh];
      t -= (int)(e >> 8);
      dst[i] = (uint8_t)e;
    }
    ok = t == low;
  }
  *status_out = ok ? ST_OK : ST_CORRUPT;
}

// =================================================================================================
// KB: FSE tables in shared memory (one warp per chunk), then one lane per sequence stream
// =================================================================================================
constexpr int KB_GROUP = 56;                                   // chunks per CTA pass: 56 x 3.75 KB of packed tables
// The 56 streams of a group are spread over KB_DEC_WARPS warps: a warp stalls whenever ANY of its lanes waits for a word
// of its bitstream, so fewer lanes per warp means fewer shared stalls at the price of more issued instructions
// (measured: see DESIGN.md 4.1).  CUDA_ZSTD_KB_WARPS selects another split for experiments.
constexpr int KB_DEC_WARPS_DEFAULT = 4;
constexpr uint32_t KB_SPLIT_DEFAULT = 2;                       // KB launches per sub-wave (see launch_decode_fast) ...
constexpr uint32_t KB_SPLIT_ONE_SUBWAVE = 4;                   // ... and when the batch is a single sub-wave
constexpr uint32_t KB_SPLIT_MIN_CHUNKS = 1;                    // (measured down to one chunk: 0.88 -> 0.60 ms; CUDA_ZSTD_MIN_SPLIT raises it)
__device__ __forceinline__ constexpr int kb_norm_off(int t) { return t == 0 ? 0 : t == 1 ? 40 : 72; }   // LL 36 | OF 32 | ML 53 normalised counts
struct __align__(16) SeqScratch {                              // per-warp scratch for the table build (432 B: 28 of them fit beside the tables)
  int16_t norm[128];
  uint16_t sym_next[64];
  int tab_log[3], tab_max[3], tab_mode[3];
  uint32_t bits_off, ok;
};
struct SeqInfo { uint32_t ready, ll_log, of_log, ml_log, bits_off, bits_len; };
constexpr size_t KB_TAB_BYTES = 1280 * 3;                      // LL 512 | ML 512 | OF 256 entries: a uint16 plane and a uint8 plane
constexpr size_t KB_SMEM = (size_t)KB_GROUP * KB_TAB_BYTES + 96 * 4;

// packed decode entry, 24 bits in two planes so that 56 chunks' tables fit in one SM's shared memory:
//   t16[state] = nextStateBase[0:9) | nbBits[9:13) | extraBits[13:16) (low 3 bits)      t8[state] = extraBits high 2 bits | symbol << 2
// (nextStateBase < table size <= 512; the base VALUE of a length code is looked up from the symbol in a shared
// 96-entry table: it only feeds the output record, not the bit-position chain)
struct SeqTab { uint16_t *t16; uint8_t *t8; };        // one chunk: t16 = [LL 512 | ML 512 | OF 256], t8 likewise
__device__ __forceinline__ void pack_entry(const SeqTab &T, uint32_t idx, int kind, uint32_t sym, uint32_t next_base, uint32_t nb) {
  const uint32_t xb = kind == 0 ? c_ll_bits[sym] : kind == 1 ? sym : c_ml_bits[sym];
  T.t16[idx] = (uint16_t)(next_base | (nb << 9) | ((xb & 7) << 13));
  T.t8[idx] = kind == 1 ? (uint8_t)xb : (uint8_t)((xb >> 3) | (sym << 2));   // an offset code IS its extra-bit count: no unpacking on the chain
}
// fse_build_warp (zstd_decode_tables.cuh) emitting packed entries in place; `off` = first entry of this table in the planes
__device__ inline void fse_build_warp_packed(const SeqTab &T, uint32_t off, const int16_t *norm, int max_sym, int log, int kind,
                                             uint16_t *sym_next, int lane) {
  const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
  uint8_t *cell = T.t8 + off;                       // the byte plane doubles as the symbol-of-cell scratch
  uint8_t *item_sym = (uint8_t *)(T.t16 + off);     // and the 16-bit plane, not written before the last pass, holds the sorted symbols
  // symbols sorted by value with multiplicity (item_sym), low-probability symbols parked at the top cells.  Lane l owns symbols
  // l and l + 32: two warp scans give every symbol its first item, the symbol is written there, and a running maximum over
  // the item array (symbols ascend) fills the rest -- ~200 warp instructions where a loop over the symbols took ~900
  const int s0 = lane, s1 = lane + 32;
  const int c0 = s0 <= max_sym ? norm[s0] : 0, c1 = s1 <= max_sym ? norm[s1] : 0;
  const uint32_t neg0 = __ballot_sync(0xffffffffu, c0 == -1), neg1 = __ballot_sync(0xffffffffu, c1 == -1);
  const int p0 = c0 > 0 ? c0 : 0, p1 = c1 > 0 ? c1 : 0;
  int i0 = p0, i1 = p1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
    if (lane >= o) { i0 += t0; i1 += t1; }
  }
  const int st0 = i0 - p0, st1 = __shfl_sync(0xffffffffu, i0, 31) + i1 - p1;
  for (int k = lane * 4; k < size; k += 128) *reinterpret_cast<uint32_t *>(item_sym + k) = 0u;
  __syncwarp();
  if (s0 <= max_sym) {
    if (c0 == -1) { cell[size - 1 - __popc(neg0 & lanemask_lt())] = (uint8_t)s0; sym_next[s0] = 1; }
    else { sym_next[s0] = (uint16_t)c0; if (c0 > 0) item_sym[st0] = (uint8_t)s0; }
  }
  if (s1 <= max_sym) {
    if (c1 == -1) { cell[size - 1 - __popc(neg0) - __popc(neg1 & lanemask_lt())] = (uint8_t)s1; sym_next[s1] = 1; }
    else { sym_next[s1] = (uint16_t)c1; if (c1 > 0) item_sym[st1] = (uint8_t)s1; }
  }
  const int high = size - 1 - __popc(neg0) - __popc(neg1);
  __syncwarp();
  {
    const int seg = size >> 5, base = lane * seg;              // size >= 32: sequence tables have accuracy log >= 5
    uint32_t m = 0;
    for (int k = 0; k < seg; k++) m = max(m, (uint32_t)item_sym[base + k]);
    uint32_t pm = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, pm, o); if (lane >= o) pm = max(pm, t); }
    uint32_t v = __shfl_up_sync(0xffffffffu, pm, 1);
    if (lane == 0) v = 0;
    for (int k = 0; k < seg; k++) { v = max(v, (uint32_t)item_sym[base + k]); item_sym[base + k] = (uint8_t)v; }
  }
  __syncwarp();
  int run = 0;
  for (int j0 = 0; j0 < size; j0 += 32) {
    const int pos = ((j0 + lane) * step) & mask;
    const bool ok = pos <= high;
    const uint32_t b = __ballot_sync(0xffffffffu, ok);
    if (ok) cell[pos] = item_sym[run + __popc(b & lanemask_lt())];
    run += __popc(b);
  }
  __syncwarp();
  for (int u0 = 0; u0 < size; u0 += 32) {
    const int u = u0 + lane;
    const uint32_t s = cell[u];
    const uint32_t m = __match_any_sync(0xffffffffu, s);
    const uint32_t x = (uint32_t)sym_next[s] + __popc(m & lanemask_lt());
    __syncwarp();
    if ((m >> lane) == 1u) sym_next[s] = (uint16_t)((uint32_t)sym_next[s] + __popc(m));
    __syncwarp();
    const uint32_t nb = (uint32_t)(log - highbit32(x));
    pack_entry(T, off + (uint32_t)u, kind, s, (x << nb) - (uint32_t)size, nb);
  }
  __syncwarp();
}

// One lane walks one interleaved sequence stream.  A single in-order warp runs this loop, so its cost
// is (instructions on the path) x (issue latency): the body is written branch-free where it can be --
// fields are cut out of the top 32 bits of the window with 32-bit shifts, base values come from a
// 96-entry shared table, validation accumulates into one flag (a bad stream keeps decoding harmlessly;
// KC never executes a chunk whose seq_status is set), and bit accounting happens once at the end.
// ---- sequence loop bit reader: an absolute bit position instead of a shifting window -------------------------------------
// The backward bitstream is addressed as bits of the 4-byte-aligned word array W that contains it; `t` is the index of the
// first bit already consumed (exclusive top), `low` the index of the stream's first bit.  Every sequence loads the three words
// under `t` (L1 hits; the loads issue together with the table lookups) and cuts all six fields out of them with funnel
// shifts.  There is no refill branch and no window state on the dependency chain: in a warp whose lanes each walk their own
// stream, the three divergent refill regions of a shifting window cost a quarter of the loop's instructions.
// select without giving the compiler the chance to turn a chain of ?: into divergent branches
__device__ __forceinline__ uint32_t sel_eq(uint32_t x, uint32_t k, uint32_t a, uint32_t b) {      // x == k ? a : b
  uint32_t r;
  asm("{\n.reg .pred p;\nsetp.eq.u32 p, %1, %2;\nselp.u32 %0, %3, %4, p;\n}" : "=r"(r) : "r"(x), "r"(k), "r"(a), "r"(b));
  return r;
}
struct SeqLane {                                      // one lane's decoder state; everything lives in registers
  const uint16_t *ll16, *ml16, *of16;
  const uint8_t *ll8, *ml8, *of8;
  const uint32_t *W, *bases;
  uint4 *out;
  uint32_t sl, so, sm, rep0, rep1, rep2, out_pos, lit_pos;
  int t;
  // LAST: the final sequence of a block is not followed by state updates
  template <bool LAST> __device__ __forceinline__ void step(uint32_t i) {
    const uint32_t eo = of16[so], em = ml16[sm], el = ll16[sl];
    const uint32_t eo8 = of8[so], em8 = ml8[sm], el8 = ll8[sl];
    const int k = max(t >> 5, 0);
    const uint32_t sh = (uint32_t)t & 31u;
    const uint32_t w0 = W[k], w1 = W[k - 1], w2 = W[k - 2];
    // lanes of a warp stall together: pull the next cache line of this lane's bitstream long before its words are needed
    // (eight sequences never consume 128 bytes, and `i` is warp-uniform: no divergence)
    if ((i & 7u) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(W + max(k - 32, 0)));
    const uint32_t A = __funnelshift_r(w1, w0, sh), B = __funnelshift_r(w2, w1, sh);     // bits [t-32,t) and [t-64,t-32)
    const int ob = (int)eo8, mb = (int)((em >> 13) | ((em8 & 3) << 3)), lb = (int)((el >> 13) | ((el8 & 3) << 3));
    const int xb = ob + mb + lb;
    uint32_t ov, mx, lx, S;                            // S = the 32 bits after the extra bits: the three state updates (<= 26 bits)
    if (xb <= 31) {                                   // the usual case: all extra bits are in the top word
      ov = top_bits(A, 0, ob); mx = top_bits(A, ob, mb); lx = top_bits(A, ob + mb, lb);
      S = __funnelshift_l(B, A, (uint32_t)xb);
    } else {
      ov = top_bits(A, 0, ob);
      const uint32_t a2 = peek32(W, t - ob);
      mx = top_bits(a2, 0, mb); lx = top_bits(a2, mb, lb);
      S = peek32(W, t - xb);
    }
    if (!LAST) {
      const int nl = (int)((el >> 9) & 15), nm = (int)((em >> 9) & 15), no = (int)((eo >> 9) & 15);   // <= 26 bits together
      sl = (el & 511) + top_bits(S, 0, nl);
      sm = (em & 511) + top_bits(S, nl, nm);
      so = (eo & 511) + top_bits(S, nl + nm, no);
      t -= xb + nl + nm + no;
    } else t -= xb;
    // ---- off the chain: values, repeat offsets (branch-free), positions.  Range checks are KC's (it has a lane per sequence) ----
    ov += 1u << ob;
    const uint32_t ll = bases[el8 >> 2] + lx, ml = bases[40 + (em8 >> 2)] + mx;
    // idx: 0 keeps the history, 1 swaps rep0/rep1, >= 2 rotates all three; a new offset rotates like idx 3
    const bool fresh = ov > 3;
    const uint32_t idx = fresh ? 3u : ov - 1 + (ll == 0);
    uint32_t cand = sel_eq(idx, 1, rep1, rep0);
    cand = sel_eq(idx, 2, rep2, cand);
    cand = sel_eq(idx, 3, fresh ? ov - 3 : rep0 - 1, cand);
    rep2 = idx >= 2 ? rep1 : rep2;
    rep1 = idx >= 1 ? rep0 : rep1;
    rep0 = cand;
    __stcs(out + i, make_uint4(out_pos, lit_pos, cand, ml));
    out_pos += ll + ml; lit_pos += ll;
  }
};

// `lead`: see fast_decode_huffman.  `unknown_history`: a block unit other than the first of its frame starts with repeat
// offsets nobody knows yet; sentinels far above any legal offset make every use of them fail KC's range check, which
// sends the whole frame to the serial decoder
// [part_lo, part_hi) (FastDecodeArgs): decode only the sequences of those KC parts, P = EXEC_PARTS in all; a launch that does not
// reach the last part leaves the chain's state in the slot (behind SeqInfo) and, at the index it stopped at, a record with the
// positions reached, which ends KC's last literal run; a launch that does not start at part 0 picks the state up.
#ifndef EXEC_PARTS_V
#define EXEC_PARTS_V 8
#endif
constexpr uint32_t EXEC_PARTS = EXEC_PARTS_V;     // KC's work items per chunk (see the queue in zstd_fast_exec_kernel)
struct SeqSave { int t; uint32_t sl, sm, so, rep0, rep1, rep2, out_pos, lit_pos, err; };
static_assert(sizeof(SeqInfo) + sizeof(SeqSave) <= 64, "the slot reserves 64 bytes for SeqInfo + SeqSave");
__device__ __forceinline__ void fast_decode_sequences(const uint8_t *src, FastDesc *D, uint4 *out, const SeqTab T, const uint32_t *bases,
                                                      const SeqInfo &I, uint32_t lead, bool unknown_history, uint32_t part_lo, uint32_t part_hi, SeqSave *save) {
  const uint32_t nseq = D->nseq;
  const uint32_t groups = (nseq + 31) >> 5;
  // first sequence of KC part p (< nseq for p < P: a segment that is not the last never reaches the last sequence)
  auto part_begin = [&](uint32_t p) { return p >= EXEC_PARTS ? nseq : min(nseq, (groups * p / EXEC_PARTS) << 5); };
  const uint32_t begin = part_begin(part_lo), end = part_begin(part_hi);
  const uint32_t seg = part_lo;                       // != 0: the chain's state is picked up from the slot
  uint32_t err = ST_OK;
  SeqLane L;
  L.out_pos = 0; L.lit_pos = 0;
  const uint8_t *const p = src + I.bits_off;
  const uint32_t nb = I.bits_len;
  // W[-1], W[-2] are read under the first stream bits: a fast-path frame has >= 12 header bytes before the bitstream
  if (nb == 0 || I.bits_off + lead < 12 || p[nb - 1] == 0 || (seg != 0 && save->err != ST_OK)) err = ST_CORRUPT;
  else {
    L.ll16 = T.t16; L.ml16 = T.t16 + 512; L.of16 = T.t16 + 1024;
    L.ll8 = T.t8; L.ml8 = T.t8 + 512; L.of8 = T.t8 + 1024;
    L.bases = bases; L.out = out;
    L.W = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    const int d = (int)((uintptr_t)p & 3), low = 8 * d;
    if (seg != 0) {
      const SeqSave s = *save;
      L.t = s.t; L.sl = s.sl; L.sm = s.sm; L.so = s.so; L.rep0 = s.rep0; L.rep1 = s.rep1; L.rep2 = s.rep2;
      L.out_pos = s.out_pos; L.lit_pos = s.lit_pos;
    } else {
      L.t = 8 * (d + (int)nb - 1) + highbit32(p[nb - 1]);       // the sentinel bit itself is not payload
      { const uint32_t a = peek32(L.W, L.t); L.sl = top_bits(a, 0, (int)I.ll_log); L.so = top_bits(a, (int)I.ll_log, (int)I.of_log);
        L.sm = top_bits(a, (int)(I.ll_log + I.of_log), (int)I.ml_log); L.t -= (int)(I.ll_log + I.of_log + I.ml_log); }
      L.rep0 = 1; L.rep1 = 4; L.rep2 = 8;
      if (unknown_history) { L.rep0 = 0xFFFFFF01u; L.rep1 = 0xFFFFFF02u; L.rep2 = 0xFFFFFF03u; }
    }
    const uint32_t stop = min(end, nseq - 1);
    for (uint32_t i = begin; i < stop; i++) L.step<false>(i);
    if (end == nseq) {
      L.step<true>(nseq - 1);
      // positions stay bounded even on garbage (they cannot wrap within 65536 sequences), and KC checks every record before using it
      if (L.t != low) err = ST_CORRUPT;                      // the stream must end exactly on its first bit
    }
  }
  if (end != nseq) {
    SeqSave s;
    s.t = L.t; s.sl = L.sl; s.sm = L.sm; s.so = L.so; s.rep0 = L.rep0; s.rep1 = L.rep1; s.rep2 = L.rep2;
    s.out_pos = L.out_pos; s.lit_pos = L.lit_pos; s.err = err;
    *save = s;
    out[end] = make_uint4(L.out_pos, L.lit_pos, 0, 0);       // (the next segment writes the whole record: same positions)
    D->seq_status = err;
    return;
  }
  out[nseq] = make_uint4(L.out_pos, L.lit_pos, 0, 0);       // sentinel: literal length of the last sequence, block totals
  D->seq_status = err; D->out_end = L.out_pos; D->lit_end = L.lit_pos;
}

// ---- KP kernel: prep_chunk (headers, Huffman table) followed by the sequence table build, one warp per chunk ----
__global__ void __launch_bounds__(KP_WARPS * 32) zstd_fast_prep_kernel(FastDecodeArgs F) {
  __shared__ __align__(16) uint8_t stage[KP_WARPS][4096];       // Huffman table (4 KB), then the packed FSE tables (3.75 KB)
  __shared__ HufScratch hscratch[KP_WARPS];
  __shared__ SeqScratch scratch[KP_WARPS];
  const DecodeArgs &A = F.base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SeqScratch &S = scratch[warp];
  const SeqTab T{reinterpret_cast<uint16_t *>(stage[warp]), stage[warp] + 2560};
  for (uint32_t chunk = blockIdx.x * KP_WARPS + warp; chunk < A.n; chunk += gridDim.x * KP_WARPS) {
    prep_chunk(F, chunk, stage[warp], hscratch[warp], lane);
    ChunkSlot slot = slot_of(F, chunk);
    FastDesc *D = slot.desc();
    SeqInfo *const info = reinterpret_cast<SeqInfo *>(slot.seq_info());
    if (lane == 0) { info->ready = 0; *reinterpret_cast<uint2 *>(slot.base + FAST_PROG_OFF) = make_uint2(0, 0); }
    if (D->state != 0) continue;                                                              // uniform per warp
    const uint8_t *const src = (const uint8_t *)A.in_ptrs[chunk];
    const uint32_t nseq = D->nseq;
    if (nseq == 0) {
      if (lane == 0) { slot.seqs()[0] = make_uint4(0, 0, 0, 0); D->seq_status = ST_OK; D->out_end = 0; D->lit_end = 0; }
      continue;
    }
    const uint8_t *const bp = src + D->tab_off;
    const uint32_t room = D->blk_end - D->tab_off, modes = D->modes;
    __syncwarp();                                                                             // previous chunk's copy-out is done
    if (lane == 0) {
      uint32_t p = 0;
      bool ok = true;
      for (int t = 0; t < 3 && ok; t++) {                                                     // stream order: LL, OF, ML
        const int mode = (modes >> (6 - 2 * t)) & 3;
        const int max_allowed = (t == 0) ? LL_MAX_SYM : (t == 1) ? OF_MAX_SYM : ML_MAX_SYM;
        S.tab_mode[t] = mode;
        if (mode == 0) {
          const int16_t *def = (t == 0) ? c_ll_def : (t == 1) ? c_of_def : c_ml_def;
          const int dmax = (t == 0) ? 35 : (t == 1) ? 28 : 52;
          for (int i = 0; i <= dmax; i++) S.norm[kb_norm_off(t) + i] = def[i];
          S.tab_max[t] = dmax; S.tab_log[t] = (t == 1) ? OF_DEF_LOG : LL_DEF_LOG;
        } else if (mode == 1) {
          if (p >= room || bp[p] > max_allowed) { ok = false; break; }
          S.tab_max[t] = bp[p]; S.tab_log[t] = 0;
          p += 1;
        } else if (mode == 2) {
          int ms = 0, al = 0;
          const int max_log = (t == 1) ? OF_MAX_LOG : LL_MAX_LOG;
          const int used = (p < room) ? read_ncount(bp + p, room - p, S.norm + kb_norm_off(t), max_allowed, max_log, &ms, &al) : -1;
          if (used < 0) { ok = false; break; }
          S.tab_max[t] = ms; S.tab_log[t] = al;
          p += (uint32_t)used;
        } else ok = false;                                                                    // Repeat needs a previous block
      }
      if (p >= room) ok = false;
      S.bits_off = p;
      S.ok = ok ? 1u : 0u;
    }
    __syncwarp();
    if (!S.ok) {
      if (lane == 0) { slot.seqs()[nseq] = make_uint4(0, 0, 0, 0); D->seq_status = ST_CORRUPT; D->out_end = 0; D->lit_end = 0; }
      continue;
    }
    for (int t = 0; t < 3; t++) {
      const uint32_t toff = (t == 0) ? 0u : (t == 1) ? 1024u : 512u;
      if (S.tab_mode[t] == 1) { if (lane == 0) pack_entry(T, toff, t, (uint32_t)S.tab_max[t], 0, 0); }
      else fse_build_warp_packed(T, toff, S.norm + kb_norm_off(t), S.tab_max[t], S.tab_log[t], t, S.sym_next, lane);
      __syncwarp();
    }
    {
      const uint4 *const from = reinterpret_cast<const uint4 *>(stage[warp]);
      uint4 *const to = reinterpret_cast<uint4 *>(slot.seq_tabs());
      for (uint32_t k = lane; k < KB_TAB_BYTES / 16; k += 32) to[k] = from[k];
    }
    if (lane == 0) {
      info->ll_log = (uint32_t)S.tab_log[0]; info->of_log = (uint32_t)S.tab_log[1]; info->ml_log = (uint32_t)S.tab_log[2];
      info->bits_off = D->tab_off + S.bits_off; info->bits_len = room - S.bits_off;
      info->ready = 1;
    }
  }
}

// ---- KB: 56 chunks' tables pulled into shared memory by bulk async copies, then one lane per sequence stream ----
template <int KB_DEC_WARPS>
__global__ void __launch_bounds__(32 * KB_DEC_WARPS, 1) zstd_fast_seq_kernel(FastDecodeArgs F) {
  constexpr int KB_THREADS = 32 * KB_DEC_WARPS, KB_LANES = KB_GROUP / KB_DEC_WARPS;
  extern __shared__ __align__(128) uint8_t kb_smem[];
  __shared__ uint32_t s_group;
  __shared__ __align__(8) uint64_t s_bar;
  uint16_t *const tab16 = reinterpret_cast<uint16_t *>(kb_smem);                             // [KB_GROUP][1280]: LL 512 | ML 512 | OF 256
  uint8_t *const tab8 = kb_smem + (size_t)KB_GROUP * 1280 * 2;                               // [KB_GROUP][1280]
  uint32_t *const bases = reinterpret_cast<uint32_t *>(kb_smem + (size_t)KB_GROUP * KB_TAB_BYTES);   // [0,36) LL bases, [40,93) ML bases
  const DecodeArgs &A = F.base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 36; i += KB_THREADS) bases[i] = c_ll_base[i];
  for (int i = threadIdx.x; i < 53; i += KB_THREADS) bases[40 + i] = c_ml_base[i];
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  for (;;) {
    if (threadIdx.x == 0) s_group = atomicAdd(F.group_counters + F.kb_queue, 1u);
    __syncthreads();                                                  // also: every lane is done with the previous group's tables
    const uint32_t g0 = F.lo + s_group * KB_GROUP;
    if (g0 >= F.hi) break;
    const uint32_t cnt = min((uint32_t)KB_GROUP, F.hi - g0);
    const uint32_t c = (uint32_t)warp * KB_LANES + (uint32_t)lane;
    const bool mine = lane < KB_LANES && c < cnt;
    ChunkSlot slot = slot_of(F, g0 + (mine ? c : 0));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the tables about to be overwritten were read through the generic proxy
    if (threadIdx.x == 0) mbar_arrive_tx(&s_bar, cnt * (uint32_t)KB_TAB_BYTES);
    if (mine) {                                                       // every lane pulls its own chunk's two planes
      bulk_g2s(tab16 + (size_t)c * 1280, slot.seq_tabs(), 2560, &s_bar);
      bulk_g2s(tab8 + (size_t)c * 1280, slot.seq_tabs() + 2560, 1280, &s_bar);
    }
    SeqInfo info;
    info.ready = 0;
    if (mine) info = *reinterpret_cast<const SeqInfo *>(slot.seq_info());
    mbar_wait(&s_bar, phase);
    phase ^= 1;
    if (mine && info.ready) {
      const SeqTab T{tab16 + (size_t)c * 1280, tab8 + (size_t)c * 1280};
      fast_decode_sequences((const uint8_t *)A.in_ptrs[g0 + c], slot.desc(), slot.seqs(), T, bases, info, F.bare_blocks ? 6u : 0u,
                            F.bare_blocks && (F.unit_base + g0 + c) != 0, F.part_lo, F.part_hi,
                            reinterpret_cast<SeqSave *>(slot.seq_info() + sizeof(SeqInfo)));
    }
  }
}

// =================================================================================================
// KC: sequence execution, one warp per chunk, one lane per sequence
// =================================================================================================
constexpr int EXEC_WARPS = 4;

struct LitSrc {
  const uint8_t *base;
  uint32_t seg, pad, mode;     // mode 0: contiguous, 1: rle (every index reads base[0]), 2: four padded segments
  __device__ __forceinline__ const uint8_t *run(uint32_t p, uint32_t len, bool *contig) const {
    if (mode == 0) { *contig = true; return base + p; }
    if (mode == 1) { *contig = false; return base; }
    const uint32_t s = (p >= seg) + (p >= 2 * seg) + (p >= 3 * seg);
    *contig = s == 3 || p + len <= (s + 1) * seg;
    return base + p + s * pad;
  }
  __device__ __forceinline__ uint8_t at(uint32_t p) const {
    if (mode == 0) return base[p];
    if (mode == 1) return base[0];
    const uint32_t s = (p >= seg) + (p >= 2 * seg) + (p >= 3 * seg);
    return base[p + s * pad];
  }
};

// Whole-warp match copy (RFC 8878 3.1.2.5 semantics incl. overlap): 32 consecutive bytes per step,
// four steps in flight when source and destination are disjoint.
__device__ __forceinline__ void warp_copy_match(uint8_t *out, uint32_t d, uint32_t offset, uint32_t ml, int lane) {
  const uint8_t *sp = out + d - offset;
  uint8_t *dp = out + d;
  if (offset >= ml) {
    uint32_t k = lane;
    for (; k + 96 < ml; k += 128) {
      const uint8_t a = sp[k], b = sp[k + 32], c = sp[k + 64], e = sp[k + 96];
      dp[k] = a; dp[k + 32] = b; dp[k + 64] = c; dp[k + 96] = e;
    }
    // (up to three more steps: their loads leave together too -- a step-by-step tail costs one L2 round trip per 32 bytes)
    const bool h0 = k < ml, h1 = k + 32 < ml, h2 = k + 64 < ml;
    uint8_t a = 0, b = 0, c = 0;
    if (h0) a = sp[k];
    if (h1) b = sp[k + 32];
    if (h2) c = sp[k + 64];
    if (h0) dp[k] = a;
    if (h1) dp[k + 32] = b;
    if (h2) dp[k + 64] = c;
    for (k += 96; k < ml; k += 32) dp[k] = sp[k];
  } else if (offset >= 32) {
    // every 32-byte step reads bytes that earlier steps wrote: keep the steps ordered
    for (uint32_t k0 = 0; k0 < ml; k0 += 32) {
      const uint32_t k = k0 + lane;
      if (k < ml) dp[k] = sp[k];
      __syncwarp();
    }
  } else {
    for (uint32_t k = lane; k < ml; k += 32) dp[k] = sp[k % offset];          // periodic extension of the last `offset` bytes
  }
}

// (9 CTAs per SM by registers = 56 per thread: eight resident KC CTAs then leave room for KB's CTA beside them)
#ifndef EXEC_OWN_V
#define EXEC_OWN_V 256
#endif
constexpr uint32_t EXEC_OWN = EXEC_OWN_V;      // pieces of a group whose owning sequence is looked up in shared memory (a group rarely has more)
__global__ void __launch_bounds__(EXEC_WARPS * 32, 9) zstd_fast_exec_kernel(FastDecodeArgs F) {
  __shared__ uint8_t s_own[EXEC_WARPS][EXEC_OWN ? EXEC_OWN : 1];
  const DecodeArgs &A = F.base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *const own = s_own[warp];
  for (uint32_t k = lane; k < EXEC_OWN; k += 32) own[k] = 0;      // stale entries are only ever read for empty pieces, but must name a lane
  __syncwarp();
  // chunks are handed out through a work counter (one per sub-wave, zeroed with the workspace header): in a mixed batch a warp
  // that draws raw / RLE chunks (finished by KP) or short ones simply takes more of them
  // A chunk is executed in EXEC_PARTS consecutive parts (quarters of its groups), each a work item of its own: items are
  // numbered part-major, so part p of a chunk is always drawn after part p - 1, by whichever warp is free.  With one item
  // per chunk a sub-wave of 8,192 chunks gave the 4,736 resident warps 1.73 items each -- two rounds, the second with a
  // quarter of the warps idle; quarters make that 6.9 -> 7 rounds of a quarter's length.  A part waits for its predecessor
  // on the chunk's hand-over word (release / acquire at gpu scope: its matches read what the predecessor wrote, possibly
  // on another SM); the predecessor is always held by a running warp, and part 0 waits for nobody.
  // (When KB decodes a sub-wave in two launches, this launch executes parts [part_lo, part_hi) only.)
  uint32_t *const queue = F.base.counter + F.kc_queue;
  const uint32_t span = F.hi - F.lo;
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(queue, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= span * (F.part_hi - F.part_lo)) break;
    const uint32_t part = F.part_lo + item / span, chunk = F.kc_order ? F.kc_order[F.lo + item % span] : F.lo + item % span;
    ChunkSlot slot = slot_of(F, chunk);
    const FastDesc *D = slot.desc();
    if (D->state != 0) continue;
    uint32_t *const prog = reinterpret_cast<uint32_t *>(slot.base + FAST_PROG_OFF);
    const uint8_t *const src = (const uint8_t *)A.in_ptrs[chunk];
    uint8_t *const out = (uint8_t *)A.out_ptrs[chunk];
    uint32_t status;
    if (part == 0) {
      status = D->seq_status;
      for (uint32_t k = 0; k < D->n_streams; k++) if (status == ST_OK) status = D->lit_status[k];
    } else {
      uint32_t seen = 0;
      if (lane == 0) {
        for (;;) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(prog) : "memory");
          if (seen >= part) break;
          __nanosleep(100);
        }
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(prog + 1) : "memory");
      }
      __syncwarp();
      status = __shfl_sync(0xffffffffu, seen, 0);
      // the first part of a later launch: KB has finished the chunk in between, its verdict on the whole stream counts now
      if (part == F.part_lo && status == ST_OK) status = D->seq_status;
    }
    uint32_t total = 0;
    if (status == ST_OK) {
      const uint32_t nseq = D->nseq, lit_size = D->lit_size, cap = D->cap;
      LitSrc L;
      if (D->lit_type == 2) { L.base = slot.lits(); L.seg = D->seg; L.pad = seg_padded(D->seg) - D->seg; L.mode = D->n_streams == 4 ? 2u : 0u; }
      else { L.base = src + D->lit_src; L.seg = 0; L.pad = 0; L.mode = D->lit_type; }
      const uint4 *__restrict__ seqs = slot.seqs();
      const uintptr_t out_addr = (uintptr_t)out;
      // records of the first group (the sentinel at index nseq is fetched like a record: it ends the last literal run)
      const uint32_t groups = (nseq + 31) >> 5;
      const uint32_t g_lo = (groups * part / EXEC_PARTS) << 5, g_hi = min(nseq, (groups * (part + 1) / EXEC_PARTS) << 5);
      uint4 r_nxt = make_uint4(0, 0, 0, 0);
      if (g_lo + (uint32_t)lane <= nseq) r_nxt = __ldcs(seqs + g_lo + lane);
      for (uint32_t g0 = g_lo; g0 < g_hi; g0 += 32) {
        const uint32_t i = g0 + (uint32_t)lane;
        const bool valid = i < nseq;
        const uint4 r = r_nxt;
        // next group's records are requested now and used one iteration later (hides one memory round trip per group)
        r_nxt = make_uint4(0, 0, 0, 0);
        if (i + 32 <= nseq) r_nxt = __ldcs(seqs + i + 32);
        // literal length = next record's literal position - mine
        uint32_t next_lit = __shfl_down_sync(0xffffffffu, r.y, 1);
        const uint32_t nl0 = __shfl_sync(0xffffffffu, r_nxt.y, 0);
        if (lane == 31) next_lit = nl0;
        const uint32_t ll = valid ? next_lit - r.y : 0;
        const uint32_t group_start = __shfl_sync(0xffffffffu, r.x, 0);
        const uint32_t d = r.x + ll;
        // range checks (KB only decodes): a record that would read before the output, past the literals or write past the
        // capacity ends the chunk before anything of its group is executed
        const bool oob = valid && (r.z == 0 || r.z > d || r.y + ll > lit_size);
        const bool over = valid && (unsigned long long)r.x + ll + r.w > cap;    // 64-bit: garbage lengths may wrap
        if (__any_sync(0xffffffffu, oob || over)) { status = __any_sync(0xffffffffu, oob) ? ST_CORRUPT : ST_BUFFER_TOO_SMALL; break; }
        const bool indep = valid && (d - r.z + r.w <= group_start);            // whole source precedes this group's output
        // ---- piece-parallel phase: the literal run of every sequence and every independent match are cut into
        // destination-aligned 16-byte pieces; pieces are dealt to lanes round-robin, so the work per lane is
        // uniform whatever the length distribution.  Two pieces per lane are in flight per round: all loads of a
        // round are issued before the first store, so a round costs one memory round trip ----
        const uint32_t mlen = indep ? r.w : 0u;
        const uint32_t ca = ll ? (uint32_t)(((out_addr + r.x + ll - 1) >> 4) - ((out_addr + r.x) >> 4)) + 1 : 0u;
        const uint32_t cb = mlen ? (uint32_t)(((out_addr + d + mlen - 1) >> 4) - ((out_addr + d) >> 4)) + 1 : 0u;
        uint32_t incl = ca + cb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const uint32_t excl = incl - (ca + cb), total = __shfl_sync(0xffffffffu, incl, 31);
        // piece -> owning sequence: every lane writes its own index over its range of pieces (a few stores), so that a piece
        // finds its sequence with one shared load instead of a five-round shuffle search; pieces past EXEC_OWN still search
        for (uint32_t k = excl, e = min(incl, EXEC_OWN); k < e; k++) own[k] = (uint8_t)lane;
        __syncwarp();
        struct Piece { uint8_t *dp; uint32_t nb, sh, w0, w1, w2, w3, w4; };
        // all 32 lanes call (shuffles inside); q >= total yields an empty piece.  q - lane is warp-uniform.
        auto fetch = [&](uint32_t q) -> Piece {
          Piece P{nullptr, 0, 0, 0, 0, 0, 0, 0};
          uint32_t j = 0;
          if (q - (uint32_t)lane + 32 <= EXEC_OWN) j = own[q];
          else {
#pragma unroll
            for (int st = 16; st; st >>= 1) { const uint32_t pj = __shfl_sync(0xffffffffu, excl, (j + st) & 31); if (j + st < 32 && pj <= q) j += st; }
          }
          const uint32_t jx = __shfl_sync(0xffffffffu, r.x, j), jy = __shfl_sync(0xffffffffu, r.y, j), jz = __shfl_sync(0xffffffffu, r.z, j);
          const uint32_t jll = __shfl_sync(0xffffffffu, ll, j), jml = __shfl_sync(0xffffffffu, mlen, j);
          const uint32_t jca = __shfl_sync(0xffffffffu, ca, j), jex = __shfl_sync(0xffffffffu, excl, j);
          if (q < total) {
            uint32_t k = q - jex;
            const bool is_match = k >= jca;
            if (is_match) k -= jca;
            const uint32_t a = is_match ? jx + jll : jx, len = is_match ? jml : jll;       // run start (output offset), run length
            const uintptr_t a0 = out_addr + a, blk = (a0 >> 4) + k;
            const uintptr_t lo = a0 > (blk << 4) ? a0 : (blk << 4);
            const uintptr_t hi = (a0 + len) < ((blk + 1) << 4) ? (a0 + len) : ((blk + 1) << 4);
            const uint32_t nb = (uint32_t)(hi - lo), rel = (uint32_t)(lo - a0);
            const uint8_t *sp = nullptr;
            if (is_match) sp = out + a - jz + rel;
            else {
              bool contig;
              sp = L.run(jy + rel, nb, &contig);
              if (!contig) sp = nullptr;
            }
            P.dp = reinterpret_cast<uint8_t *>(lo);
            P.nb = nb;
            if (sp) {
              const uintptr_t s = (uintptr_t)sp;
              const uint32_t *wbase = reinterpret_cast<const uint32_t *>(s & ~(uintptr_t)3);
              const uint32_t lastw = (uint32_t)(((s + nb - 1) >> 2) - (s >> 2));
              P.sh = (uint32_t)(s & 3) * 8;
              P.w0 = wbase[0]; P.w1 = wbase[min(1u, lastw)]; P.w2 = wbase[min(2u, lastw)]; P.w3 = wbase[min(3u, lastw)]; P.w4 = wbase[min(4u, lastw)];
            } else {
              // RLE literals or a piece straddling two Huffman segments: byte gather (already shifted into place)
              uint32_t t[4] = {0, 0, 0, 0};
              for (uint32_t u = 0; u < nb; u++) t[u >> 2] |= (uint32_t)L.at(jy + rel + u) << (8 * (u & 3));
              P.sh = 0; P.w0 = t[0]; P.w1 = t[1]; P.w2 = t[2]; P.w3 = t[3]; P.w4 = 0;
            }
          }
          return P;
        };
        auto store = [&](const Piece &P) {
          if (P.nb == 0) return;
          const uint32_t v0 = __funnelshift_r(P.w0, P.w1, P.sh), v1 = __funnelshift_r(P.w1, P.w2, P.sh), v2 = __funnelshift_r(P.w2, P.w3, P.sh),
                         v3 = __funnelshift_r(P.w3, P.w4, P.sh);
          if (P.nb == 16) *reinterpret_cast<uint4 *>(P.dp) = make_uint4(v0, v1, v2, v3);
          else {
#pragma unroll
            for (int u = 0; u < 15; u++) {
              const uint32_t word = u < 4 ? v0 : u < 8 ? v1 : u < 12 ? v2 : v3;
              if ((uint32_t)u < P.nb) P.dp[u] = (uint8_t)(word >> (8 * (u & 3)));
            }
          }
        };
        // software pipeline in units of 32 pieces: the loads of round k+1 are issued before the stores of round k, so two
        // pieces per lane are in flight as before, but a group pays for ceil(total / 32) rounds instead of 2 * ceil(total / 64)
        if (total) {
          Piece cur = fetch((uint32_t)lane);
          for (uint32_t q0 = 32; q0 < total; q0 += 32) {
            const Piece nxt = fetch(q0 + (uint32_t)lane);
            store(cur);
            cur = nxt;
          }
          store(cur);
        }
        __syncwarp();
        // ---- the next group's sources: its records have arrived by now (requested at the top of this iteration), so every lane
        // asks for the line its own next sequence will read -- literal run and match source -- one group before the piece loads
        // need them (the resident chunks' output and literals exceed the L2: without this those loads wait on HBM)
        if (g0 + 32 < nseq) {
          const uint32_t ny = __shfl_down_sync(0xffffffffu, r_nxt.y, 1);
          const uint32_t nll = lane < 31 ? ny - r_nxt.y : 0u;
          if (i + 32 < nseq) {
            if (r_nxt.y < lit_size) { bool cg; asm volatile("prefetch.global.L1 [%0];" ::"l"(L.run(r_nxt.y, 1, &cg))); }
            const uint32_t nd = r_nxt.x + nll;
            if (r_nxt.z != 0 && r_nxt.z <= nd && nd < cap) asm volatile("prefetch.global.L1 [%0];" ::"l"(out + nd - r_nxt.z));
          }
        }
        // ---- matches that read this group's own output: in sequence order, the whole warp on each ----
        uint32_t dep = __ballot_sync(0xffffffffu, valid && !indep);
        while (dep) {
          const int j = __ffs(dep) - 1;
          dep &= dep - 1;
          const uint32_t dj = __shfl_sync(0xffffffffu, d, j), oj = __shfl_sync(0xffffffffu, r.z, j), mj = __shfl_sync(0xffffffffu, r.w, j);
          warp_copy_match(out, dj, oj, mj, lane);
          __syncwarp();
        }
      }
      if (part + 1 < EXEC_PARTS) {                          // hand the chunk over: everything this warp wrote, then the word
        __syncwarp();
        if (lane == 0) {
          prog[1] = status;
          __threadfence();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(prog), "r"(part + 1) : "memory");
        }
        continue;
      }
      // trailing literals
      const uint32_t out_end = D->out_end, lit_end = D->lit_end, rest = lit_size - lit_end;
      if (status != ST_OK) {}
      else if (lit_end > lit_size) status = ST_CORRUPT;
      else if (rest > cap - out_end) status = ST_BUFFER_TOO_SMALL;
      else {
        // (a literal-only block has all of its bytes here: vector copies, per Huffman segment where the literals are split)
        if (rest < 64) { for (uint32_t k = lane; k < rest; k += 32) out[out_end + k] = L.at(lit_end + k); }
        else if (L.mode == 1) warp_fill_block(out + out_end, L.base[0], rest, lane);
        else if (L.mode == 0)
          warp_copy_block<1>(out + out_end, L.base + lit_end, rest, D->lit_type == 2 ? rest + 3 : (uint32_t)A.in_sizes[chunk] - D->lit_src - lit_end, lane);
        else {
          uint32_t p = lit_end, o = out_end;
          while (p < lit_size) {
            const uint32_t sg = (p >= L.seg) + (p >= 2 * L.seg) + (p >= 3 * L.seg);
            const uint32_t len = (sg == 3 ? lit_size : (sg + 1) * L.seg) - p;
            warp_copy_block<1>(out + o, L.base + p + sg * L.pad, len, len + 3, lane);      // segments are padded to 16 bytes in the pool
            p += len; o += len;
          }
        }
        total = out_end + rest;
        if (D->content != 0xFFFFFFFFu && D->content != total) status = ST_CORRUPT;
      }
      __syncwarp();
      if (status == ST_OK && D->ck_off && A.verify_checksum) {
        const uint64_t hs = xxh64_warp(out, total, lane);
        if ((uint32_t)hs != ld_le32(src + D->ck_off)) status = ST_CHECKSUM;
      }
    }
    if (part + 1 < EXEC_PARTS) {                            // (only reached with a failed status: pass it on)
      __syncwarp();
      if (lane == 0) {
        prog[1] = status;
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(prog), "r"(part + 1) : "memory");
      }
      continue;
    }
    if (lane == 0) {
      A.out_sizes[chunk] = status == ST_OK ? total : 0;
      if (A.statuses) A.statuses[chunk] = status;
    }
  }
}

// =================================================================================================
// Multi-block frames: cut ONE frame into block units so that its blocks decode side by side
// =================================================================================================
// One thread walks the block headers (a frame of B blocks costs B dependent 3-byte reads).  The cut is speculative: unit k
// is given the output range [k * B, +min(B, rest)), B = the frame's Block_Maximum_Size, i.e. every block but the last is
// assumed to regenerate a full block -- what one-shot compressors (libzstd's, and this library's block-parallel one) emit.  The units are then
// decoded as bare blocks; a unit that regenerates anything else, uses repeat offsets it did not establish itself,
// reaches behind its own output, or needs a previous block's tables fails, and the caller decodes the frame serially.
__global__ void zstd_split_frame_kernel(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, uint32_t max_units, const void **in_ptrs,
                                        size_t *in_sizes, void **out_ptrs, size_t *out_sizes, SplitInfo *info) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  SplitInfo R{};
  do {
    if (n < 9 || n > 0x7FFFFFFFull || ld_le32(src) != ZSTD_FRAME_MAGIC) break;
    uint32_t h = 4;
    const uint32_t fhd = src[h++];
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
    if ((fhd & 0x08) || did_flag) break;                                   // reserved bit / dictionaries: serial path reports them
    const uint32_t fcs_size = fcs_flag == 0 ? single : (1u << fcs_flag);
    if (fcs_size == 0) break;                                             // no content size: nothing to speculate on
    if (h + (single ? 0 : 1) + fcs_size + 3 > n) break;
    // Block_Maximum_Size = min(Window_Size, 128 KB) (RFC 8878 3.1.1.2.4): a frame whose window is below 128 KB -- this
    // library's block-parallel compressor writes 64 KB blocks for mid-sized buffers -- is cut at that size
    uint64_t blk = BLOCK_MAX;
    if (!single) {
      const uint32_t wd = src[h++];
      const uint64_t wbase = 1ull << (10 + (wd >> 3)), wsize = wbase + (wbase >> 3) * (wd & 7);
      if (wsize < blk) blk = wsize;
    }
    uint64_t fcs = 0;
    for (uint32_t k = 0; k < fcs_size; k++) fcs |= (uint64_t)src[h + k] << (8 * k);
    if (fcs_size == 2) fcs += 256;
    h += fcs_size;
    if (fcs <= blk || fcs > cap) break;                                   // one block: the ordinary path is as good
    const uint64_t want = (fcs + blk - 1) / blk;
    if (want > max_units) break;
    uint32_t units = 0;
    bool good = true, last = false;
    while (!last) {
      if ((uint64_t)h + 3 > n || units >= want) { good = false; break; }
      const uint32_t bh = ld_le24(src + h);
      last = bh & 1;
      const uint32_t btype = (bh >> 1) & 3, bsize = bh >> 3;
      const uint32_t body = btype == 1 ? 1u : bsize;
      if (btype == 3 || bsize > blk || (uint64_t)h + 3 + body > n) { good = false; break; }
      in_ptrs[units] = src + h;
      in_sizes[units] = 3 + (size_t)body;
      out_ptrs[units] = dst + (size_t)units * blk;
      const uint64_t rest = fcs - (uint64_t)units * blk;
      out_sizes[units] = rest < blk ? rest : blk;
      units++;
      h += 3 + body;
    }
    const uint32_t has_ck = (fhd >> 2) & 1;
    if (!good || units != want || (uint64_t)h + (has_ck ? 4 : 0) != n) break;      // trailing frames etc.: serial path
    R.ok = 1; R.units = units; R.content_size = fcs; R.has_checksum = has_ck; R.checksum_off = h;
  } while (false);
  *info = R;
}
cudaError_t launch_split_frame(const void *d_src, size_t n, void *d_dst, size_t cap, uint32_t max_units, const void **d_in_ptrs,
                               size_t *d_in_sizes, void **d_out_ptrs, size_t *d_out_sizes, SplitInfo *d_info, cudaStream_t stream) {
  zstd_split_frame_kernel<<<1, 32, 0, stream>>>((const uint8_t *)d_src, n, (uint8_t *)d_dst, cap, max_units, d_in_ptrs, d_in_sizes, d_out_ptrs,
                                                d_out_sizes, d_info);
  return cudaGetLastError();
}
__global__ void __launch_bounds__(32) zstd_verify_checksum_kernel(const uint8_t *data, size_t n, const uint8_t *expect, uint32_t *flag) {
  const uint64_t hsh = xxh64_warp(data, (uint32_t)n, threadIdx.x);
  if (threadIdx.x == 0) *flag = ((uint32_t)hsh != ld_le32(expect)) ? 1u : 0u;
}
cudaError_t launch_verify_checksum(const void *d_data, size_t n, const void *d_expect, uint32_t *d_flag, cudaStream_t stream) {
  zstd_verify_checksum_kernel<<<1, 32, 0, stream>>>((const uint8_t *)d_data, n, (const uint8_t *)d_expect, d_flag);
  return cudaGetLastError();
}

// KC grid: a bounded number of resident CTAs per SM, each striding over chunks.  Fewer chunks in flight keep the
// recently written output (the match sources) inside the 126 MB L2 instead of re-reading it from HBM.
static int g_exec_ctas_per_sm = 8;    // = what 64 registers x 128 threads leave resident: every CTA of the grid runs at once
static int g_exec_ctas_last = 0;      // KC of the last sub-wave runs alone (0 = same as the others)
extern "C" void cuda_zstd_b200_tune_exec_ctas(int v) { if (v > 0) g_exec_ctas_per_sm = v; }
extern "C" void cuda_zstd_b200_tune_exec_ctas_last(int v) { g_exec_ctas_last = v > 0 ? v : 0; }
static uint32_t exec_grid(uint32_t chunks, uint32_t sms, bool last = false) {
  const int per_sm = last && g_exec_ctas_last ? g_exec_ctas_last : g_exec_ctas_per_sm;
  const uint32_t blocks = (chunks + EXEC_WARPS - 1) / EXEC_WARPS, cap = sms * (uint32_t)per_sm;
  return blocks < cap ? blocks : cap;
}

static void launch_kb(int warps, uint32_t grid, cudaStream_t stream, const FastDecodeArgs &F) {
  switch (warps) {
    case 2: zstd_fast_seq_kernel<2><<<grid, 64, KB_SMEM, stream>>>(F); break;
    case 4: zstd_fast_seq_kernel<4><<<grid, 128, KB_SMEM, stream>>>(F); break;
    case 14: zstd_fast_seq_kernel<14><<<grid, 448, KB_SMEM, stream>>>(F); break;
    default: zstd_fast_seq_kernel<8><<<grid, 256, KB_SMEM, stream>>>(F); break;
  }
}

cudaError_t launch_decode_fast(const FastDecodeArgs &F0, cudaStream_t stream, const FastOverlap *ov, int *launches) {
  static const int kb_warps = getenv("CUDA_ZSTD_KB_WARPS") ? atoi(getenv("CUDA_ZSTD_KB_WARPS")) : KB_DEC_WARPS_DEFAULT;
  const uint32_t n = F0.base.n;
  if (launches) *launches = 0;
  if (n == 0) return cudaSuccess;
  // the opt-in to large dynamic shared memory belongs to the (function, device) pair: once per device, whichever thread
  // gets there first (a process may drive several GPUs)
  static std::atomic<unsigned long long> attr_done_mask{0};
  cudaError_t e;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if (dev >= 64 || !((attr_done_mask.load(std::memory_order_acquire) >> dev) & 1ull)) {
    if ((e = cudaFuncSetAttribute(zstd_fast_lit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KA_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(zstd_fast_seq_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KB_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(zstd_fast_seq_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KB_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(zstd_fast_seq_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KB_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(zstd_fast_seq_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KB_SMEM)) != cudaSuccess) return e;
    if (dev < 64) attr_done_mask.fetch_or(1ull << dev, std::memory_order_release);
  }
  FastDecodeArgs F = F0;
  F.lo = 0; F.hi = n; F.sub = 0;
  F.part_lo = 0; F.part_hi = EXEC_PARTS; F.kb_queue = 1; F.kc_queue = 2;
  // CUDA_ZSTD_TRACE=1: events after every launch, printed as "kernel@ms since the call began" once the call has drained
  // (a debugging aid: it synchronises the stream; this is how the KB / KC overlap in DESIGN.md 4.1 was timed)
  static const bool trace = getenv("CUDA_ZSTD_TRACE") != nullptr;
  cudaEvent_t te[40]; const char *tn[40]; int tc = 0;
  auto mark = [&](const char *name, cudaStream_t st) { if (trace && tc < 40 && cudaEventCreate(&te[tc]) == cudaSuccess) { cudaEventRecord(te[tc], st); tn[tc++] = name; } };
  mark("start", stream);
  // one memset zeroes every counter of the pipeline: general work queue, slow count, pool heads, group counters
  if ((e = cudaMemsetAsync(F.base.counter, 0, WS_HEADER_BYTES, stream)) != cudaSuccess) return e;
  const uint32_t sms = (uint32_t)(F.sm_count > 0 ? F.sm_count : 148);
  F.sub_chunks = ov ? sms * KB_GROUP : n;                                  // (without the side stream the wave is one KB / KC pass)
  if ((n + F.sub_chunks - 1) / F.sub_chunks > FAST_ORDER_SUBS) F.kc_order = nullptr;
  const uint32_t kp_blocks = (n + KP_WARPS - 1) / KP_WARPS;
  zstd_fast_prep_kernel<<<kp_blocks < 5 * sms ? kp_blocks : 5 * sms, KP_WARPS * 32, 0, stream>>>(F);
  mark("KP", stream);
  zstd_fast_order_kernel<<<(n + 255) / 256, 256, 0, stream>>>(F);
  int count = 3;
  // KB (SMEM-bound: one CTA and four busy warps per SM) and KC (no SMEM, wants many warps) run together: the batch
  // is cut into sub-waves of one full KB pass; KC of sub-wave k runs on the side stream while KB decodes k+1.
  // KA needs nothing from KB and KB nothing from KA (only KC needs both): KA goes to the side stream in front of the first
  // KC, so that KB's CTAs move in as KA's work queue drains -- on mixed batches KA ends in a long thin tail.
  const uint32_t sub_chunks = sms * KB_GROUP;
  const uint32_t nsub = ov ? (n + sub_chunks - 1) / sub_chunks : 1;
  // (a batch of one sub-wave overlaps too, through the segments below, once it is large enough to pay for the extra launches)
  static const uint32_t min_split = getenv("CUDA_ZSTD_MIN_SPLIT") ? (uint32_t)atoi(getenv("CUDA_ZSTD_MIN_SPLIT")) : KB_SPLIT_MIN_CHUNKS;
  const bool overlap = ov && (nsub > 1 || n >= min_split) && nsub <= (uint32_t)FastOverlap::MAX_SUB;
  const uint32_t ka_groups = (n + KA_GROUP - 1) / KA_GROUP;
  cudaStream_t ka_stream = stream;
  if (overlap) {
    if ((e = cudaEventRecord(ov->prep, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(ov->side, ov->prep, 0)) != cudaSuccess) return e;
    ka_stream = ov->side;
  }
  zstd_fast_lit_kernel<<<ka_groups < 2 * sms ? ka_groups : 2 * sms, KA_THREADS, KA_SMEM, ka_stream>>>(F);
  mark("KA", ka_stream);
  if (!overlap) {
    const uint32_t kb_groups = (n + KB_GROUP - 1) / KB_GROUP;
    launch_kb(kb_warps, kb_groups < sms ? kb_groups : sms, stream, F);
    zstd_fast_exec_kernel<<<exec_grid(n, sms), EXEC_WARPS * 32, 0, stream>>>(F);
    count += 2;
  } else {
    // Each sub-wave's sequences are decoded by S KB launches (the chain's state waits in the slot in between): KC starts on
    // the first parts of every chunk while KB decodes the sequences of the later ones (CUDA_ZSTD_KB_SPLIT=S; 1: one launch).
    static const uint32_t split_env = getenv("CUDA_ZSTD_KB_SPLIT") ? (uint32_t)atoi(getenv("CUDA_ZSTD_KB_SPLIT")) : 0u;
    uint32_t S = split_env ? split_env : nsub == 1 ? KB_SPLIT_ONE_SUBWAVE : KB_SPLIT_DEFAULT;
    if (S > EXEC_PARTS) S = EXEC_PARTS;
    while (S > 1 && S * nsub > (uint32_t)FastOverlap::MAX_SUB) S--;
    // part boundaries of the segments: equal shares, or CUDA_ZSTD_KB_BOUNDS="0,2,5,8" (S + 1 ascending numbers, 0 .. EXEC_PARTS)
    uint32_t bounds[EXEC_PARTS + 1];
    for (uint32_t h = 0; h <= S; h++) bounds[h] = h * EXEC_PARTS / S;
    // several sub-waves: a short first segment lets KC start early, measured best of {0,4,8}, {0,3,8}, {0,2,5,8}, {0,3,6,8},
    // {0,2,4,6,8}, {0,1,3,5,8} on the mixed batch (4.41, 4.35, 4.26, 4.44, 4.37, 4.30 ms) and level on the uniform one (3.89 - 3.96)
    if (!split_env && nsub > 1 && EXEC_PARTS == 8 && 3 * nsub <= (uint32_t)FastOverlap::MAX_SUB) { S = 3; bounds[0] = 0; bounds[1] = 2; bounds[2] = 5; bounds[3] = 8; }
    static const char *const bounds_env = getenv("CUDA_ZSTD_KB_BOUNDS");
    if (bounds_env) {
      uint32_t b[EXEC_PARTS + 1], nb = 0;
      for (const char *p = bounds_env; *p && nb <= EXEC_PARTS;) { b[nb++] = (uint32_t)strtoul(p, nullptr, 10); while (*p && *p != ',') p++; if (*p) p++; }
      bool ok = nb >= 2 && b[0] == 0 && b[nb - 1] == EXEC_PARTS && (nb - 1) * nsub <= (uint32_t)FastOverlap::MAX_SUB;
      for (uint32_t h = 1; h < nb && ok; h++) ok = b[h] > b[h - 1];
      if (ok) { S = nb - 1; for (uint32_t h = 0; h <= S; h++) bounds[h] = b[h]; }
    }
    for (uint32_t k = 0; k < nsub; k++) {
      F.lo = k * sub_chunks; F.hi = F.lo + sub_chunks < n ? F.lo + sub_chunks : n; F.sub = k;
      const uint32_t m = F.hi - F.lo, kb_groups = (m + KB_GROUP - 1) / KB_GROUP;
      for (uint32_t h = 0; h < S; h++) {
        const uint32_t q = S * k + h;
        F.part_lo = bounds[h]; F.part_hi = bounds[h + 1];
        F.kb_queue = 1 + q; F.kc_queue = S > 1 ? 20 + q : 2 + q;
        launch_kb(kb_warps, kb_groups < sms ? kb_groups : sms, stream, F);
        mark(h ? "KBb" : "KB", stream);
        if ((e = cudaEventRecord(ov->ev[q], stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(ov->side, ov->ev[q], 0)) != cudaSuccess) return e;
        zstd_fast_exec_kernel<<<exec_grid(m, sms, k + 1 == nsub && h + 1 == S), EXEC_WARPS * 32, 0, ov->side>>>(F);
        mark(h ? "KCb" : "KC", ov->side);
        count += 2;
      }
    }
    if ((e = cudaEventRecord(ov->done, ov->side)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(stream, ov->done, 0)) != cudaSuccess) return e;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  // whatever the fast path declined: the general kernel pulls it from the device-side list
  DecodeArgs G = F.base;
  G.list = F.slow_list;
  G.list_count = F.slow_count;
  e = launch_decode_batch_nomemset(G, F.general_grid, stream);
  if (trace) {
    mark("end", stream);
    cudaStreamSynchronize(stream);
    for (int i = 1; i < tc; i++) { float ms = 0; cudaEventElapsedTime(&ms, te[0], te[i]); fprintf(stderr, "%s@%.3f ", tn[i], ms); }
    fprintf(stderr, "\n");
    for (int i = 0; i < tc; i++) cudaEventDestroy(te[i]);
  }
  if (launches) *launches = count + 1;
  return e;
}

} // namespace b200zstd
