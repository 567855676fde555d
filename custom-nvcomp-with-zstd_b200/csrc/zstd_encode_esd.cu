// zstd_encode_esd.cu -- levels 1-4 of the batched Zstandard compressor for sm_100a.
//
// Replaces, for the batch path at these levels, ZstdBatchManager::compress_batch -> DefaultZstdManager::compress
// (src/cuda_zstd_manager.cu:5715-5797, 1536-3112 in the reference): find_matches_kernel + greedy_parse_kernel +
// build_sequences_gpu_kernel (src/lz77_parallel.cu:26-268), compress_literals / compress_sequences
// (manager.cu:4406-4484, 4864-4974), write_frame_header / write_block (:3998-4106, 4227-4286).
//
// Three kernels per wave of blocks, each of them data-parallel over the whole wave (zstd_encode_lz.cuh holds the
// arithmetic of the first two, shared with the host-side model tests/model/enc_model.cpp; the bytes must agree):
//
// MATCH  (zstd_lz_match_kernel)   one CTA of 1024 threads per <= 128 KB block, one CTA per SM.  The block arrives HBM ->
//        shared memory by cp.async.bulk (TMA) copies completing on an mbarrier; the hash tables (uint32 buckets, 128 KB
//        for a 64 KB block) live in shared memory.  The block is walked in windows of 256 positions, one per thread of
//        a GROUP of 8 warps; the four groups take the windows in turn.  Per window: hash, look the candidates up in the
//        state the windows before left behind, note "first position of this window with my hash" in a small side table
//        (atomicMin on a window-tagged key), group barrier, insert (atomicMax: highest position wins a bucket), tell
//        the next group to start its window (named barriers: arrive / sync), then -- off that chain -- verify the
//        candidate against the staged block and measure the match up to 16 bytes.  R[p] = offset | length << 17 goes to
//        the wave's scratch.  Nothing in this kernel depends on the parse: every position is independent work, and
//        the only serial part is the lookup -> insert hand-over from window to window.
// SELECT (zstd_lz_select_kernel)  one warp per block, one lane per 1/32 of it: each lane walks its sub-segment greedily
//        over R (repeat offsets, one-step lazy, extension of long matches), first speculatively from its first
//        position, then again from the state the lane before it really ended in until the two walks meet; the rounds
//        repeat until no lane's exit state changes.  The joined lists are exactly the serial walk of the whole block,
//        repeat-offset codes included.
// FINISH (zstd_lz_finish_kernel)  one warp per block: joins the lists, gathers the literals, runs the entropy stage
//        (zstd_encode_entropy.cuh) and writes the frame (or the bare block in block mode).
#include "zstd_common.cuh"
#include "zstd_device_api.h"
#include "zstd_encode_core.cuh"
#include "zstd_encode_entropy.cuh"
#include "zstd_encode_lz.cuh"

#include <cstdlib>
#include <type_traits>

namespace b200zstd {

using namespace enc;
using namespace lz;

namespace {

constexpr int MATCH_GROUPS = 4;                  // groups of LZ_WIN threads that take the windows in turn
constexpr int MATCH_THREADS = MATCH_GROUPS * (int)LZ_WIN;
constexpr uint32_t LZ_IN_PAD = 64;               // 16 bytes of alignment slack in front, read-ahead room behind
constexpr uint32_t KIND_PARSED = 0, KIND_RLE = 1, KIND_SKIP = 2;

// ---- scratch slot of one block: [BlockHdr 16 B | 32 x LaneHdr | bitmap "R[p] != 0": block_max bits | R: uint32 x block_max |
//      spec lists | prefix lists] ----
struct BlockHdr { uint32_t kind, pad0, pad1, pad2; };
struct LaneHdr { uint16_t spec_cnt, sync_k, pre_cnt, pad; };
__host__ __device__ constexpr size_t slot_map_off() { return 16 + LZ_LANES * sizeof(LaneHdr); }
__host__ __device__ constexpr size_t slot_r_off(uint32_t bm) { return slot_map_off() + bm / 8; }
__host__ __device__ constexpr size_t slot_spec_off(uint32_t bm) { return slot_r_off(bm) + (size_t)bm * 4; }
__host__ __device__ constexpr size_t slot_lists_bytes(uint32_t bm) { return (size_t)LZ_LANES * ((bm / LZ_LANES + 4) / 4 + 8) * sizeof(Seq); }
__host__ __device__ constexpr size_t slot_pre_off(uint32_t bm) { return slot_spec_off(bm) + slot_lists_bytes(bm); }
__host__ __device__ constexpr size_t slot_bytes_of(uint32_t bm) { return (slot_pre_off(bm) + slot_lists_bytes(bm) + 255) & ~(size_t)255; }

struct MatchCtl {
  unsigned long long mbar;
  uint32_t item;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

// 8 bytes at byte offset `at` of the staged block (three aligned shared-memory words and two funnel shifts)
__device__ __forceinline__ uint64_t lds64(const uint8_t *base16, uint32_t at) {
  const uint32_t *w = reinterpret_cast<const uint32_t *>(base16) + (at >> 2);
  const uint32_t sh = at << 3;                       // the funnel shift takes it modulo 32
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
  return ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
}

// named barriers: 1 + g = the threads of group g among themselves; 5 + g = "the window before group g's is inserted"
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

struct EsdArgs {
  EncodeArgs A;
  EsdParams E;
  const uint32_t *list;          // null: items are wave_base .. wave_base + wave_n - 1; else indices drawn from list[0 .. *list_count)
  const uint32_t *list_count;
  uint32_t *work_head;
  uint32_t *defer_list;          // items this geometry cannot take: blocks above block_max (null: none can occur) ...
  uint32_t *defer_count;
  uint32_t *big_list;            // ... and items above one block (multi-block frames go to the general kernel)
  uint32_t *big_count;
  uint8_t *slots;                // wave_n x slot_bytes
  uint32_t slot_bytes, slot_block_max;
  uint32_t wave_base, wave_n;
};

template <int DFAST, int BIG>
__global__ void __launch_bounds__(MATCH_THREADS, 1) zstd_lz_match_kernel(EsdArgs K) {
  constexpr uint32_t block_max = BIG ? 131072u : 65536u;
  extern __shared__ __align__(128) uint8_t smem[];
  MatchCtl *const ctl = reinterpret_cast<MatchCtl *>(smem);
  uint8_t *const in_base = smem + 128;                                         // 16-byte aligned
  uint32_t *const tab1 = reinterpret_cast<uint32_t *>(in_base + block_max + LZ_IN_PAD);
  uint32_t *const tab2 = tab1 + ((size_t)1 << K.E.hash_log);
  uint32_t *const first1 = tab2 + (DFAST ? ((size_t)1 << K.E.long_log) : 0);
  uint32_t *const first2 = first1 + (1u << LZ_FIRST_LOG);

  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid >> LZ_WIN_LOG;                     // warp-uniform
  const uint32_t t = (uint32_t)tid & (LZ_WIN - 1u);
  const EncodeArgs &A = K.A;
  const bool blocks_only = A.block_mode != 0;
  const int hash_log = K.E.hash_log, long_log = K.E.long_log, hash_bytes = K.E.hash_bytes;

  if (tid == 0) mbar_init(&ctl->mbar, 1);
  __syncthreads();

  uint32_t phase = 0;
  for (;;) {
    if (tid == 0) {
      const uint32_t idx = atomicAdd(K.work_head, 1u);
      uint32_t item = 0xFFFFFFFFu;
      if (K.list) { if (idx < *K.list_count) item = K.list[idx]; }
      else if (idx < K.wave_n) item = K.wave_base + idx;
      ctl->item = item;
    }
    __syncthreads();
    const uint32_t item = ctl->item;
    if (item == 0xFFFFFFFFu) break;
    uint8_t *const slot = K.slots + (size_t)(item - K.wave_base) * K.slot_bytes;
    BlockHdr *const hdr = reinterpret_cast<BlockHdr *>(slot);
    const uint8_t *const chunk = (const uint8_t *)A.in_ptrs[item];
    const size_t n = A.in_sizes[item];
    uint8_t *const dst = (uint8_t *)A.out_ptrs[item];
    const size_t cap = A.out_sizes[item];
    uint32_t status = ST_OK;
    if (!chunk || !dst) status = ST_INVALID_PARAMETER;
    else if (n == 0) status = ST_INVALID_PARAMETER;                       // reference: manager.cu:1554-1558
    else if (n > 0xFFFF0000ull) status = ST_UNSUPPORTED;
    else if (blocks_only && n > BLOCK_BYTES) status = ST_INVALID_PARAMETER;
    else {
      const size_t nblocks = (n + BLOCK_BYTES - 1) / BLOCK_BYTES;
      if (cap < (blocks_only ? 0 : (size_t)frame_header_size(n) + (A.prm.checksum ? 4 : 0)) + n + 3 * nblocks) status = ST_BUFFER_TOO_SMALL;
    }
    if (status != ST_OK || n > block_max) {
      if (tid == 0) {
        if (status != ST_OK) {
          A.out_sizes[item] = 0;
          if (A.statuses) A.statuses[item] = status;
          hdr->kind = KIND_SKIP;
        } else if (n > K.slot_block_max) {
          K.big_list[atomicAdd(K.big_count, 1u)] = item;            // multi-block (or, for 64 KB slots, > 64 KB) item: the general kernel takes it
          hdr->kind = KIND_SKIP;
        } else K.defer_list[atomicAdd(K.defer_count, 1u)] = item;   // the 128 KB geometry writes this slot later
      }
      __syncthreads();
      continue;
    }
    const uint32_t bn = (uint32_t)n;
    const uint32_t delta = (uint32_t)((uintptr_t)chunk & 15);
    // ---- load: HBM -> shared memory by bulk async copies; tables are cleared while the copy is in flight ----
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t total = (delta + bn + 15u) & ~15u;
      mbar_arrive_tx(&ctl->mbar, total);
      const uint8_t *src = chunk - delta;
      for (uint32_t o = 0; o < total; o += 16384u) bulk_g2s(in_base + o, src + o, min(16384u, total - o), &ctl->mbar);
    }
    {
      uint4 *z = reinterpret_cast<uint4 *>(tab1);
      const uint32_t vecs = (uint32_t)(((size_t)4 << hash_log) + (DFAST ? ((size_t)4 << long_log) : 0)) >> 4;
      for (uint32_t i = tid; i < vecs; i += MATCH_THREADS) z[i] = make_uint4(0, 0, 0, 0);
      for (uint32_t i = tid; i < (2u << LZ_FIRST_LOG); i += MATCH_THREADS) first1[i] = 0xFFFFFFFFu;
    }
    mbar_wait(&ctl->mbar, phase);
    phase ^= 1;
    const uint8_t *const in = in_base + delta;                              // block byte i lives at in[i]
    // ---- RLE block? (decided on the first 32 bytes in the common case) ----
    bool rle = false;
    {
      const uint8_t b0 = in[0];
      bool same = true;
      if ((uint32_t)tid < min(bn, 32u)) same = in[tid] == b0;
      if (__syncthreads_and(same)) {
        for (uint32_t i = tid; i < bn && same; i += MATCH_THREADS) same = in[i] == b0;
        rle = __syncthreads_and(same) && bn > 1;
      }
    }
    __syncthreads();                                                       // tables cleared
    if (tid == 0) hdr->kind = rle ? KIND_RLE : KIND_PARSED;
    if (!rle) {
      const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
      const uint32_t nwin = (ilimit + LZ_WIN - 1) / LZ_WIN;
      uint32_t *const Rg = reinterpret_cast<uint32_t *>(slot + slot_r_off(K.slot_block_max));
      uint32_t *const Mg = reinterpret_cast<uint32_t *>(slot + slot_map_off());
      auto rd = [&](uint32_t q) { return lds64(in_base, delta + q); };
      const int next_grp = (grp + 1) & (MATCH_GROUPS - 1);
      constexpr uint32_t KEY_HMASK = (1u << LZ_KEY_HBITS) - 1u;
      const uint32_t tkey = t << LZ_KEY_HBITS;
      for (uint32_t w = (uint32_t)grp; w < nwin; w += MATCH_GROUPS) {
        const uint32_t p = w * LZ_WIN + t;
        const bool act = p < ilimit;
        const uint64_t v = act ? rd(p) : 0ull;
        const uint32_t h1 = hash_short(v, hash_bytes, hash_log);
        const uint32_t hp1 = __shfl_up_sync(0xffffffffu, h1, 1);
        const bool ins1 = act && inserts((uint32_t)lane, h1, hp1);
        const uint32_t s1 = h1 >> (hash_log - LZ_FIRST_LOG);
        uint32_t h2 = 0, s2 = 0;
        bool ins2 = false;
        if (DFAST) {
          h2 = hash_long(v, long_log);
          const uint32_t hp2 = __shfl_up_sync(0xffffffffu, h2, 1);
          ins2 = act && inserts((uint32_t)lane, h2, hp2);
          s2 = h2 >> (long_log - LZ_FIRST_LOG);
        }
        // the window before this one must be in the tables
        if (w > 0) bar_sync(5 + grp, 2 * (int)LZ_WIN);
        // phase 1: the state before this window, first-of-window side table.  key = tag | t | low hash bits; my own
        // key without t is what a same-hash entry of this window must equal outside the t field
        const uint32_t tagw = (~w & 0x1FFu) << 23;
        const uint32_t k1 = tagw | (h1 & KEY_HMASK), k2 = tagw | (h2 & KEY_HMASK);
        const uint32_t e1 = tab1[h1];
        if (ins1) atomicMin(&first1[s1], k1 | tkey);
        uint32_t e2 = 0;
        if (DFAST) {
          e2 = tab2[h2];
          if (ins2) atomicMin(&first2[s2], k2 | tkey);
        }
        bar_sync(1 + grp, (int)LZ_WIN);
        // phase 2: inserts; the candidate is the first earlier position of this window with my hash, else the old entry
        if (ins1) atomicMax(&tab1[h1], p);
        const uint32_t x1 = first1[s1] ^ k1;
        uint32_t x2 = 0;
        if (DFAST) {
          if (ins2) atomicMax(&tab2[h2], p);
          x2 = first2[s2] ^ k2;
        }
        if (w + 1 < nwin) bar_arrive(5 + next_grp, 2 * (int)LZ_WIN);
        // off the chain: verification (x = t' << KEY_HBITS exactly when the slot holds an earlier thread t' of this
        // window with my hash bits; candidates at or beyond p never verify)
        const uint32_t base_p = p - t;
        const uint32_t c1 = (x1 < tkey && (x1 & KEY_HMASK) == 0) ? base_p + (x1 >> LZ_KEY_HBITS) : e1;
        uint32_t r = 0;
        if (act) {
          uint32_t off = 0, len = 0;
          const uint64_t y1 = rd(min(c1, p));
          if (DFAST) {
            const uint32_t c2 = (x2 < tkey && (x2 & KEY_HMASK) == 0) ? base_p + (x2 >> LZ_KEY_HBITS) : e2;
            const uint64_t y2 = rd(min(c2, p));
            if (c2 < p && y2 == v) { off = p - c2; len = 8; }
          }
          if (len == 0 && c1 < p) {
            const uint32_t c = common8(v, y1);
            if (c >= LZ_MIN_MATCH) { off = p - c1; len = c; }
          }
          if (len == 8 && p + 16 <= bn) len += common8(rd(p + 8), rd(p + 8 - off));
          r = len ? (off | (len << 17)) : 0u;
          Rg[p] = r;
        }
        const uint32_t any = __ballot_sync(0xffffffffu, r != 0);
        if (lane == 0) Mg[p >> 5] = any;
      }
    }
    __syncthreads();                              // every warp is done with the staged block and the tables
  }
}

// -----------------------------------------------------------------------------------------------------------------
// MATCH, levels 5-9: rows (zstd_encode_lz.cuh).  Same window hand-over as above; the 8-byte table holds rows of 16
// tagged entries (128 KB), the 4-byte table single tagged entries.  A position reads its whole row with four 16-byte
// loads, filters the ways by tag, and measures every surviving candidate exactly (up to LZ_QCAP bytes): the lanes of a
// warp take their hits one per round, so a round costs one measurement whatever way the hits sit in.  Blocks above
// 64 KB leave shared memory to the tables and read the block through L1 (BIG).
// -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pick16(const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d, uint32_t y) {
  const uint4 q = (y & 8) ? ((y & 4) ? d : c) : ((y & 4) ? b : a);
  return (y & 2) ? ((y & 1) ? q.w : q.z) : ((y & 1) ? q.y : q.x);
}

template <int BIG>
__global__ void __launch_bounds__(MATCH_THREADS, 1) zstd_lz_rows_kernel(EsdArgs K) {
  constexpr uint32_t block_max = BIG ? 131072u : 65536u;
  extern __shared__ __align__(128) uint8_t smem[];
  MatchCtl *const ctl = reinterpret_cast<MatchCtl *>(smem);
  uint8_t *const in_base = smem + 128;                                         // 16-byte aligned (unused when BIG)
  uint32_t *const rows = reinterpret_cast<uint32_t *>(in_base + (BIG ? 0u : block_max + LZ_IN_PAD));
  uint32_t *const tabs = rows + ((size_t)LZ_ROW_WAYS << K.E.long_log);
  uint32_t *const first1 = tabs + ((size_t)1 << K.E.hash_log);
  uint32_t *const first2 = first1 + (1u << LZ_FIRST_LOG);

  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid >> LZ_WIN_LOG;                     // warp-uniform
  const uint32_t t = (uint32_t)tid & (LZ_WIN - 1u);
  const EncodeArgs &A = K.A;
  const bool blocks_only = A.block_mode != 0;
  const int hash_log = K.E.hash_log, long_log = K.E.long_log, hash_bytes = K.E.hash_bytes;

  if (tid == 0) mbar_init(&ctl->mbar, 1);
  __syncthreads();

  uint32_t phase = 0;
  for (;;) {
    if (tid == 0) {
      const uint32_t idx = atomicAdd(K.work_head, 1u);
      uint32_t item = 0xFFFFFFFFu;
      if (K.list) { if (idx < *K.list_count) item = K.list[idx]; }
      else if (idx < K.wave_n) item = K.wave_base + idx;
      ctl->item = item;
    }
    __syncthreads();
    const uint32_t item = ctl->item;
    if (item == 0xFFFFFFFFu) break;
    uint8_t *const slot = K.slots + (size_t)(item - K.wave_base) * K.slot_bytes;
    BlockHdr *const hdr = reinterpret_cast<BlockHdr *>(slot);
    const uint8_t *const chunk = (const uint8_t *)A.in_ptrs[item];
    const size_t n = A.in_sizes[item];
    uint8_t *const dst = (uint8_t *)A.out_ptrs[item];
    const size_t cap = A.out_sizes[item];
    uint32_t status = ST_OK;
    if (!chunk || !dst) status = ST_INVALID_PARAMETER;
    else if (n == 0) status = ST_INVALID_PARAMETER;                       // reference: manager.cu:1554-1558
    else if (n > 0xFFFF0000ull) status = ST_UNSUPPORTED;
    else if (blocks_only && n > BLOCK_BYTES) status = ST_INVALID_PARAMETER;
    else {
      const size_t nblocks = (n + BLOCK_BYTES - 1) / BLOCK_BYTES;
      if (cap < (blocks_only ? 0 : (size_t)frame_header_size(n) + (A.prm.checksum ? 4 : 0)) + n + 3 * nblocks) status = ST_BUFFER_TOO_SMALL;
    }
    if (status != ST_OK || n > block_max) {
      if (tid == 0) {
        if (status != ST_OK) {
          A.out_sizes[item] = 0;
          if (A.statuses) A.statuses[item] = status;
          hdr->kind = KIND_SKIP;
        } else if (n > K.slot_block_max) {
          K.big_list[atomicAdd(K.big_count, 1u)] = item;
          hdr->kind = KIND_SKIP;
        } else K.defer_list[atomicAdd(K.defer_count, 1u)] = item;
      }
      __syncthreads();
      continue;
    }
    const uint32_t bn = (uint32_t)n;
    const uint32_t delta = (uint32_t)((uintptr_t)chunk & 15);
    if (!BIG && tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t total = (delta + bn + 15u) & ~15u;
      mbar_arrive_tx(&ctl->mbar, total);
      const uint8_t *src = chunk - delta;
      for (uint32_t o = 0; o < total; o += 16384u) bulk_g2s(in_base + o, src + o, min(16384u, total - o), &ctl->mbar);
    }
    {
      uint4 *z = reinterpret_cast<uint4 *>(rows);
      const uint32_t vecs = (uint32_t)((((size_t)4 * LZ_ROW_WAYS) << long_log) + ((size_t)4 << hash_log)) >> 4;
      for (uint32_t i = tid; i < vecs; i += MATCH_THREADS) z[i] = make_uint4(0, 0, 0, 0);
      for (uint32_t i = tid; i < (2u << LZ_FIRST_LOG); i += MATCH_THREADS) first1[i] = 0xFFFFFFFFu;
    }
    if (!BIG) { mbar_wait(&ctl->mbar, phase); phase ^= 1; }
    auto rd = [&](uint32_t q) -> uint64_t { return BIG ? rd64(chunk, q) : lds64(in_base, delta + q); };
    auto byte_at = [&](uint32_t q) -> uint32_t { return BIG ? (uint32_t)chunk[q] : (uint32_t)in_base[delta + q]; };
    bool rle = false;
    {
      const uint32_t b0 = byte_at(0);
      bool same = true;
      if ((uint32_t)tid < min(bn, 32u)) same = byte_at((uint32_t)tid) == b0;
      if (__syncthreads_and(same)) {
        for (uint32_t i = tid; i < bn && same; i += MATCH_THREADS) same = byte_at(i) == b0;
        rle = __syncthreads_and(same) && bn > 1;
      }
    }
    __syncthreads();                                                       // tables cleared
    if (tid == 0) hdr->kind = rle ? KIND_RLE : KIND_PARSED;
    if (!rle) {
      const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
      const uint32_t nwin = (ilimit + LZ_WIN - 1) / LZ_WIN;
      uint32_t *const Rg = reinterpret_cast<uint32_t *>(slot + slot_r_off(K.slot_block_max));
      uint32_t *const Mg = reinterpret_cast<uint32_t *>(slot + slot_map_off());
      const int next_grp = (grp + 1) & (MATCH_GROUPS - 1);
      constexpr uint32_t KEY_HMASK = (1u << LZ_KEY_HBITS) - 1u;
      const uint32_t tkey = t << LZ_KEY_HBITS;
      for (uint32_t w = (uint32_t)grp; w < nwin; w += MATCH_GROUPS) {
        const uint32_t p = w * LZ_WIN + t;
        const bool act = p < ilimit;
        const uint64_t v = act ? rd(p) : 0ull;
        const uint32_t h1 = hash_short(v, hash_bytes, hash_log), h2 = hash_long(v, long_log);
        const uint32_t tg1 = short_tag(v), tg2 = row_tag(v, (act && p + 12 <= bn) ? (uint32_t)(rd(p + 4) >> 32) : 0u);
        const uint32_t hp1 = __shfl_up_sync(0xffffffffu, h1, 1), hp2 = __shfl_up_sync(0xffffffffu, h2, 1);
        const bool ins1 = act && inserts((uint32_t)lane, h1, hp1), ins2 = act && inserts((uint32_t)lane, h2, hp2);
        const uint32_t s1 = h1 >> (hash_log - LZ_FIRST_LOG), s2 = h2 >> (long_log - LZ_FIRST_LOG);
        if (w > 0) bar_sync(5 + grp, 2 * (int)LZ_WIN);
        // phase 1: my row and my short entry as the windows before left them, first-of-window side tables
        const uint32_t tagw = (~w & 0x1FFu) << 23;
        const uint32_t k1 = tagw | (h1 & KEY_HMASK), k2 = tagw | (h2 & KEY_HMASK);
        const uint4 *const rowp = reinterpret_cast<const uint4 *>(rows + (size_t)h2 * LZ_ROW_WAYS);
        const uint4 ra = rowp[0], rb = rowp[1], rc = rowp[2], rdd = rowp[3];
        const uint32_t e1 = tabs[h1];
        if (ins1) atomicMin(&first1[s1], k1 | tkey);
        if (ins2) atomicMin(&first2[s2], k2 | tkey);
        bar_sync(1 + grp, (int)LZ_WIN);
        // phase 2: inserts (way = window mod 16, highest position wins), side-table candidates
        if (ins2) atomicMax(&rows[(size_t)h2 * LZ_ROW_WAYS + (w & (LZ_ROW_WAYS - 1))], row_entry(p, tg2));
        if (ins1) atomicMax(&tabs[h1], row_entry(p, tg1));
        const uint32_t x1 = first1[s1] ^ k1, x2 = first2[s2] ^ k2;
        if (w + 1 < nwin) bar_arrive(5 + next_grp, 2 * (int)LZ_WIN);
        // off the chain: hits = candidates whose tag fits, in the order they are tried (bit 0: first of window for the
        // 8-byte hash, 1: for the 4-byte hash, 2-17: strong ways 0-15, 18: the 4-byte entry, 19: the first weak way, tried
        // only while nothing of 12 bytes is in hand); every round tries one hit per lane
        uint32_t r = 0;
        uint32_t hits = 0, weak_c = 0xFFFFFFFFu;
        const uint32_t base_p = p - t;
        if (act) {
          if (x2 < tkey && (x2 & KEY_HMASK) == 0) hits |= 1u;
          if (x1 < tkey && (x1 & KEY_HMASK) == 0) hits |= 2u;
#pragma unroll
          for (int y = 15; y >= 0; y--) {
            const uint32_t e = pick16(ra, rb, rc, rdd, (uint32_t)y);
            if ((e >> 15) >= p || ((e ^ tg2) & LZ_TAG_HI) != 0) continue;
            if (((e ^ tg2) & LZ_TAG_ALL) == 0) hits |= 4u << y;              // strong: the four bytes behind the hashed ones fit too
            else weak_c = e >> 15;                                            // weak: fewer than 12 equal bytes; the first one is kept
          }
          if ((e1 & 0x7FFFu) == tg1 && (e1 >> 15) < p) hits |= 1u << 18;
          if (weak_c != 0xFFFFFFFFu) hits |= 1u << 19;
        }
        uint32_t best_len = 0, best_off = 0;
        while (__any_sync(0xffffffffu, hits != 0)) {
          if (hits) {
            const uint32_t y = (uint32_t)__ffs((int)hits) - 1;
            hits &= hits - 1;
            const uint32_t c = y == 0 ? base_p + (x2 >> LZ_KEY_HBITS) : y == 1 ? base_p + (x1 >> LZ_KEY_HBITS) : y < 18 ? pick16(ra, rb, rc, rdd, y - 2) >> 15
                               : y == 18 ? e1 >> 15 : weak_c;
            if (y < 19 || best_len < 12) try_candidate(rd, p, c, bn, best_len, best_off);
          }
        }
        if (act) {
          r = best_len >= LZ_Q_MIN_MATCH ? (best_off | (best_len << 17)) : 0u;
          Rg[p] = r;
        }
        const uint32_t any = __ballot_sync(0xffffffffu, r != 0);
        if (lane == 0) Mg[p >> 5] = any;
      }
    }
    __syncthreads();                              // every warp is done with the staged block and the tables
  }
}

// SELECT, levels 5-9: the shared walk (select_walk / select_rewalk of zstd_encode_lz.cuh) as it stands, lane by lane --
// the lazy decisions of these levels read more of R and measure repeat-offset matches exactly, and the match kernel is
// what bounds these levels
constexpr int SELQ_WARPS = 4;
// -----------------------------------------------------------------------------------------------------------------
// SELECT: one warp per block, lane j walks sub-segment j.
// -----------------------------------------------------------------------------------------------------------------
constexpr int SEL_WARPS = 4;
struct SelArgs {
  EncodeArgs A;
  uint8_t *slots;
  uint32_t slot_bytes, slot_block_max;
  uint32_t wave_base, wave_n;
  int lazy;
};
__device__ __forceinline__ State shfl_up_state(const State &s) {
  State r;
  r.ip = __shfl_up_sync(0xffffffffu, s.ip, 1); r.anchor = __shfl_up_sync(0xffffffffu, s.anchor, 1);
  r.r0 = __shfl_up_sync(0xffffffffu, s.r0, 1); r.r1 = __shfl_up_sync(0xffffffffu, s.r1, 1); r.r2 = __shfl_up_sync(0xffffffffu, s.r2, 1);
  return r;
}

// The walk of zstd_encode_lz.cuh (select_step / select_walk / select_rewalk), restated for a warp whose lanes walk their
// own sub-segments: one loop iteration takes every lane through the stages of one step -- decide, finish an open match,
// extend backwards, emit -- and the lanes that are in the same stage execute it together, with all loads of a stage in
// flight at once.  Literal skipping reads the "R != 0" bitmap, 32 positions per load.  Same sequences as the shared code.
struct Walker {
  const uint8_t *in;
  const uint32_t *R, *map;
  uint32_t n, lim;
  int lazy;
};
enum : uint32_t { W_DECIDE = 0, W_SKIP = 1, W_MATCH = 2, W_DONE = 3, W_RECODE = 4 };

template <bool REWALK>
__device__ __forceinline__ uint32_t walk_lanes(const Walker &K, State &st, bool active, Seq *out, const Seq *spec, uint32_t spec_cnt, const State &spec0,
                                               const State &spec_exit, uint32_t *sync_k) {
  uint32_t cnt = 0, k = 0;
  State sp = spec0;
  uint32_t mode = (active && st.ip < K.lim) ? W_DECIDE : W_DONE;
  uint32_t s = 0, off = 0, len = 0, q = 0, nb = 0;
  bool synced = false, open = false, back_done = false;
  while (__any_sync(0xffffffffu, mode != W_DONE)) {
    // ---- decide: every load of the decision is independent of the others (one round trip), the bitmap word that a
    //      literal skip would need included ----
    if (mode == W_DECIDE) {
      const uint32_t ip = st.ip;
      const bool has1 = ip + 1 < K.lim;
      const uint32_t e0 = K.R[ip], e1 = has1 ? K.R[ip + 1] : 0u;
      const uint32_t mw = K.map[(ip + 1) >> 5];
      const uint64_t v0 = rd64(K.in, ip);
      const uint32_t nxt = has1 ? rd8(K.in, ip + 8) : 0u;
      const bool behind = ip == st.anchor;
      const bool try1 = behind && st.r1 != 0 && ip >= st.r1;
      const bool try0a = st.r0 != 0 && ip >= st.r0 && !behind, try0b = st.r0 != 0 && ip + 1 >= st.r0 && has1;
      uint64_t y = 0, x0 = 0;
      if (try1) y = rd64(K.in, ip - st.r1);
      // one load pair serves the repeat-offset checks at ip and ip + 1 when both are in range
      if (try0a) x0 = rd64(K.in, ip - st.r0);
      uint64_t x1 = 0;
      if (try0b) x1 = try0a ? ((x0 >> 8) | ((uint64_t)rd8(K.in, ip - st.r0 + 8) << 56)) : rd64(K.in, ip + 1 - st.r0);
      const uint64_t v1 = (v0 >> 8) | ((uint64_t)nxt << 56);
      const uint32_t lt0 = e0 >> 17;
      bool found = false;
      if (try1) {
        const uint32_t c = common8(v0, y);
        if (c >= 4 && (c == 8 || c + LZ_REP_BONUS >= lt0)) { s = ip; off = st.r1; len = c; open = c == 8; found = true; }
      }
      if (!found) {
        const uint32_t rl0 = try0a ? common8(v0, x0) : 0u, rl1 = try0b ? common8(v1, x1) : 0u;
        const bool t0 = e0 != 0, r0ok = rl0 >= 4, r1ok = rl1 >= 4;
        if (!t0 && !r0ok && !r1ok) {
          // literals: on to the next position with a table candidate; the first bitmap word is already here
          q = ip + 1;
          const uint32_t word = mw & (0xFFFFFFFFu << (q & 31));
          q = word ? (q & ~31u) + (uint32_t)(__ffs((int)word) - 1) : (q | 31u) + 1;
          if (word || q >= K.lim) { st.ip = min(q, K.lim); mode = st.ip < K.lim ? W_DECIDE : W_DONE; }
          else mode = W_SKIP;
        } else {
          if (r0ok && (!t0 || rl0 == 8 || rl0 + LZ_REP_BONUS >= lt0)) { s = ip; off = st.r0; len = rl0; open = rl0 == 8; }
          else if (r1ok && (!t0 || rl1 == 8 || rl1 + LZ_REP_BONUS >= lt0)) { s = ip + 1; off = st.r0; len = rl1; open = rl1 == 8; }
          else {
            uint32_t e = e0;
            s = ip;
            if (K.lazy && (e1 >> 17) > lt0) { e = e1; s = ip + 1; }
            off = e & LZ_OFF_MASK; len = e >> 17; open = len == LZ_LCAP;
          }
          found = true;
        }
      }
      if (found) { mode = W_MATCH; back_done = false; }
    } else if (mode == W_SKIP) {
      // next position >= q with a table candidate, one bitmap word per turn
      const uint32_t word = K.map[q >> 5] & (0xFFFFFFFFu << (q & 31));
      q = word ? (q & ~31u) + (uint32_t)(__ffs((int)word) - 1) : (q | 31u) + 1;
      if (word || q >= K.lim) { st.ip = min(q, K.lim); mode = st.ip < K.lim ? W_DECIDE : W_DONE; }
    }
    // ---- a match in hand: the loads of its forward extension (16 bytes per turn) and of its backward extension leave
    //      together; the sequence is emitted in the turn its forward extension ends ----
    if (mode == W_MATCH) {
      const uint32_t a = s + len;
      const bool fast = open && a + 16 <= K.n;
      uint64_t p0 = 0, p1 = 0, c0 = 0, c1 = 0, b0 = 0, b1 = 0;
      const uint32_t room = back_done ? 0u : min(min(LZ_BACK_MAX, s - st.anchor), s >= off ? s - off : 0u);
      const bool back8 = room != 0 && s >= off + 8;
      if (fast) { p0 = rd64(K.in, a); p1 = rd64(K.in, a + 8); c0 = rd64(K.in, a - off); c1 = rd64(K.in, a + 8 - off); }
      if (back8) { b0 = rd64(K.in, s - 8); b1 = rd64(K.in, s - 8 - off); }
      if (open) {
        if (fast) {
          uint32_t c = common8(p0, c0);
          if (c == 8) c += common8(p1, c1);
          len += c;
          if (c < 16) open = false;
        } else {
          if (a + 8 <= K.n) { const uint32_t c = common8(rd64(K.in, a), rd64(K.in, a - off)); len += c; if (c == 8) { while (s + len < K.n && rd8(K.in, s + len) == rd8(K.in, s + len - off)) len++; } }
          else while (s + len < K.n && rd8(K.in, s + len) == rd8(K.in, s + len - off)) len++;
          open = false;
        }
      }
      if (!back_done) {
        nb = 0;
        if (back8) {
          const uint64_t x = b0 ^ b1;
          const uint32_t xh = (uint32_t)(x >> 32), xl = (uint32_t)x;
          nb = xh ? (uint32_t)__clz((int)xh) >> 3 : xl ? 4u + ((uint32_t)__clz((int)xl) >> 3) : 8u;
          nb = min(nb, room);
        } else while (nb < room && rd8(K.in, s - nb - 1) == rd8(K.in, s - nb - 1 - off)) nb++;
        back_done = true;
      }
      if (!open) {
        s -= nb; len += nb;
        const uint32_t code = code_offset(off, s == st.anchor, st.r0, st.r1, st.r2);
        out[cnt++] = pack_seq(s, len, code);
        st.ip = st.anchor = s + len;
        mode = st.ip < K.lim ? W_DECIDE : W_DONE;
        if (REWALK) {
          while (k < spec_cnt) {
            const Seq e = spec[k];
            const uint32_t es = seq_start(e), ee = es + seq_len(e);
            if (ee > st.anchor) break;
            decode_offset(seq_code(e), es == sp.anchor, sp.r0, sp.r1, sp.r2);
            sp.ip = sp.anchor = ee;
            k++;
            if (ee == st.anchor) break;
          }
          // Same position and the same two youngest history entries: every decision from here on is the speculative
          // walk's (decisions never read the third entry).  With the third entry equal too the lists are identical;
          // otherwise only offset codes that involve it can differ, and the lane goes on copying the speculative
          // sequences, re-coded on the true history, until a new offset pushes the differing entry out.
          if (sp.anchor == st.anchor && sp.r0 == st.r0 && sp.r1 == st.r1 && k > 0) {
            if (sp.r2 == st.r2) { synced = true; mode = W_DONE; }
            else mode = W_RECODE;
          }
        }
      }
    } else if (REWALK && mode == W_RECODE) {
      if (k == spec_cnt) { st.ip = spec_exit.ip; st.anchor = spec_exit.anchor; mode = W_DONE; }
      else {
        const Seq e = spec[k];
        const uint32_t es = seq_start(e), el = seq_len(e);
        const bool ll0 = es == st.anchor;
        const uint32_t o = decode_offset(seq_code(e), ll0, sp.r0, sp.r1, sp.r2);
        out[cnt++] = pack_seq(es, el, code_offset(o, ll0, st.r0, st.r1, st.r2));
        st.ip = st.anchor = sp.anchor = es + el;
        k++;
        if (sp.r2 == st.r2) { synced = true; mode = W_DONE; }
      }
    }
  }
  if (REWALK && active) {
    *sync_k = synced ? k : spec_cnt;
    if (synced) st = spec_exit;
  }
  return cnt;
}

__global__ void __launch_bounds__(32 * SEL_WARPS, 8) zstd_lz_select_kernel(SelArgs S) {
  const int lane = threadIdx.x & 31;
  const uint32_t idx = blockIdx.x * SEL_WARPS + (threadIdx.x >> 5);
  if (idx >= S.wave_n) return;
  const uint32_t item = S.wave_base + idx;
  uint8_t *const slot = S.slots + (size_t)idx * S.slot_bytes;
  if (reinterpret_cast<const BlockHdr *>(slot)->kind != KIND_PARSED) return;
  const uint32_t bn = (uint32_t)S.A.in_sizes[item];
  const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
  const uint32_t cap = lane_list_cap(S.slot_block_max);
  Seq *const spec = reinterpret_cast<Seq *>(slot + slot_spec_off(S.slot_block_max)) + (size_t)lane * cap;
  Seq *const prefix = reinterpret_cast<Seq *>(slot + slot_pre_off(S.slot_block_max)) + (size_t)lane * cap;
  const uint32_t span = lane_span(ilimit);
  const uint32_t B = lane_begin((uint32_t)lane, span, ilimit), E = lane_begin((uint32_t)lane + 1, span, ilimit);
  Walker K;
  K.in = (const uint8_t *)S.A.in_ptrs[item];
  K.R = reinterpret_cast<const uint32_t *>(slot + slot_r_off(S.slot_block_max));
  K.map = reinterpret_cast<const uint32_t *>(slot + slot_map_off());
  K.n = bn; K.lim = E; K.lazy = S.lazy;
  // the history a decoder holds at the start of the block: the format's initial one, unless the block is coded on its
  // own inside a larger frame (block mode), where it is unknown (0 never matches and is never written as a repeat code)
  State st{B, B, 0, 0, 0};
  if (lane == 0 && !(S.A.block_mode != 0 && item != 0)) { st.r0 = 1; st.r1 = 4; st.r2 = 8; }
  const State spec0 = st;
  uint32_t sync_k = 0;
  const uint32_t spec_cnt = walk_lanes<false>(K, st, true, spec, nullptr, 0, spec0, spec0, &sync_k);
  const State spec_exit = st;
  State exit_state = st, entry_used = spec0;
  uint32_t pre_cnt = 0;
  sync_k = 0;
  for (;;) {
    const State entry = shfl_up_state(exit_state);
    const bool need = lane > 0 && !entry.same(entry_used);
    if (!__any_sync(0xffffffffu, need)) break;
    // An entry that differs from the last one only in the third history entry leads to the same decisions: the lane's
    // lists are re-coded (prefix in place, then the kept speculative sequences are copied behind it) until the two
    // histories meet, without reading the input again (select_recode of zstd_encode_lz.cuh, written out: the call costs this kernel 3 ms per 16,384 chunks).
    const bool recode = need && entry.ip == entry_used.ip && entry.anchor == entry_used.anchor && entry.r0 == entry_used.r0 && entry.r1 == entry_used.r1;
    if (recode) {
      uint32_t o0 = entry_used.r0, o1 = entry_used.r1, o2 = entry_used.r2, n0 = entry.r0, n1 = entry.r1, n2 = entry.r2;
      uint32_t anchor = entry.anchor;
      bool met = false;
      for (uint32_t i = 0; i < pre_cnt && !met; i++) {
        const Seq e = prefix[i];
        const uint32_t es = seq_start(e), el = seq_len(e);
        const bool ll0 = es == anchor;
        const uint32_t o = decode_offset(seq_code(e), ll0, o0, o1, o2);
        const uint32_t code = code_offset(o, ll0, n0, n1, n2);
        if (code != seq_code(e)) prefix[i] = pack_seq(es, el, code);
        anchor = es + el;
        met = o2 == n2;
      }
      if (!met) {
        uint32_t k = sync_k;
        while (k < spec_cnt && !met) {
          const Seq e = spec[k];
          const uint32_t es = seq_start(e), el = seq_len(e);
          const bool ll0 = es == anchor;
          const uint32_t o = decode_offset(seq_code(e), ll0, o0, o1, o2);
          prefix[pre_cnt++] = pack_seq(es, el, code_offset(o, ll0, n0, n1, n2));
          anchor = es + el;
          k++;
          met = o2 == n2;
        }
        sync_k = k;
        if (!met) { exit_state.r0 = n0; exit_state.r1 = n1; exit_state.r2 = n2; }
      }
      entry_used = entry;
    }
    const bool walk = need && !recode;
    State s2 = entry;
    const uint32_t c = walk_lanes<true>(K, s2, walk, prefix, spec, spec_cnt, spec0, spec_exit, &sync_k);
    if (walk) { entry_used = entry; pre_cnt = c; exit_state = s2; }
  }
  LaneHdr h;
  h.spec_cnt = (uint16_t)spec_cnt; h.sync_k = (uint16_t)sync_k; h.pre_cnt = (uint16_t)pre_cnt; h.pad = 0;
  reinterpret_cast<LaneHdr *>(slot + 16)[lane] = h;
}

__global__ void __launch_bounds__(32 * SELQ_WARPS) zstd_lz_select_rows_kernel(SelArgs S) {
  const int lane = threadIdx.x & 31;
  const uint32_t idx = blockIdx.x * SELQ_WARPS + (threadIdx.x >> 5);
  if (idx >= S.wave_n) return;
  const uint32_t item = S.wave_base + idx;
  uint8_t *const slot = S.slots + (size_t)idx * S.slot_bytes;
  if (reinterpret_cast<const BlockHdr *>(slot)->kind != KIND_PARSED) return;
  const uint8_t *const in = (const uint8_t *)S.A.in_ptrs[item];
  const uint32_t bn = (uint32_t)S.A.in_sizes[item];
  const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
  const uint32_t *const R = reinterpret_cast<const uint32_t *>(slot + slot_r_off(S.slot_block_max));
  const uint32_t cap = lane_list_cap(S.slot_block_max);
  Seq *const spec = reinterpret_cast<Seq *>(slot + slot_spec_off(S.slot_block_max)) + (size_t)lane * cap;
  Seq *const prefix = reinterpret_cast<Seq *>(slot + slot_pre_off(S.slot_block_max)) + (size_t)lane * cap;
  const SelectParams SP{S.lazy, 1};
  const uint32_t span = lane_span(ilimit);
  const uint32_t B = lane_begin((uint32_t)lane, span, ilimit), E = lane_begin((uint32_t)lane + 1, span, ilimit);
  State st{B, B, 0, 0, 0};
  if (lane == 0 && !(S.A.block_mode != 0 && item != 0)) { st.r0 = 1; st.r1 = 4; st.r2 = 8; }
  const State spec0 = st;
  const uint32_t spec_cnt = select_walk(in, bn, R, E, SP, st, spec);
  const State spec_exit = st;
  State exit_state = st, entry_used = spec0;
  uint32_t pre_cnt = 0, sync_k = 0;
  __syncwarp();
  for (;;) {
    const State entry = shfl_up_state(exit_state);
    const bool need = lane > 0 && !entry.same(entry_used);
    if (!__any_sync(0xffffffffu, need)) break;
    if (need && entry_differs_in_r2_only(entry, entry_used)) {
      select_recode(prefix, pre_cnt, spec, spec_cnt, sync_k, entry_used, entry, exit_state);
      entry_used = entry;
    } else if (need) {
      entry_used = entry;
      State s2 = entry;
      pre_cnt = select_rewalk(in, bn, R, E, SP, s2, spec, spec_cnt, spec0, spec_exit, prefix, &sync_k);
      exit_state = s2;
    }
    __syncwarp();
  }
  LaneHdr h;
  h.spec_cnt = (uint16_t)spec_cnt; h.sync_k = (uint16_t)sync_k; h.pre_cnt = (uint16_t)pre_cnt; h.pad = 0;
  reinterpret_cast<LaneHdr *>(slot + 16)[lane] = h;
}

// -----------------------------------------------------------------------------------------------------------------
// FINISH: one warp per block.  Joins the lane lists into the three sequence arrays, gathers the literals from the
// input, codes the block and writes the frame (or the bare block in block mode).
// -----------------------------------------------------------------------------------------------------------------
constexpr int FIN_WARPS = 4;                     // warps per CTA (each with its own entropy workspace)
struct FinArgs {
  EncodeArgs A;
  const uint8_t *slots;
  uint32_t slot_bytes, slot_block_max;
  uint32_t wave_base, wave_n;
  uint32_t *work_head;
  uint8_t *scratch;          // per warp: literals + three uint32 sequence arrays
};
__host__ __device__ constexpr size_t fin_warp_scratch(uint32_t block_max) { return ((size_t)block_max + 64 + (size_t)3 * (block_max / 4 + 64) * 4 + 255) & ~(size_t)255; }

// copies n bytes src -> dst with the whole warp (dst and src are unrelated in alignment)
// Long runs (a literal-only block is one run of 64 KB) go as 16-byte stores fed by aligned 32-bit loads and funnel shifts,
// two vectors per lane in flight; the last vector is left to the byte loop so that no word past src + n is read.
__device__ __forceinline__ void warp_copy_bytes(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
  uint32_t done = 0;
  if (n >= 128) {
    const uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    for (uint32_t i = (uint32_t)lane; i < head; i += 32) dst[i] = src[i];
    const uint8_t *const s0 = src + head;
    const uint32_t sh = (uint32_t)((uintptr_t)s0 & 3) * 8;
    const uint32_t *const w = reinterpret_cast<const uint32_t *>((uintptr_t)s0 & ~(uintptr_t)3);
    uint4 *const d4 = reinterpret_cast<uint4 *>(dst + head);
    const uint32_t vecs = ((n - head) >> 4) - 1;
    uint32_t j = (uint32_t)lane;
    for (; j + 32 < vecs; j += 64) {
      uint32_t a[5], b[5];
#pragma unroll
      for (int q = 0; q < 5; q++) { a[q] = w[4 * j + q]; b[q] = w[4 * (j + 32) + q]; }
      d4[j] = make_uint4(__funnelshift_r(a[0], a[1], sh), __funnelshift_r(a[1], a[2], sh), __funnelshift_r(a[2], a[3], sh), __funnelshift_r(a[3], a[4], sh));
      d4[j + 32] = make_uint4(__funnelshift_r(b[0], b[1], sh), __funnelshift_r(b[1], b[2], sh), __funnelshift_r(b[2], b[3], sh), __funnelshift_r(b[3], b[4], sh));
    }
    for (; j < vecs; j += 32) {
      uint32_t a[5];
#pragma unroll
      for (int q = 0; q < 5; q++) a[q] = w[4 * j + q];
      d4[j] = make_uint4(__funnelshift_r(a[0], a[1], sh), __funnelshift_r(a[1], a[2], sh), __funnelshift_r(a[2], a[3], sh), __funnelshift_r(a[3], a[4], sh));
    }
    done = head + 16 * vecs;
  }
  for (uint32_t i = done + (uint32_t)lane; i < n; i += 32) dst[i] = src[i];
}

__global__ void __launch_bounds__(32 * FIN_WARPS, 8) zstd_lz_finish_kernel(FinArgs F) {
  __shared__ EntropyWs s_ws[FIN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  EntropyWs &W = s_ws[warp];
  const EncodeArgs &A = F.A;
  const EncodeParams P = A.prm;
  const bool blocks_only = A.block_mode != 0;
  const uint32_t max_seq = F.slot_block_max / 4 + 64;
  uint8_t *const buf = F.scratch + (size_t)(blockIdx.x * FIN_WARPS + warp) * fin_warp_scratch(F.slot_block_max);
  uint8_t *const lits = buf;
  uint32_t *const s_ll = reinterpret_cast<uint32_t *>(buf + F.slot_block_max + 64);
  uint32_t *const s_ml = s_ll + max_seq;
  uint32_t *const s_of = s_ml + max_seq;
  const uint32_t cap = lane_list_cap(F.slot_block_max);
  for (;;) {
    uint32_t idx = 0;
    if (lane == 0) idx = atomicAdd(F.work_head, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= F.wave_n) break;
    const uint32_t item = F.wave_base + idx;
    const uint8_t *const slot = F.slots + (size_t)idx * F.slot_bytes;
    const uint32_t kind = reinterpret_cast<const BlockHdr *>(slot)->kind;
    if (kind == KIND_SKIP) continue;
    const uint8_t *const chunk = (const uint8_t *)A.in_ptrs[item];
    const uint32_t bn = (uint32_t)A.in_sizes[item];
    uint8_t *const dst = (uint8_t *)A.out_ptrs[item];
    const bool last = blocks_only ? item + 1 == A.n : true;
    size_t op = 0;
    if (!blocks_only) {
      if (lane == 0) op = write_frame_header(dst, bn, P.checksum != 0);
      op = __shfl_sync(0xffffffffu, (unsigned long long)op, 0);
    }
    if (kind == KIND_RLE) {
      if (lane == 0) { write_block_header(dst + op, last, 1, bn); dst[op + 3] = chunk[0]; }
      op += 4;
    } else {
      // ---- join the lists: lane j contributed prefix[0 .. pre_cnt) then spec[sync_k .. spec_cnt) ----
      const LaneHdr mine = reinterpret_cast<const LaneHdr *>(slot + 16)[lane];
      const Seq *const spec_all = reinterpret_cast<const Seq *>(slot + slot_spec_off(F.slot_block_max));
      const Seq *const pre_all = reinterpret_cast<const Seq *>(slot + slot_pre_off(F.slot_block_max));
      uint32_t nseq = 0, nlit = 0, prev_end = 0;
      for (int j = 0; j < 32; j++) {
        const uint32_t pc = __shfl_sync(0xffffffffu, (uint32_t)mine.pre_cnt, j), sk = __shfl_sync(0xffffffffu, (uint32_t)mine.sync_k, j),
                       sc = __shfl_sync(0xffffffffu, (uint32_t)mine.spec_cnt, j);
        const uint32_t cnt = pc + (sc - sk);
        const Seq *const pl = pre_all + (size_t)j * cap, *const sl = spec_all + (size_t)j * cap + sk;
        for (uint32_t b = 0; b < cnt; b += 32) {
          const uint32_t k = b + (uint32_t)lane;
          const bool valid = k < cnt;
          uint32_t s = 0, ml = 0, code = 0;
          if (valid) {
            const Seq q = k < pc ? pl[k] : sl[k - pc];
            s = seq_start(q); ml = seq_len(q); code = seq_code(q);
          }
          const uint32_t e = s + ml;
          uint32_t pe = __shfl_up_sync(0xffffffffu, e, 1);
          if (lane == 0) pe = prev_end;
          const uint32_t ll = valid ? s - pe : 0u;
          uint32_t sl_incl = ll;                       // inclusive prefix sum of the literal runs of this round
          for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, sl_incl, o); if (lane >= o) sl_incl += t; }
          const uint32_t my_lit = nlit + sl_incl - ll;
          if (valid) { s_ll[nseq + lane] = ll; s_ml[nseq + lane] = ml; s_of[nseq + lane] = code; }
          // literal runs: short ones by their own lane (one 8-byte read when the block has the room), long ones by the whole warp
          if (valid && ll != 0 && ll <= 8) {
            if (pe + 8 <= bn) {
              const uint64_t v = rd64(chunk, pe);
              for (uint32_t i = 0; i < ll; i++) lits[my_lit + i] = (uint8_t)(v >> (8 * i));
            } else for (uint32_t i = 0; i < ll; i++) lits[my_lit + i] = chunk[pe + i];
          }
          uint32_t big = __ballot_sync(0xffffffffu, valid && ll > 8);
          while (big) {
            const int src = __ffs((int)big) - 1;
            big &= big - 1;
            warp_copy_bytes(lits + __shfl_sync(0xffffffffu, my_lit, src), chunk + __shfl_sync(0xffffffffu, pe, src), __shfl_sync(0xffffffffu, ll, src), lane);
          }
          const uint32_t nv = min(32u, cnt - b);
          nlit += __shfl_sync(0xffffffffu, sl_incl, 31);
          prev_end = __shfl_sync(0xffffffffu, e, (int)nv - 1);
          nseq += nv;
        }
      }
      // literals after the last sequence (the whole block when nothing was parsed)
      warp_copy_bytes(lits + nlit, chunk + prev_end, bn - prev_end, lane);
      nlit += bn - prev_end;
      __syncwarp();
      const uint32_t payload = entropy_stage_warp(W, lits, nlit, s_ll, s_ml, s_of, nseq, dst + op + 3, bn - 1, lane);
      if (payload == 0 || payload >= bn) {
        if (lane == 0) write_block_header(dst + op, last, 0, bn);
        for (uint32_t i = lane; i < bn; i += 32) dst[op + 3 + i] = chunk[i];
        op += 3 + (size_t)bn;
      } else {
        if (lane == 0) write_block_header(dst + op, last, 2, payload);
        op += 3 + (size_t)payload;
      }
    }
    __syncwarp();
    if (P.checksum && !blocks_only) {
      const uint64_t h = xxh64_warp(chunk, bn, lane);
      if (lane == 0) { dst[op] = (uint8_t)h; dst[op + 1] = (uint8_t)(h >> 8); dst[op + 2] = (uint8_t)(h >> 16); dst[op + 3] = (uint8_t)(h >> 24); }
      op += 4;
    }
    if (lane == 0) {
      A.out_sizes[item] = op;
      if (A.statuses) A.statuses[item] = ST_OK;
    }
    __syncwarp();
  }
}

size_t match_smem_bytes(const EsdParams &e, int big) {
  const size_t in = (big ? 131072u : 65536u) + LZ_IN_PAD;
  const size_t tabs = ((size_t)4 << e.hash_log) + (e.dfast ? ((size_t)4 << e.long_log) : 0);
  return 128 + in + tabs + ((size_t)8 << LZ_FIRST_LOG);
}

size_t rows_smem_bytes(const EsdParams &e, int big) {
  const size_t in = big ? 0 : 65536u + LZ_IN_PAD;
  return 128 + in + (((size_t)4 * LZ_ROW_WAYS) << e.long_log) + ((size_t)4 << e.hash_log) + ((size_t)8 << LZ_FIRST_LOG);
}
template <int BIG> cudaError_t rows_launch_one(const EsdArgs &k, int grid, cudaStream_t stream) {
  const size_t smem = rows_smem_bytes(k.E, BIG);
  cudaError_t e = cudaFuncSetAttribute(zstd_lz_rows_kernel<BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  zstd_lz_rows_kernel<BIG><<<grid, MATCH_THREADS, smem, stream>>>(k);
  return cudaGetLastError();
}

template <int DFAST, int BIG> cudaError_t match_launch_one(const EsdArgs &k, int grid, cudaStream_t stream) {
  const size_t smem = match_smem_bytes(k.E, BIG);
  cudaError_t e = cudaFuncSetAttribute(zstd_lz_match_kernel<DFAST, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  zstd_lz_match_kernel<DFAST, BIG><<<grid, MATCH_THREADS, smem, stream>>>(k);
  return cudaGetLastError();
}

constexpr int FIN_CTAS_PER_SM = 8;               // 32 finish warps per SM

// blocks per wave: bounds the scratch (8 bytes per input byte of the wave)
uint32_t esd_wave(size_t n, uint32_t block_max) { return (uint32_t)min(n, (size_t)(block_max > 65536 ? 4096 : 8192)); }

} // namespace

size_t esd_scratch_bytes(size_t n_items, uint32_t bm, int sm_count) {
  const size_t lists = (size_t)esd_wave(n_items, bm) * slot_bytes_of(bm);
  const size_t fin_warps = min((size_t)sm_count * FIN_CTAS_PER_SM * FIN_WARPS, ((n_items + FIN_WARPS - 1) / FIN_WARPS) * FIN_WARPS);
  return lists + fin_warps * fin_warp_scratch(bm);
}
size_t esd_counter_words(size_t n_items, uint32_t bm) {
  const uint32_t wave = esd_wave(n_items, bm);
  return 4 * ((n_items + wave - 1) / wave) + 2;
}

// Levels 1-4.  Per wave: the 64 KB-block match kernel, the 128 KB-block match kernel (half the table size)
// over what the first handed on, the select kernel, the finish kernel; after the last wave the general kernel
// (zstd_encode.cu) for multi-block items.  max_item_bytes != 0 (the host knows the sizes) skips the launches that cannot have work.
cudaError_t launch_encode_esd(const EncodeArgs &args, const EsdLaunch &L, cudaStream_t stream, int *launches) {
  if (launches) *launches = 0;
  if (args.n == 0) return cudaSuccess;
  const bool known = L.max_item_bytes != 0;
  const bool any64 = !known || L.min_item_bytes <= 65536;
  const uint32_t bm = L.block_max;
  const bool any128 = bm > 65536 && (!known || (L.max_item_bytes > 65536 && L.min_item_bytes <= bm));
  const bool any_big = !known || L.max_item_bytes > bm;
  const bool only_big = known && L.min_item_bytes > bm;
  const uint32_t wave = esd_wave(args.n, bm);
  const uint32_t slot_bytes = (uint32_t)slot_bytes_of(bm);
  const size_t lists_bytes = (size_t)wave * slot_bytes;
  if (L.scratch_bytes < esd_scratch_bytes(args.n, bm, L.sm_count)) return cudaErrorInvalidValue;
  const size_t nwaves = only_big ? 0 : (args.n + wave - 1) / wave;
  // counters: per wave {match64 head, match128 head, finish head, defer count}; then general-kernel head and big count
  const size_t ctr_words = 4 * nwaves + 2;
  if (ctr_words > L.counter_words) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(L.counters, 0, ctr_words * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  uint32_t *const big_count = L.counters + 4 * nwaves + 1;
  uint32_t *const list128 = L.lists, *const list_big = L.lists + args.n;
  int nl = 0;
  for (size_t w = 0; w < nwaves; w++) {
    uint32_t *const c = L.counters + 4 * w;
    EsdArgs k{};
    k.A = args;
    k.slots = L.scratch; k.slot_bytes = slot_bytes; k.slot_block_max = bm;
    k.wave_base = (uint32_t)(w * wave); k.wave_n = (uint32_t)min((size_t)wave, args.n - w * wave);
    k.big_list = list_big; k.big_count = big_count;
    if (any64) {
      k.E = esd_params_for_level(args.prm.level, 0);
      k.list = nullptr; k.list_count = nullptr;
      k.work_head = c + 0;
      k.defer_list = list128 + k.wave_base; k.defer_count = c + 3;
      const int grid = (int)min((size_t)k.wave_n, (size_t)L.sm_count);
      e = k.E.rows ? rows_launch_one<0>(k, grid, stream) : k.E.dfast ? match_launch_one<1, 0>(k, grid, stream) : match_launch_one<0, 0>(k, grid, stream);
      if (e != cudaSuccess) return e;
      nl++;
    }
    if (any128) {
      k.E = esd_params_for_level(args.prm.level, 1);
      if (any64) { k.list = list128 + k.wave_base; k.list_count = c + 3; } else { k.list = nullptr; k.list_count = nullptr; }
      k.work_head = c + 1;
      k.defer_list = nullptr; k.defer_count = nullptr;
      const int grid = (int)min((size_t)k.wave_n, (size_t)L.sm_count);
      e = k.E.rows ? rows_launch_one<1>(k, grid, stream) : k.E.dfast ? match_launch_one<1, 1>(k, grid, stream) : match_launch_one<0, 1>(k, grid, stream);
      if (e != cudaSuccess) return e;
      nl++;
    }
    SelArgs s{};
    s.A = args;
    s.slots = L.scratch; s.slot_bytes = slot_bytes; s.slot_block_max = bm;
    s.wave_base = k.wave_base; s.wave_n = k.wave_n;
    s.lazy = esd_params_for_level(args.prm.level, 0).lazy;
    if (esd_params_for_level(args.prm.level, 0).rows) zstd_lz_select_rows_kernel<<<(k.wave_n + SELQ_WARPS - 1) / SELQ_WARPS, 32 * SELQ_WARPS, 0, stream>>>(s);
    else zstd_lz_select_kernel<<<(k.wave_n + SEL_WARPS - 1) / SEL_WARPS, 32 * SEL_WARPS, 0, stream>>>(s);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    nl++;
    FinArgs f{};
    f.A = args;
    f.slots = L.scratch; f.slot_bytes = slot_bytes; f.slot_block_max = bm;
    f.wave_base = k.wave_base; f.wave_n = k.wave_n;
    f.work_head = c + 2;
    f.scratch = L.scratch + lists_bytes;
    const int fgrid = (int)min((size_t)L.sm_count * FIN_CTAS_PER_SM, ((size_t)k.wave_n + FIN_WARPS - 1) / FIN_WARPS);
    zstd_lz_finish_kernel<<<fgrid, 32 * FIN_WARPS, 0, stream>>>(f);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    nl++;
  }
  if (any_big) {
    EncodeArgs g = args;
    g.scratch = L.scratch;
    g.counter = L.counters + 4 * nwaves;
    if (!only_big) { g.list = list_big; g.list_count = big_count; }
    const size_t per = encode_cta_scratch_bytes(g.prm);
    const int grid = (int)min(min((size_t)args.n, (size_t)L.sm_count * 4), L.scratch_bytes / per);
    if (grid < 1) return cudaErrorInvalidValue;
    e = launch_encode_batch_nomemset(g, grid, stream);
    if (e != cudaSuccess) return e;
    nl++;
  }
  if (launches) *launches = nl;
  return cudaSuccess;
}

} // namespace b200zstd
