// zstd_encode_esd.cu -- levels 1-4 of the batched Zstandard compressor for sm_100a.
//
// Replaces, for the batch path at these levels, ZstdBatchManager::compress_batch -> DefaultZstdManager::compress
// (src/cuda_zstd_manager.cu:5715-5797, 1536-3112 in the reference): find_matches_kernel + greedy_parse_kernel +
// build_sequences_gpu_kernel (src/lz77_parallel.cu:26-268), compress_literals / compress_sequences
// (manager.cu:4406-4484, 4864-4974), write_frame_header / write_block (:3998-4106, 4227-4286).
//
// Two kernels per wave of blocks.
//
// PARSE (zstd_encode_esd_kernel): one CTA per <= 128 KB block, the block resident in shared memory, specialised warps
// running as a pipeline:
//   load     the block arrives HBM -> shared memory by cp.async.bulk (TMA) copies completing on an mbarrier
//   H warp   walks the block in fixed windows of 32 positions: hashes, reads the candidate position(s) from the
//            shared-memory table(s), resolves equal hashes inside the window (match_any) and inserts -- the table
//            state never depends on the parse, so this warp only waits for ring room
//   V warps  take windows in turn: compare each position with its candidate(s) and measure the match up to 32 bytes
//   S warps  the serial greedy parse, one SUB-SEGMENT of 1024 positions per warp at a time: 32 positions from the
//            parse position per step, repeat-offset matches by a byte compare + ballot (exact lengths as a bit mask),
//            measured table matches from the V warps, cooperative extension of the rare longer match, backward
//            extension, repeat-offset coding.  Sub-segments are independent (a match ends at the sub-segment's end, the
//            repeat-offset history is "unknown" at its start, which only ever costs a full offset code), so several S
//            warps work on one block; all of them read candidates that H / V computed once, in block order.
//   output   per sub-segment a list of {literal run, match length, offset code} in the wave's scratch in HBM
// FINISH (zstd_encode_finish_kernel): one warp per block, many resident per SM: joins the lists (a sub-segment's
// trailing literals go to the next sequence), gathers the literals from the input, runs the entropy stage
// (zstd_encode_entropy.cuh) and writes the frame.
//
// A lone warp issues a dependent instruction every 5-10 cycles, so the per-sequence chain is kept short (no literal
// copies) and four of them run per block; everything that can be computed per position is computed by the other
// warps.  tests/model/enc_model.cpp (parse_block_esd) restates the stages on the host; the bytes must agree.
#include "zstd_common.cuh"
#include "zstd_device_api.h"
#include "zstd_encode_core.cuh"
#include "zstd_encode_entropy.cuh"

#include <cstdlib>

namespace b200zstd {

using namespace enc;

namespace {

constexpr int ESD_NV = 8;                       // verify warps (power of two)
constexpr int ESD_NS = 4;                       // select warps
constexpr int ESD_NSLOT = 5;                    // ring capacity in sub-segments: NS being parsed + one being filled
// Warp roles by warp id.  The SM's schedulers favour the highest warp id among the eligible warps, so the serial chains
// sit on top and the warps with slack at the bottom.
constexpr int ESD_W_V0 = 0;
constexpr int ESD_W_S0 = ESD_NV;
constexpr int ESD_W_H = ESD_NV + ESD_NS;        // every other warp waits for these two (one per table): top priority
constexpr int ESD_WARPS = ESD_NV + ESD_NS + 2;
constexpr int ESD_THREADS = 32 * ESD_WARPS;
constexpr uint32_t ESD_SUB_WINDOWS = ESD_SUB / 32;
constexpr uint32_t ESD_SUB_SEQ = ESD_SUB / 4;   // every sequence covers >= 4 positions of its sub-segment
constexpr uint32_t ESD_RING_POS = ESD_NSLOT * ESD_SUB;
constexpr uint32_t ESD_IN_PAD = 64;             // 16 bytes of alignment slack in front, read-ahead room behind
constexpr uint32_t ESD_KIND_PARSED = 0, ESD_KIND_RLE = 1, ESD_KIND_SKIP = 2;

// ---- optional cycle accounting per warp role (build with -DESD_PROF; tools/esd_prof.py reads it) ----
#ifdef ESD_PROF
__device__ unsigned long long g_esd_prof[16];
#define PROF_T0(v) const long long v = clock64()
#define PROF_ADD(slot, v) do { if (lane == 0) atomicAdd(&g_esd_prof[slot], (unsigned long long)(v)); } while (0)
#define PROF_SINCE(slot, v) PROF_ADD(slot, clock64() - (v))
#else
#define PROF_T0(v) do { } while (0)
#define PROF_ADD(slot, v) do { } while (0)
#define PROF_SINCE(slot, v) do { } while (0)
#endif
// slots: 0 S busy (summed over sub-segments), 1 S waits for V, 2 block total (CTA), 3 finish busy, 4 H total, 5 H waits for ring room,
//        6 V total, 7 V waits for H, 8 blocks, 9 load wait, 10 S steps, 11 sequences, 13 S open extensions

// per-block record in the wave's scratch: header, then the sub-segment lists
struct EsdBlockHdr { uint32_t kind, nsub, pad0, pad1; };
// scratch slot of one block: [EsdBlockHdr | nsub_max x {uint16 count, uint16 tail} | nsub_max x ESD_SUB_SEQ x uint2]
__host__ __device__ constexpr uint32_t esd_nsub_max(uint32_t block_max) { return block_max / ESD_SUB; }
__host__ __device__ constexpr size_t esd_slot_bytes(uint32_t block_max) {
  return 16 + (size_t)esd_nsub_max(block_max) * 4 + (size_t)esd_nsub_max(block_max) * ESD_SUB_SEQ * 8;
}

struct EsdCtl {
  unsigned long long mbar;
  volatile uint32_t h_done[2];              // windows hashed, per table
  uint32_t item;                            // broadcast of the work-queue draw
  uint32_t pad;
  volatile uint32_t v_done[ESD_NV];         // per V warp: the next window it will publish
  volatile uint32_t s_done[ESD_NS];         // per S warp: sub-segments finished
  uint32_t flag;
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
__device__ __forceinline__ void cc_barrier() { asm volatile("" ::: "memory"); }

// 8 bytes at byte offset `at` of the staged block (three aligned shared-memory words and two funnel shifts)
__device__ __forceinline__ uint64_t lds64(const uint8_t *base16, uint32_t at) {
  const uint32_t *w = reinterpret_cast<const uint32_t *>(base16) + (at >> 2);
  const uint32_t sh = (at & 3) * 8;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
  return ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
}
__device__ __forceinline__ uint32_t common8(uint64_t a, uint64_t b) {
  const uint32_t xl = (uint32_t)a ^ (uint32_t)b, xh = (uint32_t)(a >> 32) ^ (uint32_t)(b >> 32);
  if (xl) return (uint32_t)(__ffs((int)xl) - 1) >> 3;
  if (xh) return 4u + ((uint32_t)(__ffs((int)xh) - 1) >> 3);
  return 8u;
}
__device__ __forceinline__ uint32_t ones_from(uint32_t m, uint32_t j) {          // length of the run of 1 bits of m starting at bit j
  const uint32_t z = ~(m >> j);
  return z ? (uint32_t)(__ffs((int)z) - 1) : 32u;
}

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
  uint32_t dbg;                  // ESD_PROF builds: 1 = V warps publish without working, 2 = S warps skip their sub-segments
};

template <int DFAST, int BIG>
__global__ void __launch_bounds__(ESD_THREADS, BIG ? 1 : 2) zstd_encode_esd_kernel(EsdArgs K) {
  constexpr uint32_t block_max = BIG ? 131072u : 65536u;
  extern __shared__ __align__(128) uint8_t smem[];
  EsdCtl *const ctl = reinterpret_cast<EsdCtl *>(smem);
  uint8_t *const in_base = smem + 128;                                         // 16-byte aligned
  uint16_t *const tab1 = reinterpret_cast<uint16_t *>(in_base + block_max + ESD_IN_PAD);
  uint16_t *const tab2 = tab1 + ((size_t)1 << K.E.hash_log);
  uint32_t *const ring = reinterpret_cast<uint32_t *>(tab2 + (DFAST ? ((size_t)1 << K.E.long_log) : 0));

  // (the warp id is broadcast from lane 0 so that the compiler treats the role branches as warp-uniform)
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const EncodeArgs &A = K.A;
  const bool blocks_only = A.block_mode != 0;

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
    PROF_T0(tb0);
    uint8_t *const slot = K.slots + (size_t)(item - K.wave_base) * K.slot_bytes;
    EsdBlockHdr *const hdr = reinterpret_cast<EsdBlockHdr *>(slot);
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
          hdr->kind = ESD_KIND_SKIP;
        } else if (n > K.slot_block_max) {
          K.big_list[atomicAdd(K.big_count, 1u)] = item;            // multi-block (or, for 64 KB slots, > 64 KB) item: the general kernel takes it
          hdr->kind = ESD_KIND_SKIP;
        } else K.defer_list[atomicAdd(K.defer_count, 1u)] = item;   // the 128 KB geometry writes this slot later
      }
      __syncthreads();
      continue;
    }
    const uint32_t bn = (uint32_t)n;
    const uint32_t delta = (uint32_t)((uintptr_t)chunk & 15);
    const uint8_t *const in = in_base + delta;                              // block byte i lives at in[i]
    // ---- load: HBM -> shared memory by bulk async copies; tables are cleared while the copy is in flight ----
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t total = (delta + bn + 15u) & ~15u;
      mbar_arrive_tx(&ctl->mbar, total);
      const uint8_t *src = chunk - delta;
      for (uint32_t o = 0; o < total; o += 16384u) bulk_g2s(in_base + o, src + o, min(16384u, total - o), &ctl->mbar);
      ctl->h_done[0] = 0; ctl->h_done[1] = 0;
      for (int j = 0; j < ESD_NV; j++) ctl->v_done[j] = (uint32_t)j;
      for (int j = 0; j < ESD_NS; j++) ctl->s_done[j] = 0;
    }
    {
      uint4 *z = reinterpret_cast<uint4 *>(tab1);
      const uint32_t vecs = (uint32_t)((((size_t)2 << K.E.hash_log) + (DFAST ? ((size_t)2 << K.E.long_log) : 0)) >> 4);
      for (uint32_t i = tid; i < vecs; i += ESD_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    PROF_T0(tl0);
    mbar_wait(&ctl->mbar, phase);
    phase ^= 1;
    if (warp == 0) PROF_SINCE(9, tl0);
    // ---- RLE block? (decided on the first 32 bytes in the common case) ----
    bool rle = false;
    {
      const uint8_t b0 = in[0];
      bool same = true;
      if ((uint32_t)tid < min(bn, 32u)) same = in[tid] == b0;
      if (__syncthreads_and(same)) {
        for (uint32_t i = tid; i < bn && same; i += ESD_THREADS) same = in[i] == b0;
        rle = __syncthreads_and(same) && bn > 1;
      }
    }
    __syncthreads();                                                       // tables cleared, control words set
    const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
    const uint32_t nwin = rle ? 0 : (ilimit + 31) >> 5;
    const uint32_t nsub = rle ? 0 : (ilimit + ESD_SUB - 1) / ESD_SUB;
    if (tid == 0) { hdr->kind = rle ? ESD_KIND_RLE : ESD_KIND_PARSED; hdr->nsub = nsub; }
    uint16_t *const sub_hdr = reinterpret_cast<uint16_t *>(slot + 16);
    uint2 *const lists = reinterpret_cast<uint2 *>(slot + 16 + (size_t)esd_nsub_max(K.slot_block_max) * 4);

    if (warp >= ESD_W_H) {
      // ============================ H warps: one per table (the second only for DFAST) ============================
      // Sequential semantics inside a window ("nearest lower inserting lane with my hash", "highest inserting lane
      // writes") need to know which lanes share a hash.  MATCH.ANY answers that but takes ~300 cycles on this part, so:
      // the inserting lanes store, every lane reads its bucket back, and a lane that does not find what it expects (its
      // own position if it inserted, the old entry if not) has met another lane of the window in its bucket.  Each
      // such bucket is then settled by one broadcast of its hash and a ballot.
      const int T = warp - ESD_W_H;                                       // 0: primary table, 1: long table
      if (T == 0 || DFAST) {
        uint16_t *const tab = T ? tab2 : tab1;
        uint16_t *const ring16 = reinterpret_cast<uint16_t *>(ring) + T;
        uint32_t prev = 0xFFFFFFFFu;
        const uint32_t lt = lanemask_lt();
        PROF_T0(th0);
        uint32_t slot_base = 0;
        uint64_t v = (uint32_t)lane < ilimit ? lds64(in_base, delta + (uint32_t)lane) : 0ull;
        for (uint32_t k = 0; k < nwin; k++) {
          if ((k & (ESD_SUB_WINDOWS - 1)) == 0) {
            const uint32_t j = k / ESD_SUB_WINDOWS;
            slot_base = (j % ESD_NSLOT) * ESD_SUB;
            if (j >= ESD_NSLOT) {                           // ring room: the sub-segment that held this slot must be parsed
              const uint32_t o = j - ESD_NSLOT;
              PROF_T0(th1);
              while (ctl->s_done[o % ESD_NS] <= o / ESD_NS) __nanosleep(64);
              if (T == 0) PROF_SINCE(5, th1);
            }
            cc_barrier();
          }
          const uint32_t p0 = k << 5, p = p0 + (uint32_t)lane;
          const bool act = p < ilimit;
          const uint32_t h = T ? hash_long(v, K.E.long_log) : hash_short(v, K.E.hash_bytes, K.E.hash_log);
          v = p + 32 < ilimit ? lds64(in_base, delta + p + 32) : 0ull;     // next window's bytes, in flight during the table work
          uint32_t e = act ? tab[h] : 0u;
          uint32_t hp = __shfl_up_sync(0xffffffffu, h, 1);
          if (lane == 0) hp = prev;
          const bool ins = act && hp != h;
          prev = __shfl_sync(0xffffffffu, h, 31);
          __syncwarp();                                     // every lookup of the window precedes its inserts
          if (ins) tab[h] = (uint16_t)p;
          __syncwarp();
          uint32_t cm = __ballot_sync(0xffffffffu, act && tab[h] != (ins ? (p & 0xFFFFu) : e));
          if (cm) {
            if (T == 0) PROF_ADD(12, 1);
            const uint32_t insmask = __ballot_sync(0xffffffffu, ins);
            do {
              const uint32_t ht = __shfl_sync(0xffffffffu, h, __ffs((int)cm) - 1);
              const bool mine = act && h == ht;
              const uint32_t g = __ballot_sync(0xffffffffu, mine), gi = g & insmask;
              if (mine) {
                const uint32_t lower = gi & lt;
                if (lower) e = (p0 + (uint32_t)(31 - __clz(lower))) & 0xFFFFu;
                if (ins && (gi >> lane) == 1u) tab[h] = (uint16_t)p;
              }
              cm &= ~g;
            } while (cm);
          }
          ring16[2 * (slot_base + (p & (ESD_SUB - 1)))] = (uint16_t)e;
          __syncwarp();
          cc_barrier();
          if (lane == 0) ctl->h_done[T] = k + 1;
        }
        if (T == 0) PROF_SINCE(4, th0);
      }
    } else if (warp < ESD_W_S0) {
      // ======================================= V warps =======================================
      const int jv = warp - ESD_W_V0;
      PROF_T0(tv0);
      for (uint32_t k = (uint32_t)jv; k < nwin; k += ESD_NV) {
        PROF_T0(tv1);
        while (ctl->h_done[0] <= k || (DFAST && ctl->h_done[1] <= k)) __nanosleep(32);
        PROF_SINCE(7, tv1);
        cc_barrier();
        const uint32_t p = (k << 5) + (uint32_t)lane;
        const uint32_t ri = ((k / ESD_SUB_WINDOWS) % ESD_NSLOT) * ESD_SUB + (p & (ESD_SUB - 1));
        uint32_t res = 0;
        if (p < ilimit && !(K.dbg & 1)) {
          const uint32_t e = ring[ri];
          const uint64_t v = lds64(in_base, delta + p);
          int32_t c1 = (int32_t)((BIG ? (p & ~0xFFFFu) : 0u) | (e & 0xFFFFu));
          if (c1 >= (int32_t)p) c1 -= 0x10000;
          uint32_t off = 0, len = 0;
          if (DFAST) {
            int32_t c2 = (int32_t)((BIG ? (p & ~0xFFFFu) : 0u) | (e >> 16));
            if (c2 >= (int32_t)p) c2 -= 0x10000;
            if (c2 >= 0 && lds64(in_base, delta + (uint32_t)c2) == v) { off = p - (uint32_t)c2; len = 8; }
          }
          if (len == 0 && c1 >= 0) {
            const uint32_t c = common8(v, lds64(in_base, delta + (uint32_t)c1));
            if (c >= ESD_MIN_MATCH) { off = p - (uint32_t)c1; len = c; }
          }
          if (len == 8) {
            while (len < ESD_LCAP && p + len < bn) {
              uint32_t c = common8(lds64(in_base, delta + p + len), lds64(in_base, delta + p + len - off));
              const uint32_t room = bn - (p + len);
              if (c > room) c = room;
              len += c;
              if (c < 8) break;
            }
            if (len > ESD_LCAP) len = ESD_LCAP;
          }
          res = off | (len << 17);
        }
        ring[ri] = res;
        __syncwarp();
        cc_barrier();
        if (lane == 0) ctl->v_done[jv] = k + ESD_NV;
      }
      PROF_SINCE(6, tv0);
    } else {
      // ======================================= S warps =======================================
      const int ws = warp - ESD_W_S0;
      for (uint32_t j = (uint32_t)ws; j < nsub; j += ESD_NS) {
        PROF_T0(ts0);
        long long waited = 0;
        const uint32_t B = j * ESD_SUB, E = min(B + ESD_SUB, bn), lim = min(E, ilimit);
        const uint32_t slot_base = (j % ESD_NSLOT) * ESD_SUB;
        uint2 *const list = lists + (size_t)j * ESD_SUB_SEQ;
        // a sub-segment parsed on its own does not know the repeat offsets the decoder will hold when it gets there:
        // 0 = unknown, never matched against and never equal to a real offset, so its first sequences simply carry full
        // offset codes (RFC 8878 3.1.1.5 lets any offset be written that way); the block's first sub-segment starts
        // from the frame's initial history unless the block is itself encoded on its own (block mode)
        uint32_t r0 = 0, r1 = 0, r2 = 0;
        if (j == 0 && !(blocks_only && item != 0)) { r0 = 1; r1 = 4; r2 = 8; }
        uint32_t ip = B, anchor = B, rep0 = r0, nseq = 0, ready_end = B;
        while (ip < lim && !(K.dbg & 2)) {
          const uint32_t need = min(ip + 32, lim);
          if (ready_end < need) {
            PROF_T0(ts2);
            while (ready_end < need) {
              const uint32_t kk = ready_end >> 5;
              while (ctl->v_done[kk & (ESD_NV - 1)] <= kk) __nanosleep(20);
              ready_end += 32;
            }
            cc_barrier();
#ifdef ESD_PROF
            waited += clock64() - ts2;
#endif
          }
          PROF_ADD(10, 1);
          const uint32_t p = ip + (uint32_t)lane;
          const bool inb = p < lim && p + 4 <= E;
          const uint32_t r = inb ? ring[slot_base + (p & (ESD_SUB - 1))] : 0u;
          const bool e = rep0 != 0 && p >= rep0 && p < E && in[p] == in[p - rep0];
          const uint32_t eq = __ballot_sync(0xffffffffu, e);
          const uint32_t ok = __ballot_sync(0xffffffffu, r != 0);
          const uint32_t rp = eq & (eq >> 1) & (eq >> 2) & (eq >> 3) & __ballot_sync(0xffffffffu, inb);
          const uint32_t cand = ok | rp;
          if (cand == 0) { ip += 32; continue; }
          uint32_t f = (uint32_t)__ffs((int)cand) - 1;
          const uint32_t rlen = r >> 17;
          const uint32_t len_f = __shfl_sync(0xffffffffu, rlen, f), len_g = __shfl_sync(0xffffffffu, rlen, (f + 1) & 31);
          bool use_rep = false;
          if ((rp >> f) & 1) {
            const uint32_t rl = ones_from(eq, f);
            if (!((ok >> f) & 1) || f + rl == 32 || rl + ESD_REP_BONUS >= len_f) use_rep = true;
          } else if (f + 1 < 32 && ((rp >> (f + 1)) & 1)) {
            const uint32_t rl = ones_from(eq, f + 1);
            if (f + 1 + rl == 32 || rl + ESD_REP_BONUS >= len_f) { f = f + 1; use_rep = true; }
          }
          uint32_t off, len;
          bool open;
          if (use_rep) { len = ones_from(eq, f); off = rep0; open = f + len == 32; }
          else {
            if (K.E.lazy && f + 1 < 32 && ((ok >> (f + 1)) & 1) && len_g > len_f) f = f + 1;
            const uint32_t rf = __shfl_sync(0xffffffffu, r, f);
            off = rf & 0x1FFFFu; len = rf >> 17; open = len == ESD_LCAP;
          }
          uint32_t s = ip + f;
          if (open) {
            PROF_ADD(13, 1);
            for (;;) {
              const uint32_t q = s + len + (uint32_t)lane;
              const uint32_t m = __ballot_sync(0xffffffffu, q < E && in[q] == in[q - off]);
              const uint32_t nn = ones_from(m, 0);
              len += nn;
              if (nn < 32) break;
            }
          }
          if (s + len > E) len = E - s;
          {
            const uint32_t jb = (uint32_t)lane;
            const bool mb = jb < s - anchor && s - 1 - jb >= off && in[s - 1 - jb] == in[s - 1 - jb - off];
            const uint32_t nb = ones_from(__ballot_sync(0xffffffffu, mb), 0);
            s -= nb; len += nb;
          }
          const uint32_t llen = s - anchor;
          // repeat-offset code and history update (enc::offset_to_code, RFC 8878 3.1.2.5) without branches
          uint32_t code;
          {
            const bool l0 = llen == 0;
            const uint32_t ca = l0 ? r1 : r0, cb = l0 ? r2 : r1, cc = l0 ? (r0 > 1 ? r0 - 1 : 0u) : r2;
            code = off == ca ? 1u : off == cb ? 2u : off == cc ? 3u : off + 3u;
            const bool same = !l0 && off == r0;
            const uint32_t n2 = same ? r2 : (off == r1 ? r2 : r1), n1 = same ? r1 : r0;
            r2 = n2; r1 = n1; r0 = off;
          }
          if (lane == 0) list[nseq] = make_uint2(llen | (len << 12), code);
          nseq++;
          ip = anchor = s + len; rep0 = off;
        }
        if (lane == 0) { sub_hdr[2 * j] = (uint16_t)nseq; sub_hdr[2 * j + 1] = (uint16_t)(E - anchor); }
        __syncwarp();
        cc_barrier();
        if (lane == 0) ctl->s_done[ws] = j / ESD_NS + 1;
#ifdef ESD_PROF
        PROF_ADD(0, clock64() - ts0 - waited);
        PROF_ADD(1, waited);
        PROF_ADD(11, nseq);
#endif
      }
    }
    __syncthreads();                              // every warp is done with the staged block and the tables
    if (warp == 0) { PROF_SINCE(2, tb0); PROF_ADD(8, 1); }
  }
}

// -----------------------------------------------------------------------------------------------------------------
// FINISH: one warp per block.  Joins the sub-segment lists into the three sequence arrays, gathers the literals from
// the input, codes the block and writes the frame (or the bare block in block mode).
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

__global__ void __launch_bounds__(32 * FIN_WARPS) zstd_encode_finish_kernel(FinArgs F) {
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
  const uint32_t nsub_max = esd_nsub_max(F.slot_block_max);
  for (;;) {
    uint32_t idx = 0;
    if (lane == 0) idx = atomicAdd(F.work_head, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= F.wave_n) break;
    const uint32_t item = F.wave_base + idx;
    const uint8_t *const slot = F.slots + (size_t)idx * F.slot_bytes;
    const EsdBlockHdr H = *reinterpret_cast<const EsdBlockHdr *>(slot);
    if (H.kind == ESD_KIND_SKIP) continue;
    PROF_T0(tf0);
    const uint8_t *const chunk = (const uint8_t *)A.in_ptrs[item];
    const uint32_t bn = (uint32_t)A.in_sizes[item];
    uint8_t *const dst = (uint8_t *)A.out_ptrs[item];
    const bool last = blocks_only ? item + 1 == A.n : true;
    size_t op = 0;
    if (!blocks_only) {
      if (lane == 0) op = write_frame_header(dst, bn, P.checksum != 0);
      op = __shfl_sync(0xffffffffu, (unsigned long long)op, 0);
    }
    if (H.kind == ESD_KIND_RLE) {
      if (lane == 0) { write_block_header(dst + op, last, 1, bn); dst[op + 3] = chunk[0]; }
      op += 4;
    } else {
      // ---- join the lists; every sequence learns where its literal run starts in the block and in the literal buffer ----
      const uint16_t *const sub_hdr = reinterpret_cast<const uint16_t *>(slot + 16);
      const uint2 *const lists = reinterpret_cast<const uint2 *>(slot + 16 + (size_t)nsub_max * 4);
      uint32_t nseq = 0, nlit = 0, pos = 0, carry = 0;          // pos: block position where the pending literal run starts
      for (uint32_t j = 0; j < H.nsub; j++) {
        const uint32_t cnt = sub_hdr[2 * j], tail = sub_hdr[2 * j + 1];
        const uint2 *const list = lists + (size_t)j * ESD_SUB_SEQ;
        for (uint32_t b = 0; b < cnt; b += 32) {
          const uint32_t k = b + (uint32_t)lane;
          uint32_t ll = 0, ml = 0, code = 0;
          if (k < cnt) {
            const uint2 q = list[k];
            ll = q.x & 0xFFFu; ml = q.x >> 12; code = q.y;
            if (k == 0) ll += carry;
          }
          // inclusive prefix sums of ll and ll + ml over the 32 sequences of this round
          uint32_t sl = ll, sb = ll + ml;
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tl = __shfl_up_sync(0xffffffffu, sl, o), tb = __shfl_up_sync(0xffffffffu, sb, o);
            if (lane >= o) { sl += tl; sb += tb; }
          }
          const uint32_t my_lit = nlit + sl - ll, my_pos = pos + sb - (ll + ml);
          if (k < cnt) {
            s_ll[nseq + lane] = ll; s_ml[nseq + lane] = ml; s_of[nseq + lane] = code;
            for (uint32_t i = 0; i < ll; i++) lits[my_lit + i] = chunk[my_pos + i];
          }
          nlit += __shfl_sync(0xffffffffu, sl, 31);
          pos += __shfl_sync(0xffffffffu, sb, 31);
          nseq += min(32u, cnt - b);
        }
        carry = (cnt ? 0u : carry) + tail;
      }
      // literals after the last sequence (the whole block when nothing was parsed)
      for (uint32_t i = pos + lane; i < bn; i += 32) lits[nlit + i - pos] = chunk[i];
      nlit += bn - pos;
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
    PROF_SINCE(3, tf0);
  }
}

size_t esd_smem_bytes(const EsdParams &e, int big) {
  const size_t in = (big ? 131072u : 65536u) + ESD_IN_PAD;
  const size_t tabs = ((size_t)2 << e.hash_log) + (e.dfast ? ((size_t)2 << e.long_log) : 0);
  return 128 + in + tabs + (size_t)ESD_RING_POS * 4 + 16;
}

template <int DFAST, int BIG> cudaError_t esd_launch_one(const EsdArgs &k, int grid, cudaStream_t stream) {
  const size_t smem = esd_smem_bytes(k.E, BIG);
  cudaError_t e = cudaFuncSetAttribute(zstd_encode_esd_kernel<DFAST, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  zstd_encode_esd_kernel<DFAST, BIG><<<grid, ESD_THREADS, smem, stream>>>(k);
  return cudaGetLastError();
}

constexpr int FIN_CTAS_PER_SM = 6;               // 24 finish warps per SM

// blocks per wave: bounds the list scratch (2 bytes per input byte of the wave)
uint32_t esd_wave(size_t n, uint32_t block_max) { return (uint32_t)min(n, (size_t)(block_max > 65536 ? 8192 : 16384)); }
size_t esd_slot_stride(uint32_t bm) { return (esd_slot_bytes(bm) + 255) & ~(size_t)255; }

} // namespace

#ifdef ESD_PROF
extern "C" void cuda_zstd_b200_esd_prof(unsigned long long *out16, int reset) {
  cudaMemcpyFromSymbol(out16, g_esd_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_esd_prof, z, sizeof z); }
}
#endif

size_t esd_scratch_bytes(size_t n_items, uint32_t bm, int sm_count) {
  const size_t lists = (size_t)esd_wave(n_items, bm) * esd_slot_stride(bm);
  const size_t fin_warps = min((size_t)sm_count * FIN_CTAS_PER_SM * FIN_WARPS, ((n_items + FIN_WARPS - 1) / FIN_WARPS) * FIN_WARPS);
  return lists + fin_warps * fin_warp_scratch(bm);
}
size_t esd_counter_words(size_t n_items, uint32_t bm) {
  const uint32_t wave = esd_wave(n_items, bm);
  return 4 * ((n_items + wave - 1) / wave) + 2;
}

// Levels 1-4.  Per wave: the 64 KB-block parse kernel (two CTAs per SM), the 128 KB-block parse kernel (one CTA per SM)
// over what the first handed on, the finish kernel; after the last wave the general kernel (zstd_encode.cu) for
// multi-block items.  max_item_bytes != 0 (the host knows the sizes) skips the launches that cannot have work.
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
  const uint32_t slot_bytes = (uint32_t)esd_slot_stride(bm);
  const size_t lists_bytes = (size_t)wave * slot_bytes;
  if (L.scratch_bytes < esd_scratch_bytes(args.n, bm, L.sm_count)) return cudaErrorInvalidValue;
  const size_t nwaves = only_big ? 0 : (args.n + wave - 1) / wave;
  // counters: per wave {parse64 head, parse128 head, finish head, defer count}; then general-kernel head and big count
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
    k.E = esd_params_for_level(args.prm.level);
    k.slots = L.scratch; k.slot_bytes = slot_bytes; k.slot_block_max = bm;
    k.wave_base = (uint32_t)(w * wave); k.wave_n = (uint32_t)min((size_t)wave, args.n - w * wave);
    k.big_list = list_big; k.big_count = big_count;
#ifdef ESD_PROF
    if (const char *d = getenv("ESD_DBG")) k.dbg = (uint32_t)atoi(d);
#endif
    if (any64) {
      k.list = nullptr; k.list_count = nullptr;
      k.work_head = c + 0;
      k.defer_list = list128 + k.wave_base; k.defer_count = c + 3;
      const int grid = (int)min((size_t)k.wave_n, (size_t)L.sm_count * 2);
      e = k.E.dfast ? esd_launch_one<1, 0>(k, grid, stream) : esd_launch_one<0, 0>(k, grid, stream);
      if (e != cudaSuccess) return e;
      nl++;
    }
    if (any128) {
      if (any64) { k.list = list128 + k.wave_base; k.list_count = c + 3; } else { k.list = nullptr; k.list_count = nullptr; }
      k.work_head = c + 1;
      k.defer_list = nullptr; k.defer_count = nullptr;
      const int grid = (int)min((size_t)k.wave_n, (size_t)L.sm_count);
      e = k.E.dfast ? esd_launch_one<1, 1>(k, grid, stream) : esd_launch_one<0, 1>(k, grid, stream);
      if (e != cudaSuccess) return e;
      nl++;
    }
    FinArgs f{};
    f.A = args;
    f.slots = L.scratch; f.slot_bytes = slot_bytes; f.slot_block_max = bm;
    f.wave_base = k.wave_base; f.wave_n = k.wave_n;
    f.work_head = c + 2;
    f.scratch = L.scratch + lists_bytes;
    const int fgrid = (int)min((size_t)L.sm_count * FIN_CTAS_PER_SM, ((size_t)k.wave_n + FIN_WARPS - 1) / FIN_WARPS);
    zstd_encode_finish_kernel<<<fgrid, 32 * FIN_WARPS, 0, stream>>>(f);
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
