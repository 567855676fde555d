// zstd_encode_esd.cu -- levels 1-4 of the batched Zstandard compressor for sm_100a: one CTA per <= 128 KB block,
// the block resident in shared memory, the work of a block split over specialised warps that run as a pipeline.
//
// Replaces, for the batch path at these levels, ZstdBatchManager::compress_batch -> DefaultZstdManager::compress
// (src/cuda_zstd_manager.cu:5715-5797, 1536-3112 in the reference): find_matches_kernel + greedy_parse_kernel +
// build_sequences_gpu_kernel (src/lz77_parallel.cu:26-268), compress_literals / compress_sequences
// (manager.cu:4406-4484, 4864-4974), write_frame_header / write_block (:3998-4106, 4227-4286).
//
//   load     the block arrives HBM -> shared memory by cp.async.bulk (TMA) copies completing on an mbarrier
//   H warp   walks the block in fixed windows of 32 positions: hashes, reads the candidate position(s) from the
//            shared-memory table(s), resolves equal hashes inside the window (match_any) and inserts -- the table
//            state never depends on the parse, so this warp runs ahead of everything else
//   V warps  take windows in turn: compare each position with its candidate(s) and measure the match up to 32 bytes
//   S warp   the serial greedy parse: 32 positions from the parse position per step, repeat-offset matches by a
//            byte-compare + ballot (exact lengths as a bit mask), measured table matches from the V warps, cooperative
//            extension of the rare longer match, backward extension, literals and sequences streamed to the CTA's
//            scratch in HBM
//   X warps  the entropy stage (zstd_encode_entropy.cuh) and frame assembly of the PREVIOUS block, while the next
//            one is being parsed
// The serial chain of a block (one step per sequence) is short and runs at shared-memory latency; everything that can
// be computed per position is computed by the other warps.  tests/model/enc_model.cpp (parse_block_esd) restates the
// three parse stages on the host; the bytes must agree.
#include "zstd_common.cuh"
#include "zstd_device_api.h"
#include "zstd_encode_core.cuh"
#include "zstd_encode_entropy.cuh"

namespace b200zstd {

using namespace enc;

namespace {

constexpr int ESD_NV = 3;                       // verify warps
constexpr int ESD_NX = 2;                       // entropy warps
constexpr int ESD_NBUF = 3;                     // scratch buffers per CTA: one being parsed into, two being coded
constexpr int ESD_PARSE_WARPS = 2 + ESD_NV;     // S, H, V...
constexpr int ESD_WARPS = ESD_PARSE_WARPS + ESD_NX;
constexpr int ESD_THREADS = 32 * ESD_WARPS;
constexpr int ESD_PARSE_THREADS = 32 * ESD_PARSE_WARPS;
constexpr uint32_t ESD_RING_WINDOWS = 32, ESD_RING_POS = 32 * ESD_RING_WINDOWS;
constexpr uint32_t ESD_PARSE_OVER = 0xFFFFFFE0u;
constexpr uint32_t ESD_IN_PAD = 64;             // 16 bytes of alignment slack in front, read-ahead room behind

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
// slots: 0 S total, 1 S waits for V, 2 S waits for a scratch buffer, 3 X busy, 4 H total, 5 H waits for ring room, 6 V total,
//        7 V waits for H, 8 blocks, 9 load wait, 10 S steps, 11 sequences, 12 X waits, 13 S open extensions

struct EsdRecord { uint32_t item, kind, nlit, nseq; };           // S -> X hand-over; kind 0 parsed, 1 RLE block
struct EsdCtl {
  unsigned long long mbar;
  volatile uint32_t h_done;                 // windows hashed
  volatile uint32_t s_pos;                  // parse position (ring slots below it are free)
  volatile uint32_t v_done[4];              // per V warp: the next window it will publish
  volatile uint32_t pub;                    // records published to the X warps
  volatile uint32_t stop;
  volatile uint32_t x_done[4];              // per X warp: the next record it will take
  uint32_t item;                            // broadcast of the work-queue draw
  uint32_t pad0;
  EsdRecord rec[ESD_NBUF];
};

template <int BIG> struct EsdGeom {
  static constexpr uint32_t block_max = BIG ? 131072u : 65536u;
  static constexpr uint32_t max_seq = block_max / 4 + 64;
  static constexpr size_t lits_bytes = block_max + 64;
  static constexpr size_t buf_bytes = lits_bytes + (size_t)3 * max_seq * 4;
  static constexpr size_t cta_scratch = ((size_t)ESD_NBUF * buf_bytes + 255) & ~(size_t)255;
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
// barrier over the parse warps only (the X warps run a block behind and never join it)
__device__ __forceinline__ void parse_bar() { asm volatile("bar.sync 1, %0;" ::"n"(ESD_PARSE_THREADS) : "memory"); }
__device__ __forceinline__ bool parse_bar_and(bool pred) {
  uint32_t r;
  asm volatile("{\n.reg .pred p, q;\nsetp.ne.u32 q, %1, 0;\nbar.red.and.pred p, 1, %2, q;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(r) : "r"((uint32_t)pred), "n"(ESD_PARSE_THREADS) : "memory");
  return r != 0;
}
__device__ __forceinline__ void cc_barrier() { asm volatile("" ::: "memory"); }

// 8 bytes at byte offset `at` of the staged block (three aligned shared-memory words and two funnel shifts)
__device__ __forceinline__ uint64_t lds64(const uint8_t *in, uint32_t at) {
  const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (at >> 2);
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
  const uint32_t *list;          // null: items are 0 .. A.n-1; else indices drawn from list[0 .. *list_count)
  const uint32_t *list_count;
  uint32_t *work_head;
  uint32_t *defer_list;          // items this geometry cannot take: blocks above block_max (null: none can occur) ...
  uint32_t *defer_count;
  uint32_t *big_list;            // ... and items above one block (multi-block frames go to the general kernel)
  uint32_t *big_count;
};

template <int DFAST, int BIG>
__global__ void __launch_bounds__(ESD_THREADS, BIG ? 1 : 2) zstd_encode_esd_kernel(EsdArgs K) {
  using G = EsdGeom<BIG>;
  extern __shared__ __align__(128) uint8_t smem[];
  EsdCtl *const ctl = reinterpret_cast<EsdCtl *>(smem);
  uint8_t *const in_base = smem + 128;                                         // 16-byte aligned
  uint16_t *const tab1 = reinterpret_cast<uint16_t *>(in_base + G::block_max + ESD_IN_PAD);
  uint16_t *const tab2 = tab1 + ((size_t)1 << K.E.hash_log);
  uint32_t *const ring = reinterpret_cast<uint32_t *>(tab2 + (DFAST ? ((size_t)1 << K.E.long_log) : 0));
  EntropyWs *const ews = reinterpret_cast<EntropyWs *>(reinterpret_cast<uint8_t *>(ring) + ESD_RING_POS * 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const EncodeArgs &A = K.A;
  const EncodeParams P = A.prm;
  const bool blocks_only = A.block_mode != 0;
  uint8_t *const scratch = A.scratch + (size_t)blockIdx.x * G::cta_scratch;

  if (tid == 0) {
    mbar_init(&ctl->mbar, 1);
    ctl->pub = 0; ctl->stop = 0;
    for (int x = 0; x < ESD_NX; x++) ctl->x_done[x] = (uint32_t)x;
  }
  __syncthreads();

  if (warp >= ESD_PARSE_WARPS) {
    // =============================== X warps: entropy stage + frame assembly ===============================
    const int x = warp - ESD_PARSE_WARPS;
    EntropyWs &W = ews[x];
    for (uint32_t k = (uint32_t)x;; k += ESD_NX) {
      bool got = false;
      PROF_T0(tx0);
      for (;;) {
        if (ctl->pub > k) { got = true; break; }
        if (ctl->stop) { got = ctl->pub > k; break; }
        __nanosleep(200);
      }
      if (!got) break;
      PROF_SINCE(12, tx0);
      PROF_T0(tx1);
      __threadfence_block();
      const EsdRecord R = ctl->rec[k % ESD_NBUF];
      uint8_t *const buf = scratch + (size_t)(k % ESD_NBUF) * G::buf_bytes;
      uint8_t *const lits = buf;
      uint32_t *const s_ll = reinterpret_cast<uint32_t *>(buf + G::lits_bytes);
      uint32_t *const s_ml = s_ll + G::max_seq;
      uint32_t *const s_of = s_ml + G::max_seq;
      const uint8_t *const chunk = (const uint8_t *)A.in_ptrs[R.item];
      const uint32_t bn = (uint32_t)A.in_sizes[R.item];
      uint8_t *const dst = (uint8_t *)A.out_ptrs[R.item];
      const bool last = blocks_only ? R.item + 1 == A.n : true;
      size_t op = 0;
      if (!blocks_only) {
        if (lane == 0) op = write_frame_header(dst, bn, P.checksum != 0);
        op = __shfl_sync(0xffffffffu, (unsigned long long)op, 0);
      }
      if (R.kind == 1) {
        if (lane == 0) { write_block_header(dst + op, last, 1, bn); dst[op + 3] = chunk[0]; }
        op += 4;
      } else {
        const uint32_t payload = entropy_stage_warp(W, lits, R.nlit, s_ll, s_ml, s_of, R.nseq, dst + op + 3, bn - 1, lane);
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
        A.out_sizes[R.item] = op;
        if (A.statuses) A.statuses[R.item] = ST_OK;
      }
      __syncwarp();
      __threadfence_block();
      if (lane == 0) ctl->x_done[x] = k + ESD_NX;
      PROF_SINCE(3, tx1);
    }
    return;
  }

  // ======================================= parse warps (S, H, V) =======================================
  uint32_t phase = 0, nrec = 0;                  // mbarrier parity; records published by this CTA
  for (;;) {
    if (tid == 0) {
      const uint32_t idx = atomicAdd(K.work_head, 1u);
      uint32_t item = 0xFFFFFFFFu;
      if (K.list) { if (idx < *K.list_count) item = K.list[idx]; }
      else if (idx < A.n) item = idx;
      ctl->item = item;
    }
    parse_bar();
    const uint32_t item = ctl->item;
    if (item == 0xFFFFFFFFu) break;
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
      if (cap < (blocks_only ? 0 : (size_t)frame_header_size(n) + (P.checksum ? 4 : 0)) + n + 3 * nblocks) status = ST_BUFFER_TOO_SMALL;
    }
    if (status != ST_OK) {
      if (tid == 0) { A.out_sizes[item] = 0; if (A.statuses) A.statuses[item] = status; }
      parse_bar();
      continue;
    }
    if (n > G::block_max) {
      // not this geometry's: hand the item to the kernel that takes it
      if (tid == 0) {
        if (n > BLOCK_BYTES) K.big_list[atomicAdd(K.big_count, 1u)] = item;
        else K.defer_list[atomicAdd(K.defer_count, 1u)] = item;
      }
      parse_bar();
      continue;
    }
    const uint32_t bn = (uint32_t)n;
    const uint32_t delta = (uint32_t)((uintptr_t)chunk & 15);
    const uint8_t *const in = in_base + delta;                              // block byte i lives at in[i]
    // ---- load: HBM -> shared memory by bulk async copies ----
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t total = (delta + bn + 15u) & ~15u;
      mbar_arrive_tx(&ctl->mbar, total);
      const uint8_t *src = chunk - delta;
      for (uint32_t o = 0; o < total; o += 16384u) bulk_g2s(in_base + o, src + o, min(16384u, total - o), &ctl->mbar);
    }
    // tables are cleared while the copy is in flight
    {
      uint4 *z = reinterpret_cast<uint4 *>(tab1);
      const uint32_t vecs = (uint32_t)((((size_t)2 << K.E.hash_log) + (DFAST ? ((size_t)2 << K.E.long_log) : 0)) >> 4);
      for (uint32_t i = tid; i < vecs; i += ESD_PARSE_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
      ctl->h_done = 0; ctl->s_pos = 0;
      for (int j = 0; j < ESD_NV; j++) ctl->v_done[j] = (uint32_t)j;
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
      // (cheap uniform decision first: almost every block differs inside its first 32 bytes)
      if (parse_bar_and(same)) {
        for (uint32_t i = tid; i < bn && same; i += ESD_PARSE_THREADS) same = in[i] == b0;
        rle = parse_bar_and(same) && bn > 1;
      }
    }
    parse_bar();                                                           // tables cleared, control words set
    const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
    const uint32_t nwin = (ilimit + 31) >> 5;

    if (warp == 1 && !rle) {
      // ======================================= H warp =======================================
      uint32_t prev1 = 0xFFFFFFFFu, prev2 = 0xFFFFFFFFu;
      const uint32_t lt = lanemask_lt();
      PROF_T0(th0);
      for (uint32_t k = 0; k < nwin; k++) {
        PROF_T0(th1);
        while (k >= (ctl->s_pos >> 5) + ESD_RING_WINDOWS) __nanosleep(100);
        PROF_SINCE(5, th1);
        if (ctl->s_pos == ESD_PARSE_OVER) break;
        cc_barrier();
        const uint32_t p0 = k << 5, p = p0 + (uint32_t)lane;
        const bool act = p < ilimit;
        uint32_t h1 = 0, h2 = 0, e1 = 0, e2 = 0;
        if (act) {
          const uint64_t v = lds64(in_base, delta + p);
          h1 = hash_short(v, K.E.hash_bytes, K.E.hash_log);
          e1 = tab1[h1];
          if (DFAST) { h2 = hash_long(v, K.E.long_log); e2 = tab2[h2]; }
        }
        bool w1, w2 = false;
        {
          uint32_t hp = __shfl_up_sync(0xffffffffu, h1, 1);
          if (lane == 0) hp = prev1;
          const bool ins = act && hp != h1;
          const uint32_t g = __match_any_sync(0xffffffffu, act ? h1 : 0x80000000u + (uint32_t)lane) & __ballot_sync(0xffffffffu, ins);
          const uint32_t lower = g & lt;
          if (lower) e1 = (p0 + (uint32_t)(31 - __clz(lower))) & 0xFFFFu;
          w1 = ins && (g >> lane) == 1u;
          prev1 = __shfl_sync(0xffffffffu, h1, 31);
        }
        if (DFAST) {
          uint32_t hp = __shfl_up_sync(0xffffffffu, h2, 1);
          if (lane == 0) hp = prev2;
          const bool ins = act && hp != h2;
          const uint32_t g = __match_any_sync(0xffffffffu, act ? h2 : 0x80000000u + (uint32_t)lane) & __ballot_sync(0xffffffffu, ins);
          const uint32_t lower = g & lt;
          if (lower) e2 = (p0 + (uint32_t)(31 - __clz(lower))) & 0xFFFFu;
          w2 = ins && (g >> lane) == 1u;
          prev2 = __shfl_sync(0xffffffffu, h2, 31);
        }
        __syncwarp();                                       // every lookup of the window precedes its inserts
        if (w1) tab1[h1] = (uint16_t)p;
        if (DFAST && w2) tab2[h2] = (uint16_t)p;
        ring[p & (ESD_RING_POS - 1)] = e1 | (e2 << 16);
        __syncwarp();
        cc_barrier();
        if (lane == 0) ctl->h_done = k + 1;
      }
      PROF_SINCE(4, th0);
    } else if (warp >= 2 && !rle) {
      // ======================================= V warps =======================================
      const int j = warp - 2;
      PROF_T0(tv0);
      for (uint32_t k = (uint32_t)j; k < nwin; k += ESD_NV) {
        bool over = false;
        PROF_T0(tv1);
        while (ctl->h_done <= k) { if (ctl->s_pos == ESD_PARSE_OVER) { over = true; break; } }
        PROF_SINCE(7, tv1);
        if (over) break;
        cc_barrier();
        const uint32_t p = (k << 5) + (uint32_t)lane;
        if (((k + 1) << 5) > ctl->s_pos) {                  // else: the parse is already past this window
          uint32_t res = 0;
          if (p < ilimit) {
            const uint32_t e = ring[p & (ESD_RING_POS - 1)];
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
          ring[p & (ESD_RING_POS - 1)] = res;
        }
        __syncwarp();
        cc_barrier();
        if (lane == 0) ctl->v_done[j] = k + ESD_NV;
      }
      PROF_SINCE(6, tv0);
    } else if (warp == 0) {
      // ======================================= S warp =======================================
      // scratch buffer of this record: wait until the X warp that coded its previous occupant is done
      PROF_T0(ts0);
      if (nrec >= ESD_NBUF) { const uint32_t o = nrec - ESD_NBUF; while (ctl->x_done[o % ESD_NX] <= o) __nanosleep(100); }
      PROF_SINCE(2, ts0);
      PROF_T0(ts1);
      __threadfence_block();
      uint8_t *const buf = scratch + (size_t)(nrec % ESD_NBUF) * G::buf_bytes;
      uint8_t *const lits = buf;
      uint32_t *const seqs = reinterpret_cast<uint32_t *>(buf + G::lits_bytes);          // [ll | ml | of] x max_seq
      uint32_t nseq = 0, nlit = 0;
      if (!rle) {
        uint32_t rep[3] = {1, 4, 8};
        // a block encoded on its own does not know the repeat offsets the decoder will hold when it gets there: 0 =
        // unknown, never matched against and never equal to a real offset (RFC 8878 3.1.1.5)
        if (blocks_only && item != 0) { rep[0] = 0; rep[1] = 0; rep[2] = 0; }
        uint32_t ip = 0, anchor = 0, rep0 = rep[0];
        while (ip < ilimit) {
          if (lane == 0) ctl->s_pos = ip;
          const uint32_t k0 = ip >> 5, k1 = min((ip + 31) >> 5, nwin - 1);
          PROF_T0(ts2);
          while (ctl->v_done[k0 % ESD_NV] <= k0 || ctl->v_done[k1 % ESD_NV] <= k1) { }
          PROF_SINCE(1, ts2);
          PROF_ADD(10, 1);
          cc_barrier();
          const uint32_t p = ip + (uint32_t)lane;
          const bool inb = p < ilimit;
          const uint32_t r = inb ? ring[p & (ESD_RING_POS - 1)] : 0u;
          const bool e = rep0 != 0 && p >= rep0 && p < bn && in[p] == in[p - rep0];
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
              const uint32_t m = __ballot_sync(0xffffffffu, q < bn && in[q] == in[q - off]);
              const uint32_t nn = ones_from(m, 0);
              len += nn;
              if (nn < 32) break;
            }
          }
          {
            const uint32_t jb = (uint32_t)lane;
            const bool mb = jb < s - anchor && s - 1 - jb >= off && in[s - 1 - jb] == in[s - 1 - jb - off];
            const uint32_t nb = ones_from(__ballot_sync(0xffffffffu, mb), 0);
            s -= nb; len += nb;
          }
          const uint32_t llen = s - anchor;
          for (uint32_t i = lane; i < llen; i += 32) lits[nlit + i] = in[anchor + i];
          const uint32_t code = offset_to_code(off, llen, rep);
          if (lane < 3) seqs[(uint32_t)lane * G::max_seq + nseq] = lane == 0 ? llen : lane == 1 ? len : code;
          nlit += llen; nseq++;
          ip = anchor = s + len; rep0 = off;
        }
        if (lane == 0) ctl->s_pos = ESD_PARSE_OVER;         // releases the H and V warps wherever they are
        for (uint32_t i = anchor + lane; i < bn; i += 32) lits[nlit + i - anchor] = in[i];
        nlit += bn - anchor;
      }
      __syncwarp();
      __threadfence_block();
      if (lane == 0) {
        EsdRecord R;
        R.item = item; R.kind = rle ? 1u : 0u; R.nlit = nlit; R.nseq = nseq;
        ctl->rec[nrec % ESD_NBUF] = R;
        __threadfence_block();
        ctl->pub = nrec + 1;
      }
      PROF_SINCE(0, ts1);
      PROF_ADD(8, 1);
      PROF_ADD(11, nseq);
    }
    nrec++;
    parse_bar();                                  // all parse warps are done with the staged block and the tables
  }
  if (tid == 0) ctl->stop = 1;
}

size_t esd_smem_bytes(const EsdParams &e, int big) {
  const size_t in = (big ? 131072u : 65536u) + ESD_IN_PAD;
  const size_t tabs = ((size_t)2 << e.hash_log) + (e.dfast ? ((size_t)2 << e.long_log) : 0);
  return 128 + in + tabs + ESD_RING_POS * 4 + ESD_NX * ((sizeof(EntropyWs) + 15) & ~(size_t)15) + 16;
}

template <int DFAST, int BIG> cudaError_t esd_launch_one(const EsdArgs &k, int grid, cudaStream_t stream) {
  const size_t smem = esd_smem_bytes(k.E, BIG);
  cudaError_t e = cudaFuncSetAttribute(zstd_encode_esd_kernel<DFAST, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  zstd_encode_esd_kernel<DFAST, BIG><<<grid, ESD_THREADS, smem, stream>>>(k);
  return cudaGetLastError();
}

} // namespace

#ifdef ESD_PROF
extern "C" void cuda_zstd_b200_esd_prof(unsigned long long *out16, int reset) {
  cudaMemcpyFromSymbol(out16, g_esd_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_esd_prof, z, sizeof z); }
}
#endif
size_t esd_cta_scratch_bytes(int big) { return big ? EsdGeom<1>::cta_scratch : EsdGeom<0>::cta_scratch; }
int esd_ctas_per_sm(int big) { return big ? 1 : 2; }

// Levels 1-4.  Up to three launches on `stream`: blocks <= 64 KB (two CTAs per SM), blocks <= 128 KB (one CTA per
// SM) over the items the first kernel handed on, and the general kernel (zstd_encode.cu) for multi-block items.
// max_item_bytes != 0 (the host knows the sizes) skips the launches that cannot have work.
cudaError_t launch_encode_esd(const EncodeArgs &args, const EsdLaunch &L, cudaStream_t stream, int *launches) {
  if (launches) *launches = 0;
  if (args.n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(L.counters, 0, 8 * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  EsdArgs k{};
  k.A = args;
  k.E = esd_params_for_level(args.prm.level);
  uint32_t *const list128 = L.lists, *const list_big = L.lists + args.n;
  const bool known = L.max_item_bytes != 0;
  const bool any64 = !known || L.min_item_bytes <= 65536;
  const bool any128 = !known || (L.max_item_bytes > 65536 && L.min_item_bytes <= BLOCK_BYTES);
  const bool any_big = !known || L.max_item_bytes > BLOCK_BYTES;
  const bool only_big = known && L.min_item_bytes > BLOCK_BYTES;
  int nl = 0;
  if (any64) {
    k.A.scratch = L.scratch;
    k.list = nullptr; k.list_count = nullptr;
    k.work_head = L.counters + 0;
    k.defer_list = list128; k.defer_count = L.counters + 3;
    k.big_list = list_big; k.big_count = L.counters + 4;
    const int grid = (int)min((size_t)args.n, (size_t)L.sm_count * 2);
    e = k.E.dfast ? esd_launch_one<1, 0>(k, grid, stream) : esd_launch_one<0, 0>(k, grid, stream);
    if (e != cudaSuccess) return e;
    nl++;
  }
  if (any128) {
    k.A.scratch = L.scratch;
    if (any64) { k.list = list128; k.list_count = L.counters + 3; } else { k.list = nullptr; k.list_count = nullptr; }
    k.work_head = L.counters + 1;
    k.defer_list = nullptr; k.defer_count = nullptr;
    k.big_list = list_big; k.big_count = L.counters + 4;
    const int grid = (int)min((size_t)args.n, (size_t)L.sm_count);
    e = k.E.dfast ? esd_launch_one<1, 1>(k, grid, stream) : esd_launch_one<0, 1>(k, grid, stream);
    if (e != cudaSuccess) return e;
    nl++;
  }
  if (any_big) {
    EncodeArgs g = args;
    g.scratch = L.scratch;
    g.counter = L.counters + 2;
    if (!only_big) { g.list = list_big; g.list_count = L.counters + 4; }
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
