// zstd_decode.cu -- batched Zstandard frame decoder for sm_100a: one persistent CTA per in-flight chunk.
//
// Replaces, for the batch path, the reference's chain of <<<1,1>>> kernels and host-side header
// parsing: ZstdBatchManager::decompress_batch (src/cuda_zstd_manager.cu:5799-5859) ->
// DefaultZstdManager::decompress (:3194-3700) -> decompress_block (:4292-4404) ->
// decompress_literals (:4981-5104) / huffman::decode_huffman_rfc8878 (src/cuda_zstd_huffman.cu:
// 2204-2378) / decompress_sequences (:5106-5531) / k_decode_sequences_interleaved
// (src/cuda_zstd_fse.cu:3839-4062) / sequence::execute_sequences (src/cuda_zstd_sequence.cu:459-568).
// Semantics follow RFC 8878 / libzstd where the reference deviates.
//
// Work split inside a CTA of 4 warps, per compressed block:
//   warp 0          Huffman tree (weights, DTable in SMEM) then the 1 or 4 literal streams (lanes 0-3)
//   warps 1..3      NCount headers (lane 0 of warp 1) then one FSE decode table each (LL, OF, ML)
//   warp 1 lane 0   interleaved FSE sequence decode, repcode resolution, running output offsets,
//                   written to a double-buffered SMEM ring of SEQ_BATCH records
//   warps 0,2,3     sequence execution (literal + match copies) of the previous batch while warp 1
//                   decodes the next; cross-warp match dependencies are resolved with a published
//                   "oldest byte still in flight" word per warp instead of a barrier per sequence.
#include "zstd_common.cuh"
#include "zstd_decode_tables.cuh"
#include "zstd_device_api.h"

namespace b200zstd {

constexpr int DEC_THREADS = 128;
constexpr int SEQ_BATCH = 256;
constexpr int N_EXEC = 3;                       // executing warps: 0, 2, 3

struct __align__(16) SeqRec { uint32_t out_pos, lit_pos, offset, ml; };

struct __align__(16) DecSmem {
  uint2 ll_tab[512];
  uint2 ml_tab[512];
  uint2 of_tab[256];
  uint16_t huf[1 << HUF_MAX_LOG];
  SeqRec seq[2][SEQ_BATCH];
  uint32_t seq_ll[2][SEQ_BATCH];
  int16_t norm[3][64];
  uint8_t item_sym[3][512];
  uint16_t sym_next[3][64];
  uint8_t weights[256];
  uint32_t huf_ft[64];           // FSE table of the Huffman weight stream (log <= 6)
  int16_t huf_norm[16];
  uint32_t rank_cnt[16];
  uint32_t rank_start[16];
  uint32_t cur_start[2][4];      // per batch buffer, per executing warp: out_pos of the sequence it works on
  // control words (written by one thread, read by all after a barrier)
  uint32_t chunk;
  uint32_t status;
  int tab_log[3];                // LL, OF, ML accuracy logs
  int tab_max[3];
  int tab_valid[3];
  int huf_log;
  int huf_valid;
  uint32_t seq_bits_off;         // offset of the sequence bitstream inside the block
  uint32_t batch_cnt[2];
  uint32_t blk_out_end;          // output position after the last decoded sequence
  uint32_t blk_lit_end;          // literals consumed by sequences
};

// One Huffman stream -> dst[0..count).  Single thread.  Returns false on malformed stream.
__device__ bool huf_decode_stream(const uint16_t *tab, int log, const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t count) {
  BackBits b;
  if (!b.init(src, n)) return false;
  const int sh = 64 - log;
  uint32_t i = 0;
  // after refill() at least 33 bits are buffered: three symbols (<= 11 bits each) per refill
  for (; i + 3 <= count; i += 3) {
    b.refill();
    uint32_t e0 = tab[b.win >> sh]; b.skip((int)(e0 >> 8));
    uint32_t e1 = tab[b.win >> sh]; b.skip((int)(e1 >> 8));
    uint32_t e2 = tab[b.win >> sh]; b.skip((int)(e2 >> 8));
    dst[i] = (uint8_t)e0; dst[i + 1] = (uint8_t)e1; dst[i + 2] = (uint8_t)e2;
  }
  for (; i < count; i++) {
    b.refill();
    uint32_t e = tab[b.win >> sh];
    b.skip((int)(e >> 8));
    dst[i] = (uint8_t)e;
  }
  return b.left == 0;
}

// ---------------------------------------------------------------------------------------------
// Sequence execution by one warp: literal run then match (RFC 8878 3.1.2.5 copy semantics)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_literals_warp(uint8_t *dst, const uint8_t *lit, uint32_t lit_pos, uint32_t lit_stride1,
                                                   uint32_t len, int lane) {
  // lit_stride1 = 1 for a real literal buffer, 0 for RLE literals (every read hits byte 0)
  for (uint32_t k = lane; k < len; k += 32) dst[k] = lit[(lit_pos + k) * lit_stride1];
}

__device__ __forceinline__ void copy_match_warp(uint8_t *out, uint32_t d, uint32_t offset, uint32_t ml, int lane) {
  const uint8_t *src = out + d - offset;
  uint8_t *dst = out + d;
  if (offset >= ml) {
    for (uint32_t k = lane; k < ml; k += 32) dst[k] = src[k];
  } else {
    // overlapping copy == periodic extension of the last `offset` bytes
    for (uint32_t k = lane; k < ml; k += 32) dst[k] = src[k % offset];
  }
}

struct SeqDecodeState {
  BackBits b;
  uint32_t sl, so, sm;
  uint32_t rep0, rep1, rep2;
  uint32_t out_pos, lit_pos;
  uint32_t remaining;      // sequences not yet decoded
};

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_THREADS) zstd_decode_batch_kernel(DecodeArgs A) {
  __shared__ DecSmem S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t *const lit_buf = A.lit_scratch + (size_t)blockIdx.x * LIT_SCRATCH_BYTES;
  const int exec_id = (warp == 0) ? 0 : warp - 1;      // warps 0,2,3 -> 0,1,2 (warp 1 never executes)

  for (;;) {
    if (tid == 0) {
      // work queue: either every chunk of the batch, or the list the fast path left for this kernel
      const uint32_t q = atomicAdd(A.counter, 1u);
      const uint32_t limit = A.list ? *A.list_count : A.n;
      S.chunk = q < limit ? (A.list ? A.list[q] : q) : 0xFFFFFFFFu;
    }
    __syncthreads();
    const uint32_t chunk = S.chunk;
    if (chunk == 0xFFFFFFFFu) break;

    const uint8_t *const src = (const uint8_t *)A.in_ptrs[chunk];
    const size_t src_size = A.in_sizes[chunk];
    uint8_t *const dst = (uint8_t *)A.out_ptrs[chunk];
    const size_t dst_cap = A.out_sizes[chunk];
    uint32_t status = ST_OK;
    size_t ip = 0, op = 0;
    int frames = 0;
    if (src == nullptr || (dst == nullptr && dst_cap != 0)) status = ST_INVALID_PARAMETER;
    else if (src_size < 4) status = ST_INVALID_PARAMETER;          // reference: manager.cu:3202-3206

    // ---- frames (uniform control flow: every thread parses the same header bytes) ----
    while (status == ST_OK && ip < src_size) {
      if (src_size - ip < 4) { status = frames ? ST_CORRUPT : ST_INVALID_MAGIC; break; }
      const uint32_t magic = ld_le32(src + ip);
      if ((magic & 0xFFFFFFF0u) == ZSTD_SKIP_MAGIC) {
        if (src_size - ip < 8) { status = ST_CORRUPT; break; }
        const uint64_t sz = ld_le32(src + ip + 4);
        if (sz + 8 > src_size - ip) { status = ST_CORRUPT; break; }
        ip += 8 + (size_t)sz;
        continue;
      }
      if (magic != ZSTD_FRAME_MAGIC) { status = frames ? ST_CORRUPT : ST_INVALID_MAGIC; break; }
      size_t h = ip + 4;
      if (h >= src_size) { status = ST_CORRUPT; break; }
      const uint32_t fhd = src[h++];
      const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, has_ck = (fhd >> 2) & 1, did_flag = fhd & 3;
      if (fhd & 0x08) { status = ST_UNSUPPORTED; break; }
      const int did_size = did_flag == 3 ? 4 : did_flag;
      const int fcs_size = fcs_flag == 0 ? single : (1 << fcs_flag);
      if (h + (single ? 0 : 1) + did_size + fcs_size > src_size) { status = ST_CORRUPT; break; }
      if (!single) { if ((src[h++] >> 3) > 21) { status = ST_UNSUPPORTED; break; } }
      uint32_t dict_id = 0;
      for (int k = 0; k < did_size; k++) dict_id |= (uint32_t)src[h + k] << (8 * k);
      h += did_size;
      uint64_t fcs = ~0ull;
      if (fcs_size) {
        fcs = 0;
        for (int k = 0; k < fcs_size; k++) fcs |= (uint64_t)src[h + k] << (8 * k);
        if (fcs_size == 2) fcs += 256;
        h += fcs_size;
      }
      if (dict_id != 0) { status = ST_DICT_MISMATCH; break; }
      if (fcs != ~0ull && fcs > dst_cap - op) { status = ST_BUFFER_TOO_SMALL; break; }
      ip = h;
      const size_t frame_op = op;                    // offsets may reach back to here only
      uint8_t *const fout = dst + frame_op;
      const uint64_t fcap64 = dst_cap - frame_op;
      const uint32_t fcap = fcap64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)fcap64;
      uint32_t fop = 0;                              // output position inside this frame
      uint32_t rep0 = 1, rep1 = 4, rep2 = 8;         // meaningful in warp 1 lane 0 only
      __syncthreads();
      if (tid == 0) { S.huf_valid = 0; S.tab_valid[0] = S.tab_valid[1] = S.tab_valid[2] = 0; }
      __syncthreads();

      // ---- blocks ----
      for (;;) {
        if (src_size - ip < 3) { status = ST_CORRUPT; break; }
        const uint32_t bh = ld_le24(src + ip);
        ip += 3;
        const int last = bh & 1, btype = (bh >> 1) & 3;
        const uint32_t bsize = bh >> 3;
        if (btype == 3 || bsize > BLOCK_MAX) { status = ST_CORRUPT; break; }
        if (btype == 0) {                                            // Raw
          if (bsize > src_size - ip) { status = ST_CORRUPT; break; }
          if (bsize > fcap - fop) { status = ST_BUFFER_TOO_SMALL; break; }
          for (uint32_t k = tid; k < bsize; k += DEC_THREADS) fout[fop + k] = src[ip + k];
          ip += bsize; fop += bsize;
        } else if (btype == 1) {                                     // RLE
          if (src_size - ip < 1) { status = ST_CORRUPT; break; }
          if (bsize > fcap - fop) { status = ST_BUFFER_TOO_SMALL; break; }
          const uint8_t v = src[ip];
          for (uint32_t k = tid; k < bsize; k += DEC_THREADS) fout[fop + k] = v;
          ip += 1; fop += bsize;
        } else {                                                     // Compressed
          if (bsize > src_size - ip || bsize < 2) { status = ST_CORRUPT; break; }
          const uint8_t *const bp = src + ip;
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
          uint32_t sp = lhs + lit_comp;                              // cursor inside the block
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

          const uint8_t *lit_ptr = lit_buf;
          uint32_t lit_stride1 = 1;
          if (ltype == 0) lit_ptr = bp + lhs;
          else if (ltype == 1) { lit_ptr = bp + lhs; lit_stride1 = 0; }

          __syncthreads();
          if (tid == 0) S.status = ST_OK;
          __syncthreads();

          // ================= phase 1: tables + literals (warp 0)  ||  FSE tables + first batch =================
          if (warp == 0) {
            if (ltype >= 2) {
              const uint8_t *hp = bp + lhs;
              uint32_t rem = lit_comp;
              bool ok = true;
              if (ltype == 2) {
                int used = huf_read_table_warp(S, hp, rem, lane);
                if (used < 0) ok = false; else { hp += used; rem -= (uint32_t)used; }
              } else if (!S.huf_valid) ok = false;
              if (ok) {
                const int hlog = S.huf_log;
                if (nstreams == 1) {
                  if (lane == 0) ok = huf_decode_stream(S.huf, hlog, hp, rem, lit_buf, lit_size);
                } else {
                  const uint32_t seg = (lit_size + 3) >> 2;
                  if (rem < 6 || seg * 3 > lit_size) ok = false;
                  else {
                    const uint32_t s1 = ld_le16(hp), s2 = ld_le16(hp + 2), s3 = ld_le16(hp + 4);
                    if (6 + s1 + s2 + s3 > rem) ok = false;
                    else if (lane < 4) {
                      const uint32_t s4 = rem - 6 - s1 - s2 - s3;
                      const uint32_t off = (lane == 0) ? 0 : (lane == 1) ? s1 : (lane == 2) ? s1 + s2 : s1 + s2 + s3;
                      const uint32_t len = (lane == 0) ? s1 : (lane == 1) ? s2 : (lane == 2) ? s3 : s4;
                      const uint32_t cnt = (lane == 3) ? lit_size - 3 * seg : seg;
                      ok = huf_decode_stream(S.huf, hlog, hp + 6 + off, len, lit_buf + lane * seg, cnt);
                    }
                  }
                }
              }
              if (__any_sync(0xffffffffu, !ok) && lane == 0) S.status = ST_CORRUPT;
            }
          } else if (nseq) {
            // warps 1..3: table descriptions are serial in the byte stream (lane 0 of warp 1) ...
            if (warp == 1 && lane == 0) {
              uint32_t p = sp;
              bool ok = true;
              for (int t = 0; t < 3 && ok; t++) {                    // order in the stream: LL, OF, ML
                const int mode = (modes >> (6 - 2 * t)) & 3;
                const int max_allowed = (t == 0) ? LL_MAX_SYM : (t == 1) ? OF_MAX_SYM : ML_MAX_SYM;
                if (mode == 0) {
                  const int16_t *def = (t == 0) ? c_ll_def : (t == 1) ? c_of_def : c_ml_def;
                  const int dmax = (t == 0) ? 35 : (t == 1) ? 28 : 52;
                  for (int i = 0; i <= dmax; i++) S.norm[t][i] = def[i];
                  S.tab_max[t] = dmax; S.tab_log[t] = (t == 1) ? OF_DEF_LOG : LL_DEF_LOG; S.tab_valid[t] = 2;   // 2 = (re)build
                } else if (mode == 1) {
                  if (p >= bsize || bp[p] > max_allowed) { ok = false; break; }
                  uint2 *tab = (t == 0) ? S.ll_tab : (t == 1) ? S.of_tab : S.ml_tab;
                  tab[0] = seq_entry(t, bp[p], 0, 0);
                  S.tab_log[t] = 0; S.tab_valid[t] = 1;
                  p += 1;
                } else if (mode == 2) {
                  int ms = 0, al = 0;
                  const int max_log = (t == 1) ? OF_MAX_LOG : LL_MAX_LOG;
                  int used = (p < bsize) ? read_ncount(bp + p, bsize - p, S.norm[t], max_allowed, max_log, &ms, &al) : -1;
                  if (used < 0) { ok = false; break; }
                  S.tab_max[t] = ms; S.tab_log[t] = al; S.tab_valid[t] = 2;
                  p += (uint32_t)used;
                } else if (!S.tab_valid[t]) ok = false;                // Repeat needs a previous table
              }
              if (!ok || p >= bsize) S.status = ST_CORRUPT;
              S.seq_bits_off = p;
            }
            asm volatile("bar.sync 1, 96;" ::: "memory");              // warps 1..3 only
            // ... then one table per warp, in parallel
            const int t = warp - 1;
            if (S.status == ST_OK && S.tab_valid[t] == 2) {
              uint2 *tab = (t == 0) ? S.ll_tab : (t == 1) ? S.of_tab : S.ml_tab;
              fse_build_warp(tab, S.norm[t], S.tab_max[t], S.tab_log[t], t, S.item_sym[t], S.sym_next[t], lane);
              if (lane == 0) S.tab_valid[t] = 1;
            }
            asm volatile("bar.sync 1, 96;" ::: "memory");
          }

          // sequence decoder state lives in warp 1 lane 0's registers for the whole block
          SeqDecodeState D;
          D.remaining = 0;
          const bool is_decoder = (warp == 1 && lane == 0);
          if (is_decoder && nseq && S.status == ST_OK) {
            const uint32_t p = S.seq_bits_off;
            if (!D.b.init(bp + p, bsize - p)) S.status = ST_CORRUPT;
            else {
              D.b.refill();
              D.sl = D.b.read(S.tab_log[0]); D.so = D.b.read(S.tab_log[1]);
              D.b.refill();
              D.sm = D.b.read(S.tab_log[2]);
              D.rep0 = rep0; D.rep1 = rep1; D.rep2 = rep2;
              D.out_pos = fop; D.lit_pos = 0; D.remaining = nseq;
            }
          }

          // decode one batch into ring slot `slot` (warp 1 lane 0)
          auto decode_batch = [&](int slot) {
            uint32_t cnt = 0;
            uint32_t err = ST_OK;
            while (cnt < SEQ_BATCH && D.remaining) {
              D.b.refill();
              const uint2 eo = S.of_tab[D.so], em = S.ml_tab[D.sm], el = S.ll_tab[D.sl];
              const uint32_t ov = eo.y + D.b.read((int)(eo.x >> 24));
              D.b.refill();
              const uint32_t ml = em.y + D.b.read((int)(em.x >> 24));
              const uint32_t ll = el.y + D.b.read((int)(el.x >> 24));
              D.b.refill();
              uint32_t offset;
              if (ov > 3) { offset = ov - 3; D.rep2 = D.rep1; D.rep1 = D.rep0; D.rep0 = offset; }
              else {
                const uint32_t idx = ov - 1 + (ll == 0);
                if (idx == 0) offset = D.rep0;
                else {
                  offset = (idx == 3) ? D.rep0 - 1 : (idx == 1) ? D.rep1 : D.rep2;
                  if (idx != 1) D.rep2 = D.rep1;
                  D.rep1 = D.rep0; D.rep0 = offset;
                }
              }
              D.remaining--;
              if (D.remaining) {
                D.sl = (el.x & 0xFFFF) + D.b.read((int)((el.x >> 16) & 0xFF));
                D.sm = (em.x & 0xFFFF) + D.b.read((int)((em.x >> 16) & 0xFF));
                D.so = (eo.x & 0xFFFF) + D.b.read((int)((eo.x >> 16) & 0xFF));
              }
              if (D.b.left < 0 || offset == 0 || D.lit_pos + ll > lit_size || offset > D.out_pos + ll) {
                err = ST_CORRUPT; D.remaining = 0; break;
              }
              if ((uint64_t)D.out_pos + ll + ml > fcap) { err = ST_BUFFER_TOO_SMALL; D.remaining = 0; break; }
              S.seq[slot][cnt] = SeqRec{D.out_pos, D.lit_pos, offset, ml};
              S.seq_ll[slot][cnt] = ll;
              D.out_pos += ll + ml; D.lit_pos += ll;
              cnt++;
            }
            if (err == ST_OK && D.remaining == 0 && D.b.left != 0) err = ST_CORRUPT;    // stream must end exactly
            if (err != ST_OK) S.status = err;
            S.batch_cnt[slot] = cnt;
            for (int w = 0; w < N_EXEC; w++) S.cur_start[slot][w] = ((uint32_t)w < cnt) ? S.seq[slot][w].out_pos : 0xFFFFFFFFu;
            S.blk_out_end = D.out_pos; S.blk_lit_end = D.lit_pos;
          };

          if (is_decoder) {
            if (nseq && S.status == ST_OK) decode_batch(0);
            else { S.batch_cnt[0] = 0; S.blk_out_end = fop; S.blk_lit_end = 0; }
          }
          __syncthreads();                                              // literals + tables + batch 0 ready

          // ================= phase 2: execute batch b  ||  decode batch b+1 =================
          uint32_t slot = 0;
          for (;;) {
            const uint32_t cnt = S.batch_cnt[slot];
            if (cnt == 0) break;
            if (warp == 1) {
              if (lane == 0) {
                if (D.remaining && S.status == ST_OK) decode_batch(slot ^ 1);
                else S.batch_cnt[slot ^ 1] = 0;
              }
            } else {
              volatile uint32_t *cs = S.cur_start[slot];
              for (uint32_t i = exec_id; i < cnt; i += N_EXEC) {
                const SeqRec r = S.seq[slot][i];
                const uint32_t ll = S.seq_ll[slot][i];
                copy_literals_warp(fout + r.out_pos, lit_ptr, r.lit_pos, lit_stride1, ll, lane);
                __syncwarp();
                const uint32_t d = r.out_pos + ll;
                const uint32_t need = min(d - r.offset + r.ml, d);      // bytes below `need` must be final
                // every other executing warp must have moved past `need`
                for (;;) {
                  uint32_t a = cs[(exec_id + 1) % N_EXEC], c = cs[(exec_id + 2) % N_EXEC];
                  if (min(a, c) >= need) break;
                }
                __threadfence_block();
                copy_match_warp(fout, d, r.offset, r.ml, lane);
                __syncwarp();
                if (lane == 0) {
                  __threadfence_block();
                  cs[exec_id] = (i + N_EXEC < cnt) ? S.seq[slot][i + N_EXEC].out_pos : 0xFFFFFFFFu;
                }
                __syncwarp();
              }
            }
            __syncthreads();
            slot ^= 1;
          }
          // trailing literals + bookkeeping
          if (S.status != ST_OK) { status = S.status; break; }
          {
            const uint32_t out_end = S.blk_out_end, lit_end = S.blk_lit_end;
            const uint32_t rest = lit_size - lit_end;
            if (rest > fcap - out_end) { status = ST_BUFFER_TOO_SMALL; break; }
            for (uint32_t k = tid; k < rest; k += DEC_THREADS) fout[out_end + k] = lit_ptr[(lit_end + k) * lit_stride1];
            fop = out_end + rest;
          }
          if (is_decoder && nseq) { rep0 = D.rep0; rep1 = D.rep1; rep2 = D.rep2; }
          ip += bsize;
          __syncthreads();                                             // output of this block visible to all warps
        }
        if (last) break;
      }
      if (status != ST_OK) break;
      if (fcs != ~0ull && (uint64_t)fop != fcs) { status = ST_CORRUPT; break; }
      if (has_ck) {
        if (src_size - ip < 4) { status = ST_CORRUPT; break; }
        if (A.verify_checksum) {
          __syncthreads();
          if (warp == 0) {
            const uint64_t hsh = xxh64_warp(fout, fop, lane);
            if (lane == 0) S.status = ((uint32_t)hsh == ld_le32(src + ip)) ? ST_OK : ST_CHECKSUM;
          }
          __syncthreads();
          if (S.status != ST_OK) { status = S.status; break; }
        }
        ip += 4;
      }
      op += fop;
      frames++;
    }
    if (status == ST_OK && frames == 0) status = ST_INVALID_MAGIC;
    __syncthreads();
    if (tid == 0) {
      A.out_sizes[chunk] = (status == ST_OK) ? op : 0;
      if (A.statuses) A.statuses[chunk] = status;
    }
  }
}

cudaError_t launch_decode_batch_nomemset(const DecodeArgs &args, int grid, cudaStream_t stream) {
  if (args.n == 0) return cudaSuccess;
  zstd_decode_batch_kernel<<<grid, DEC_THREADS, 0, stream>>>(args);
  return cudaGetLastError();
}
cudaError_t launch_decode_batch(const DecodeArgs &args, int grid, cudaStream_t stream) {
  if (args.n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(args.counter, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  return launch_decode_batch_nomemset(args, grid, stream);
}

int decode_ctas_per_sm() {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, zstd_decode_batch_kernel, DEC_THREADS, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}

} // namespace b200zstd
