// zstd_encode_entropy.cuh -- warp-parallel entropy stage of one block (literals section + sequences section),
// shared by the two encoder kernels (zstd_encode.cu: chain levels and large items; zstd_encode_esd.cu: levels 1-4).
// Byte-identical to enc::encode_block_payload, the serial form the host model runs (tests/model/enc_model.cpp).
// Reference counterparts: compress_literals / compress_sequences, src/cuda_zstd_manager.cu:4406-4484, 4864-4974.
#pragma once
#include "zstd_common.cuh"
#include "zstd_encode_core.cuh"

namespace b200zstd {
using namespace enc;

// ---------------------------------------------------------------------------------------------------
// Entropy stage of one block, warp-parallel.  Produces exactly the bytes of enc::encode_block_payload
// (the serial form the host model runs): histograms by shared-memory atomics across 32 lanes, the 4
// Huffman streams cut into one piece per lane at bit offsets known from a bit-count pre-pass, the three FSE state
// chains on 3 lanes fed by shuffles (state bits parked in the spare high bits of the sequence arrays), and the interleaved
// sequence bitstream assembled by all lanes: each lane packs a contiguous run of sequences into
// 32-bit words at a bit offset known from a prefix sum (atomicOr only for the two words it shares).
// The table builders (Huffman lengths, FSE normalisation / CTable) stay on lane 0.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) {
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
constexpr uint32_t SEQ_VAL_MASK = (1u << 18) - 1;     // value | state bits << 18 | nb << 27

static __device__ uint32_t entropy_stage_warp(EntropyWs &W, const uint8_t *lits, uint32_t nlit, uint32_t *sll, uint32_t *sml, uint32_t *sofv,
                                       uint32_t nseq, uint8_t *dst, uint32_t cap, int lane) {
  uint32_t op = 0;
  bool done = false;
  // ---------------- literals ----------------
  if (nlit >= 64) {
    for (int i = lane; i < 256; i += 32) W.count[i] = 0;
    __syncwarp();
    {
      // four literals per lane and step through aligned 32-bit loads (lits is 4-byte aligned: the kernels' own buffer)
      const uint32_t head = min(nlit, (uint32_t)((4 - ((uintptr_t)lits & 3)) & 3));
      if ((uint32_t)lane < head) atomicAdd(&W.count[lits[lane]], 1u);
      const uint32_t *w = reinterpret_cast<const uint32_t *>(lits + head);
      const uint32_t words = (nlit - head) >> 2;
      for (uint32_t i = lane; i < words; i += 32) {
        const uint32_t x = w[i];
        atomicAdd(&W.count[x & 0xFF], 1u); atomicAdd(&W.count[(x >> 8) & 0xFF], 1u);
        atomicAdd(&W.count[(x >> 16) & 0xFF], 1u); atomicAdd(&W.count[x >> 24], 1u);
      }
      const uint32_t done4 = head + (words << 2);
      if (done4 + (uint32_t)lane < nlit) atomicAdd(&W.count[lits[done4 + lane]], 1u);
    }
    __syncwarp();
    uint32_t ms = 0, mc = 0;
    for (int s = lane; s < 256; s += 32) { const uint32_t c = W.count[s]; if (c) ms = (uint32_t)s; mc = max(mc, c); }
    const int max_sym = (int)warp_max_u32(ms);
    const uint32_t maxc = warp_max_u32(mc);
    if (maxc == nlit) {
      if (cap < 4) return 0;
      const uint32_t hdr = nlit < 32 ? 1 : nlit < 4096 ? 2 : 3;
      if (lane == 0) { write_lit_header_raw_rle(dst, 1, nlit); dst[hdr] = lits[0]; }
      op = hdr + 1;
      done = true;
    } else if (maxc <= (nlit >> 7) + 4) {
      // flat histogram: raw
    } else {
      const uint32_t hs = lit_header_size_compressed(nlit);
      const bool single = nlit < 256;
      uint32_t budget = nlit - ((nlit >> 6) + 2);
      if (budget + hs > cap) budget = cap > hs ? cap - hs : 0;
      uint8_t *body = dst + hs;
      int tl = 0;
      uint32_t t = 0;
      if (lane == 0) {
        tl = huf_build_lengths(W.count, max_sym, 11, W.huflen, W.order, W.ncount, W.parent);
        if (tl > 0) {
          uint16_t *codes = W.order;
          huf_assign_codes(W.huflen, max_sym, tl, codes);
          for (int s = 0; s <= max_sym; s++) W.hufc[s] = W.huflen[s] ? ((uint32_t)codes[s] | ((uint32_t)W.huflen[s] << 16)) : 0u;
          t = huf_write_table(W, max_sym, tl, body, budget);
        }
      }
      tl = __shfl_sync(0xffffffffu, tl, 0);
      t = __shfl_sync(0xffffffffu, t, 0);
      __syncwarp();
      if (tl > 0) {
        bool ok = t != 0;
        uint32_t used = t;
        if (ok) {
          const uint32_t nstreams = single ? 1 : 4, seg = (nlit + 3) / 4;
          // Every stream is cut into pieces, one per lane (8 lanes per stream, 32 for a single stream); a piece is a
          // contiguous run of symbols.  A stream is written from its LAST symbol to its first, so a piece starts at the
          // bit count of the pieces behind it: pass 1 counts bits, a suffix sum inside the stream's lane group gives
          // every piece its bit offset and every stream its size, pass 2 writes.
          const uint32_t per = 32 / nstreams;                       // lanes per stream
          const uint32_t k = (uint32_t)lane / per, m = (uint32_t)lane % per;
          const uint32_t b0 = single ? 0 : k * seg, cnt = single ? nlit : (k < 3 ? seg : nlit - 3 * seg);
          const uint32_t piece = (cnt + per - 1) / per;
          const uint32_t lo = min(cnt, m * piece), hi = min(cnt, lo + piece);
          uint32_t bits = 0;
          for (uint32_t i = lo; i < hi; i++) bits += W.hufc[lits[b0 + i]] >> 16;
          // suffix sum over the lanes of my stream that hold later symbols (higher m)
          uint32_t after = 0, total = bits;
          for (uint32_t o = 1; o < per; o <<= 1) {
            const uint32_t t = __shfl_down_sync(0xffffffffu, total, o);
            if (m + o < per) total += t;
          }
          after = total - bits;
          const uint32_t stream_bits = __shfl_sync(0xffffffffu, total, (int)(k * per));
          uint32_t sz[4];
          for (uint32_t q = 0; q < 4; q++) sz[q] = q < nstreams ? (__shfl_sync(0xffffffffu, stream_bits, (int)(q * per)) + 1 + 7) >> 3 : 0u;
          uint32_t off[4] = {0, 0, 0, 0};
          if (single) {
            if (sz[0] > budget - used) ok = false; else { off[0] = used; used += sz[0]; }
          } else {
            if (used + 6 > budget) ok = false;
            else {
              const uint32_t jt = used;
              used += 6;
              for (int q = 0; q < 4 && ok; q++) {
                if (sz[q] > budget - used || sz[q] > 0xFFFF) { ok = false; break; }
                off[q] = used;
                used += sz[q];
              }
              if (ok && lane == 0)
                for (int q = 0; q < 3; q++) { body[jt + 2 * q] = (uint8_t)sz[q]; body[jt + 2 * q + 1] = (uint8_t)(sz[q] >> 8); }
            }
          }
          if (ok) {
            // zero the words the streams cover (not the bytes before them in the first word), then OR / store the pieces
            uint8_t *const sp = body + off[0];
            const uint32_t span_bytes = (single ? sz[0] : off[3] + sz[3] - off[0]);
            uint32_t *const w32 = reinterpret_cast<uint32_t *>((uintptr_t)sp & ~(uintptr_t)3);
            const uint32_t lead = (uint32_t)((uintptr_t)sp & 3);
            {
              const uint32_t last_word = (lead + span_bytes - 1) >> 2;
              const uint32_t headb = (4 - lead) & 3;
              if ((uint32_t)lane < min(headb, span_bytes)) sp[lane] = 0;
              uint32_t *z = w32 + (lead ? 1 : 0);
              const uint32_t words = lead ? last_word : last_word + 1;
              for (uint32_t i = lane; i < words; i += 32) z[i] = 0;
            }
            __syncwarp();
            const uint32_t o0 = k == 0 ? off[0] : k == 1 ? off[1] : k == 2 ? off[2] : off[3];
            uint32_t bitpos = (lead + (o0 - off[0])) * 8 + after;
            uint32_t wi = bitpos >> 5, nacc = bitpos & 31;
            uint64_t acc = 0;
            bool first = true;
            if (k < nstreams) {
              for (uint32_t i = hi; i > lo; i--) {
                const uint32_t c = W.hufc[lits[b0 + i - 1]];
                acc |= (uint64_t)(c & 0xFFFF) << nacc;
                nacc += c >> 16;
                if (nacc >= 32) {
                  if (first) { atomicOr(w32 + wi, (uint32_t)acc); first = false; } else w32[wi] = (uint32_t)acc;
                  wi++; acc >>= 32; nacc -= 32;
                }
              }
              if (m == 0) { acc |= 1ull << nacc; nacc++; }           // end mark behind the stream's first symbol
              if (nacc >= 32) {
                if (first) { atomicOr(w32 + wi, (uint32_t)acc); first = false; } else w32[wi] = (uint32_t)acc;
                wi++; acc >>= 32; nacc -= 32;
              }
              if (nacc > 0 && (uint32_t)acc != 0) atomicOr(w32 + wi, (uint32_t)acc);
            }
            __syncwarp();
          }
        }
        if (ok && used < budget) {
          if (lane == 0) write_lit_header_compressed(dst, hs, single, nlit, used);
          op = hs + used;
          done = true;
        }
      }
    }
  }
  if (!done) {
    if (cap < nlit + 3) return 0;
    const uint32_t hdr = nlit < 32 ? 1 : nlit < 4096 ? 2 : 3;
    if (lane == 0) write_lit_header_raw_rle(dst, 0, nlit);
    __syncwarp();
    for (uint32_t i = lane; i < nlit; i += 32) dst[hdr + i] = lits[i];
    op = hdr + nlit;
  }
  // ---------------- sequences ----------------
  if (op + 4 > cap) return 0;
  if (lane == 0) seq_count_header(dst + op, nseq);
  op += nseq < 128 ? 1 : nseq < 0x7F00 ? 2 : 3;
  if (nseq == 0) return op;
  const uint32_t modes_pos = op++;
  uint32_t modes = 0;
  for (int kind = 0; kind < 3; kind++) {
    for (int i = lane; i < 64; i += 32) W.count[i] = 0;
    __syncwarp();
    uint32_t mx = 0;
    for (uint32_t i = lane; i < nseq; i += 32) {
      const uint32_t c = kind == 0 ? ll_code(sll[i]) : kind == 1 ? (uint32_t)hb32(sofv[i]) : ml_code(sml[i]);
      atomicAdd(&W.count[c], 1u);
      mx = max(mx, c);
    }
    const int maxc = (int)warp_max_u32(mx);
    __syncwarp();
    int mode = 0;
    uint32_t desc = 0;
    if (lane == 0) mode = seq_table_prepare(W, kind, W.count, maxc, nseq, dst + op, cap - op, &desc);
    mode = __shfl_sync(0xffffffffu, mode, 0);
    desc = __shfl_sync(0xffffffffu, desc, 0);
    __syncwarp();
    if (mode < 0) return 0;
    op += desc;
    modes |= (uint32_t)mode << (6 - 2 * kind);
  }
  if (lane == 0) dst[modes_pos] = (uint8_t)modes;
  // ---- the three FSE state chains, one lane each: state bits parked above the 18-bit values ----
  const int log_ll = W.tab_log[0], log_of = W.tab_log[1], log_ml = W.tab_log[2];
  // Rounds of 32 sequences, last sequence first: every lane loads one sequence and computes its three codes; the codes
  // are handed to the chain lanes (0: LL, 1: OF, 2: ML) one sequence at a time by shuffle, so the serial part of a step is
  // two shared-memory lookups and no global access; what a chain emits for a sequence (bits | count << 9) comes back
  // through W.count, which is free by now.
  uint32_t fin = 0;
  {
    const int t = lane < 3 ? lane : 0;
    const int log = W.tab_log[t];
    const uint16_t *st = W.state_tab(t);
    const SymTT *tt = W.tt[t];
    uint32_t *const park = W.count;                    // [3][32]
    uint32_t state = 0;
    for (uint32_t base = 0; base < nseq; base += 32) {
      const uint32_t kk = base + (uint32_t)lane;       // stream order: k = 0 is the last sequence
      const bool valid = kk < nseq;
      const uint32_t i = valid ? nseq - 1 - kk : 0;
      uint32_t a = 0, b = 0, c = 1;
      if (valid) { a = sll[i]; b = sml[i]; c = sofv[i]; }
      const uint32_t pk = ll_code(a) | ((uint32_t)hb32(c) << 8) | (ml_code(valid ? b : 3u) << 16);
      const uint32_t nv = min(32u, nseq - base);
      for (uint32_t j = 0; j < nv; j++) {
        const uint32_t x = __shfl_sync(0xffffffffu, pk, (int)j);
        if (lane < 3 && log) {
          const uint32_t code = (x >> (8 * t)) & 0xFF;
          if (base + j == 0) state = fse_init_state(st, tt, code);
          else {
            const SymTT e = tt[code];
            const uint32_t nb = (uint32_t)((int32_t)state + e.delta_nb) >> 16;
            park[t * 32 + j] = (state & ((1u << nb) - 1u)) | (nb << 9);
            state = st[(int32_t)(state >> nb) + e.delta_state];
          }
        }
      }
      __syncwarp();
      if (valid && kk > 0) {
        if (log_ll) { const uint32_t v = park[lane]; sll[i] = a | ((v & 0x1FF) << 18) | ((v >> 9) << 27); }
        if (log_of) { const uint32_t v = park[32 + lane]; sofv[i] = c | ((v & 0x1FF) << 18) | ((v >> 9) << 27); }
        if (log_ml) { const uint32_t v = park[64 + lane]; sml[i] = b | ((v & 0x1FF) << 18) | ((v >> 9) << 27); }
      }
      __syncwarp();
    }
    if (lane < 3) fin = state;
  }
  const uint32_t fin_ll = __shfl_sync(0xffffffffu, fin, 0), fin_of = __shfl_sync(0xffffffffu, fin, 1), fin_ml = __shfl_sync(0xffffffffu, fin, 2);
  __syncwarp();
  // ---- bit budget of every lane's run (stream order k = nseq-1-i), prefix sum, capacity check ----
  const uint32_t R = (nseq + 31) / 32;
  const uint32_t k_lo = min(nseq, (uint32_t)lane * R), k_hi = min(nseq, k_lo + R);
  uint32_t my_bits = 0;
  for (uint32_t k = k_lo; k < k_hi; k++) {
    const uint32_t i = nseq - 1 - k;
    const uint32_t a = sll[i], b = sml[i], c = sofv[i];
    my_bits += ll_xbits(ll_code(a & SEQ_VAL_MASK)) + ml_xbits(ml_code(b & SEQ_VAL_MASK)) + (uint32_t)hb32(c & SEQ_VAL_MASK);
    my_bits += (a >> 27) + (b >> 27) + (c >> 27);                   // zero for the first sequence in stream order
  }
  uint32_t incl = my_bits;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  const uint32_t start_bit = incl - my_bits, body_bits = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t total_bits = body_bits + (uint32_t)(log_ml + log_of + log_ll) + 1;
  const uint32_t nbytes = (total_bits + 7) >> 3;
  if (nbytes > (cap - op >= 3 ? cap - op - 3 : 0)) return 0;
  // zero the words the stream will be OR-ed / stored into (not the bytes before it in the first word)
  uint8_t *sp = dst + op;
  uint32_t *const w32 = reinterpret_cast<uint32_t *>((uintptr_t)sp & ~(uintptr_t)3);
  const uint32_t lead = (uint32_t)((uintptr_t)sp & 3);
  {
    // exactly the words the stream touches: word 0 only from byte `lead` on (earlier bytes hold table descriptions)
    const uint32_t last_word = (lead * 8 + total_bits - 1) >> 5;
    const uint32_t head = (4 - lead) & 3;
    if ((uint32_t)lane < head) sp[lane] = 0;
    uint32_t *z = w32 + (lead ? 1 : 0);
    const uint32_t words = lead ? last_word : last_word + 1;
    for (uint32_t k = lane; k < words; k += 32) z[k] = 0;
  }
  __syncwarp();
  // ---- every lane packs its run ----
  {
    uint32_t bitpos = lead * 8 + start_bit;
    uint32_t wi = bitpos >> 5;
    uint64_t acc = 0;
    uint32_t nacc = bitpos & 31;
    bool first = true;
    auto put = [&](uint32_t v, uint32_t k) {
      acc |= (uint64_t)v << nacc;
      nacc += k;
      if (nacc >= 32) {
        if (first) { atomicOr(w32 + wi, (uint32_t)acc); first = false; } else w32[wi] = (uint32_t)acc;
        wi++; acc >>= 32; nacc -= 32;
      }
    };
    for (uint32_t k = k_lo; k < k_hi; k++) {
      const uint32_t i = nseq - 1 - k;
      const uint32_t a = sll[i], b = sml[i], c = sofv[i];
      const uint32_t ll = a & SEQ_VAL_MASK, ml = b & SEQ_VAL_MASK, ofv = c & SEQ_VAL_MASK;
      const uint32_t llc = ll_code(ll), mlc = ml_code(ml), ofc = (uint32_t)hb32(ofv);
      if (k > 0) {
        put((c >> 18) & 0x1FF, c >> 27);
        put((b >> 18) & 0x1FF, b >> 27);
        put((a >> 18) & 0x1FF, a >> 27);
      }
      put(ll - ll_base(llc), ll_xbits(llc));
      put(ml - ml_base(mlc), ml_xbits(mlc));
      put(ofv - (1u << ofc), ofc);
    }
    if (nacc > 0 && k_hi > k_lo) atomicOr(w32 + wi, (uint32_t)acc);
  }
  __syncwarp();
  if (lane == 0) {
    // final states (ML, OF, LL) and the end mark
    uint64_t tail = 0;
    uint32_t nb = 0;
    tail |= (uint64_t)(fin_ml & ((1u << log_ml) - 1u)) << nb; nb += (uint32_t)log_ml;
    tail |= (uint64_t)(fin_of & ((1u << log_of) - 1u)) << nb; nb += (uint32_t)log_of;
    tail |= (uint64_t)(fin_ll & ((1u << log_ll) - 1u)) << nb; nb += (uint32_t)log_ll;
    tail |= 1ull << nb; nb += 1;
    const uint32_t bitpos = lead * 8 + body_bits;
    const uint32_t sh = bitpos & 31;
    uint32_t wi = bitpos >> 5;
    // up to 28 + 31 bits: at most two words
    const uint64_t lo = tail << sh;
    atomicOr(w32 + wi, (uint32_t)lo);
    if ((lo >> 32) != 0) atomicOr(w32 + wi + 1, (uint32_t)(lo >> 32));
  }
  __syncwarp();
  return op + nbytes;
}

} // namespace b200zstd
