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
// Huffman streams on 4 lanes at offsets known from a bit-count pre-pass, the three FSE state chains on
// 3 lanes (state bits parked in the spare high bits of the sequence arrays), and the interleaved
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
    for (uint32_t i = lane; i < nlit; i += 32) atomicAdd(&W.count[lits[i]], 1u);
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
          // bit count of every stream -> byte sizes and offsets before any stream is written
          uint32_t sz[4] = {0, 0, 0, 0};
          for (uint32_t k = 0; k < nstreams; k++) {
            const uint32_t b0 = single ? 0 : k * seg, cnt = single ? nlit : (k < 3 ? seg : nlit - 3 * seg);
            uint32_t bits = 0;
            for (uint32_t i = lane; i < cnt; i += 32) bits += W.hufc[lits[b0 + i]] >> 16;
            sz[k] = (warp_sum_u32(bits) + 1 + 7) >> 3;           // + end mark, rounded up to bytes
          }
          uint32_t off[4] = {0, 0, 0, 0};
          if (single) {
            if (sz[0] > budget - used) ok = false; else { off[0] = used; used += sz[0]; }
          } else {
            if (used + 6 > budget) ok = false;
            else {
              const uint32_t jt = used;
              used += 6;
              for (int k = 0; k < 4 && ok; k++) {
                if (sz[k] > budget - used || sz[k] > 0xFFFF) { ok = false; break; }
                off[k] = used;
                used += sz[k];
              }
              if (ok && lane == 0)
                for (int k = 0; k < 3; k++) { body[jt + 2 * k] = (uint8_t)sz[k]; body[jt + 2 * k + 1] = (uint8_t)(sz[k] >> 8); }
            }
          }
          if (ok && (uint32_t)lane < nstreams) {
            const uint32_t k = (uint32_t)lane;
            const uint32_t b0 = single ? 0 : k * seg, cnt = single ? nlit : (k < 3 ? seg : nlit - 3 * seg);
            const uint32_t s0 = k == 0 ? sz[0] : k == 1 ? sz[1] : k == 2 ? sz[2] : sz[3];
            const uint32_t o0 = k == 0 ? off[0] : k == 1 ? off[1] : k == 2 ? off[2] : off[3];
            huf_encode_stream(lits + b0, cnt, W.hufc, body + o0, s0);
          }
          __syncwarp();
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
  uint32_t fin = 0;
  if (lane < 3) {
    const int t = lane;
    const int log = W.tab_log[t];
    const uint16_t *st = W.state_tab(t);
    const SymTT *tt = W.tt[t];
    uint32_t *arr = t == 0 ? sll : t == 1 ? sofv : sml;
    uint32_t state = 0;
    if (log) {
      for (uint32_t i = nseq; i-- > 0;) {
        const uint32_t v = arr[i];
        const uint32_t code = t == 0 ? ll_code(v) : t == 1 ? (uint32_t)hb32(v) : ml_code(v);
        if (i == nseq - 1) state = fse_init_state(st, tt, code);
        else {
          const uint32_t nb = (uint32_t)((int32_t)state + tt[code].delta_nb) >> 16;
          arr[i] = v | ((state & ((1u << nb) - 1u)) << 18) | (nb << 27);
          state = st[(int32_t)(state >> nb) + tt[code].delta_state];
        }
      }
    }
    fin = state;
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
