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
// chains on 3 lanes, 32 sequences per round with the codes staged in shared memory, and the interleaved sequence
// bitstream assembled by all lanes: each lane ORs the bits of its own sequence into a shared-memory staging buffer at a
// bit offset known from a prefix sum, completed words leave with coalesced stores.
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

// enc::fse_build_ctable with the whole warp; produces the same tables.  norm[0 .. max_sym] (-1 = low-probability symbol,
// one cell at the top of the table), 2^log cells.  cumul and cell are scratch (as in the serial form).
static __device__ void fse_build_ctable_warp(const int16_t *norm, int max_sym, int log, uint16_t *state_tab, SymTT *tt, uint8_t *cell, uint16_t *cumul,
                                             uint16_t *spread_first /* max_sym + 2 entries of scratch */, int lane) {
  const uint32_t size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
  const uint32_t lt = lanemask_lt();
  // per symbol (two per lane): cells it owns, cells below it in the spread, low-probability symbols below it
  uint32_t run_all = 0, run_spread = 0, run_low = 0;
  uint32_t my_all[2], my_spread[2], my_low[2];
  int my_c[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int s = lane + 32 * r;
    const int c = s <= max_sym ? (int)norm[s] : 0;
    my_c[r] = c;
    const uint32_t all = c == -1 ? 1u : (uint32_t)max(c, 0), spread = (uint32_t)max(c, 0), low = c == -1 ? 1u : 0u;
    uint32_t ia = all, is = spread, il = low;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), ts = __shfl_up_sync(0xffffffffu, is, o), tl = __shfl_up_sync(0xffffffffu, il, o);
      if (lane >= o) { ia += ta; is += ts; il += tl; }
    }
    my_all[r] = run_all + ia - all; my_spread[r] = run_spread + is - spread; my_low[r] = run_low + il - low;
    run_all += __shfl_sync(0xffffffffu, ia, 31); run_spread += __shfl_sync(0xffffffffu, is, 31); run_low += __shfl_sync(0xffffffffu, il, 31);
  }
  const uint32_t nlow = run_low, high = size - 1 - nlow;
  // cumul[] = first slot of every symbol in the state table, spread_first[] = its first occurrence in the spread (a binary
  // search over it finds the owner of an occurrence); the per-symbol transform; the low-probability cells
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int s = lane + 32 * r;
    if (s > max_sym) continue;
    const int c = my_c[r];
    cumul[s] = (uint16_t)my_all[r];
    spread_first[s] = (uint16_t)my_spread[r];
    if (c == -1) cell[size - 1 - my_low[r]] = (uint8_t)s;
    SymTT e;
    if (c == 0) { e.delta_nb = ((log + 1) << 16) - (1 << log); e.delta_state = 0; }
    else if (c == -1 || c == 1) { e.delta_nb = (log << 16) - (1 << log); e.delta_state = (int32_t)my_all[r] - 1; }
    else {
      const int max_bits_out = log - hb32((uint32_t)(c - 1));
      e.delta_nb = (max_bits_out << 16) - (c << max_bits_out);
      e.delta_state = (int32_t)my_all[r] - c;
    }
    tt[s] = e;
  }
  if (lane == 0) spread_first[max_sym + 1] = (uint16_t)run_spread;
  __syncwarp();
  // spread: the i-th position of the walk 0, step, 2 step, ... (mod size) that is not above `high` takes occurrence i
  uint32_t seen = 0;
  for (uint32_t k0 = 0; k0 < size; k0 += 32) {
    const uint32_t pos = ((k0 + (uint32_t)lane) * step) & mask;
    const bool valid = pos <= high;
    const uint32_t vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t i = seen + (uint32_t)__popc(vm & lt);
      int lo = 0, hi = max_sym;                                     // last symbol whose first occurrence is <= i and that owns cells
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (spread_first[mid] <= i) lo = mid; else hi = mid - 1; }
      cell[pos] = (uint8_t)lo;
    }
    seen += (uint32_t)__popc(vm);
  }
  __syncwarp();
  // state table: the cells of a symbol, in cell order, take consecutive slots from cumul[symbol]
  for (uint32_t u0 = 0; u0 < size; u0 += 32) {
    const uint32_t u = u0 + (uint32_t)lane;
    const uint32_t s = cell[u];
    const uint32_t same = __match_any_sync(0xffffffffu, s);
    const uint32_t base = cumul[s];
    __syncwarp();
    state_tab[base + (uint32_t)__popc(same & lt)] = (uint16_t)(size + u);
    if ((same & lt) == 0) cumul[s] = (uint16_t)(base + (uint32_t)__popc(same));
    __syncwarp();
  }
}

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
          {
            // (four literals per load: every lane walks its own piece, and byte loads from 128 pieces per CTA thrash the L1)
            uint32_t i = lo;
            for (; i < hi && ((uintptr_t)(lits + b0 + i) & 3u); i++) bits += W.hufc[lits[b0 + i]] >> 16;
            for (; i + 4 <= hi; i += 4) {
              const uint32_t w = *reinterpret_cast<const uint32_t *>(lits + b0 + i);
              bits += (W.hufc[w & 0xFF] >> 16) + (W.hufc[(w >> 8) & 0xFF] >> 16) + (W.hufc[(w >> 16) & 0xFF] >> 16) + (W.hufc[w >> 24] >> 16);
            }
            for (; i < hi; i++) bits += W.hufc[lits[b0 + i]] >> 16;
          }
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
              auto emit = [&](uint32_t sym) { const uint32_t c = W.hufc[sym]; acc |= (uint64_t)(c & 0xFFFF) << nacc; nacc += c >> 16; };
              auto flush = [&]() {
                if (nacc >= 32) {
                  if (first) { atomicOr(w32 + wi, (uint32_t)acc); first = false; } else w32[wi] = (uint32_t)acc;
                  wi++; acc >>= 32; nacc -= 32;
                }
              };
              // last symbol first; four literals per load, a word check every two symbols (31 + 2 x 11 bits fit the accumulator)
              // (eight literals per load, the next load in flight while these are coded)
              uint32_t i = hi;
              for (; i > lo && ((uintptr_t)(lits + b0 + i) & 7u); i--) { emit(lits[b0 + i - 1]); flush(); }
              if (i >= lo + 8) {
                uint2 cur = *reinterpret_cast<const uint2 *>(lits + b0 + i - 8);
                for (; i >= lo + 8; i -= 8) {
                  uint2 nxt = cur;
                  if (i >= lo + 16) nxt = *reinterpret_cast<const uint2 *>(lits + b0 + i - 16);
                  emit(cur.y >> 24); emit((cur.y >> 16) & 0xFF); flush();
                  emit((cur.y >> 8) & 0xFF); emit(cur.y & 0xFF); flush();
                  emit(cur.x >> 24); emit((cur.x >> 16) & 0xFF); flush();
                  emit((cur.x >> 8) & 0xFF); emit(cur.x & 0xFF); flush();
                  cur = nxt;
                }
              }
              for (; i > lo; i--) { emit(lits[b0 + i - 1]); flush(); }
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
  // one pass over the sequences for the three code histograms (W.count[0..63] LL, [64..127] OF, [128..191] ML)
  for (int i = lane; i < 192; i += 32) W.count[i] = 0;
  __syncwarp();
  uint32_t mx = 0;
  for (uint32_t i = lane; i < nseq; i += 32) {
    const uint32_t c0 = ll_code(sll[i]), c1 = (uint32_t)hb32(sofv[i]), c2 = ml_code(sml[i]);
    atomicAdd(&W.count[c0], 1u); atomicAdd(&W.count[64 + c1], 1u); atomicAdd(&W.count[128 + c2], 1u);
    mx = max(mx & 0xFF, c0) | (max((mx >> 8) & 0xFF, c1) << 8) | (max(mx >> 16, c2) << 16);
  }
  uint32_t mx0 = warp_max_u32(mx & 0xFF), mx1 = warp_max_u32((mx >> 8) & 0xFF), mx2 = warp_max_u32(mx >> 16);
  __syncwarp();
  for (int kind = 0; kind < 3; kind++) {
    const int maxc = (int)(kind == 0 ? mx0 : kind == 1 ? mx1 : mx2);
    int mode = 0;
    uint32_t desc = 0;
    int build_max = -1;
    if (lane == 0) mode = seq_table_prepare(W, kind, W.count + 64 * kind, maxc, nseq, dst + op, cap - op, &desc, &build_max);
    mode = __shfl_sync(0xffffffffu, mode, 0);
    desc = __shfl_sync(0xffffffffu, desc, 0);
    build_max = __shfl_sync(0xffffffffu, build_max, 0);
    __syncwarp();
    if (mode < 0) return 0;
    if (build_max >= 0) fse_build_ctable_warp(W.norm, build_max, W.tab_log[kind], W.state_tab(kind), W.tt[kind], W.cell, W.cumul,
                                              reinterpret_cast<uint16_t *>(W.count + 192), lane);      // W.count[192..255] is free
    __syncwarp();
    op += desc;
    modes |= (uint32_t)mode << (6 - 2 * kind);
  }
  if (lane == 0) dst[modes_pos] = (uint8_t)modes;
  __syncwarp();
  // ---- the interleaved bitstream, 32 sequences per round, last sequence first ----
  // Every lane loads one sequence and computes its codes; the codes are staged in shared memory and the three FSE state
  // chains (lanes 0: LL, 1: OF, 2: ML) walk them, leaving "bits | count << 9" per sequence; then every lane assembles the
  // bits of its own sequence, a prefix sum places them, they are OR-ed into a staging buffer in shared memory and the
  // completed words leave with coalesced stores.  W.count is free by now: [0..95] chain output, [96..127] codes,
  // [128..223] staging.
  const int log_ll = W.tab_log[0], log_of = W.tab_log[1], log_ml = W.tab_log[2];
  uint32_t *const park = W.count, *const codes = W.count + 96, *const stage = W.count + 128;
  // (the table builds are over: their scratch -- cell, norm, cumul -- stages the transforms: LL, OF in cell, ML in norm / cumul)
  uint32_t *const est01 = reinterpret_cast<uint32_t *>(W.cell), *const est2 = reinterpret_cast<uint32_t *>(W.norm);
  uint8_t *const sp = dst + op;
  uint32_t *const w32 = reinterpret_cast<uint32_t *>((uintptr_t)sp & ~(uintptr_t)3);
  const uint32_t lead = (uint32_t)((uintptr_t)sp & 3);
  const uint32_t cap_stream = cap - op >= 3 ? cap - op - 3 : 0;
  const uint32_t word_limit = (lead + cap_stream + 3) >> 2;            // words that may be written (the caller's buffer has this slack)
  uint32_t carry = lead ? (*w32 & ((1u << (8 * lead)) - 1u)) : 0u;     // bytes in front of the stream inside its first word
  uint32_t carry_bits = 8 * lead, wbase = 0;
  unsigned long long total_bits = 0;
  bool overflow = false;
  uint32_t fin = 0;
  {
    const int t = lane < 3 ? lane : 0;
    const int log = W.tab_log[t];
    const uint16_t *st = W.state_tab(t);
    const SymTT *tt = W.tt[t];
    uint32_t state = 0;
    uint32_t na = 0, nb_ = 0, nc = 1;
    if ((uint32_t)lane < nseq) { const uint32_t i = nseq - 1 - (uint32_t)lane; na = sll[i]; nb_ = sml[i]; nc = sofv[i]; }
    for (uint32_t base = 0; base < nseq; base += 32) {
      const uint32_t kk = base + (uint32_t)lane;       // stream order: k = 0 is the last sequence
      const bool valid = kk < nseq;
      const uint32_t a = na, b = nb_, c = nc;
      {
        const uint32_t kn = kk + 32;                   // next round's sequence, in flight during this round
        if (kn < nseq) { const uint32_t i = nseq - 1 - kn; na = sll[i]; nb_ = sml[i]; nc = sofv[i]; }
      }
      const uint32_t llc = ll_code(a), ofc = (uint32_t)hb32(c), mlc = ml_code(valid ? b : 3u);
      // the per-symbol transforms of my three codes: the chain lanes read them from shared memory (so a chain step is the
      // recurrence only: add, shift, shift, add, table load), and I need them again to cut my bits out of the parked states
      const SymTT e_ll = W.tt[0][llc], e_of = W.tt[1][ofc], e_ml = W.tt[2][mlc];
      codes[lane] = llc | (ofc << 8) | (mlc << 16);
      est01[lane] = (uint32_t)e_ll.delta_nb; est01[32 + lane] = (uint32_t)e_ll.delta_state;
      est01[64 + lane] = (uint32_t)e_of.delta_nb; est01[96 + lane] = (uint32_t)e_of.delta_state;
      est2[lane] = (uint32_t)e_ml.delta_nb; est2[32 + lane] = (uint32_t)e_ml.delta_state;
      stage[lane] = 0; stage[32 + lane] = 0; stage[64 + lane] = 0;
      __syncwarp();
      const uint32_t nv = min(32u, nseq - base);
      if (lane < 3 && log) {
        const uint32_t *const dn = t < 2 ? est01 + 64 * t : est2, *const ds = dn + 32;
        uint32_t j = 0;
        if (base == 0) { state = fse_init_state(st, tt, (codes[0] >> (8 * t)) & 0xFF); j = 1; }
        for (; j < nv; j++) {
          const uint32_t nb = (uint32_t)((int32_t)state + (int32_t)dn[j]) >> 16;
          park[t * 32 + j] = state;                                         // the state BEFORE the step: its low nb bits are the output
          state = st[(int32_t)(state >> nb) + (int32_t)ds[j]];
        }
      }
      if (lane == 0) stage[0] = carry;
      __syncwarp();
      // bits of my sequence: [OF state][ML state][LL state] (not for the very first), LL extra, ML extra, OF extra
      uint64_t lo = 0, hi = 0;
      uint32_t nlo = 0, nhi = 0;
      if (valid) {
        if (kk > 0) {
          if (log_of) { const uint32_t v = park[32 + lane], nb = (uint32_t)((int32_t)v + e_of.delta_nb) >> 16; lo |= (uint64_t)(v & ((1u << nb) - 1u)) << nlo; nlo += nb; }
          if (log_ml) { const uint32_t v = park[64 + lane], nb = (uint32_t)((int32_t)v + e_ml.delta_nb) >> 16; lo |= (uint64_t)(v & ((1u << nb) - 1u)) << nlo; nlo += nb; }
          if (log_ll) { const uint32_t v = park[lane], nb = (uint32_t)((int32_t)v + e_ll.delta_nb) >> 16; lo |= (uint64_t)(v & ((1u << nb) - 1u)) << nlo; nlo += nb; }
        }
        lo |= (uint64_t)(a - ll_base(llc)) << nlo; nlo += ll_xbits(llc);
        hi = (uint64_t)(b - ml_base(mlc)); nhi = ml_xbits(mlc);
        hi |= (uint64_t)(c - (1u << ofc)) << nhi; nhi += ofc;
      }
      const uint32_t mine = nlo + nhi;
      uint32_t incl = mine;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
      const uint32_t round_bits = __shfl_sync(0xffffffffu, incl, 31);
      if (valid) {
        uint32_t pos = carry_bits + incl - mine;
        if (nlo) {
          const uint32_t wi = pos >> 5, sh = pos & 31;
          const uint64_t x = lo << sh;                                    // nlo <= 42: up to 73 bits
          atomicOr(&stage[wi], (uint32_t)x);
          if (sh + nlo > 32) atomicOr(&stage[wi + 1], (uint32_t)(x >> 32));
          if (sh + nlo > 64) atomicOr(&stage[wi + 2], (uint32_t)(lo >> (64 - sh)));
        }
        pos += nlo;
        if (nhi) {
          const uint32_t wi = pos >> 5, sh = pos & 31;
          const uint64_t x = hi << sh;                                    // nhi <= 33: up to 64 bits
          atomicOr(&stage[wi], (uint32_t)x);
          if (sh + nhi > 32) atomicOr(&stage[wi + 1], (uint32_t)(x >> 32));
        }
      }
      __syncwarp();
      const uint32_t have = carry_bits + round_bits, full = have >> 5;
      if (wbase + full > word_limit) overflow = true;
      else for (uint32_t j = lane; j < full; j += 32) w32[wbase + j] = stage[j];
      carry = stage[full];
      carry_bits = have & 31;
      wbase += full;
      total_bits += round_bits;
      __syncwarp();
      if (overflow) break;
    }
    if (lane < 3) fin = state;
  }
  const uint32_t fin_ll = __shfl_sync(0xffffffffu, fin, 0), fin_of = __shfl_sync(0xffffffffu, fin, 1), fin_ml = __shfl_sync(0xffffffffu, fin, 2);
  if (overflow) return 0;
  const uint32_t tail_bits = (uint32_t)(log_ml + log_of + log_ll) + 1;
  const unsigned long long all_bits = total_bits + tail_bits;
  const uint32_t nbytes = (uint32_t)((all_bits + 7) >> 3);
  if (nbytes > cap_stream) return 0;
  if (lane == 0) {
    // final states (ML, OF, LL) and the end mark behind the carried partial word
    uint64_t tail = 0;
    uint32_t nb = 0;
    tail |= (uint64_t)(fin_ml & ((1u << log_ml) - 1u)) << nb; nb += (uint32_t)log_ml;
    tail |= (uint64_t)(fin_of & ((1u << log_of) - 1u)) << nb; nb += (uint32_t)log_of;
    tail |= (uint64_t)(fin_ll & ((1u << log_ll) - 1u)) << nb; nb += (uint32_t)log_ll;
    tail |= 1ull << nb; nb += 1;
    // carry_bits + nb <= 31 + 28 bits: the bytes that hold them are written one by one (nothing behind the stream is touched)
    const uint64_t v = (uint64_t)carry | (tail << carry_bits);
    const uint32_t end_byte = lead + nbytes;                           // offset of the stream's end from w32's first byte
    uint8_t *const wb = reinterpret_cast<uint8_t *>(w32 + wbase);
    for (uint32_t j = (wbase ? 0u : lead); wbase * 4 + j < end_byte; j++) wb[j] = (uint8_t)(v >> (8 * j));
  }
  __syncwarp();
  return op + nbytes;
}

} // namespace b200zstd
