// tests/model/enc_model.cpp -- TEST INFRASTRUCTURE: a host-side, lane-by-lane model of the sm_100a
// compressor (custom-nvcomp-with-zstd_b200/csrc/zstd_encode.cu).  It shares the entropy-stage
// arithmetic with the kernel through zstd_encode_core.cuh and restates the warp-synchronous parse
// as explicit loops over 32 lanes, so that on the same input it must produce the SAME BYTES as the
// GPU.  Uses: (1) CPU tests prove every frame the algorithm emits decodes in stock libzstd and in
// oracle/zstd_oracle.c without needing a GPU; (2) GPU tests compare kernel output with this model
// bit for bit.  It is never linked into the product library.
#include "../../custom-nvcomp-with-zstd_b200/csrc/zstd_encode_core.cuh"
#include "../../custom-nvcomp-with-zstd_b200/csrc/zstd_encode_params.h"
#include "../../custom-nvcomp-with-zstd_b200/csrc/zstd_encode_lz.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace b200zstd;
using namespace b200zstd::enc;

namespace {

struct XX {
  static uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
  static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
  static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
  static uint64_t round(uint64_t a, uint64_t in) { return rotl(a + in * 0xC2B2AE3D27D4EB4FULL, 31) * 0x9E3779B185EBCA87ULL; }
  static uint64_t merge(uint64_t h, uint64_t v) { return (h ^ round(0, v)) * 0x9E3779B185EBCA87ULL + 0x85EBCA77C2B2AE63ULL; }
  static uint64_t hash(const uint8_t *p, size_t len) {
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL,
                   P5 = 0x27D4EB2F165667C5ULL;
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
      uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
      do { v1 = round(v1, rd64(p)); v2 = round(v2, rd64(p + 8)); v3 = round(v3, rd64(p + 16)); v4 = round(v4, rd64(p + 24)); p += 32; } while (p + 32 <= end);
      h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
      h = merge(h, v1); h = merge(h, v2); h = merge(h, v3); h = merge(h, v4);
    } else h = P5;
    h += len;
    while (p + 8 <= end) { h ^= round(0, rd64(p)); h = rotl(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * P5; h = rotl(h, 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
  }
};

// 8 bytes at position pos of the block, clamped exactly like the kernel's word loads: bytes at or
// beyond `n` are whatever the clamped last word holds -- the parser never uses them (ilimit).
inline uint64_t read64(const uint8_t *b, uint32_t pos, uint32_t n) {
  uint64_t v = 0;
  for (int k = 0; k < 8; k++) { uint32_t q = pos + k; v |= (uint64_t)(q < n ? b[q] : 0) << (8 * k); }
  return v;
}
inline uint32_t common8(uint64_t a, uint64_t b) { uint64_t x = a ^ b; return x ? (uint32_t)(__builtin_ctzll(x) >> 3) : 8u; }

struct BlockOut {
  std::vector<uint8_t> lits;
  std::vector<uint32_t> ll, ml, ofv;
};

// ------------------------------------------------------------------------------------------------
// Parse of one block, window of 32 positions per step (see zstd_encode.cu "parse").
// ------------------------------------------------------------------------------------------------
void parse_block(const uint8_t *chunk, uint32_t blk_off, uint32_t bn, const EncodeParams &P, uint32_t rep[3], BlockOut &out,
                 uint32_t max_seq) {
  const uint8_t *b = chunk + blk_off;
  std::vector<uint16_t> tab1((size_t)1 << P.hash_log, 0), tab2(P.long_log ? (size_t)1 << P.long_log : 0, 0);
  std::vector<uint16_t> chain(P.chain_depth > 0 ? bn : 0, 0);
  uint32_t ip = 0, anchor = 0;
  const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
  auto recon = [](uint32_t pos, uint16_t e) -> int64_t {
    int64_t c = (int64_t)((pos & ~0xFFFFu) | e);
    if (c >= (int64_t)pos) c -= 0x10000;
    return c;
  };
  auto emit = [&](uint32_t start, uint32_t len, uint32_t offset) {
    uint32_t llen = start - anchor;
    out.lits.insert(out.lits.end(), b + anchor, b + start);
    out.ll.push_back(llen); out.ml.push_back(len);
    out.ofv.push_back(offset_to_code(offset, llen, rep));
  };
  // full forward length of a match starting at (s, s-off), already known to be >= have bytes
  auto extend = [&](uint32_t s, uint32_t off, uint32_t have) {
    uint32_t len = have;
    while (s + len < bn && b[s + len] == b[(int64_t)s + len - off]) len++;
    return len;
  };
  auto insert_pos = [&](uint32_t pos) {
    if (pos >= ilimit) return;
    uint64_t v = read64(b, pos, bn);
    uint32_t h1 = hash_short(v, P.hash_bytes, P.hash_log);
    if (P.chain_depth > 0) { uint32_t d = (pos - (uint32_t)tab1[h1]) & 0xFFFF; chain[pos] = (uint16_t)d; }
    tab1[h1] = (uint16_t)pos;
    if (P.long_log) tab2[hash_long(v, P.long_log)] = (uint16_t)pos;
  };
  while (ip < ilimit && out.ll.size() < max_seq) {
    uint32_t blen[32], boff[32];
    bool has[32];
    for (int lane = 0; lane < 32; lane++) {
      uint32_t pos = ip + lane;
      has[lane] = false; blen[lane] = 0; boff[lane] = 0;
      if (pos >= ilimit) continue;
      uint64_t v = read64(b, pos, bn);
      uint32_t best = 0, bo = 0;
      if (P.long_log) {
        int64_t c = recon(pos, tab2[hash_long(v, P.long_log)]);
        if (c >= 0) { uint32_t l = common8(v, read64(b, (uint32_t)c, bn)); if (l == 8) { best = 8; bo = pos - (uint32_t)c; } }
      }
      {
        uint32_t h1 = hash_short(v, P.hash_bytes, P.hash_log);
        int64_t c = recon(pos, tab1[h1]);
        int depth = P.chain_depth > 0 ? P.chain_depth : 1;
        while (depth-- > 0 && c >= 0) {
          uint32_t l = common8(v, read64(b, (uint32_t)c, bn));
          if (P.chain_depth > 0 && l == 8) l = extend(pos, pos - (uint32_t)c, 8) > P.lane_cap ? P.lane_cap : extend(pos, pos - (uint32_t)c, 8);
          if (l >= (uint32_t)P.min_match && l > best) { best = l; bo = pos - (uint32_t)c; }
          if (P.chain_depth == 0) break;
          uint16_t d = chain[(uint32_t)c];
          if (d == 0) break;
          c -= d;
        }
      }
      // repeat offset 0 (absolute position inside the chunk decides validity)
      if ((uint64_t)blk_off + pos >= rep[0]) {
        uint32_t l = common8(v, read64(chunk, blk_off + pos - rep[0], blk_off + bn));
        if (P.chain_depth > 0 && l == 8) {
          uint32_t e = 8;
          while (pos + e < bn && e < P.lane_cap && chunk[blk_off + pos + e] == chunk[blk_off + pos + e - rep[0]]) e++;
          l = e;
        }
        if (l >= 4 && l + P.rep_bonus > best) { best = l; bo = rep[0]; }
      }
      if (best >= (uint32_t)P.min_match || (bo == rep[0] && best >= 4)) { has[lane] = true; blen[lane] = best; boff[lane] = bo; }
    }
    int f = -1;
    for (int lane = 0; lane < 32; lane++) if (has[lane]) { f = lane; break; }
    if (f < 0) {
      for (int lane = 0; lane < 32; lane++) insert_pos(ip + lane);
      ip += 32;
      continue;
    }
    // lazy evaluation among the next P.lazy lanes
    uint32_t s = ip + f, off = boff[f];
    uint32_t len = extend(s, off, blen[f] < 8 ? blen[f] : 8);
    for (int step = 1; step <= P.lazy; step++) {
      int g = f + step;
      if (g >= 32 || !has[g]) continue;
      uint32_t s2 = ip + g, off2 = boff[g];
      if (blen[g] < 8 && blen[g] <= len) continue;
      uint32_t len2 = extend(s2, off2, blen[g] < 8 ? blen[g] : 8);
      int gain1 = (int)len * 4 - hb32(off + 1) + 3 * step + (off == rep[0] ? hb32(off + 1) : 0);
      int gain2 = (int)len2 * 4 - hb32(off2 + 1) + (off2 == rep[0] ? hb32(off2 + 1) : 0);
      if (gain2 > gain1) { s = s2; off = off2; len = len2; }
    }
    // backward extension into pending literals
    while (s > anchor && (uint64_t)blk_off + s > off && b[s - 1] == chunk[(uint64_t)blk_off + s - 1 - off]) { s--; len++; }
    // table updates: window positions before the match, then positions inside it
    for (uint32_t p = ip; p < s && p < ip + 32; p++) insert_pos(p);
    emit(s, len, off);
    {
      uint32_t from = s > ip ? s : ip;
      uint32_t end = s + len;
      if (P.insert_all) { for (uint32_t p = from; p < end; p++) insert_pos(p); }
      else { insert_pos(from); if (end >= 2) insert_pos(end - 2); }
    }
    ip = anchor = s + len;
  }
  // trailing literals
  out.lits.insert(out.lits.end(), b + anchor, b + bn);
  (void)ip;
}


// ------------------------------------------------------------------------------------------------
// Levels 1-4, blocks <= 128 KB: the match / select parse of zstd_encode_esd.cu, stage by stage (zstd_encode_lz.cuh
// holds the per-position and per-lane arithmetic, shared with the kernels).
//   match stage  : windows of LZ_WIN positions; a position's candidates are the table state left by the windows before
//                  its own; after the lookups the window is inserted, the highest position winning a bucket.
//   select stage : LZ_LANES lanes walk their sub-segments speculatively, then re-walk from the true entry state until
//                  they meet their speculative walk; repeated until no lane's exit state changes.
// ------------------------------------------------------------------------------------------------
void parse_block_lz(const uint8_t *b, uint32_t bn, const EsdParams &EP_, bool known_history, BlockOut &out) {
  EsdParams EPm = EP_;
  if (getenv("HLOG")) EPm.hash_log = atoi(getenv("HLOG"));
  if (getenv("LLOG") && EPm.dfast) EPm.long_log = atoi(getenv("LLOG"));
  const EsdParams &EP = EPm;
  using namespace b200zstd::lz;
  const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
  std::vector<uint32_t> R((size_t)bn + 8, 0);
  if (EP.rows) {
    // levels 5+: rows of tagged entries for the 8-byte hash, single tagged entries for the 4-byte hash
    std::vector<uint32_t> rows((size_t)LZ_ROW_WAYS << EP.long_log, 0), tabs((size_t)1 << EP.hash_log, 0);
    std::vector<uint32_t> first1((size_t)1 << LZ_FIRST_LOG, 0xFFFFFFFFu), first2((size_t)1 << LZ_FIRST_LOG, 0xFFFFFFFFu);
    auto rd = [&](uint32_t p) { return rd64(b, p); };
    std::vector<uint32_t> h1(LZ_WIN + 1, 0), h2(LZ_WIN + 1, 0);
    std::vector<uint32_t> snap((size_t)LZ_WIN * (LZ_ROW_WAYS + 1));
    for (uint32_t w = 0; w * LZ_WIN < ilimit; w++) {
      const uint32_t w0 = w * LZ_WIN, w1 = std::min(w0 + LZ_WIN, ilimit);
      for (uint32_t p = w0; p < w1; p++) {              // phase 1: state before the window, side tables
        const uint32_t t = p - w0;
        const uint64_t v = rd64(b, p);
        h1[t + 1] = hash_short(v, EP.hash_bytes, EP.hash_log);
        h2[t + 1] = hash_long(v, EP.long_log);
        for (int y = 0; y < LZ_ROW_WAYS; y++) snap[(size_t)t * (LZ_ROW_WAYS + 1) + y] = rows[(size_t)h2[t + 1] * LZ_ROW_WAYS + y];
        snap[(size_t)t * (LZ_ROW_WAYS + 1) + LZ_ROW_WAYS] = tabs[h1[t + 1]];
        if (inserts(p, h1[t + 1], h1[t])) { uint32_t &f = first1[h1[t + 1] >> (EP.hash_log - LZ_FIRST_LOG)]; f = std::min(f, first_key(w, t, h1[t + 1])); }
        if (inserts(p, h2[t + 1], h2[t])) { uint32_t &f = first2[h2[t + 1] >> (EP.long_log - LZ_FIRST_LOG)]; f = std::min(f, first_key(w, t, h2[t + 1])); }
      }
      for (uint32_t p = w0; p < w1; p++) {              // phase 2: inserts, candidates, exact measurement
        const uint32_t t = p - w0;
        const uint64_t v = rd64(b, p);
        const uint32_t tg2 = row_tag(v, p + 12 <= bn ? (uint32_t)(rd64(b, p + 4) >> 32) : 0u), tg1 = short_tag(v);
        if (inserts(p, h2[t + 1], h2[t])) { uint32_t &e = rows[(size_t)h2[t + 1] * LZ_ROW_WAYS + (w & (LZ_ROW_WAYS - 1))]; e = std::max(e, row_entry(p, tg2)); }
        if (inserts(p, h1[t + 1], h1[t])) { uint32_t &e = tabs[h1[t + 1]]; e = std::max(e, row_entry(p, tg1)); }
        uint32_t best_len = 0, best_off = 0;
        const int32_t a2 = first_candidate(first2[h2[t + 1] >> (EP.long_log - LZ_FIRST_LOG)], w, t, h2[t + 1]);
        if (a2 >= 0) try_candidate(rd, p, (uint32_t)a2, bn, best_len, best_off);
        const int32_t a1 = first_candidate(first1[h1[t + 1] >> (EP.hash_log - LZ_FIRST_LOG)], w, t, h1[t + 1]);
        if (a1 >= 0) try_candidate(rd, p, (uint32_t)a1, bn, best_len, best_off);
        int weak = -1;                                   // first way that shares the hashed bytes only
        for (int y = 0; y < LZ_ROW_WAYS; y++) {
          const uint32_t e = snap[(size_t)t * (LZ_ROW_WAYS + 1) + y], c = e >> 15;
          if (c >= p || ((e ^ tg2) & LZ_TAG_HI) != 0) continue;
          if (((e ^ tg2) & LZ_TAG_ALL) == 0) try_candidate(rd, p, c, bn, best_len, best_off);
          else if (weak < 0) weak = y;
        }
        { const uint32_t e = snap[(size_t)t * (LZ_ROW_WAYS + 1) + LZ_ROW_WAYS], c = e >> 15;
          if ((e & 0x7FFFu) == tg1 && c < p) try_candidate(rd, p, c, bn, best_len, best_off); }
        if (weak >= 0 && best_len < 12) try_candidate(rd, p, snap[(size_t)t * (LZ_ROW_WAYS + 1) + weak] >> 15, bn, best_len, best_off);
        R[p] = best_len >= LZ_Q_MIN_MATCH ? (best_off | (best_len << 17)) : 0u;
      }
      h1[0] = h1[w1 - w0]; h2[0] = h2[w1 - w0];
    }
  } else {
    // table entries are positions; a never-written bucket reads as position 0 (a candidate like any other: the bytes decide)
    std::vector<uint32_t> tab1((size_t)1 << EP.hash_log, 0), tab2(EP.dfast ? (size_t)1 << EP.long_log : 0, 0);
    std::vector<uint32_t> first1((size_t)1 << LZ_FIRST_LOG, 0xFFFFFFFFu), first2((size_t)1 << LZ_FIRST_LOG, 0xFFFFFFFFu);
    auto rd = [&](uint32_t p) { return rd64(b, p); };
    std::vector<uint32_t> h1(LZ_WIN + 1, 0), h2(LZ_WIN + 1, 0);      // [0] = hash of the position before the window
    std::vector<int32_t> c1(LZ_WIN), c2(LZ_WIN);
    for (uint32_t w = 0; w * LZ_WIN < ilimit; w++) {
      const uint32_t w0 = w * LZ_WIN, w1 = std::min(w0 + LZ_WIN, ilimit);
      // phase 1: hashes, lookups of the state before the window, first-of-window side table
      for (uint32_t p = w0; p < w1; p++) {
        const uint32_t t = p - w0;
        const uint64_t v = rd64(b, p);
        h1[t + 1] = hash_short(v, EP.hash_bytes, EP.hash_log);
        c1[t] = (int32_t)tab1[h1[t + 1]];
        if (inserts(p, h1[t + 1], h1[t])) { uint32_t &f = first1[h1[t + 1] >> (EP.hash_log - LZ_FIRST_LOG)]; f = std::min(f, first_key(w, t, h1[t + 1])); }
        if (EP.dfast) {
          h2[t + 1] = hash_long(v, EP.long_log);
          c2[t] = (int32_t)tab2[h2[t + 1]];
          if (inserts(p, h2[t + 1], h2[t])) { uint32_t &f = first2[h2[t + 1] >> (EP.long_log - LZ_FIRST_LOG)]; f = std::min(f, first_key(w, t, h2[t + 1])); }
        } else c2[t] = -1;
      }
      // phase 2: inserts (highest position wins), candidates, verification
      for (uint32_t p = w0; p < w1; p++) {
        const uint32_t t = p - w0;
        if (inserts(p, h1[t + 1], h1[t])) tab1[h1[t + 1]] = std::max(tab1[h1[t + 1]], p);
        if (EP.dfast && inserts(p, h2[t + 1], h2[t])) tab2[h2[t + 1]] = std::max(tab2[h2[t + 1]], p);
        int32_t a1 = first_candidate(first1[h1[t + 1] >> (EP.hash_log - LZ_FIRST_LOG)], w, t, h1[t + 1]);
        static int lzv2 = getenv("LZV2") ? atoi(getenv("LZV2")) : 0;
        if (lzv2 == 1) { a1 = -1; for (uint32_t q = t; q-- > 0;) if (h1[q + 1] == h1[t + 1] && inserts(w0 + q, h1[q + 1], h1[q])) { a1 = (int32_t)(w0 + q); break; } }
        if (lzv2 == 2) { a1 = -1; for (uint32_t q = 0; q < t; q++) if (h1[q + 1] == h1[t + 1] && inserts(w0 + q, h1[q + 1], h1[q])) { a1 = (int32_t)(w0 + q); break; } }
        if (lzv2 == 3) { a1 = -1; }
        if (a1 < 0) a1 = c1[t] < (int32_t)p ? c1[t] : -1;
        int32_t a2 = -1;
        if (EP.dfast) {
          a2 = first_candidate(first2[h2[t + 1] >> (EP.long_log - LZ_FIRST_LOG)], w, t, h2[t + 1]);
          if (a2 < 0) a2 = c2[t] < (int32_t)p ? c2[t] : -1;
        }
        R[p] = match_verify(rd, p, rd64(b, p), a2, a1, bn);
      }
      h1[0] = h1[w1 - w0]; h2[0] = h2[w1 - w0];
    }
  }
  SelectParams SP{EP.lazy, EP.rows};
  if (const char *e = getenv("ENC_MODEL_EXACT_STITCH")) SP.exact_stitch = atoi(e);      // cross-check of the stitching rule (tests/test_model.py)
  const uint32_t span = lane_span(ilimit), cap = lane_list_cap(BLOCK_BYTES);
  std::vector<Seq> spec((size_t)LZ_LANES * cap), prefix((size_t)LZ_LANES * cap);
  State spec0[LZ_LANES], spec_exit[LZ_LANES], exit_[LZ_LANES], entry_used[LZ_LANES];
  uint32_t spec_cnt[LZ_LANES], pre_cnt[LZ_LANES], sync_k[LZ_LANES];
  for (uint32_t j = 0; j < LZ_LANES; j++) {
    const uint32_t B = lane_begin(j, span, ilimit), E = lane_begin(j + 1, span, ilimit);
    State st{B, B, 0, 0, 0};
    if (j == 0 && known_history) { st.r0 = 1; st.r1 = 4; st.r2 = 8; }
    spec0[j] = entry_used[j] = st;
    spec_cnt[j] = select_walk(b, bn, R.data(), E, SP, st, &spec[(size_t)j * cap]);
    spec_exit[j] = exit_[j] = st;
    pre_cnt[j] = 0; sync_k[j] = 0;
  }
  for (;;) {
    bool changed = false;
    State entry[LZ_LANES];
    for (uint32_t j = 1; j < LZ_LANES; j++) entry[j] = exit_[j - 1];          // (a shuffle on the GPU: all lanes see the same round)
    for (uint32_t j = 1; j < LZ_LANES; j++) {
      if (entry[j].same(entry_used[j])) continue;
      if (!SP.exact_stitch && entry_differs_in_r2_only(entry[j], entry_used[j])) {
        State ex = exit_[j];
        select_recode(&prefix[(size_t)j * cap], pre_cnt[j], &spec[(size_t)j * cap], spec_cnt[j], sync_k[j], entry_used[j], entry[j], ex);
        entry_used[j] = entry[j];
        if (!ex.same(exit_[j])) { exit_[j] = ex; changed = true; }
        continue;
      }
      entry_used[j] = entry[j];
      const uint32_t E = lane_begin(j + 1, span, ilimit);
      State st = entry[j];
      pre_cnt[j] = select_rewalk(b, bn, R.data(), E, SP, st, &spec[(size_t)j * cap], spec_cnt[j], spec0[j], spec_exit[j], &prefix[(size_t)j * cap], &sync_k[j]);
      if (!st.same(exit_[j])) { exit_[j] = st; changed = true; }
    }
    if (!changed) break;
  }
  uint32_t anchor = 0;
  auto take = [&](const Seq &q) {
    const uint32_t s = seq_start(q), len = seq_len(q);
    out.lits.insert(out.lits.end(), b + anchor, b + s);
    out.ll.push_back(s - anchor); out.ml.push_back(len); out.ofv.push_back(seq_code(q));
    anchor = s + len;
  };
  for (uint32_t j = 0; j < LZ_LANES; j++) {
    for (uint32_t i = 0; i < pre_cnt[j]; i++) take(prefix[(size_t)j * cap + i]);
    for (uint32_t i = sync_k[j]; i < spec_cnt[j]; i++) take(spec[(size_t)j * cap + i]);
  }
  out.lits.insert(out.lits.end(), b + anchor, b + bn);
}

} // namespace

extern "C" {

// Compress one chunk into one frame.  Returns frame bytes, 0 if dst is too small.
size_t model_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level, int checksum) {
  EncodeParams P = encode_params_for_level(level, checksum);
  if (n == 0) return 0;
  const uint32_t fh = frame_header_size(n);
  const size_t nblocks = (n + BLOCK_BYTES - 1) / BLOCK_BYTES;
  if (cap < fh + n + 3 * nblocks + (checksum ? 4 : 0)) return 0;      // the kernel requires room for the stored form
  size_t op = write_frame_header(dst, n, checksum != 0);
  uint32_t rep[3] = {1, 4, 8};
  static thread_local EntropyWs W;
  size_t ip = 0;
  std::vector<uint8_t> tmp(BLOCK_BYTES + 64);
  while (ip < n) {
    uint32_t bn = (uint32_t)std::min<size_t>(BLOCK_BYTES, n - ip);
    const bool last = ip + bn == n;
    // RLE block?
    bool same = true;
    for (uint32_t i = 1; i < bn && same; i++) same = src[ip + i] == src[ip];
    if (same && bn > 1) {
      write_block_header(dst + op, last, 1, bn); op += 3;
      dst[op++] = src[ip];
      ip += bn;
      continue;
    }
    BlockOut B;
    uint32_t rep_save[3] = {rep[0], rep[1], rep[2]};
    if (esd_level(P.level) && n <= BLOCK_BYTES) parse_block_lz(src, bn, esd_params_for_level(P.level, bn > 65536), true, B);
    else parse_block(src, (uint32_t)ip, bn, P, rep, B, MAX_SEQ_PER_BLOCK);
    uint32_t payload = B.ll.size() >= MAX_SEQ_PER_BLOCK ? 0 : encode_block_payload(W, B.lits.data(), (uint32_t)B.lits.size(), B.ll.data(), B.ml.data(), B.ofv.data(), (uint32_t)B.ll.size(), tmp.data(), bn - 1);
    if (payload == 0 || payload >= bn) {
      rep[0] = rep_save[0]; rep[1] = rep_save[1]; rep[2] = rep_save[2];
      write_block_header(dst + op, last, 0, bn); op += 3;
      memcpy(dst + op, src + ip, bn); op += bn;
    } else {
      write_block_header(dst + op, last, 2, payload); op += 3;
      memcpy(dst + op, tmp.data(), payload); op += payload;
    }
    ip += bn;
  }
  if (checksum) { uint32_t c = (uint32_t)XX::hash(src, n); memcpy(dst + op, &c, 4); op += 4; }
  return op;
}

}
