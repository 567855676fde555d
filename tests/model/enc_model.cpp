// tests/model/enc_model.cpp -- TEST INFRASTRUCTURE: a host-side, lane-by-lane model of the sm_100a
// compressor (custom-nvcomp-with-zstd_b200/csrc/zstd_encode.cu).  It shares the entropy-stage
// arithmetic with the kernel through zstd_encode_core.cuh and restates the warp-synchronous parse
// as explicit loops over 32 lanes, so that on the same input it must produce the SAME BYTES as the
// GPU.  Uses: (1) CPU tests prove every frame the algorithm emits decodes in stock libzstd and in
// oracle/zstd_oracle.c without needing a GPU; (2) GPU tests compare kernel output with this model
// bit for bit.  It is never linked into the product library.
#include "../../custom-nvcomp-with-zstd_b200/csrc/zstd_encode_core.cuh"
#include "../../custom-nvcomp-with-zstd_b200/csrc/zstd_encode_params.h"

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
// Levels 1-4, blocks <= 128 KB: the decoupled parse of zstd_encode_esd.cu, stage by stage.
//   hash stage   : fixed windows of 32 positions; every position looks its candidate(s) up in the table state left by
//                  the windows before it, or takes the nearest lower position of its own window that has the same hash
//                  and was inserted; then the window is inserted (a position whose hash equals its predecessor's is
//                  not: runs keep their first position).  uint16 entries, zero = "position 0" (never-written buckets).
//   verify stage : long candidate needs 8 equal bytes, the short one ESD_MIN_MATCH; an 8-byte match is measured up to
//                  ESD_LCAP bytes.
//   select stage : windows of 32 positions from the parse position: repeat-offset matches (>= 4 bytes, byte-exact
//                  inside the window) and measured table matches; first candidate wins (optionally displaced by its
//                  right neighbour), open matches are finished, then extended backwards into pending literals.
// ------------------------------------------------------------------------------------------------
void parse_block_esd(const uint8_t *b, uint32_t bn, const EsdParams &EP, uint32_t rep[3], BlockOut &out) {
  const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
  std::vector<uint32_t> Roff((size_t)bn + 64, 0), Rlen((size_t)bn + 64, 0);
  std::vector<uint16_t> tab1((size_t)1 << EP.hash_log, 0), tab2(EP.dfast ? (size_t)1 << EP.long_log : 0, 0);
  auto recon = [](uint32_t pos, uint16_t e) -> int64_t {
    int64_t c = (int64_t)((pos & ~0xFFFFu) | e);
    if (c >= (int64_t)pos) c -= 0x10000;
    return c;
  };
  uint32_t prev_h1 = 0xFFFFFFFFu, prev_h2 = 0xFFFFFFFFu;
  for (uint32_t p0 = 0; p0 < ilimit; p0 += 32) {
    uint32_t h1[32], h2[32];
    int64_t c1[32], c2[32];
    bool act[32], ins1[32], ins2[32];
    for (int l = 0; l < 32; l++) {
      const uint32_t p = p0 + l;
      act[l] = p < ilimit; ins1[l] = ins2[l] = false; h1[l] = h2[l] = 0; c1[l] = c2[l] = -1;
      if (!act[l]) continue;
      const uint64_t v = read64(b, p, bn);
      h1[l] = hash_short(v, EP.hash_bytes, EP.hash_log);
      c1[l] = recon(p, tab1[h1[l]]);
      ins1[l] = (l > 0 ? h1[l - 1] : prev_h1) != h1[l];
      for (int m = l - 1; m >= 0; m--) if (ins1[m] && h1[m] == h1[l]) { c1[l] = (int64_t)(p0 + m); break; }
      if (EP.dfast) {
        h2[l] = hash_long(v, EP.long_log);
        c2[l] = recon(p, tab2[h2[l]]);
        ins2[l] = (l > 0 ? h2[l - 1] : prev_h2) != h2[l];
        for (int m = l - 1; m >= 0; m--) if (ins2[m] && h2[m] == h2[l]) { c2[l] = (int64_t)(p0 + m); break; }
      }
    }
    // (inactive lanes only exist in the last window; lane 31's hash of a full window feeds the next window's lane 0)
    prev_h1 = h1[31]; prev_h2 = h2[31];
    for (int l = 0; l < 32; l++) if (act[l]) {
      const uint32_t p = p0 + l;
      if (ins1[l]) tab1[h1[l]] = (uint16_t)p;
      if (ins2[l]) tab2[h2[l]] = (uint16_t)p;
    }
    for (int l = 0; l < 32; l++) if (act[l]) {
      const uint32_t p = p0 + l;
      const uint64_t v = read64(b, p, bn);
      uint32_t off = 0, len = 0;
      if (c2[l] >= 0 && read64(b, (uint32_t)c2[l], bn) == v) { off = p - (uint32_t)c2[l]; len = 8; }
      else if (c1[l] >= 0) {
        const uint32_t c = common8(v, read64(b, (uint32_t)c1[l], bn));
        if (c >= ESD_MIN_MATCH) { off = p - (uint32_t)c1[l]; len = c; }
      }
      if (len == 8) {
        while (len < ESD_LCAP && p + len < bn) {
          uint32_t c = common8(read64(b, p + len, bn), read64(b, p + len - off, bn));
          const uint32_t room = bn - (p + len);
          if (c > room) c = room;
          len += c;
          if (c < 8) break;
        }
        if (len > ESD_LCAP) len = ESD_LCAP;
      }
      Roff[p] = off; Rlen[p] = len;
    }
  }
  // select stage: independent sub-segments of ESD_SUB positions
  uint32_t anchor = 0;                                   // end of the last match anywhere in the block
  for (uint32_t B = 0; B < ilimit; B += ESD_SUB) {
    const uint32_t E = std::min(B + ESD_SUB, bn), lim = std::min(E, ilimit);
    uint32_t rs[3] = {0, 0, 0};                          // repeat-offset history as the sub-segment knows it (0 = unknown)
    if (B == 0) { rs[0] = rep[0]; rs[1] = rep[1]; rs[2] = rep[2]; }
    uint32_t ip = B, lanchor = B, rep0 = rs[0];
    while (ip < lim) {
      uint32_t eq = 0, ok = 0, inb = 0;
      for (int l = 0; l < 32; l++) {
        const uint32_t p = ip + l;
        if (rep0 && p >= rep0 && p < E && b[p] == b[p - rep0]) eq |= 1u << l;
        if (p < lim && p + 4 <= E) { inb |= 1u << l; if (Roff[p]) ok |= 1u << l; }
      }
      const uint32_t rp = eq & (eq >> 1) & (eq >> 2) & (eq >> 3) & inb;
      const uint32_t cand = ok | rp;
      if (!cand) { ip += 32; continue; }
      int f = __builtin_ctz(cand);
      auto replen = [&](int j) { const uint32_t m = ~(eq >> j); return m ? (uint32_t)__builtin_ctz(m) : 32u; };
      auto rlen = [&](int j) { return ((inb >> j) & 1) ? Rlen[ip + j] : 0u; };
      bool use_rep = false;
      if ((rp >> f) & 1) {
        const uint32_t rl = replen(f);
        if (!((ok >> f) & 1) || f + rl == 32 || rl + ESD_REP_BONUS >= rlen(f)) use_rep = true;
      } else if (f + 1 < 32 && ((rp >> (f + 1)) & 1)) {
        const uint32_t rl = replen(f + 1);
        if (f + 1 + rl == 32 || rl + ESD_REP_BONUS >= rlen(f)) { f = f + 1; use_rep = true; }
      }
      if (!use_rep && EP.lazy && f + 1 < 32 && ((ok >> (f + 1)) & 1) && rlen(f + 1) > rlen(f)) f = f + 1;
      uint32_t s = ip + f, off, len;
      bool open;
      if (use_rep) { len = replen(f); off = rep0; open = f + len == 32; }
      else { off = Roff[s]; len = Rlen[s]; open = len == ESD_LCAP; }
      if (open) while (s + len < E && b[s + len] == b[s + len - off]) len++;
      if (s + len > E) len = E - s;
      uint32_t nb = 0;
      while (nb < 32 && s - nb > lanchor && s - nb - 1 >= off && b[s - nb - 1] == b[s - nb - 1 - off]) nb++;
      s -= nb; len += nb;
      out.lits.insert(out.lits.end(), b + anchor, b + s);
      out.ll.push_back(s - anchor); out.ml.push_back(len);
      out.ofv.push_back(offset_to_code(off, s - lanchor, rs));     // coded with the literal run the sub-segment sees
      ip = lanchor = anchor = s + len; rep0 = off;
    }
    if (B == 0) { rep[0] = rs[0]; rep[1] = rs[1]; rep[2] = rs[2]; }
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
    if (esd_level(P.level) && n <= BLOCK_BYTES) parse_block_esd(src, bn, esd_params_for_level(P.level), rep, B);
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
