// tools/esd_experiment.cpp -- CPU-side ratio experiments for the decoupled (hash / verify / select) parse of
// zstd_encode_esd.cu.  TEST/DEV TOOL ONLY: links the oracle's generators and libzstd to compare sizes.
//   g++ -O2 -std=c++17 tools/esd_experiment.cpp -o /tmp/esd_exp -L oracle -l:liboracle_zstd.so -l:libzstd.so.1 -Wl,-rpath,$PWD/oracle
#include "../custom-nvcomp-with-zstd_b200/csrc/zstd_encode_core.cuh"
#include "../custom-nvcomp-with-zstd_b200/csrc/zstd_encode_params.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace b200zstd;
using namespace b200zstd::enc;

extern "C" {
void orc_gen_batch(uint8_t *dst, size_t U, uint64_t first_idx, size_t n, int kind, uint32_t P);
void orc_gen_textlike(uint8_t *dst, size_t size);
size_t ZSTD_compress(void *dst, size_t cap, const void *src, size_t n, int level);
size_t ZSTD_decompress(void *dst, size_t cap, const void *src, size_t n);
size_t ZSTD_compressBound(size_t n);
unsigned ZSTD_isError(size_t c);
}

static int envi(const char *k, int d) { const char *v = getenv(k); return v ? atoi(v) : d; }

struct Knobs {
  int hash_log, long_log, hash_bytes, min_match;
  int intra;        // intra-window candidates (match_any)
  int lcap;         // verify-stage extension cap
  int long_stride;  // insert every k-th position into the long table
  int rep_bonus;    // rep wins if replen + bonus >= table len
  int lazy;         // 1: lane f+1 may replace f
  int rep_next;     // 1: rep at f+1 beats a table match at f
  int short_stride;
  int runskip;
};

static inline uint64_t read64(const uint8_t *b, uint32_t pos, uint32_t n) {
  uint64_t v = 0;
  for (int k = 0; k < 8; k++) { uint32_t q = pos + k; v |= (uint64_t)(q < n ? b[q] : 0) << (8 * k); }
  return v;
}
static inline uint32_t common8(uint64_t a, uint64_t b) { uint64_t x = a ^ b; return x ? (uint32_t)(__builtin_ctzll(x) >> 3) : 8u; }

struct BlockOut { std::vector<uint8_t> lits; std::vector<uint32_t> ll, ml, ofv; };

static void parse_esd(const uint8_t *b, uint32_t bn, const Knobs &K, uint32_t rep[3], BlockOut &out, long *stats) {
  const uint32_t ilimit = bn > 8 ? bn - 8 : 0;
  std::vector<uint32_t> Roff(bn + 64, 0), Rlen(bn + 64, 0);
  // ---- H + V stages ----
  std::vector<int32_t> tab1((size_t)1 << K.hash_log, -1), tab2(K.long_log ? (size_t)1 << K.long_log : 0, -1);
  uint32_t prev_h1 = 0xFFFFFFFFu, prev_h2 = 0xFFFFFFFFu;
  for (uint32_t p0 = 0; p0 < ilimit; p0 += 32) {
    uint32_t h1[32], h2[32];
    int32_t c1[32], c2[32];
    bool act[32];
    bool ins1[32], ins2[32];
    for (int l = 0; l < 32; l++) {
      uint32_t p = p0 + l;
      act[l] = p < ilimit;
      ins1[l] = ins2[l] = false;
      if (!act[l]) continue;
      uint64_t v = read64(b, p, bn);
      h1[l] = hash_short(v, K.hash_bytes, K.hash_log);
      c1[l] = tab1[h1[l]];
      if (K.long_log) { h2[l] = hash_long(v, K.long_log); c2[l] = tab2[h2[l]]; } else c2[l] = -1;
      ins1[l] = !(K.runskip && (l > 0 ? h1[l - 1] == h1[l] : prev_h1 == h1[l]));
      ins2[l] = K.long_log && !(K.runskip && (l > 0 ? h2[l - 1] == h2[l] : prev_h2 == h2[l]));
      if (K.intra) {
        for (int m = l - 1; m >= 0; m--) if (ins1[m] && h1[m] == h1[l]) { c1[l] = (int32_t)(p0 + m); break; }
        if (K.long_log) for (int m = l - 1; m >= 0; m--) if (ins2[m] && h2[m] == h2[l]) { c2[l] = (int32_t)(p0 + m); break; }
      }
      if (l == 31) { prev_h1 = h1[l]; prev_h2 = K.long_log ? h2[l] : 0; }
    }
    for (int l = 0; l < 32; l++) if (act[l]) {
      uint32_t p = p0 + l;
      if (ins1[l]) tab1[h1[l]] = (int32_t)p;
      if (ins2[l]) tab2[h2[l]] = (int32_t)p;
    }
    for (int l = 0; l < 32; l++) if (act[l]) {
      uint32_t p = p0 + l;
      uint64_t v = read64(b, p, bn);
      uint32_t off = 0, len = 0;
      if (c2[l] >= 0 && read64(b, (uint32_t)c2[l], bn) == v) { off = p - (uint32_t)c2[l]; len = 8; }
      else if (c1[l] >= 0) { uint32_t c = common8(v, read64(b, (uint32_t)c1[l], bn)); if (c >= (uint32_t)K.min_match) { off = p - (uint32_t)c1[l]; len = c; } }
      if (len == 8) {
        while (len < (uint32_t)K.lcap && p + len < bn) {
          uint32_t c = common8(read64(b, p + len, bn), read64(b, p + len - off, bn));
          uint32_t room = bn - (p + len);
          if (c > room) c = room;
          len += c;
          if (c < 8) break;
        }
        if (len > (uint32_t)K.lcap) len = K.lcap;
      }
      Roff[p] = off; Rlen[p] = len;
    }
  }
  // ---- S stage, lane-serial semantics: visited positions are the one after each match and the next position that has a table candidate ----
  const uint32_t SUB = (uint32_t)envi("SUB", 1024);
  const uint32_t NBMAX = (uint32_t)envi("NBMAX", 8);
  uint32_t anchor = 0;
  auto emit = [&](uint32_t s, uint32_t len, uint32_t off, uint32_t llc, uint32_t *rp3) {
    uint32_t ll = s - anchor;
    out.lits.insert(out.lits.end(), b + anchor, b + s);
    out.ll.push_back(ll); out.ml.push_back(len);
    out.ofv.push_back(offset_to_code(off, llc, rp3));
  };
  for (uint32_t B = 0; B < ilimit; B += SUB) {
    const uint32_t E = std::min(B + SUB, bn), lim = std::min(E, ilimit);
    uint32_t rs[3] = {0, 0, 0};
    if (B == 0) { rs[0] = rep[0]; rs[1] = rep[1]; rs[2] = rep[2]; }
    uint32_t ip = B, rep0 = rs[0], lanchor = B;
    auto tab_at = [&](uint32_t p) { return p < lim && p + 4 <= E && Roff[p] != 0; };
    auto rep_len = [&](uint32_t p, uint32_t cap) {      // equal bytes at p vs p - rep0, up to cap, inside the sub-segment
      if (!rep0 || p < rep0) return 0u;
      uint32_t n = 0;
      while (n < cap && p + n < E && b[p + n] == b[p + n - rep0]) n++;
      return n;
    };
    while (ip < lim) {
      const bool t0 = tab_at(ip);
      const uint32_t rl0 = (ip + 4 <= E) ? rep_len(ip, 8) : 0, rl1 = (ip + 1 < lim && ip + 5 <= E) ? rep_len(ip + 1, 7) : 0;
      const bool r0ok = rl0 >= 4, r1ok = rl1 >= 4;
      if (!t0 && !r0ok && !r1ok) {
        uint32_t q = ip + 1;
        while (q < lim && !tab_at(q)) q++;
        ip = q;
        continue;
      }
      uint32_t f = ip, off, len;
      bool use_rep = false, open = false;
      const uint32_t lt0 = t0 ? Rlen[ip] : 0;
      if (r0ok) { if (!t0 || rl0 == 8 || rl0 + K.rep_bonus >= lt0) { use_rep = true; len = rl0; open = rl0 == 8; } }
      if (!use_rep && r1ok && K.rep_next) { if (!t0 || rl1 == 7 || rl1 + K.rep_bonus >= lt0) { use_rep = true; f = ip + 1; len = rl1; open = rl1 == 7; } }
      if (use_rep) { off = rep0; stats[0]++; }
      else {
        if (K.lazy && tab_at(ip + 1) && Rlen[ip + 1] > lt0) f = ip + 1;
        off = Roff[f]; len = Rlen[f]; open = len == (uint32_t)K.lcap; stats[1]++;
      }
      uint32_t s = f;
      if (open) { while (s + len < E && b[s + len] == b[s + len - off]) len++; stats[2]++; }
      if (s + len > E) len = E - s;
      uint32_t nb = 0;
      while (nb < NBMAX && s - nb > lanchor && s - nb - 1 >= off && b[s - nb - 1] == b[s - nb - 1 - off]) nb++;
      s -= nb; len += nb;
      emit(s, len, off, s - lanchor, rs);
      ip = lanchor = anchor = s + len; rep0 = off;
    }
    if (B == 0 && SUB >= bn) { rep[0] = rs[0]; rep[1] = rs[1]; rep[2] = rs[2]; }
  }
  out.lits.insert(out.lits.end(), b + anchor, b + bn);
}

static size_t compress_esd(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, const Knobs &K, long *stats) {
  size_t op = write_frame_header(dst, n, false);
  uint32_t rep[3] = {1, 4, 8};
  static EntropyWs W;
  std::vector<uint8_t> tmp(BLOCK_BYTES + 64);
  uint32_t bn = (uint32_t)n;
  bool same = true;
  for (uint32_t i = 1; i < bn && same; i++) same = src[i] == src[0];
  if (same && bn > 1) { write_block_header(dst + op, true, 1, bn); op += 3; dst[op++] = src[0]; return op; }
  BlockOut B;
  parse_esd(src, bn, K, rep, B, stats);
  stats[3] += (long)B.ll.size(); stats[4] += (long)B.lits.size();
  uint32_t payload = encode_block_payload(W, B.lits.data(), (uint32_t)B.lits.size(), B.ll.data(), B.ml.data(), B.ofv.data(), (uint32_t)B.ll.size(), tmp.data(), bn - 1);
  if (payload == 0 || payload >= bn) { write_block_header(dst + op, true, 0, bn); op += 3; memcpy(dst + op, src, bn); op += bn; }
  else { write_block_header(dst + op, true, 2, payload); op += 3; memcpy(dst + op, tmp.data(), payload); op += payload; }
  (void)cap;
  return op;
}

int main(int argc, char **argv) {
  const int level = argc > 1 ? atoi(argv[1]) : 3;
  const size_t chunk = argc > 2 ? (size_t)atoi(argv[2]) : 65536;
  const int nchunks = argc > 3 ? atoi(argv[3]) : 24;
  Knobs K;
  const bool fast = level <= 2;
  K.hash_log = envi("HLOG", fast ? 13 : 12);
  K.long_log = envi("LLOG", fast ? 0 : 13);
  K.hash_bytes = envi("HBYTES", 5);
  K.min_match = envi("MINM", 5);
  K.intra = envi("INTRA", 1);
  K.lcap = envi("LCAP", 32);
  K.long_stride = envi("LSTRIDE", 1);
  K.short_stride = envi("SSTRIDE", 1);
  K.rep_bonus = envi("REPB", 2);
  K.lazy = envi("LAZY", 0);
  K.rep_next = envi("REPNEXT", 1);
  K.runskip = envi("RUNSKIP", 0);
  struct Cls { const char *name; int kind; uint32_t P; };
  const Cls classes[] = {{"p0", 0, 0}, {"p25", 0, 16384}, {"p50", 0, 32768}, {"p75", 0, 49152}, {"p90", 0, 58982}, {"mixed", 2, 0}, {"text", -1, 0}};
  std::vector<uint8_t> data(chunk * nchunks), out(chunk + 1024), zo(ZSTD_compressBound(chunk)), back(chunk);
  printf("level %d chunk %zu  HLOG %d LLOG %d INTRA %d LCAP %d LSTRIDE %d REPB %d LAZY %d REPNEXT %d\n", level, chunk, K.hash_log, K.long_log, K.intra, K.lcap,
         K.long_stride, K.rep_bonus, K.lazy, K.rep_next);
  for (const Cls &c : classes) {
    if (c.kind >= 0) orc_gen_batch(data.data(), chunk, 0, nchunks, c.kind, c.P);
    else orc_gen_textlike(data.data(), data.size());
    size_t mine = 0, theirs = 0;
    long stats[8] = {0};
    for (int i = 0; i < nchunks; i++) {
      const uint8_t *src = data.data() + (size_t)i * chunk;
      size_t m = compress_esd(src, chunk, out.data(), out.size(), K, stats);
      size_t d = ZSTD_decompress(back.data(), chunk, out.data(), m);
      if (ZSTD_isError(d) || d != chunk || memcmp(back.data(), src, chunk)) { printf("ROUNDTRIP FAILED class %s chunk %d\n", c.name, i); return 1; }
      mine += m;
      theirs += ZSTD_compress(zo.data(), zo.size(), src, chunk, level);
    }
    printf("%-6s mine %8zu  libzstd %8zu  ratio %.4f  | rep %ld tab %ld open %ld  nseq/chunk %ld nlit/chunk %ld\n", c.name, mine, theirs, (double)mine / theirs,
           stats[0] / nchunks, stats[1] / nchunks, stats[2] / nchunks, stats[3] / nchunks, stats[4] / nchunks);
  }
  return 0;
}
