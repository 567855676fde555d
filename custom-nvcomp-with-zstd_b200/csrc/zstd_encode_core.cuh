// zstd_encode_core.cuh -- single-thread building blocks of the Zstandard encoder back end, written
// as __host__ __device__ code so that the exact same arithmetic runs in the sm_100a kernel (one lane
// executes these; the warp-parallel parts live in zstd_encode.cu) and in the host-side model used
// by the CPU tests (tests/model/enc_model.cpp).  None of this is a CPU fallback: the product library
// only ever calls these from device code.
//
// Format: RFC 8878.  Reference counterparts (behavioural spec only): FSE CTable build
// src/cuda_zstd_fse_encoding_kernel.cu:199-325, interleaved sequence encode order :71-178, sequence
// code tables include/cuda_zstd_internal.h:235-449, literals header src/cuda_zstd_manager.cu:4437-4457,
// sequences header :4511-4523, frame/block headers :3998-4106, 4227-4286.  Unlike the reference this
// encoder emits Huffman-compressed literals and FSE-compressed (not only predefined) sequence tables,
// and it uses repeat-offset codes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZHD __host__ __device__ __forceinline__
#define ZHDN __host__ __device__ inline
#else
#define ZHD inline
#define ZHDN inline
#endif

namespace b200zstd {
namespace enc {

ZHD int hb32(uint32_t v) {          // floor(log2(v)), v != 0
#if defined(__CUDA_ARCH__)
  return 31 - __clz(v);
#else
  return 31 - __builtin_clz(v);
#endif
}

// ---- sequence code mapping -----------------------------------------------------------------------
ZHD uint32_t ll_code(uint32_t ll) {
  if (ll < 16) return ll;
  if (ll < 64) {                    // 16..63: table region
    // codes 16..24: bases 16,18,20,22,24,28,32,40,48
    if (ll < 24) return 16 + ((ll - 16) >> 1);
    if (ll < 32) return 20 + ((ll - 24) >> 2);
    if (ll < 48) return 22 + ((ll - 32) >> 3);
    return 24;
  }
  return (uint32_t)hb32(ll) + 19;   // 64 -> 25, 128 -> 26, ...
}
ZHD uint32_t ml_code(uint32_t ml) {   // ml >= 3
  uint32_t m = ml - 3;
  if (m < 32) return m;
  if (m < 128) {
    // bases (ml): 35,37,39,41 (1 bit) | 43,47 (2) | 51,59 (3) | 67,83 (4) | 99 (5)
    if (ml < 43) return 32 + ((ml - 35) >> 1);
    if (ml < 51) return 36 + ((ml - 43) >> 2);
    if (ml < 67) return 38 + ((ml - 51) >> 3);
    if (ml < 99) return 40 + ((ml - 67) >> 4);
    return 42;
  }
  return (uint32_t)hb32(m) + 36;    // ml-3 in [128,256) -> 43 ...
}

struct SymTT { int32_t delta_nb; int32_t delta_state; };

// base value and number of extra bits of a length code, computed (no table lookups in the hot loops)
ZHD uint32_t ll_xbits(uint32_t c) { return c < 16 ? 0u : c < 20 ? 1u : c < 22 ? 2u : c < 24 ? 3u : c == 24 ? 4u : c - 19; }
ZHD uint32_t ll_base(uint32_t c) {
  return c < 16 ? c : c < 20 ? 16 + 2 * (c - 16) : c < 22 ? 24 + 4 * (c - 20) : c < 24 ? 32 + 8 * (c - 22) : c == 24 ? 48u : 1u << (c - 19);
}
ZHD uint32_t ml_xbits(uint32_t c) { return c < 32 ? 0u : c < 36 ? 1u : c < 38 ? 2u : c < 40 ? 3u : c < 42 ? 4u : c == 42 ? 5u : c - 36; }
ZHD uint32_t ml_base(uint32_t c) {
  return c < 32 ? c + 3 : c < 36 ? 35 + 2 * (c - 32) : c < 38 ? 43 + 4 * (c - 36) : c < 40 ? 51 + 8 * (c - 38)
                                                   : c < 42 ? 67 + 16 * (c - 40) : c == 42 ? 99u : (1u << (c - 36)) + 3;
}

ZHD void default_norm(int kind, int16_t *norm, int *max_sym, int *log) {
  const int16_t ll[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
  const int16_t ml[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
  const int16_t of[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
  if (kind == 0) { for (int i = 0; i < 36; i++) norm[i] = ll[i]; *max_sym = 35; *log = 6; }
  else if (kind == 1) { for (int i = 0; i < 29; i++) norm[i] = of[i]; *max_sym = 28; *log = 5; }
  else { for (int i = 0; i < 53; i++) norm[i] = ml[i]; *max_sym = 52; *log = 6; }
}

// ---- forward bit writer (little-endian, LSB first) -------------------------------------------------
struct BitW {
  uint8_t *p;        // next byte to write
  uint8_t *end;      // capacity limit
  uint64_t acc;
  int n;             // bits in acc (< 32 after flush)
  bool ovf;
  ZHD void init(uint8_t *dst, uint8_t *lim) { p = dst; end = lim; acc = 0; n = 0; ovf = false; }
  ZHD void add(uint32_t v, int k) { acc |= (uint64_t)(v & (k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u))) << n; n += k; }   // k <= 32, n+k <= 64
  ZHD void flush() {
    while (n >= 8) {
      if (p < end) *p++ = (uint8_t)acc; else ovf = true;
      acc >>= 8; n -= 8;
    }
  }
  // end mark + final partial byte; returns total bytes or 0 on overflow
  ZHD uint32_t close(uint8_t *start) {
    add(1, 1);
    flush();
    if (n > 0) { if (p < end) *p++ = (uint8_t)acc; else ovf = true; acc = 0; n = 0; }
    return ovf ? 0u : (uint32_t)(p - start);
  }
};

// ---- FSE: table-log choice, normalisation, header, compression table -------------------------------
ZHD int fse_optimal_log(int max_log, uint32_t total, int max_sym) {
  int max_bits_src = hb32(total - 1 ? total - 1 : 1) - 2;
  int min_bits_src = hb32(total) + 1, min_bits_sym = hb32((uint32_t)max_sym ? (uint32_t)max_sym : 1u) + 2;
  int min_bits = min_bits_src < min_bits_sym ? min_bits_src : min_bits_sym;
  int log = max_log;
  if (max_bits_src < log) log = max_bits_src;
  if (min_bits > log) log = min_bits;
  if (log < 5) log = 5;
  if (log > max_log) log = max_log;
  return log;
}

// Normalises count[0..max_sym] (sum == total, at least two distinct symbols present) to sum 1<<log,
// every present symbol >= 1.  Largest-remainder style: floor share, minimum 1, then the surplus or
// deficit is settled on the most frequent symbols.  Any such table is a valid zstd table; it need
// not equal libzstd's choice.  Returns false if it cannot (more present symbols than slots).
ZHDN bool fse_normalize(const uint32_t *count, int max_sym, uint32_t total, int log, int16_t *norm) {
  const uint32_t size = 1u << log;
  uint32_t sum = 0;
  int present = 0;
  for (int s = 0; s <= max_sym; s++) {
    uint32_t c = count[s];
    if (!c) { norm[s] = 0; continue; }
    present++;
    uint64_t sc = ((uint64_t)c << log);
    uint32_t q = (uint32_t)(sc / total);
    uint32_t r = (uint32_t)(sc % total);
    if (q == 0) q = 1;
    else if (q < 8 && (uint64_t)r * 2 > total) q++;      // round small probabilities to nearest
    norm[s] = (int16_t)q;
    sum += q;
  }
  if ((uint32_t)present > size) return false;
  // settle the difference
  while (sum != size) {
    if (sum < size) {
      // give the whole deficit to the most frequent symbol
      int best = -1;
      for (int s = 0; s <= max_sym; s++) if (count[s] && (best < 0 || count[s] > count[best])) best = s;
      norm[best] = (int16_t)(norm[best] + (int)(size - sum));
      sum = size;
    } else {
      // take from the symbol whose cost increase is smallest: the one with the largest norm (> 1)
      int best = -1;
      for (int s = 0; s <= max_sym; s++) if (norm[s] > 1 && (best < 0 || norm[s] > norm[best])) best = s;
      if (best < 0) return false;
      uint32_t over = sum - size;
      uint32_t can = (uint32_t)(norm[best] - 1);
      uint32_t take = over < can ? over : can;
      // do not take more than a quarter at once from one symbol unless necessary
      uint32_t quarter = (uint32_t)norm[best] >> 2;
      if (take > quarter && quarter > 0) take = quarter;
      norm[best] = (int16_t)(norm[best] - (int)take);
      sum -= take;
    }
  }
  return true;
}

// Writes the normalised-count header; returns bytes written or 0 on overflow.
ZHDN uint32_t fse_write_ncount(uint8_t *dst, uint32_t cap, const int16_t *norm, int max_sym, int log) {
  const int table_size = 1 << log;
  uint64_t bits = (uint64_t)(log - 5);
  int nb = 4;
  int remaining = table_size + 1, threshold = table_size, nbits = log + 1;
  uint32_t out = 0;
  const int alphabet = max_sym + 1;
  int sym = 0;
  bool prev0 = false;
  while (sym < alphabet && remaining > 1) {
    if (prev0) {
      int start = sym;
      while (sym < alphabet && norm[sym] == 0) sym++;
      if (sym == alphabet) break;
      while (sym >= start + 3) { start += 3; bits |= (uint64_t)3 << nb; nb += 2; if (nb >= 16) { if (out + 2 > cap) return 0; dst[out++] = (uint8_t)bits; dst[out++] = (uint8_t)(bits >> 8); bits >>= 16; nb -= 16; } }
      bits |= (uint64_t)(sym - start) << nb; nb += 2;
      if (nb >= 16) { if (out + 2 > cap) return 0; dst[out++] = (uint8_t)bits; dst[out++] = (uint8_t)(bits >> 8); bits >>= 16; nb -= 16; }
    }
    int count = norm[sym++];
    const int mx = (2 * threshold - 1) - remaining;
    remaining -= count < 0 ? -count : count;
    count++;
    if (count >= threshold) count += mx;
    bits |= (uint64_t)(uint32_t)count << nb;
    nb += nbits;
    nb -= (count < mx);
    prev0 = (count == 1);
    if (remaining < 1) return 0;
    while (remaining < threshold) { nbits--; threshold >>= 1; }
    while (nb >= 16) { if (out + 2 > cap) return 0; dst[out++] = (uint8_t)bits; dst[out++] = (uint8_t)(bits >> 8); bits >>= 16; nb -= 16; }
  }
  if (remaining != 1) return 0;
  while (nb > 0) { if (out + 1 > cap) return 0; dst[out++] = (uint8_t)bits; bits >>= 8; nb -= 8; }
  return out;
}

// Compression table (state table + per-symbol transform).  state_tab has 1<<log entries.
// scratch: sym_of_cell[1<<log] bytes, cumul[max_sym+2] uint16.
ZHDN void fse_build_ctable(const int16_t *norm, int max_sym, int log, uint16_t *state_tab, SymTT *tt, uint8_t *sym_of_cell,
                           uint16_t *cumul) {
  const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
  int high = size - 1;
  cumul[0] = 0;
  for (int s = 0; s <= max_sym; s++) {
    if (norm[s] == -1) { cumul[s + 1] = (uint16_t)(cumul[s] + 1); sym_of_cell[high--] = (uint8_t)s; }
    else cumul[s + 1] = (uint16_t)(cumul[s] + norm[s]);
  }
  int pos = 0;
  for (int s = 0; s <= max_sym; s++)
    for (int i = 0; i < norm[s]; i++) {
      sym_of_cell[pos] = (uint8_t)s;
      do { pos = (pos + step) & mask; } while (pos > high);
    }
  for (int u = 0; u < size; u++) { int s = sym_of_cell[u]; state_tab[cumul[s]++] = (uint16_t)(size + u); }
  int total = 0;
  for (int s = 0; s <= max_sym; s++) {
    int c = norm[s];
    if (c == 0) { tt[s].delta_nb = ((log + 1) << 16) - (1 << log); tt[s].delta_state = 0; }
    else if (c == -1 || c == 1) { tt[s].delta_nb = (log << 16) - (1 << log); tt[s].delta_state = total - 1; total++; }
    else {
      int max_bits_out = log - hb32((uint32_t)(c - 1));
      int min_state_plus = c << max_bits_out;
      tt[s].delta_nb = (max_bits_out << 16) - min_state_plus;
      tt[s].delta_state = total - c;
      total += c;
    }
  }
}

// encoder state helpers (FSE_initCState2 / FSE_encodeSymbol / flush semantics)
ZHD uint32_t fse_init_state(const uint16_t *state_tab, const SymTT *tt, uint32_t sym) {
  const int32_t dnb = tt[sym].delta_nb;
  const uint32_t nb_out = (uint32_t)(dnb + (1 << 15)) >> 16;
  const uint32_t v = (nb_out << 16) - (uint32_t)dnb;
  return state_tab[(int32_t)(v >> nb_out) + tt[sym].delta_state];
}

// bit cost (in 1/256 bit) of coding `count` with a table of normalised counts: sum c * -log2(n/size)
ZHD uint32_t frac_log2_cost(uint32_t norm, int log) {
  // -log2(norm / 2^log) in 1/256 bits, piecewise-linear between powers of two
  int h = hb32(norm);
  uint32_t frac = ((norm << 8) >> h) - 256;             // 0..255 position inside [2^h, 2^(h+1))
  return (uint32_t)((log - h) << 8) - frac;
}

// ---- Huffman ---------------------------------------------------------------------------------------
// Length-limited code lengths for count[0..max_sym]: classic two-queue Huffman on symbols sorted by
// count, then a Kraft-sum repair that clamps to max_bits and keeps the code complete.
// work: order[256] (uint16), node arrays sized 512.  Returns max code length used (0 on failure).
ZHDN int huf_build_lengths(const uint32_t *count, int max_sym, int max_bits, uint8_t *len, uint16_t *order, uint32_t *ncount,
                           uint16_t *parent) {
  int n = 0;
  for (int s = 0; s <= max_sym; s++) { len[s] = 0; if (count[s]) order[n++] = (uint16_t)s; }
  if (n < 2) return 0;
  // insertion sort by (count asc, symbol asc): n <= 256, typically <= 64
  for (int i = 1; i < n; i++) {
    uint16_t k = order[i];
    uint32_t ck = count[k];
    int j = i - 1;
    while (j >= 0 && (count[order[j]] > ck)) { order[j + 1] = order[j]; j--; }
    order[j + 1] = k;
  }
  // nodes 0..n-1 leaves (sorted), n.. internal
  for (int i = 0; i < n; i++) ncount[i] = count[order[i]];
  int leaf = 0, inode = n, next = n;
  const int total_nodes = 2 * n - 1;
  while (next < total_nodes) {
    int a, b;
    if (leaf < n && (inode >= next || ncount[leaf] <= ncount[inode])) a = leaf++; else a = inode++;
    if (leaf < n && (inode >= next || ncount[leaf] <= ncount[inode])) b = leaf++; else b = inode++;
    ncount[next] = ncount[a] + ncount[b];
    parent[a] = (uint16_t)next; parent[b] = (uint16_t)next;
    next++;
  }
  // depths: root = total_nodes-1 has depth 0; walk down in node order (parents have larger indices)
  // reuse ncount as depth storage for internal nodes
  ncount[total_nodes - 1] = 0;
  for (int i = total_nodes - 2; i >= n; i--) ncount[i] = ncount[parent[i]] + 1;
  int maxlen = 0;
  for (int i = 0; i < n; i++) {
    int d = (int)ncount[parent[i]] + 1;
    if (d > max_bits) d = max_bits;
    len[order[i]] = (uint8_t)d;
    if (d > maxlen) maxlen = d;
  }
  // Kraft repair in units of 2^-max_bits
  uint32_t kraft = 0;
  for (int i = 0; i < n; i++) kraft += 1u << (max_bits - len[order[i]]);
  const uint32_t full = 1u << max_bits;
  // over-subscribed: lengthen the least frequent symbols that are still shorter than max_bits
  while (kraft > full) {
    int i = 0;
    // order[] is ascending by count: scan from the rarest for a symbol with len < max_bits
    for (i = 0; i < n; i++) if (len[order[i]] < max_bits) break;
    if (i == n) return 0;
    // prefer the longest such code among the rare ones (smallest Kraft step) when the excess is small
    int pick = i;
    uint32_t excess = kraft - full;
    for (int j = i; j < n; j++) {
      int lj = len[order[j]];
      if (lj < max_bits && (1u << (max_bits - lj - 1)) <= excess && lj > len[order[pick]]) { pick = j; }
      if (j - i > 16) break;
    }
    kraft -= 1u << (max_bits - len[order[pick]] - 1);
    len[order[pick]]++;
  }
  // under-subscribed: shorten the most frequent symbols whose step fits
  while (kraft < full) {
    uint32_t room = full - kraft;
    int pick = -1;
    for (int i = n - 1; i >= 0; i--) {
      int l = len[order[i]];
      if (l > 1 && (1u << (max_bits - l)) <= room) { pick = i; break; }
    }
    if (pick < 0) return 0;
    kraft += 1u << (max_bits - len[order[pick]]);
    len[order[pick]]--;
  }
  maxlen = 0;
  for (int i = 0; i < n; i++) if (len[order[i]] > maxlen) maxlen = len[order[i]];
  return maxlen;
}

// Canonical code values as libzstd's decoder expects them (RFC 8878 4.2.1.3): symbols sorted by
// (length desc == weight asc, symbol asc) take consecutive values starting from 0 at the longest
// length; value of the first code of length L-1 = (next value after length L) >> 1.
ZHDN void huf_assign_codes(const uint8_t *len, int max_sym, int table_log, uint16_t *code) {
  uint32_t nb_per_len[16];
  uint32_t start[16];
  for (int i = 0; i < 16; i++) nb_per_len[i] = 0;
  for (int s = 0; s <= max_sym; s++) nb_per_len[len[s]]++;
  uint32_t v = 0;
  for (int l = table_log; l >= 1; l--) { start[l] = v; v = (v + nb_per_len[l]) >> 1; }
  for (int s = 0; s <= max_sym; s++) if (len[s]) code[s] = (uint16_t)start[len[s]]++;
}


// ---- entropy-stage workspace (lives in shared memory in the kernel) --------------------------------
struct EntropyWs {
  // 6.0 KB, laid out by lifetime so that 32 one-warp CTAs fit an SM.  Three phases share the storage:
  //   tree build   : count, ncount, parent, order -> huflen
  //   table write  : huflen, weights, hufc + the FSE set of the weights (st_ll, tt[0], cell, norm, cumul)
  //   sequences    : count[0..63], st_ll/st_of/st_ml, tt, cell, norm, cumul
  uint32_t count[256];        // literal histogram, then per-alphabet sequence histograms
  union {
    struct { uint32_t ncount[512]; uint16_t parent[512]; };            // tree-build node arrays
    struct { uint16_t st_ll[512]; SymTT tt[3][64]; uint8_t cell[512]; };
  };
  union {
    struct { uint32_t hufc[256];                                       // Huffman code | length << 16
             uint8_t huflen[256]; uint8_t weights[256];
             uint16_t order[256]; };                                   // dead before the table is written
    struct { uint16_t st_ml[512];
             uint16_t st_of[256];                                      // offset tables never exceed 2^8 states
             int16_t norm[64]; uint16_t cumul[64]; };
  };
  int tab_log[3];
  ZHD uint16_t *state_tab(int kind) { return kind == 0 ? st_ll : kind == 1 ? st_of : st_ml; }   // LL, OF, ML
};
static_assert(sizeof(EntropyWs) <= 6160, "EntropyWs must stay small enough for 32 CTAs per SM");

// Huffman tree description (RFC 8878 4.2.1): FSE-compressed weights when that is smaller, else
// direct 4-bit weights.  Returns bytes written, 0 when the table cannot be represented.
ZHDN uint32_t huf_write_table(EntropyWs &W, int max_sym, int table_log, uint8_t *dst, uint32_t cap) {
  const int n = max_sym;                                   // weights for symbols 0..max_sym-1; the last is implied
  if (n < 1) return 0;
  uint32_t wcount[16];
  for (int i = 0; i < 16; i++) wcount[i] = 0;
  int max_w = 0;
  for (int s = 0; s < n; s++) {
    uint8_t w = W.huflen[s] ? (uint8_t)(table_log + 1 - W.huflen[s]) : 0;
    W.weights[s] = w;
    wcount[w]++;
    if (w > max_w) max_w = w;
  }
  // try FSE (two interleaved states, accuracy log <= 6)
  uint32_t fse_size = 0;
  if (n >= 2 && cap > 1) {
    uint32_t maxc = 0;
    for (int i = 0; i <= max_w; i++) if (wcount[i] > maxc) maxc = wcount[i];
    if (maxc != (uint32_t)n && maxc > 1) {
      int log = fse_optimal_log(6, (uint32_t)n, max_w);
      int16_t *norm = W.norm;
      if (fse_normalize(wcount, max_w, (uint32_t)n, log, norm)) {
        uint8_t *body = dst + 1;
        const uint32_t body_cap = cap - 1 < 127 ? cap - 1 : 127;
        uint32_t h = fse_write_ncount(body, body_cap, norm, max_w, log);
        if (h) {
          uint16_t *st = W.state_tab(0);
          SymTT *tt = W.tt[0];
          fse_build_ctable(norm, max_w, log, st, tt, W.cell, W.cumul);
          BitW bw;
          bw.init(body + h, body + body_cap);
          uint32_t s1 = 0, s2 = 0;
          bool i1 = false, i2 = false;
          for (int i = n - 1; i >= 0; i--) {
            const uint32_t sym = W.weights[i];
            uint32_t &st_v = (i & 1) ? s2 : s1;
            bool &inited = (i & 1) ? i2 : i1;
            if (!inited) { st_v = fse_init_state(st, tt, sym); inited = true; }
            else {
              const uint32_t nb = (uint32_t)((int32_t)st_v + tt[sym].delta_nb) >> 16;
              bw.add(st_v, (int)nb);
              bw.flush();
              st_v = st[(int32_t)(st_v >> nb) + tt[sym].delta_state];
            }
          }
          bw.add(s2, log); bw.flush();
          bw.add(s1, log); bw.flush();
          uint32_t b = bw.close(body + h);
          if (b) fse_size = h + b;
        }
      }
    }
  }
  if (fse_size > 1 && fse_size < (uint32_t)(n / 2) && fse_size < 128) { dst[0] = (uint8_t)fse_size; return 1 + fse_size; }
  if (n > 128) return 0;
  const uint32_t direct = 1 + (uint32_t)(n + 1) / 2;
  if (direct > cap) return 0;
  dst[0] = (uint8_t)(127 + n);
  for (int i = 0; i < n; i += 2) dst[1 + i / 2] = (uint8_t)((W.weights[i] << 4) | (i + 1 < n ? W.weights[i + 1] : 0));
  return direct;
}

// One Huffman stream: symbols are emitted last-to-first so that the backward reader sees them in order.
ZHDN uint32_t huf_encode_stream(const uint8_t *lit, uint32_t n, const uint32_t *hufc, uint8_t *dst, uint32_t cap) {
  BitW bw;
  bw.init(dst, dst + cap);
  for (uint32_t i = n; i > 0; i--) {
    const uint32_t c = hufc[lit[i - 1]];
    bw.add(c & 0xFFFF, (int)(c >> 16));
    if (bw.n >= 32) bw.flush();
  }
  bw.flush();
  return bw.close(dst);
}

ZHD uint32_t write_lit_header_raw_rle(uint8_t *dst, int type, uint32_t regen) {
  if (regen < 32) { dst[0] = (uint8_t)(type | (regen << 3)); return 1; }
  if (regen < 4096) { uint32_t v = (uint32_t)type | (1u << 2) | (regen << 4); dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); return 2; }
  uint32_t v = (uint32_t)type | (3u << 2) | (regen << 4);
  dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16);
  return 3;
}
ZHD uint32_t lit_header_size_compressed(uint32_t regen) { return 3 + (regen >= 1024) + (regen >= 16384); }
ZHD void write_lit_header_compressed(uint8_t *dst, uint32_t hs, bool single, uint32_t regen, uint32_t comp) {
  if (hs == 3) {
    uint32_t v = 2u | ((single ? 0u : 1u) << 2) | (regen << 4) | (comp << 14);
    dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16);
  } else if (hs == 4) {
    uint32_t v = 2u | (2u << 2) | (regen << 4) | (comp << 18);
    dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24);
  } else {
    uint64_t v = 2u | (3u << 2) | ((uint64_t)regen << 4) | ((uint64_t)comp << 22);
    for (int i = 0; i < 5; i++) dst[i] = (uint8_t)(v >> (8 * i));
  }
}

// Decides the coding mode of one sequence alphabet and prepares its compression table.
// mode: 0 predefined, 1 RLE, 2 FSE-compressed.  Writes the table description (RLE byte or NCount) to
// dst and returns its size through *desc_bytes.  count[] holds the histogram of `nseq` codes.
// build_max_sym == nullptr: the compression table is built here (serial).  Otherwise the normalised counts to build from
// are left in W.norm, *build_max_sym receives the last symbol (-1: nothing to build, RLE) and the caller builds the table
// (the kernels do that with the whole warp: fse_build_ctable_warp).
ZHDN int seq_table_prepare(EntropyWs &W, int kind, const uint32_t *count, int max_code_present, uint32_t nseq, uint8_t *dst,
                           uint32_t cap, uint32_t *desc_bytes, int *build_max_sym = nullptr) {
  if (build_max_sym) *build_max_sym = -1;
  *desc_bytes = 0;
  const int max_log = (kind == 1) ? 8 : 9;
  int16_t dnorm[64];
  int dmax, dlog;
  default_norm(kind, dnorm, &dmax, &dlog);
  // single symbol -> RLE
  uint32_t maxc = 0;
  int only = 0;
  for (int s = 0; s <= max_code_present; s++) if (count[s] > maxc) { maxc = count[s]; only = s; }
  if (maxc == nseq && nseq > 2) {
    if (cap < 1) return -1;
    dst[0] = (uint8_t)only;
    *desc_bytes = 1;
    W.tab_log[kind] = 0;
    W.state_tab(kind)[0] = 0; W.state_tab(kind)[1] = 0;
    for (int s = 0; s < 64; s++) { W.tt[kind][s].delta_nb = 0; W.tt[kind][s].delta_state = 0; }
    return 1;
  }
  // cost of the predefined table (1/256 bit); impossible when a present code has no default slot
  bool def_ok = max_code_present <= dmax;
  uint64_t def_cost = 0;
  if (def_ok)
    for (int s = 0; s <= max_code_present; s++) if (count[s]) {
      int nv = dnorm[s] == -1 ? 1 : dnorm[s];
      if (nv == 0) { def_ok = false; break; }
      def_cost += (uint64_t)count[s] * frac_log2_cost((uint32_t)nv, dlog);
    }
  // candidate compressed table
  bool cmp_ok = false;
  uint64_t cmp_cost = 0;
  int log = 0;
  uint32_t hdr = 0;
  if (nseq >= 8 || !def_ok) {
    log = fse_optimal_log(max_log, nseq, max_code_present);
    if (fse_normalize(count, max_code_present, nseq, log, W.norm)) {
      hdr = fse_write_ncount(dst, cap, W.norm, max_code_present, log);
      if (hdr) {
        cmp_ok = true;
        cmp_cost = (uint64_t)hdr * 8 * 256;
        for (int s = 0; s <= max_code_present; s++) if (count[s]) cmp_cost += (uint64_t)count[s] * frac_log2_cost((uint32_t)W.norm[s], log);
      }
    }
  }
  if (cmp_ok && (!def_ok || cmp_cost < def_cost)) {
    if (build_max_sym) *build_max_sym = max_code_present;
    else fse_build_ctable(W.norm, max_code_present, log, W.state_tab(kind), W.tt[kind], W.cell, W.cumul);
    W.tab_log[kind] = log;
    *desc_bytes = hdr;
    return 2;
  }
  if (!def_ok) return -1;
  if (build_max_sym) { for (int s = 0; s <= dmax; s++) W.norm[s] = dnorm[s]; *build_max_sym = dmax; }
  else fse_build_ctable(dnorm, dmax, dlog, W.state_tab(kind), W.tt[kind], W.cell, W.cumul);
  W.tab_log[kind] = dlog;
  return 0;
}

// Sequence i as stored by the parser: literal length, match length, offset code value
// ("offBase": 1..3 = repeat offsets, otherwise real offset + 3).
struct SeqStore { const uint32_t *ll; const uint32_t *ml; const uint32_t *ofv; };

ZHD uint32_t seq_count_header(uint8_t *dst, uint32_t nseq) {
  if (nseq < 128) { dst[0] = (uint8_t)nseq; return 1; }
  if (nseq < 0x7F00) { dst[0] = (uint8_t)((nseq >> 8) + 0x80); dst[1] = (uint8_t)nseq; return 2; }
  dst[0] = 0xFF; dst[1] = (uint8_t)(nseq - 0x7F00); dst[2] = (uint8_t)((nseq - 0x7F00) >> 8);
  return 3;
}

// Interleaved FSE bitstream of all sequences (ZSTD_encodeSequences order, RFC 8878 3.1.1.3.2.1).
// Returns bytes written or 0 on overflow.
ZHDN uint32_t seq_encode_stream(const EntropyWs &W, const SeqStore &S, uint32_t nseq, uint8_t *dst, uint32_t cap) {
  BitW bw;
  bw.init(dst, dst + cap);
  const uint16_t *st_ll = W.st_ll, *st_of = W.st_of, *st_ml = W.st_ml;
  const SymTT *tt_ll = W.tt[0], *tt_of = W.tt[1], *tt_ml = W.tt[2];
  uint32_t i = nseq - 1;
  uint32_t llc = ll_code(S.ll[i]), mlc = ml_code(S.ml[i]), ofc = (uint32_t)hb32(S.ofv[i]);
  uint32_t s_ml = W.tab_log[2] ? fse_init_state(st_ml, tt_ml, mlc) : 0;
  uint32_t s_of = W.tab_log[1] ? fse_init_state(st_of, tt_of, ofc) : 0;
  uint32_t s_ll = W.tab_log[0] ? fse_init_state(st_ll, tt_ll, llc) : 0;
  bw.add(S.ll[i] - ll_base(llc), (int)ll_xbits(llc));
  bw.add(S.ml[i] - ml_base(mlc), (int)ml_xbits(mlc));
  bw.flush();
  bw.add(S.ofv[i] - (1u << ofc), (int)ofc);
  bw.flush();
  while (i > 0) {
    i--;
    llc = ll_code(S.ll[i]); mlc = ml_code(S.ml[i]); ofc = (uint32_t)hb32(S.ofv[i]);
    if (W.tab_log[1]) { const uint32_t nb = (uint32_t)((int32_t)s_of + tt_of[ofc].delta_nb) >> 16; bw.add(s_of, (int)nb); s_of = st_of[(int32_t)(s_of >> nb) + tt_of[ofc].delta_state]; }
    if (W.tab_log[2]) { const uint32_t nb = (uint32_t)((int32_t)s_ml + tt_ml[mlc].delta_nb) >> 16; bw.add(s_ml, (int)nb); s_ml = st_ml[(int32_t)(s_ml >> nb) + tt_ml[mlc].delta_state]; }
    bw.flush();
    if (W.tab_log[0]) { const uint32_t nb = (uint32_t)((int32_t)s_ll + tt_ll[llc].delta_nb) >> 16; bw.add(s_ll, (int)nb); s_ll = st_ll[(int32_t)(s_ll >> nb) + tt_ll[llc].delta_state]; }
    bw.add(S.ll[i] - ll_base(llc), (int)ll_xbits(llc));
    bw.flush();
    bw.add(S.ml[i] - ml_base(mlc), (int)ml_xbits(mlc));
    bw.flush();
    bw.add(S.ofv[i] - (1u << ofc), (int)ofc);
    bw.flush();
  }
  bw.add(s_ml, W.tab_log[2]); bw.flush();
  bw.add(s_of, W.tab_log[1]); bw.flush();
  bw.add(s_ll, W.tab_log[0]); bw.flush();
  return bw.close(dst);
}

// ---- frame / block headers (reference write_frame_header, src/cuda_zstd_manager.cu:3998-4106) --------
// Always a single-segment frame with the content size, like libzstd's one-shot ZSTD_compress.
ZHD uint32_t write_frame_header(uint8_t *dst, uint64_t content, bool checksum) {
  dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD;
  const int code = content < 256 ? 0 : content < 65536 + 256 ? 1 : content <= 0xFFFFFFFFull ? 2 : 3;
  dst[4] = (uint8_t)((code << 6) | 0x20 | (checksum ? 4 : 0));
  uint32_t n = 5;
  if (code == 0) dst[n++] = (uint8_t)content;
  else if (code == 1) { uint32_t v = (uint32_t)content - 256; dst[n++] = (uint8_t)v; dst[n++] = (uint8_t)(v >> 8); }
  else { const int bytes = code == 2 ? 4 : 8; for (int i = 0; i < bytes; i++) dst[n++] = (uint8_t)(content >> (8 * i)); }
  return n;
}
ZHD uint32_t frame_header_size(uint64_t content) { return 5 + (content < 256 ? 1 : content < 65536 + 256 ? 2 : content <= 0xFFFFFFFFull ? 4 : 8); }
ZHD void write_block_header(uint8_t *dst, bool last, int type, uint32_t size) {
  const uint32_t v = (last ? 1u : 0u) | ((uint32_t)type << 1) | (size << 3);
  dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16);
}

// repeat-offset coding of one sequence (RFC 8878 3.1.2.5); updates rep[] like the decoder will
ZHD uint32_t offset_to_code(uint32_t offset, uint32_t ll, uint32_t *rep) {
  uint32_t code;
  if (ll != 0) {
    if (offset == rep[0]) return 1;
    if (offset == rep[1]) { rep[1] = rep[0]; rep[0] = offset; return 2; }
    if (offset == rep[2]) code = 3; else code = offset + 3;
  } else {
    if (offset == rep[1]) { rep[1] = rep[0]; rep[0] = offset; return 1; }
    if (offset == rep[2]) code = 2;
    else if (rep[0] > 1 && offset == rep[0] - 1) code = 3;
    else code = offset + 3;
  }
  rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = offset;
  return code;
}

// ---- entropy stage of one block: literals section + sequences section --------------------------------
// Returns the block payload size, 0 = does not fit / not worth it (caller stores the block raw).
ZHDN uint32_t encode_block_payload(EntropyWs &W, const uint8_t *lits, uint32_t nlit, const uint32_t *sll, const uint32_t *sml,
                                   const uint32_t *sofv, uint32_t nseq, uint8_t *dst, uint32_t cap) {
  uint32_t op = 0;
  // ---- literals ----
  bool done = false;
  if (nlit >= 64) {
    for (int i = 0; i < 256; i++) W.count[i] = 0;
    for (uint32_t i = 0; i < nlit; i++) W.count[lits[i]]++;
    int max_sym = 255;
    while (max_sym > 0 && !W.count[max_sym]) max_sym--;
    uint32_t maxc = 0;
    for (int s = 0; s <= max_sym; s++) if (W.count[s] > maxc) maxc = W.count[s];
    if (maxc == nlit) {                                    // RLE literals
      if (cap < 4) return 0;
      op = write_lit_header_raw_rle(dst, 1, nlit);
      dst[op++] = lits[0];
      done = true;
    } else if (maxc <= (nlit >> 7) + 4) {
      // nearly flat histogram: not worth a Huffman table (libzstd applies the same early exit)
    } else {
      int tl = huf_build_lengths(W.count, max_sym, 11, W.huflen, W.order, W.ncount, W.parent);
      if (tl > 0) {
        uint16_t *codes = W.order;                           // free again after the length build
        huf_assign_codes(W.huflen, max_sym, tl, codes);
        for (int s = 0; s <= max_sym; s++) W.hufc[s] = W.huflen[s] ? ((uint32_t)codes[s] | ((uint32_t)W.huflen[s] << 16)) : 0u;
        const uint32_t hs = lit_header_size_compressed(nlit);
        const bool single = nlit < 256;
        uint32_t budget = nlit - ((nlit >> 6) + 2);       // must beat raw by libzstd's minimum gain
        if (budget + hs > cap) budget = cap > hs ? cap - hs : 0;
        uint8_t *body = dst + hs;
        uint32_t t = huf_write_table(W, max_sym, tl, body, budget);
        bool ok = t != 0;
        uint32_t used = t;
        if (ok) {
          if (single) {
            uint32_t n = huf_encode_stream(lits, nlit, W.hufc, body + used, budget - used);
            if (!n) ok = false; else used += n;
          } else {
            if (used + 6 > budget) ok = false;
            else {
              const uint32_t seg = (nlit + 3) / 4;
              uint32_t jt = used;
              used += 6;
              for (int k = 0; k < 4 && ok; k++) {
                const uint32_t cnt = k < 3 ? seg : nlit - 3 * seg;
                uint32_t n = huf_encode_stream(lits + k * seg, cnt, W.hufc, body + used, budget - used);
                if (!n || n > 0xFFFF) { ok = false; break; }
                if (k < 3) { body[jt + 2 * k] = (uint8_t)n; body[jt + 2 * k + 1] = (uint8_t)(n >> 8); }
                used += n;
              }
            }
          }
        }
        if (ok && used < budget) {
          write_lit_header_compressed(dst, hs, single, nlit, used);
          op = hs + used;
          done = true;
        }
      }
    }
  }
  if (!done) {                                             // raw literals
    if (cap < nlit + 3) return 0;
    op = write_lit_header_raw_rle(dst, 0, nlit);
    for (uint32_t i = 0; i < nlit; i++) dst[op + i] = lits[i];
    op += nlit;
  }
  // ---- sequences ----
  if (op + 4 > cap) return 0;
  op += seq_count_header(dst + op, nseq);
  if (nseq == 0) return op;
  uint8_t *modes = dst + op++;
  int mode[3];
  for (int kind = 0; kind < 3; kind++) {
    for (int i = 0; i < 64; i++) W.count[i] = 0;
    int maxc = 0;
    for (uint32_t i = 0; i < nseq; i++) {
      uint32_t c = kind == 0 ? ll_code(sll[i]) : kind == 1 ? (uint32_t)hb32(sofv[i]) : ml_code(sml[i]);
      W.count[c]++;
      if ((int)c > maxc) maxc = (int)c;
    }
    uint32_t desc = 0;
    mode[kind] = seq_table_prepare(W, kind, W.count, maxc, nseq, dst + op, cap - op, &desc);
    if (mode[kind] < 0) return 0;
    op += desc;
  }
  *modes = (uint8_t)((mode[0] << 6) | (mode[1] << 4) | (mode[2] << 2));
  SeqStore S{sll, sml, sofv};
  // three spare bytes: the kernel assembles this stream with 32-bit word writes (zstd_encode.cu)
  uint32_t n = seq_encode_stream(W, S, nseq, dst + op, cap - op >= 3 ? cap - op - 3 : 0);
  if (!n) return 0;
  return op + n;
}


} // namespace enc
} // namespace b200zstd
