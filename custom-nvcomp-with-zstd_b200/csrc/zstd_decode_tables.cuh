// zstd_decode_tables.cuh -- warp-cooperative construction of the decoder's entropy tables
// (FSE sequence tables, Huffman literal table).  Shared by the general decoder (zstd_decode.cu)
// and the batch fast path (zstd_decode_fast.cu).
//
// Reference counterparts (behavioural spec): FSE_buildDTable_Host src/cuda_zstd_fse.cu:1260-1389,
// read_fse_header :4712-4885, Huffman weights src/cuda_zstd_huffman.cu:233-257 (direct), :270-780
// (FSE-compressed), DTable order :1301-1330.  RFC 8878 is followed where those deviate.
#pragma once
#include "zstd_common.cuh"

namespace b200zstd {

// ---------------------------------------------------------------------------------------------
// FSE decode table (packed 8-byte entries):  x = nextStateBase | nbBits << 16 | extraBits << 24, y = baseValue
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 seq_entry(int kind, uint32_t sym, uint32_t next_base, uint32_t nb) {
  uint32_t base, bits;
  if (kind == 0) { base = c_ll_base[sym]; bits = c_ll_bits[sym]; }
  else if (kind == 1) { base = 1u << sym; bits = sym; }
  else { base = c_ml_base[sym]; bits = c_ml_bits[sym]; }
  return make_uint2(next_base | (nb << 16) | (bits << 24), base);
}

// Warp-cooperative table build from normalised counts (RFC 8878 4.1.1).  All 32 lanes call.
// The spread walk pos -> (pos + step) & mask visits every cell exactly once (step is odd), so cell
// j*step & mask receives the r-th symbol instance where r = number of earlier cells that are not in
// the low-probability area: a ballot prefix replaces the serial walk.  The per-symbol state
// numbering (ascending cell index) uses match_any to rank equal symbols inside a 32-cell stripe.
__device__ inline void fse_build_warp(uint2 *tab, const int16_t *norm, int max_sym, int log, int kind, uint8_t *item_sym,
                               uint16_t *sym_next, int lane) {
  const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
  int high = size - 1, acc = 0;
  for (int s = 0; s <= max_sym; s++) {         // uniform loop, <= 53 trips
    int c = norm[s];
    if (c == -1) { if (lane == 0) { tab[high].y = (uint32_t)s; sym_next[s] = 1; } high--; }
    else {
      if (lane == 0) sym_next[s] = (uint16_t)c;
      for (int k = lane; k < c; k += 32) item_sym[acc + k] = (uint8_t)s;
      acc += c;
    }
  }
  __syncwarp();
  int run = 0;
  for (int j0 = 0; j0 < size; j0 += 32) {
    int pos = ((j0 + lane) * step) & mask;
    bool ok = pos <= high;
    uint32_t b = __ballot_sync(0xffffffffu, ok);
    if (ok) tab[pos].y = item_sym[run + __popc(b & lanemask_lt())];
    run += __popc(b);
  }
  __syncwarp();
  for (int u0 = 0; u0 < size; u0 += 32) {
    int u = u0 + lane;
    uint32_t s = tab[u].y;
    uint32_t m = __match_any_sync(0xffffffffu, s);
    uint32_t x = (uint32_t)sym_next[s] + __popc(m & lanemask_lt());
    __syncwarp();
    if ((m >> lane) == 1u) sym_next[s] = (uint16_t)((uint32_t)sym_next[s] + __popc(m));   // highest lane of the group
    __syncwarp();
    uint32_t nb = (uint32_t)(log - highbit32(x));
    tab[u] = seq_entry(kind, s, (x << nb) - (uint32_t)size, nb);
  }
  __syncwarp();
}

// Small serial FSE table for Huffman weights (log <= 6).  One thread.  Entry: next | nb<<8 | sym<<16.
__device__ inline void fse_build_small(uint32_t *tab, const int16_t *norm, int max_sym, int log) {
  const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
  uint16_t next[16];
  int high = size - 1, pos = 0;
  for (int s = 0; s <= max_sym; s++) {
    if (norm[s] == -1) { tab[high--] = (uint32_t)s << 16; next[s] = 1; }
    else next[s] = (uint16_t)norm[s];
  }
  for (int s = 0; s <= max_sym; s++)
    for (int i = 0; i < norm[s]; i++) {
      tab[pos] = (uint32_t)s << 16;
      do { pos = (pos + step) & mask; } while (pos > high);
    }
  for (int u = 0; u < size; u++) {
    uint32_t s = tab[u] >> 16;
    uint32_t x = next[s]++;
    uint32_t nb = (uint32_t)(log - highbit32(x));
    tab[u] = (((x << nb) - (uint32_t)size) & 0xFF) | (nb << 8) | (s << 16);
  }
}

// ---------------------------------------------------------------------------------------------
// Huffman tree description -> DTable in SMEM.  Called by all lanes of warp 0.
// Returns bytes consumed (>0) or -1.
// ---------------------------------------------------------------------------------------------
template <class SM>
__device__ inline int huf_read_table_warp(SM &S, const uint8_t *src, uint32_t n, int lane) {
  if (n < 1) return -1;
  const uint32_t hb = src[0];
  int nsym, used;
  if (hb >= 128) {
    nsym = (int)hb - 127;
    used = 1 + (nsym + 1) / 2;
    if ((uint32_t)used > n) return -1;
    for (int i = lane; i < nsym; i += 32) {
      uint32_t b = src[1 + (i >> 1)];
      S.weights[i] = (uint8_t)((i & 1) ? (b & 15) : (b >> 4));
    }
  } else {
    used = 1 + (int)hb;
    if ((uint32_t)used > n || hb < 1) return -1;
    int res = 0;
    if (lane == 0) {
      // serial: NCount, 64-cell table, two interleaved states
      int16_t *norm = S.huf_norm;
      uint32_t *ft = S.huf_ft;
      int max_sym = 0, al = 0;
      int hdr = read_ncount(src + 1, hb, norm, 12, 6, &max_sym, &al);
      res = -1;
      if (hdr > 0 && (uint32_t)hdr < hb) {
        fse_build_small(ft, norm, max_sym, al);
        BackBits b;
        if (b.init(src + 1 + hdr, hb - (uint32_t)hdr)) {
          b.refill();
          uint32_t s1 = b.read(al), s2 = b.read(al);
          int cnt = 0;
          bool bad = b.left < 0;
          while (!bad) {
            b.refill();
            uint32_t e1 = ft[s1];
            if (cnt >= 254) { bad = true; break; }
            S.weights[cnt++] = (uint8_t)(e1 >> 16);
            int nb1 = (int)((e1 >> 8) & 0xFF);
            if (b.left < nb1) { S.weights[cnt++] = (uint8_t)(ft[s2] >> 16); break; }
            s1 = (e1 & 0xFF) + b.read(nb1);
            uint32_t e2 = ft[s2];
            if (cnt >= 254) { bad = true; break; }
            S.weights[cnt++] = (uint8_t)(e2 >> 16);
            int nb2 = (int)((e2 >> 8) & 0xFF);
            if (b.left < nb2) { S.weights[cnt++] = (uint8_t)(ft[s1] >> 16); break; }
            s2 = (e2 & 0xFF) + b.read(nb2);
          }
          if (!bad) res = cnt;
        }
      }
    }
    res = __shfl_sync(0xffffffffu, res, 0);
    if (res < 0) return -1;
    nsym = res;
  }
  __syncwarp();
  // weight statistics
  if (lane < 16) S.rank_cnt[lane] = 0;
  __syncwarp();
  uint32_t sum = 0;
  bool bad = false;
  for (int i = lane; i < nsym; i += 32) {
    uint32_t w = S.weights[i];
    if (w > HUF_MAX_LOG) bad = true;
    else { if (w) sum += 1u << (w - 1); atomicAdd(&S.rank_cnt[w], 1u); }
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (__any_sync(0xffffffffu, bad) || sum == 0) return -1;
  const int log = highbit32(sum) + 1;
  if (log > HUF_MAX_LOG) return -1;
  const uint32_t rest = (1u << log) - sum;
  if (rest & (rest - 1)) return -1;
  const uint32_t lastw = (uint32_t)highbit32(rest) + 1;
  __syncwarp();
  if (lane == 0) {
    S.weights[nsym] = (uint8_t)lastw;
    S.rank_cnt[lastw] += 1;
    uint32_t a = 0;
    for (int w = 1; w <= log; w++) { S.rank_start[w] = a; a += S.rank_cnt[w] << (w - 1); }
  }
  nsym += 1;
  __syncwarp();
  if (S.rank_cnt[1] < 2 || (S.rank_cnt[1] & 1)) return -1;
  // cell ranges in symbol order inside each weight class; short ranges filled by the owning lane,
  // long ones (>= 32 cells) by the whole warp
  for (int i0 = 0; i0 < nsym; i0 += 32) {
    int i = i0 + lane;
    uint32_t w = (i < nsym) ? S.weights[i] : 0;
    uint32_t m = __match_any_sync(0xffffffffu, w);
    uint32_t start = 0, len = 0;
    if (w) {
      len = 1u << (w - 1);
      start = S.rank_start[w] + (uint32_t)__popc(m & lanemask_lt()) * len;
    }
    __syncwarp();
    if (w && (m >> lane) == 1u) S.rank_start[w] += (uint32_t)__popc(m) * len;
    __syncwarp();
    uint16_t e = (uint16_t)((uint32_t)i | ((uint32_t)(log + 1 - (int)w) << 8));
    if (w && len < 32) for (uint32_t k = 0; k < len; k++) S.huf[start + k] = e;
    uint32_t big = __ballot_sync(0xffffffffu, w && len >= 32);
    while (big) {
      int src_lane = __ffs(big) - 1;
      big &= big - 1;
      uint32_t st = __shfl_sync(0xffffffffu, start, src_lane), ln = __shfl_sync(0xffffffffu, len, src_lane);
      uint32_t ee = __shfl_sync(0xffffffffu, (uint32_t)e, src_lane);
      for (uint32_t k = lane; k < ln; k += 32) S.huf[st + k] = (uint16_t)ee;
    }
  }
  if (lane == 0) { S.huf_log = log; S.huf_valid = 1; }
  __syncwarp();
  return used;
}


} // namespace b200zstd
