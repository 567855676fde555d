// zstd_encode_lz.cuh -- the parse of levels 1-4 (zstd_encode_esd.cu), written as __host__ __device__ code: the SELECT
// stage a lane runs on the GPU and the restatement of the MATCH stage that the host-side model runs
// (tests/model/enc_model.cpp) are the same functions, so the two must produce the same sequences.
//
// Reference counterparts (behavioural spec only): find_matches_kernel / greedy parse, src/lz77_parallel.cu:26-268;
// repeat-offset coding as RFC 8878 3.1.1.5 / 3.1.2.5.
//
// MATCH   every position p < n - 8 of a block looks its candidates up in hash tables that hold, per bucket, the highest
//         position of all WINDOWS BEFORE p's own (LZ_WIN positions per window; lookup of a window, then insert of the
//         window: "highest position wins" makes the insert order irrelevant) or, through a small side table, the first
//         position of its own window with the same hash; verifies them and measures the match up to LZ_LCAP bytes:
//         R[p] = offset | length << 17  (0: nothing).
// SELECT  the greedy walk over R.  A block is cut into LZ_LANES sub-segments; lane j first walks its sub-segment as if
//         the parse entered it at its first position with an unknown repeat-offset history (speculation), then re-walks
//         from the state the lane before it really left behind until it meets its own speculative walk (same match end,
//         same history) -- after that point the two walks are identical, so the speculative tail is kept.  The result is
//         exactly the serial walk of the whole block; a lane whose exit state changed makes the next lane re-walk.
#pragma once
#include <stdint.h>

#include "zstd_encode_core.cuh"
#include "zstd_encode_params.h"

namespace b200zstd {
namespace lz {

constexpr int LZ_WIN_LOG = 8;
constexpr uint32_t LZ_WIN = 1u << LZ_WIN_LOG;   // positions per match-stage window (= threads of one group of the match CTA)
constexpr uint32_t LZ_LANES = 32;         // sub-segments per block (= lanes of the select warp)
constexpr uint32_t LZ_LCAP = 16;          // the match stage measures a match up to this length; the select stage finishes longer ones
constexpr uint32_t LZ_MIN_MATCH = 5;      // shortest table match; repeat-offset matches need 4
constexpr uint32_t LZ_REP_BONUS = 2;      // a repeat-offset match wins when its length + bonus reaches the table match
constexpr uint32_t LZ_BACK_MAX = 8;       // backward extension into pending literals, bytes
constexpr uint32_t LZ_OFF_MASK = (1u << 17) - 1;
constexpr int LZ_FIRST_LOG = 9;            // slots of the "first position of this window" side table, per hash table

// The side table that lets a position see the EARLIER positions of its own window: slot = top LZ_FIRST_LOG bits of the
// hash; key = window tag | position inside the window | low hash bits, combined by atomicMin.  Tags fall from window
// to window, so a slot never needs clearing and the smallest key of the current window is its first inserted position.
// (9 tag bits cover the 128 windows of a 128 KB block many times over.)
constexpr int LZ_KEY_HBITS = 23 - LZ_WIN_LOG;
ZHD uint32_t first_key(uint32_t window, uint32_t t, uint32_t h) { return ((~window & 0x1FFu) << 23) | (t << LZ_KEY_HBITS) | (h & ((1u << LZ_KEY_HBITS) - 1u)); }
// candidate from the side table for thread t of `window` (hash h), or -1
ZHD int32_t first_candidate(uint32_t key, uint32_t window, uint32_t t, uint32_t h) {
  const uint32_t kt = (key >> LZ_KEY_HBITS) & (LZ_WIN - 1u), hm = (1u << LZ_KEY_HBITS) - 1u;
  const bool mine = (key >> 23) == (~window & 0x1FFu) && (key & hm) == (h & hm) && kt < t;
  return mine ? (int32_t)(window * LZ_WIN + kt) : -1;
}
// a position enters the tables unless its hash equals that of the position before it inside its group of 32 (runs keep
// their first position, so that a later occurrence of the run finds the aligned candidate)
ZHD bool inserts(uint32_t p, uint32_t h, uint32_t h_prev) { return (p & 31u) == 0 || h != h_prev; }

// ---- byte access.  `in` is the block (device: global memory, any alignment).  Only bytes [0, n) are ever read. ----
ZHD uint32_t rd8(const uint8_t *in, uint32_t p) { return in[p]; }
// 8 little-endian bytes at p; requires p + 8 <= n
ZHD uint64_t rd64(const uint8_t *in, uint32_t p) {
#if defined(__CUDA_ARCH__)
  const uintptr_t a = (uintptr_t)(in + p);
  const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  const uint32_t w0 = w[0], w1 = w[1];
  if (sh == 0) return ((uint64_t)w1 << 32) | w0;
  const uint32_t w2 = w[2];                  // holds byte p + 7, which is inside the block
  return ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
#else
  uint64_t v = 0;
  for (int k = 0; k < 8; k++) v |= (uint64_t)in[p + k] << (8 * k);
  return v;
#endif
}
ZHD uint32_t min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }
ZHD uint32_t common8(uint64_t a, uint64_t b) {
  const uint32_t xl = (uint32_t)a ^ (uint32_t)b, xh = (uint32_t)(a >> 32) ^ (uint32_t)(b >> 32);
#if defined(__CUDA_ARCH__)
  if (xl) return (uint32_t)(__ffs((int)xl) - 1) >> 3;
  if (xh) return 4u + ((uint32_t)(__ffs((int)xh) - 1) >> 3);
#else
  if (xl) return (uint32_t)__builtin_ctz(xl) >> 3;
  if (xh) return 4u + ((uint32_t)__builtin_ctz(xh) >> 3);
#endif
  return 8u;
}
// number of equal bytes of in[a ..] and in[a - off ..], a + result <= n
ZHD uint32_t count_fwd(const uint8_t *in, uint32_t a, uint32_t off, uint32_t n) {
  uint32_t len = 0;
  while (a + len + 8 <= n) {
    const uint32_t c = common8(rd64(in, a + len), rd64(in, a + len - off));
    len += c;
    if (c < 8) return len;
  }
  while (a + len < n && rd8(in, a + len) == rd8(in, a + len - off)) len++;
  return len;
}

// ---- levels 5+: rows.  A bucket of the 8-byte table is a row of LZ_ROW_WAYS entries "position << 15 | tag"; window w
// writes way (w mod LZ_ROW_WAYS) of a row with atomicMax, so a row keeps, per way, the highest position among the windows
// that map to it: up to 16 recent occurrences from different neighbourhoods, deterministic, no counters.  The tag (15 more
// hash bits) filters the ways before any byte is compared. ----
constexpr int LZ_ROW_WAYS = 16;
constexpr uint32_t LZ_QCAP = 64;         // rows: matches are measured up to this length (the walk finishes longer ones)
constexpr uint32_t LZ_Q_MIN_MATCH = 4;
#ifndef LZ_QTARGET_V
#define LZ_QTARGET_V 48
#endif
constexpr uint32_t LZ_QTARGET = LZ_QTARGET_V;     // a match this long is good enough: the remaining candidates are not tried
// tag of a row entry: 8 bits of the 8 hashed bytes (filters hash collisions) | 7 bits of the 4 bytes behind them.  A way
// whose high part fits is a candidate; it is STRONG when the low part fits too (12 equal bytes but for a 1 / 128 chance) and
// WEAK otherwise -- a weak way matches fewer than 12 bytes for certain.  next4: bytes 8..11 from p (0 near the block's end).
ZHD uint32_t row_tag(uint64_t v, uint32_t next4) {
  const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  return ((((lo * 3266489917u) ^ (hi * 668265263u)) >> 24) << 7) | ((next4 * 2246822519u) >> 25);
}
constexpr uint32_t LZ_TAG_HI = 0x7F80u, LZ_TAG_ALL = 0x7FFFu;
ZHD uint32_t row_entry(uint32_t p, uint32_t tag) { return (p << 15) | tag; }
ZHD uint32_t short_tag(uint64_t v) { return (((uint32_t)v) * 3266489917u) >> 17; }          // of the 4 hashed bytes
// exact common length of the bytes at p and at c (< p), up to LZ_QCAP and the end of the block
template <typename Rd64>
ZHD uint32_t match_len_q(const Rd64 &rd, uint32_t p, uint32_t c, uint32_t n) {
  uint32_t len = 0;
  while (len < LZ_QCAP && p + len + 8 <= n) {
    const uint32_t k = common8(rd(p + len), rd(c + len));
    len += k;
    if (k < 8) break;
  }
  return len > LZ_QCAP ? LZ_QCAP : len;
}
// Candidates are tried in a fixed order (first of window for the 8-byte and the 4-byte hash, ways 0..15, the 4-byte entry);
// one replaces the best so far only when it is LONGER, and it is measured only when the 8 bytes that end one past the best
// length already agree -- most ways of a row share the hashed bytes and little more, and fail that test after two loads.
// can candidate c still be longer than best_len?  (false is final: a larger best_len only makes it harder)
template <typename Rd64>
ZHD bool may_improve(const Rd64 &rd, uint32_t p, uint32_t c, uint32_t n, uint32_t best_len) {
  if (best_len >= LZ_QTARGET || p + best_len + 1 > n) return false;
  if (best_len >= 7) return rd(p + best_len - 7) == rd(c + best_len - 7);
  if (best_len > 0) return p + 8 <= n && common8(rd(p), rd(c)) > best_len;
  return true;
}
template <typename Rd64>
ZHD void try_candidate(const Rd64 &rd, uint32_t p, uint32_t c, uint32_t n, uint32_t &best_len, uint32_t &best_off) {
  if (!may_improve(rd, p, c, n, best_len)) return;
  const uint32_t len = match_len_q(rd, p, c, n);
  if (len > best_len) { best_len = len; best_off = p - c; }
}

// ---- MATCH stage, per position (the kernel runs the same arithmetic on the block staged in shared memory) ----
// c_long / c_short: candidate positions (< p) or -1.  v = rd64(p).  Returns R[p].
template <typename Rd64>
ZHD uint32_t match_verify(const Rd64 &rd, uint32_t p, uint64_t v, int32_t c_long, int32_t c_short, uint32_t n) {
  uint32_t off = 0, len = 0;
  if (c_long >= 0 && rd((uint32_t)c_long) == v) { off = p - (uint32_t)c_long; len = 8; }
  else if (c_short >= 0) {
    const uint32_t c = common8(v, rd((uint32_t)c_short));
    if (c >= LZ_MIN_MATCH) { off = p - (uint32_t)c_short; len = c; }
  }
  if (len == 8) {
    while (len < LZ_LCAP && p + len + 8 <= n) {
      const uint32_t c = common8(rd(p + len), rd(p + len - off));
      len += c;
      if (c < 8) break;
    }
    if (len > LZ_LCAP) len = LZ_LCAP;
  }
  return len ? (off | (len << 17)) : 0u;
}

// ---- SELECT stage ----
struct State {
  uint32_t ip, anchor;        // next decision position; end of the last match (start of the pending literals)
  uint32_t r0, r1, r2;        // repeat-offset history as the decoder will hold it (0 = unknown)
  ZHD bool same(const State &o) const { return ip == o.ip && anchor == o.anchor && r0 == o.r0 && r1 == o.r1 && r2 == o.r2; }
};

// repeat-offset coding of one sequence on a three-word history (enc::offset_to_code without the array)
ZHD uint32_t code_offset(uint32_t off, bool ll0, uint32_t &r0, uint32_t &r1, uint32_t &r2) {
  uint32_t code;
  if (!ll0) {
    if (off == r0) return 1;
    if (off == r1) { r1 = r0; r0 = off; return 2; }
    code = off == r2 ? 3u : off + 3u;
  } else {
    if (off == r1) { r1 = r0; r0 = off; return 1; }
    if (off == r2) code = 2;
    else if (r0 > 1 && off == r0 - 1) code = 3;
    else code = off + 3;
  }
  r2 = r1; r1 = r0; r0 = off;
  return code;
}
// the decoder's side of the same rule: offset code -> offset, history updated (RFC 8878 3.1.1.5); used to replay a list
ZHD uint32_t decode_offset(uint32_t code, bool ll0, uint32_t &r0, uint32_t &r1, uint32_t &r2) {
  uint32_t off;
  if (code > 3) { off = code - 3; r2 = r1; r1 = r0; r0 = off; return off; }
  const uint32_t idx = code - 1 + (ll0 ? 1u : 0u);
  if (idx == 0) return r0;
  if (idx == 1) { off = r1; r1 = r0; r0 = off; return off; }
  off = idx == 2 ? r2 : r0 - 1;
  r2 = r1; r1 = r0; r0 = off;
  return off;
}

// list entry: start | (length & 0x7FFF) << 17 ,  offset code | (length >> 15) << 18
struct Seq { uint32_t x, y; };
ZHD Seq pack_seq(uint32_t s, uint32_t len, uint32_t code) { return Seq{s | ((len & 0x7FFFu) << 17), code | ((len >> 15) << 18)}; }
ZHD uint32_t seq_start(const Seq &q) { return q.x & LZ_OFF_MASK; }
ZHD uint32_t seq_len(const Seq &q) { return (q.x >> 17) | ((q.y >> 18) << 15); }
ZHD uint32_t seq_code(const Seq &q) { return q.y & ((1u << 18) - 1); }

struct SelectParams { int lazy; int rows; int exact_stitch = 0; };       // exact_stitch: tests only (see select_rewalk)

// One decision of the walk at st.ip (< lim).  Either emits one sequence (returns true, *out filled) or skips literals.
// n: block bytes; lim: first position this lane does not decide.
ZHD bool select_step(const uint8_t *in, uint32_t n, const uint32_t *R, uint32_t lim, const SelectParams &P, State &st, Seq *out) {
  const uint32_t ip = st.ip;
  uint32_t s, off, len;
  bool found = false;
  // (a) right behind a match: the second repeat offset with no literals in between (libzstd does the same after every match)
  if (ip == st.anchor && st.r1 != 0 && ip >= st.r1) {
    const uint64_t a = rd64(in, ip), b = rd64(in, ip - st.r1);
    const uint32_t c = common8(a, b);
    if (c >= 4 && (c == 8 || c + LZ_REP_BONUS >= (R[ip] >> 17))) { s = ip; off = st.r1; len = c == 8 ? 8 + count_fwd(in, ip + 8, off, n) : c; found = true; }
  }
  if (!found) {
    const uint32_t e0 = R[ip], e1 = ip + 1 < lim ? R[ip + 1] : 0u;
    const uint32_t lt0 = e0 >> 17;
    uint32_t rl0 = 0, rl1 = 0;
    if (st.r0 != 0) {
      if (ip >= st.r0 && ip != st.anchor) rl0 = common8(rd64(in, ip), rd64(in, ip - st.r0));
      if (ip + 1 >= st.r0 && ip + 1 < lim) rl1 = common8(rd64(in, ip + 1), rd64(in, ip + 1 - st.r0));
    }
    if (P.rows) {           // levels 5+ compare exact lengths: finish the measurement of a repeat-offset match that filled its 8 bytes
      if (rl0 == 8) rl0 += count_fwd(in, ip + 8, st.r0, n);
      if (rl1 == 8) rl1 += count_fwd(in, ip + 9, st.r0, n);
    }
    const bool t0 = e0 != 0, r0ok = rl0 >= 4, r1ok = rl1 >= 4;
    if (!t0 && !r0ok && !r1ok) {
      uint32_t q = ip + 1;
      while (q < lim && R[q] == 0) q++;
      st.ip = q;
      return false;
    }
    bool open;
    if (P.rows) {
      // levels 5+: compare by gain (4 bits per matched byte minus the offset's cost; a repeat offset costs nothing),
      // then up to P.lazy positions further on may displace the choice when they gain more than the delay costs
      const uint32_t cap = LZ_QCAP;
      auto gain = [](uint32_t l, uint32_t o, bool rep) { return (int)(l * 4) - (rep ? 0 : enc::hb32(o + 1)); };
      s = ip;
      open = false;                                     // every length compared here is exact
      // a table match that filled the match stage's cap is measured to its end before it is compared
      auto full_len = [&](uint32_t e, uint32_t q) { uint32_t l = e >> 17; if (l == cap) l += count_fwd(in, q + l, e & LZ_OFF_MASK, n); return l; };
      const uint32_t fl0 = t0 ? full_len(e0, ip) : 0u;
      if (r0ok && (!t0 || gain(rl0, st.r0, true) + 4 >= gain(fl0, e0 & LZ_OFF_MASK, false))) { off = st.r0; len = rl0; }
      else if (t0) { off = e0 & LZ_OFF_MASK; len = fl0; }
      else { off = 0; len = 0; }
      int g0 = len ? gain(len, off, off == st.r0) : -1000;
      for (int step = 1; step <= P.lazy; step++) {
        const uint32_t q = ip + (uint32_t)step;
        if (q >= lim) break;
        const uint32_t eq = step == 1 ? e1 : R[q];
        uint32_t o2 = eq & LZ_OFF_MASK, l2 = eq ? full_len(eq, q) : 0u;
        bool rep2 = false;
        uint32_t rq = 0;
        if (st.r0 != 0 && q >= st.r0) {
          if (step == 1) rq = rl1;
          else { rq = common8(rd64(in, q), rd64(in, q - st.r0)); if (rq == 8) rq += count_fwd(in, q + 8, st.r0, n); }
        }
        if (rq >= 4 && (l2 == 0 || gain(rq, st.r0, true) + 4 >= gain(l2, o2, false))) { o2 = st.r0; l2 = rq; rep2 = true; }
        if (l2 == 0) continue;
        const int g2 = gain(l2, o2, rep2);
        if (g2 > g0 + 3 * step) { s = q; off = o2; len = l2; g0 = g2; }
      }
      if (len == 0) { st.ip = ip + 1; return false; }       // only a too-short candidate at ip + 1 / ip + 2 brought us here
    }
    else if (r0ok && (!t0 || rl0 == 8 || rl0 + LZ_REP_BONUS >= lt0)) { s = ip; off = st.r0; len = rl0; open = rl0 == 8; }
    else if (r1ok && (!t0 || rl1 == 8 || rl1 + LZ_REP_BONUS >= lt0)) { s = ip + 1; off = st.r0; len = rl1; open = rl1 == 8; }
    else {
      uint32_t e = e0;
      s = ip;
      if (P.lazy && (e1 >> 17) > lt0) { e = e1; s = ip + 1; }
      off = e & LZ_OFF_MASK; len = e >> 17; open = len == LZ_LCAP;
    }
    if (open) len += count_fwd(in, s + len, off, n);
    uint32_t nb = 0;
    while (nb < LZ_BACK_MAX && s - nb > st.anchor && s - nb - 1 >= off && rd8(in, s - nb - 1) == rd8(in, s - nb - 1 - off)) nb++;
    s -= nb; len += nb;
  }
  const uint32_t code = code_offset(off, s == st.anchor, st.r0, st.r1, st.r2);
  *out = pack_seq(s, len, code);
  st.ip = st.anchor = s + len;
  return true;
}

// Speculative walk of one lane: from st until st.ip >= lim.  Returns the number of sequences written to list.
ZHD uint32_t select_walk(const uint8_t *in, uint32_t n, const uint32_t *R, uint32_t lim, const SelectParams &P, State &st, Seq *list) {
  uint32_t cnt = 0;
  while (st.ip < lim) {
    Seq q;
    if (select_step(in, n, R, lim, P, st, &q)) list[cnt++] = q;
  }
  return cnt;
}

// Re-walk of a lane from its true entry state until it meets the speculative walk `spec` (spec_cnt sequences that
// started from spec0 and ended in spec_exit).  prefix receives the sequences of the re-walk.  On return:
//   *sync_k   index of the first speculative sequence that is kept (spec_cnt: none)
//   st        the lane's true exit state
// returns the number of prefix sequences.
// The two walks have met when they stand at the same position with the same two youngest history entries: no decision
// reads the third one (select_step), so from there on the true walk takes the speculative walk's sequences.  While the
// third entries differ an offset code that involves it can differ, so the speculative sequences are copied to prefix,
// re-coded on the true history, until a new offset has pushed the differing entry out; the rest is kept as it is.
// (P.exact_stitch: keep walking until all three entries agree -- the plain rule, used by the tests as a cross-check.)
ZHD uint32_t select_rewalk(const uint8_t *in, uint32_t n, const uint32_t *R, uint32_t lim, const SelectParams &P, State &st, const Seq *spec,
                           uint32_t spec_cnt, const State &spec0, const State &spec_exit, Seq *prefix, uint32_t *sync_k) {
  uint32_t cnt = 0, k = 0;
  State sp = spec0;                       // speculative walk replayed up to (not including) sequence k
  while (st.ip < lim) {
    Seq q;
    if (!select_step(in, n, R, lim, P, st, &q)) continue;
    prefix[cnt++] = q;
    // advance the replay to the first speculative sequence that ends at or beyond this one
    while (k < spec_cnt) {
      const Seq e = spec[k];
      const uint32_t es = seq_start(e), ee = es + seq_len(e);
      if (ee > st.anchor) break;
      decode_offset(seq_code(e), es == sp.anchor, sp.r0, sp.r1, sp.r2);
      sp.ip = sp.anchor = ee;
      k++;
      if (ee == st.anchor) break;
    }
    if (sp.anchor == st.anchor && sp.r0 == st.r0 && sp.r1 == st.r1 && k > 0 && (sp.r2 == st.r2 || !P.exact_stitch)) {
      while (sp.r2 != st.r2 && k < spec_cnt) {
        const Seq e = spec[k];
        const uint32_t es = seq_start(e), el = seq_len(e);
        const bool ll0 = es == st.anchor;
        const uint32_t o = decode_offset(seq_code(e), ll0, sp.r0, sp.r1, sp.r2);
        prefix[cnt++] = pack_seq(es, el, code_offset(o, ll0, st.r0, st.r1, st.r2));
        st.anchor = sp.anchor = es + el;
        k++;
      }
      *sync_k = k;
      if (sp.r2 == st.r2) st = spec_exit;
      else { st.ip = spec_exit.ip; st.anchor = spec_exit.anchor; }      // the list ended first: the true history stays
      return cnt;
    }
  }
  *sync_k = spec_cnt;
  return cnt;
}

// The lane's entry changed in the third history entry only (`from` -> `to`): same decisions as its last walk, so its
// lists are re-coded -- prefix in place, then kept speculative sequences are copied behind it -- until the two
// histories meet, without reading the input.  exit_state keeps its position and takes the new history if they never meet.
ZHD bool entry_differs_in_r2_only(const State &a, const State &b) {
  return a.ip == b.ip && a.anchor == b.anchor && a.r0 == b.r0 && a.r1 == b.r1 && a.r2 != b.r2;
}
ZHD void select_recode(Seq *prefix, uint32_t &pre_cnt, const Seq *spec, uint32_t spec_cnt, uint32_t &sync_k, const State &from, const State &to,
                       State &exit_state) {
  uint32_t o0 = from.r0, o1 = from.r1, o2 = from.r2, n0 = to.r0, n1 = to.r1, n2 = to.r2;
  uint32_t anchor = to.anchor;
  for (uint32_t i = 0; i < pre_cnt; i++) {
    const Seq e = prefix[i];
    const uint32_t es = seq_start(e), el = seq_len(e);
    const bool ll0 = es == anchor;
    const uint32_t o = decode_offset(seq_code(e), ll0, o0, o1, o2);
    const uint32_t code = code_offset(o, ll0, n0, n1, n2);
    if (code != seq_code(e)) prefix[i] = pack_seq(es, el, code);
    anchor = es + el;
    if (o2 == n2) return;
  }
  uint32_t k = sync_k;
  while (k < spec_cnt) {
    const Seq e = spec[k];
    const uint32_t es = seq_start(e), el = seq_len(e);
    const bool ll0 = es == anchor;
    const uint32_t o = decode_offset(seq_code(e), ll0, o0, o1, o2);
    prefix[pre_cnt++] = pack_seq(es, el, code_offset(o, ll0, n0, n1, n2));
    anchor = es + el;
    k++;
    if (o2 == n2) { sync_k = k; return; }
  }
  sync_k = k;
  exit_state.r0 = n0; exit_state.r1 = n1; exit_state.r2 = n2;
}

// sub-segment geometry: lane j decides positions [lane_begin(j), lane_begin(j + 1)) of [0, ilimit)
ZHD uint32_t lane_span(uint32_t ilimit) { return ((ilimit + LZ_LANES - 1) / LZ_LANES + 3u) & ~3u; }
ZHD uint32_t lane_begin(uint32_t j, uint32_t span, uint32_t ilimit) { const uint32_t b = j * span; return b < ilimit ? b : ilimit; }
// sequences a lane can emit: they start inside its span (+ the backward extension) and cover >= 4 bytes each
ZHD uint32_t lane_list_cap(uint32_t block_max) { return (block_max / LZ_LANES + 4) / 4 + 8; }

} // namespace lz
} // namespace b200zstd
