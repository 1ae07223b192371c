// ht_tables.cpp -- derive every HT coder table from the normative CxtVLC rows.
// See ht_tables.h for the reference file:line each table mirrors.
#include "ht_tables.h"
#include <cstring>
#include <mutex>

namespace ojb {

#include "ht_cxtvlc_rows.inc"

namespace {

struct Row { int c_q, rho, u_off, e_k, e_1, cwd, len; };

inline Row unpack(uint32_t v) {
  Row r;
  r.len = v & 7; r.cwd = (v >> 3) & 0x7F; r.e_1 = (v >> 10) & 15; r.e_k = (v >> 14) & 15;
  r.u_off = (v >> 18) & 1; r.rho = (v >> 19) & 15; r.c_q = (v >> 23) & 7;
  return r;
}

// Encoder side: for every (context, significance pattern, EMB pattern) choose the codeword.
// With u_off = 1 several rows can signal the same EMB pattern; the reference keeps the row
// whose e_k has the most known bits, the LAST one on ties (ojph_block_encoder.cpp:99-118);
// with emb == 0 it is the first u_off = 0 row (:119-130).
void build_enc(const uint32_t* rows, int n, uint16_t* tbl) {
  for (int i = 0; i < 2048; ++i) {
    int c_q = i >> 8, rho = (i >> 4) & 15, emb = i & 15;
    tbl[i] = 0;
    if ((emb & rho) != emb || (rho == 0 && c_q == 0)) continue;
    int best = -1, best_pop = -1;
    for (int j = 0; j < n; ++j) {
      Row r = unpack(rows[j]);
      if (r.c_q != c_q || r.rho != rho) continue;
      if (emb) {
        if (r.u_off != 1 || (emb & r.e_k) != r.e_1) continue;
        int pop = __builtin_popcount((unsigned)r.e_k);
        if (pop >= best_pop) { best = j; best_pop = pop; }
      } else {
        if (r.u_off != 0) continue;
        best = j; break;
      }
    }
    if (best >= 0) {
      Row r = unpack(rows[best]);
      tbl[i] = (uint16_t)((r.cwd << 8) | (r.len << 4) | r.e_k);
    }
  }
}

// Decoder side: index = context (3 bits) | next 7 stream bits; every row whose codeword is a
// prefix (LSB first) of those 7 bits claims the entry (ojph_block_common.cpp:155-170).
void build_dec(const uint32_t* rows, int n, uint16_t* tbl) {
  memset(tbl, 0, 1024 * sizeof(uint16_t));
  for (int i = 0; i < 1024; ++i) {
    int head = i & 0x7F, c_q = i >> 7;
    for (int j = 0; j < n; ++j) {
      Row r = unpack(rows[j]);
      if (r.c_q == c_q && r.cwd == (head & ((1 << r.len) - 1)))
        tbl[i] = (uint16_t)((r.e_k << 12) | (r.e_1 << 8) | (r.rho << 4) | (r.u_off << 3) | r.len);
    }
  }
}

// U-VLC prefix code (T.814 Table 3), keyed by the 3 LSBs of the stream:
//  "1" -> u_pfx 1 ; "01" -> 2 ; "001" -> 3 (+1-bit suffix) ; "000" -> 5 (+5-bit suffix)
struct Pfx { int plen, slen, u; };
inline Pfx uvlc_prefix(unsigned bits3) {
  if (bits3 & 1) return Pfx{1, 0, 1};
  if (bits3 & 2) return Pfx{2, 0, 2};
  if (bits3 & 4) return Pfx{3, 1, 3};
  return Pfx{3, 5, 5};
}
inline uint16_t pack_uvlc(int tp, int ts, int s0, int u0, int u1) {
  return (uint16_t)(tp | (ts << 3) | (s0 << 7) | (u0 << 10) | (u1 << 13));
}

void build_dec_uvlc(uint16_t* t0, uint16_t* t1) {
  // initial quad-row (kappa = 1; modes 3 and 4 distinguish the MEL event for a pair with
  // both u_off set: event 0 -> one of the two u is <= 2; event 1 -> both > 2 and coded -2)
  for (unsigned i = 0; i < 320; ++i) {
    unsigned mode = i >> 6, v = i & 0x3F;
    if (mode == 0) { t0[i] = 0; continue; }
    if (mode <= 2) {
      Pfx d = uvlc_prefix(v & 7);
      t0[i] = pack_uvlc(d.plen, d.slen, mode == 1 ? d.slen : 0,
                        mode == 1 ? d.u : 0, mode == 1 ? 0 : d.u);
    } else if (mode == 3) {
      Pfx d0 = uvlc_prefix(v & 7);
      unsigned v1 = v >> d0.plen;
      Pfx d1 = uvlc_prefix(v1 & 7);
      if (d0.plen == 3)   // u_q0 > 2: u_q1 is 1 or 2, sent as a single bit
        t0[i] = pack_uvlc(d0.plen + 1, d0.slen, d0.slen, d0.u, (int)(v1 & 1) + 1);
      else
        t0[i] = pack_uvlc(d0.plen + d1.plen, d0.slen + d1.slen, d0.slen, d0.u, d1.u);
    } else {
      Pfx d0 = uvlc_prefix(v & 7);
      Pfx d1 = uvlc_prefix((v >> d0.plen) & 7);
      t0[i] = pack_uvlc(d0.plen + d1.plen, d0.slen + d1.slen, d0.slen, d0.u + 2, d1.u + 2);
    }
  }
  for (unsigned i = 0; i < 256; ++i) {
    unsigned mode = i >> 6, v = i & 0x3F;
    if (mode == 0) { t1[i] = 0; continue; }
    if (mode <= 2) {
      Pfx d = uvlc_prefix(v & 7);
      t1[i] = pack_uvlc(d.plen, d.slen, mode == 1 ? d.slen : 0,
                        mode == 1 ? d.u : 0, mode == 1 ? 0 : d.u);
    } else {
      Pfx d0 = uvlc_prefix(v & 7);
      Pfx d1 = uvlc_prefix((v >> d0.plen) & 7);
      t1[i] = pack_uvlc(d0.plen + d1.plen, d0.slen + d1.slen, d0.slen, d0.u, d1.u);
    }
  }
}

void build_enc_uvlc(uint16_t* t) {
  // u -> prefix (LSB first) / suffix; u = 0 emits nothing
  for (int u = 0; u <= 32; ++u) {
    int pre, pl, suf, sl;
    if (u == 0)      { pre = 0; pl = 0; suf = 0; sl = 0; }
    else if (u == 1) { pre = 1; pl = 1; suf = 0; sl = 0; }
    else if (u == 2) { pre = 2; pl = 2; suf = 0; sl = 0; }
    else if (u <= 4) { pre = 4; pl = 3; suf = u - 3; sl = 1; }
    else             { pre = 0; pl = 3; suf = u - 5; sl = 5; }
    t[u] = (uint16_t)(pre | (pl << 3) | (suf << 6) | (sl << 11));
  }
}

HtTables g_tables;
std::once_flag g_once;

} // namespace

const HtTables& ht_tables() {
  std::call_once(g_once, [] {
    build_enc(kCxtVlcRows0, 444, g_tables.enc_vlc[0]);
    build_enc(kCxtVlcRows1, 358, g_tables.enc_vlc[1]);
    build_dec(kCxtVlcRows0, 444, g_tables.dec_vlc[0]);
    build_dec(kCxtVlcRows1, 358, g_tables.dec_vlc[1]);
    build_dec_uvlc(g_tables.dec_uvlc0, g_tables.dec_uvlc1);
    build_enc_uvlc(g_tables.enc_uvlc);
  });
  return g_tables;
}

} // namespace ojb
