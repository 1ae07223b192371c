// ht_encode.cu -- HTJ2K cleanup-pass block encoder for sm_100a: one warp per code-block.
//
// What it computes is what the reference's ojph_encode_codeblock32 computes
// (src/core/coding/ojph_block_encoder.cpp:542-1016): the HT cleanup pass of one code-block
// as MagSgn | MEL | VLC byte segments with the 12-bit Scup locator -- byte for byte.
// How it computes it is different: the reference walks quads serially (or 16/32 columns at
// a time with AVX); here lane i of the warp owns quad column i (samples 2i, 2i+1) of every
// quad-row, so
//   * significance / exponents / EMB patterns / VLC codewords are computed by all 32 lanes
//     at once, the neighbourhood (left quad, the four samples above) comes from
//     __shfl_up/_down of per-lane registers instead of e_val[]/cx_val[] line buffers
//     (ojph_block_encoder.cpp:568-580,873-879);
//   * the three bit-streams are packed with ONE packed warp scan per quad-row (bit offsets of
//     every lane's MagSgn bits, every quad pair's VLC bits and every MEL event), raw bits go
//     to a small per-warp shared-memory row buffer, and the data-dependent byte stuffing
//     (MagSgn: 7-bit byte after 0xFF, :483-488; VLC: 7-bit byte 0x7F after a byte > 0x8F,
//     :393-404) is applied as a parallel "find first violation, insert one bit, repeat"
//     fix-up that almost always finds nothing;
//   * MEL (13-state adaptive run-length coder, :321-348) consumes a compacted event word per
//     quad-row, whole zero-runs per step.
// Output goes to a per-block slot: MagSgn (+ MEL at the end) grows forward from the slot
// start, VLC grows backward from the slot end exactly as it will sit in the codestream; a
// later kernel closes the gap while gathering blocks into packets.
#include "ojb_device.h"
#include "ojb_kernels.h"

namespace ojb {

#define ENC_WARPS 4
#define ENC_MS_WORDS 160      // MagSgn row buffer: 32 quads x 4 x 31 bits + carry
#define ENC_VLC_WORDS 32      // VLC row buffer: 16 pairs x 30 bits + carry
#define ENC_MEL_BYTES 256

#define FULL 0xFFFFFFFFu

// MEL exponent table {0,0,0,1,1,1,2,2,2,3,3,4,5} packed 4 bits per state
#define MEL_EXP(k) ((uint32_t)((0x5433222111000ull >> (4 * (k))) & 7ull))

struct MelState {
  uint32_t k, run, tmp, rem, pos;    // rem = free bits in the current byte
};

__device__ __forceinline__ void mel_emit_bit(MelState& m, uint32_t v, uint8_t* buf, uint32_t lane) {
  m.tmp = (m.tmp << 1) | v;
  if (--m.rem == 0) {
    if (lane == 0 && m.pos < ENC_MEL_BYTES) buf[m.pos] = (uint8_t)m.tmp;
    m.pos++;
    m.rem = (m.tmp == 0xFF) ? 7 : 8;
    m.tmp = 0;
  }
}

// consume n events (LSB first in ev); executed uniformly by all lanes
__device__ __forceinline__ void mel_encode_events(MelState& m, unsigned long long ev, uint32_t n,
                                                  uint8_t* buf, uint32_t lane) {
  while (n > 0) {
    if (ev & 1ull) {                       // a "1" event: 0 bit + run in mel_exp[k] bits
      mel_emit_bit(m, 0, buf, lane);
      uint32_t t = MEL_EXP(m.k);
      while (t > 0) { --t; mel_emit_bit(m, (m.run >> t) & 1u, buf, lane); }
      m.run = 0;
      m.k = m.k > 0 ? m.k - 1 : 0;
      ev >>= 1; --n;
    } else {                               // a run of "0" events: consume it thresholds at a time
      uint32_t z = ev ? (uint32_t)(__ffsll((long long)ev) - 1) : 64u;
      z = min(z, n);
      ev = (z >= 64) ? 0ull : (ev >> z);
      n -= z;
      while (z > 0) {
        uint32_t need = (1u << MEL_EXP(m.k)) - m.run;
        if (z >= need) {
          mel_emit_bit(m, 1, buf, lane);
          m.run = 0; m.k = m.k < 12 ? m.k + 1 : 12; z -= need;
        } else { m.run += z; z = 0; }
      }
    }
  }
}

__device__ __forceinline__ void put_bits(uint32_t* buf, uint32_t o, unsigned long long v, uint32_t n) {
  if (n == 0) return;
  uint32_t wi = o >> 5, sh = o & 31;
  unsigned long long lo = v << sh;
  uint32_t a = (uint32_t)lo, b = (uint32_t)(lo >> 32);
  if (a) atomicOr(&buf[wi], a);
  if (b) atomicOr(&buf[wi + 1], b);
  if (sh) { uint32_t c = (uint32_t)(v >> (64 - sh)); if (c) atomicOr(&buf[wi + 2], c); }
}

// per-byte "== 0xFF" flags at bit 0 of each byte
__device__ __forceinline__ uint32_t ff_bytes(uint32_t w) {
  uint32_t t = w & (w >> 1); t &= t >> 2; t &= t >> 4; return t & 0x01010101u;
}

// Insert a zero bit at bit position P of the row buffer (all lanes cooperate; in step k lane l
// handles word 32k + l).  Bits at positions >= P move up by one.
__device__ __forceinline__ void insert_zero_bit(uint32_t* buf, uint32_t P, uint32_t nwords, uint32_t lane) {
  const uint32_t pw = P >> 5, pb = P & 31;
  uint32_t carry = 0;                                  // word that precedes this step's 32 words
  for (uint32_t wb = pw & ~31u; wb < nwords; wb += 32) {
    const uint32_t wi = wb + lane;
    const uint32_t w = (wi < nwords) ? buf[wi] : 0u;
    uint32_t prev = __shfl_up_sync(FULL, w, 1);
    if (lane == 0) prev = carry;
    carry = __shfl_sync(FULL, w, 31);
    if (wi < nwords && wi >= pw) {
      uint32_t nw;
      if (wi > pw) nw = (w << 1) | (prev >> 31);
      else { const uint32_t lowmask = (1u << pb) - 1u; nw = (w & lowmask) | ((w & ~lowmask) << 1); }
      buf[wi] = nw;
    }
  }
  __syncwarp();
}

// MagSgn stuffing fix-up on the row buffer: after a 0xFF byte the next byte carries 7 bits.
// `start` (persistent across quad-rows, relative to the buffer) is the index of the first
// byte whose 0xFF has not yet been answered by a stuffed bit; -1 stands for the byte that
// precedes the buffer (prev_ff tells whether that byte is 0xFF).  Returns the new bit count.
__device__ __forceinline__ uint32_t ms_stuff(uint32_t* buf, uint32_t nbits, int& start, bool prev_ff,
                                             uint32_t lane) {
  if (start < 0) {
    if (!prev_ff) start = 0;
    else if (nbits >= 7) {
      insert_zero_bit(buf, 7, (nbits >> 5) + 2, lane);
      nbits++; start = 1;
    } else return nbits;                   // still pending
  }
  for (;;) {
    if (nbits < 15) break;
    const uint32_t jlim = (nbits - 7) / 8 - 1;   // last byte index whose successor has its 7 bits
    // first 0xFF byte j with start <= j <= jlim: 32 words per step, lowest lane wins
    uint32_t j = 0xFFFFFFFFu;
    for (uint32_t wb = ((uint32_t)start >> 2) & ~31u; wb * 4 <= jlim; wb += 32) {
      const uint32_t wi = wb + lane, lo = wi * 4;
      uint32_t f = (lo <= jlim) ? ff_bytes(buf[wi]) : 0u;
      if ((uint32_t)start > lo) f &= ((uint32_t)start - lo >= 4) ? 0u : (0xFFFFFFFFu << (8 * ((uint32_t)start - lo)));
      if (jlim < lo + 3 && lo <= jlim) f &= 0xFFFFFFFFu >> (8 * (lo + 3 - jlim));
      const uint32_t vote = __ballot_sync(FULL, f != 0);
      if (vote) {
        const uint32_t srcl = (uint32_t)__ffs((int)vote) - 1u;
        const uint32_t fw = __shfl_sync(FULL, f, srcl);
        j = (wb + srcl) * 4 + (((uint32_t)__ffs((int)fw) - 1u) >> 3);
        break;
      }
    }
    if (j == 0xFFFFFFFFu) break;
    insert_zero_bit(buf, 8 * (j + 1) + 7, ((nbits + 31) >> 5) + 1, lane);
    nbits++; start = (int)j + 2;
  }
  if (nbits >= 7) start = max(start, (int)((nbits - 7) / 8));   // bytes <= jlim are settled
  return nbits;
}

// VLC stuffing fix-up: a byte whose predecessor is > 0x8F and whose low 7 bits are 0x7F
// carries only those 7 bits.  Lane i owns word i.  prev = byte preceding the buffer;
// `start` (persistent) = first byte index not yet settled.
__device__ __forceinline__ uint32_t vlc_stuff(uint32_t* buf, uint32_t nbits, int& start, uint32_t prev,
                                              uint32_t lane) {
  for (;;) {
    if (nbits < 7) break;
    const uint32_t jlim = (nbits - 7) / 8;       // last byte index that has 7 bits
    const uint32_t w = buf[lane];
    uint32_t below = __shfl_up_sync(FULL, w, 1);
    if (lane == 0) below = prev << 24;
    // all four bytes at once: flag (bit 7 of the byte) where the low 7 bits are 0x7F and the
    // preceding byte is > 0x8F (bit 7 set and any of bits 6..4 set)
    const uint32_t pv = (w << 8) | (below >> 24);
    uint32_t f = (((w & 0x7F7F7F7Fu) + 0x01010101u) & pv & (((pv & 0x70707070u) + 0x70707070u))) & 0x80808080u;
    const uint32_t lo = lane * 4;
    if ((uint32_t)start > lo) f &= ((uint32_t)start - lo >= 4) ? 0u : (0xFFFFFFFFu << (8 * ((uint32_t)start - lo)));
    if (jlim < lo + 3) f &= (jlim < lo) ? 0u : (0xFFFFFFFFu >> (8 * (lo + 3 - jlim)));
    const uint32_t vote = __ballot_sync(FULL, f != 0);
    if (vote == 0) break;
    const uint32_t srcl = (uint32_t)__ffs((int)vote) - 1u;
    const uint32_t fw = __shfl_sync(FULL, f, srcl);
    const uint32_t j = srcl * 4 + (((uint32_t)__ffs((int)fw) - 1u) >> 3);
    insert_zero_bit(buf, 8 * j + 7, ENC_VLC_WORDS, lane);
    nbits++; start = (int)j + 1;
  }
  if (nbits >= 7) start = max(start, (int)((nbits - 7) / 8) + 1);
  return nbits;
}

__device__ __forceinline__ uint32_t bytes_rev(uint32_t w) { return __byte_perm(w, 0, 0x0123); }

__global__ void __launch_bounds__(ENC_WARPS * 32)
ht_encode_kernel(const EncBlock* __restrict__ blocks, uint32_t nblocks,
                 const uint32_t* __restrict__ coef, uint8_t* __restrict__ slots,
                 EncResult* __restrict__ results, const uint16_t* __restrict__ tables,
                 uint32_t* __restrict__ status)
{
  __shared__ uint16_t s_vlc[2 * 2048];
  __shared__ uint16_t s_uvlc[36];
  __shared__ uint32_t s_ms[ENC_WARPS][ENC_MS_WORDS];
  __shared__ uint32_t s_vl[ENC_WARPS][ENC_VLC_WORDS];
  __shared__ uint8_t s_mel[ENC_WARPS][ENC_MEL_BYTES];

  for (uint32_t i = threadIdx.x; i < 2 * 2048; i += blockDim.x) s_vlc[i] = tables[i];
  if (threadIdx.x < 33) s_uvlc[threadIdx.x] = tables[2 * 2048 + threadIdx.x];
  __syncthreads();

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bidx = blockIdx.x * ENC_WARPS + warp;
  if (bidx >= nblocks) return;

  const EncBlock blk = blocks[bidx];
  const uint32_t width = blk.w, height = blk.h, stride = blk.stride, p = blk.p;
  const uint32_t* __restrict__ src = coef + blk.src_off;
  uint8_t* slot = slots + blk.slot_off;
  uint32_t* ms_buf = s_ms[warp];
  uint32_t* vl_buf = s_vl[warp];
  uint8_t* mel_buf = s_mel[warp];

  for (uint32_t i = lane; i < ENC_MS_WORDS; i += 32) ms_buf[i] = 0;
  vl_buf[lane] = 0;
  __syncwarp();

  // stream states (uniform across lanes)
  MelState mel; mel.k = 0; mel.run = 0; mel.tmp = 0; mel.rem = 8; mel.pos = 0;
  uint32_t ms_words = 0, ms_cbits = 0; bool ms_prev_ff = false; int ms_start = 0;
  uint32_t ms_last = 0;                 // last flushed MagSgn byte
  // VLC starts as if byte 0xFF was written and 4 bits 0xF are pending (vlc_init, :365-375)
  uint32_t vl_words = 0, vl_cbits = 12, vl_prev = 0; int vl_start = 0;
  if (lane == 0) vl_buf[0] = 0xFFFu;      // byte 0 = 0xFF, then the 4 bits 0xF
  __syncwarp();
  // bytes available: MagSgn grows up from 0, VLC grows down from slot_cap
  const uint32_t slot_words = blk.slot_cap >> 2;
  uint32_t any_sig = 0, negzero = 0;
  bool overflow = false;

  const uint32_t x = 2 * lane;
  const bool has0 = x < width, has1 = x + 1 < width;
  const uint32_t nquads = (width + 1) >> 1;
  // previous quad-row state of this lane
  uint32_t prev_rho = 0, prev_e1 = 0, prev_e3 = 0;

  // software prefetch of the first quad-row
  uint32_t n00 = 0, n01 = 0, n10 = 0, n11 = 0;
  // a lane's two samples of a row are one 8-byte load when the block rows are 8-byte aligned (always
  // so inside the codec: band planes are 16-byte aligned per block row)
  const bool vec2 = ((blk.src_off | stride) & 1u) == 0;
  auto load_pair = [&](const uint32_t* row, uint32_t& a, uint32_t& b) {
    if (vec2) { if (has0) { const uint2 t = *reinterpret_cast<const uint2*>(row + x); a = t.x; b = has1 ? t.y : 0u; } }
    else { if (has0) a = row[x]; if (has1) b = row[x + 1]; }
  };
  {
    load_pair(src, n00, n01);
    if (height > 1) load_pair(src + stride, n10, n11);
  }

  for (uint32_t y = 0; y < height; y += 2) {
    const uint32_t t0 = n00, t1 = n10, t2 = n01, t3 = n11;   // quad order: TL, BL, TR, BR
    n00 = n01 = n10 = n11 = 0;
    if (y + 2 < height) {
      const uint32_t* r0 = src + (size_t)(y + 2) * stride;
      load_pair(r0, n00, n01);
      if (y + 3 < height) load_pair(r0 + stride, n10, n11);
    }

    // ---- per-sample quantities (ojph_block_encoder.cpp:591-643)
    uint32_t rho = 0, e0 = 0, e1 = 0, e2 = 0, e3 = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    {
      uint32_t v;
      v = ((t0 + t0) >> p) & ~1u; if (v) { rho |= 1; --v; e0 = 32 - __clz((int)v); s0 = --v + (t0 >> 31); }
      v = ((t1 + t1) >> p) & ~1u; if (v) { rho |= 2; --v; e1 = 32 - __clz((int)v); s1 = --v + (t1 >> 31); }
      v = ((t2 + t2) >> p) & ~1u; if (v) { rho |= 4; --v; e2 = 32 - __clz((int)v); s2 = --v + (t2 >> 31); }
      v = ((t3 + t3) >> p) & ~1u; if (v) { rho |= 8; --v; e3 = 32 - __clz((int)v); s3 = --v + (t3 >> 31); }
    }
    any_sig |= rho;
    if (blk.flags & ENC_CHECK_NEGZERO)              // magnitude-overflow words keep the block coded (ojb_device.h)
      negzero |= (t0 == 0x80000000u) | (t1 == 0x80000000u) | (t2 == 0x80000000u) | (t3 == 0x80000000u);
    // nothing significant in this quad-row nor in the one above: every context is 0, every quad is one
    // MEL "0" event and no VLC / MagSgn bits (quantised high bands are mostly such rows)
    if (!__any_sync(FULL, (rho | prev_rho) != 0)) {
      mel_encode_events(mel, 0ull, nquads, mel_buf, lane);
      continue;
    }
    const uint32_t emax = max(max(e0, e1), max(e2, e3));

    // ---- neighbourhood: left quad of this row, the four samples above
    uint32_t rho_left = __shfl_up_sync(FULL, rho, 1); if (lane == 0) rho_left = 0;
    uint32_t pr_l = __shfl_up_sync(FULL, prev_rho | (prev_e3 << 8), 1); if (lane == 0) pr_l = 0;
    uint32_t pr_r = __shfl_down_sync(FULL, prev_rho | (prev_e1 << 8), 1); if (lane == 31) pr_r = 0;
    uint32_t kappa = 1, cq;
    if (y == 0) {
      cq = (rho_left >> 1) | (rho_left & 1);
    } else {
      // exponent predictor from max(E) of the 4 samples above, minus 1 (:862,:950)
      uint32_t me = max(max(pr_l >> 8, prev_e1), max(prev_e3, pr_r >> 8));
      uint32_t mem1 = me > 0 ? me - 1 : 0;
      kappa = (rho & (rho - 1)) ? max(1u, mem1) : 1u;
      // context: sigma(nw)|sigma(n-left) , sigma(w)|sigma(sw) , sigma(n-right)|sigma(ne)
      uint32_t a = ((pr_l >> 3) | (prev_rho >> 1)) & 1u;
      uint32_t b = ((rho_left >> 2) | (rho_left >> 3)) & 1u;
      uint32_t c = ((prev_rho >> 3) | (pr_r >> 1)) & 1u;
      cq = a | (b << 1) | (c << 2);
    }
    const uint32_t Uq = max(emax, kappa);
    const uint32_t uq = Uq - kappa;
    uint32_t eps = 0;
    if (uq > 0) eps = (e0 == emax ? 1u : 0u) | (e1 == emax ? 2u : 0u) | (e2 == emax ? 4u : 0u) | (e3 == emax ? 8u : 0u);
    const bool active = lane < nquads;
    const uint32_t tuple = s_vlc[(y ? 2048u : 0u) + (cq << 8) + (rho << 4) + eps];
    const uint32_t cwd = tuple >> 8, cwd_len = (tuple >> 4) & 7u, ek = tuple & 15u;

    // ---- MagSgn code of this quad (up to 4 x 31 bits) (:667-674)
    uint32_t m0 = (rho & 1) ? Uq - (ek & 1) : 0;
    uint32_t m1 = (rho & 2) ? Uq - ((ek >> 1) & 1) : 0;
    uint32_t m2 = (rho & 4) ? Uq - ((ek >> 2) & 1) : 0;
    uint32_t m3 = (rho & 8) ? Uq - ((ek >> 3) & 1) : 0;
    unsigned long long msA = (unsigned long long)(s0 & ((1u << m0) - 1u)) |
                             ((unsigned long long)(s1 & ((1u << m1) - 1u)) << m0);
    unsigned long long msB = (unsigned long long)(s2 & ((1u << m2) - 1u)) |
                             ((unsigned long long)(s3 & ((1u << m3) - 1u)) << m2);
    const uint32_t nA = m0 + m1, nB = m2 + m3;

    // ---- VLC bits of the quad pair (assembled on the even lane) (:661-662,763-785,985-988)
    const uint32_t o_uq = __shfl_down_sync(FULL, uq, 1);
    const uint32_t o_cw = __shfl_down_sync(FULL, cwd | (cwd_len << 8), 1);
    const uint32_t l_uq = __shfl_up_sync(FULL, uq, 1);
    uint32_t vbits = 0, vlen = 0;
    if (!(lane & 1) && active) {
      const bool has_q1 = lane + 1 < nquads;
      const uint32_t u0 = uq, u1 = has_q1 ? o_uq : 0;
      vbits = cwd; vlen = cwd_len;
      if (has_q1) { vbits |= (o_cw & 0xFF) << vlen; vlen += o_cw >> 8; }
      uint32_t c0, c1;
      if (y == 0 && u0 > 2 && u1 > 2) { c0 = s_uvlc[u0 - 2]; c1 = s_uvlc[u1 - 2]; }
      else if (y == 0 && u0 > 2 && u1 > 0) { c0 = s_uvlc[u0]; c1 = (u1 - 1) | (1u << 3); }   // 1-bit u1
      else { c0 = s_uvlc[u0]; c1 = s_uvlc[u1]; }
      uint32_t pl;
      pl = (c0 >> 3) & 7; vbits |= (c0 & 7) << vlen; vlen += pl;
      pl = (c1 >> 3) & 7; vbits |= (c1 & 7) << vlen; vlen += pl;
      pl = (c0 >> 11) & 31; vbits |= ((c0 >> 6) & 31) << vlen; vlen += pl;
      pl = (c1 >> 11) & 31; vbits |= ((c1 >> 6) & 31) << vlen; vlen += pl;
    }

    // ---- MEL events of this lane: own quad (context 0), then the pair event on odd lanes
    uint32_t nev = 0, evb = 0;
    if (active && cq == 0) { evb = (rho != 0) ? 1u : 0u; nev = 1; }
    if (y == 0 && (lane & 1) && active && l_uq > 0 && uq > 0) {
      evb |= (min(l_uq, uq) > 2 ? 1u : 0u) << nev; ++nev;
    }

    // ---- one packed exclusive scan: MagSgn bits | VLC bits << 12 | MEL events << 22
    uint32_t mine = (active ? (nA + nB) : 0) | (vlen << 12) | (nev << 22);
    uint32_t incl = mine;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(FULL, incl, d);
      if ((int)lane >= d) incl += o;
    }
    const uint32_t excl = incl - mine;
    const uint32_t tot = __shfl_sync(FULL, incl, 31);
    const uint32_t ms_tot = tot & 0xFFF, vl_tot = (tot >> 12) & 0x3FF, ev_tot = tot >> 22;

    // ---- scatter raw bits to the row buffers
    if (active) {
      uint32_t o = ms_cbits + (excl & 0xFFF);
      put_bits(ms_buf, o, msA, nA);
      put_bits(ms_buf, o + nA, msB, nB);
    }
    if (vlen) put_bits(vl_buf, vl_cbits + ((excl >> 12) & 0x3FF), vbits, vlen);
    {
      unsigned long long evs = (unsigned long long)evb << (excl >> 22);
      uint32_t lo = __reduce_or_sync(FULL, (uint32_t)evs);
      uint32_t hi = __reduce_or_sync(FULL, (uint32_t)(evs >> 32));
      mel_encode_events(mel, ((unsigned long long)hi << 32) | lo, ev_tot, mel_buf, lane);
    }
    __syncwarp();

    // ---- MagSgn: stuffing fix-up, flush complete words, keep the carry
    if (ms_tot) {
      uint32_t nbits = ms_stuff(ms_buf, ms_cbits + ms_tot, ms_start, ms_prev_ff, lane);
      uint32_t nw = nbits >> 5;
      if (ms_words + nw + vl_words + 24 >= slot_words) { overflow = true; break; }
      uint32_t* dst = reinterpret_cast<uint32_t*>(slot) + ms_words;
      const uint32_t carry = ms_buf[nw];
      const uint32_t lastw = nw ? ms_buf[nw - 1] : 0;
      __syncwarp();
      const uint32_t used = (nbits + 31) >> 5;
      ms_cbits = nbits & 31;
      // flush the complete words and clear the buffer in one sweep; word 0 restarts with the carry
      for (uint32_t i = lane; i <= used + 1 && i < ENC_MS_WORDS; i += 32) {
        if (i < nw) dst[i] = ms_buf[i];
        ms_buf[i] = (i == 0) ? (carry & ((ms_cbits ? (1u << ms_cbits) : 1u) - 1u)) : 0u;
      }
      if (nw) {
        ms_last = lastw >> 24; ms_words += nw;
        ms_start -= (int)(4 * nw);
        if (ms_start < 0) { ms_start = -1; ms_prev_ff = (ms_last == 0xFF); }
      }
      __syncwarp();
    }
    // ---- VLC: same, written backward from the end of the slot
    if (vl_tot) {
      uint32_t nbits = vlc_stuff(vl_buf, vl_cbits + vl_tot, vl_start, vl_prev, lane);
      uint32_t nw = nbits >> 5;
      if (ms_words + vl_words + nw + 24 >= slot_words) { overflow = true; break; }
      uint32_t* end = reinterpret_cast<uint32_t*>(slot) + slot_words;
      const uint32_t mine = vl_buf[lane];
      const uint32_t carry = __shfl_sync(FULL, mine, nw & 31);             // nw <= 31 here
      const uint32_t lastw = __shfl_sync(FULL, mine, (nw ? nw - 1 : 0) & 31);
      if (lane < nw) end[-(int)(vl_words + lane) - 1] = bytes_rev(mine);
      vl_cbits = nbits & 31;
      vl_buf[lane] = (lane == 0) ? (carry & ((vl_cbits ? (1u << vl_cbits) : 1u) - 1u)) : 0u;
      if (nw) { vl_prev = lastw >> 24; vl_words += nw; vl_start -= (int)(4 * nw); }
      __syncwarp();
    }

    prev_rho = rho; prev_e1 = e1; prev_e3 = e3;
  }

  any_sig = __reduce_or_sync(FULL, any_sig | negzero);
  if (overflow) {
    if (lane == 0) { atomicOr(status, 1u); results[bidx].len_head = 0; results[bidx].len_tail = 0; }
    return;
  }
  if (any_sig == 0) {                      // nothing significant: block is not included
    if (lane == 0) { results[bidx].len_head = 0; results[bidx].len_tail = 0; }
    return;
  }

  // ---- termination: a few bytes, done by lane 0 (terminate_mel_vlc :413-441, ms_terminate :517-533)
  __syncwarp();
  if (lane == 0) {
    // MagSgn: the carry holds < 32 bits; its complete bytes already carry their stuffing
    uint32_t ms_pos = ms_words * 4;
    {
      uint32_t acc = ms_buf[0], nb = ms_cbits, last = ms_last;
      bool have_last = ms_words > 0;
      while (nb >= 8) { last = acc & 0xFF; have_last = true; slot[ms_pos++] = (uint8_t)last; acc >>= 8; nb -= 8; }
      uint32_t cap = (have_last && last == 0xFF) ? 7u : 8u;
      if (nb) {
        uint32_t t = cap - nb;
        uint32_t byte = acc | ((0xFFu & ((1u << t) - 1u)) << nb);
        if (byte != 0xFF) slot[ms_pos++] = (uint8_t)byte;
      } else if (cap == 7) ms_pos--;
    }
    // MEL flush of a pending run
    if (mel.run > 0) mel_emit_bit(mel, 1, mel_buf, 0);
    uint32_t mel_tmp = (mel.tmp << mel.rem) & 0xFFu;
    uint32_t mel_mask = (0xFFu << mel.rem) & 0xFFu;
    // VLC: complete bytes of the carry are settled (stuffing applied); the rest is < 8 bits
    uint32_t vl_pos = vl_words * 4;
    uint8_t* vend = slot + blk.slot_cap;
    uint32_t acc = vl_buf[0], nb = vl_cbits;
    while (nb >= 8) { vend[-(int)(++vl_pos)] = (uint8_t)(acc & 0xFF); acc >>= 8; nb -= 8; }
    uint32_t vl_tmp = acc & 0xFFu;
    uint32_t vl_mask = 0xFFu >> (8 - nb);
    if ((mel_mask | vl_mask) != 0) {
      uint32_t fuse = mel_tmp | vl_tmp;
      if ((((fuse ^ mel_tmp) & mel_mask) | ((fuse ^ vl_tmp) & vl_mask)) == 0 && fuse != 0xFF && vl_pos > 1) {
        if (mel.pos < ENC_MEL_BYTES) mel_buf[mel.pos] = (uint8_t)fuse;
        mel.pos++;
      } else {
        if (mel.pos < ENC_MEL_BYTES) mel_buf[mel.pos] = (uint8_t)mel_tmp;
        mel.pos++;
        vend[-(int)(++vl_pos)] = (uint8_t)vl_tmp;
      }
    }
    if (mel.pos > 192 || ms_pos + mel.pos + vl_pos + 8 > blk.slot_cap) {
      atomicOr(status, mel.pos > 192 ? 2u : 1u);    // the reference errors out on MEL > 192 bytes
      results[bidx].len_head = 0; results[bidx].len_tail = 0;
    } else {
      for (uint32_t i = 0; i < mel.pos; ++i) slot[ms_pos + i] = mel_buf[i];
      uint32_t scup = mel.pos + vl_pos;
      vend[-1] = (uint8_t)(scup >> 4);
      vend[-2] = (uint8_t)((vend[-2] & 0xF0) | (scup & 0xF));
      results[bidx].len_head = ms_pos + mel.pos;
      results[bidx].len_tail = vl_pos;
    }
  }
}

void launch_ht_encode(const EncBlock* blocks, uint32_t nblocks, const uint32_t* coef, uint8_t* slots,
                      EncResult* results, const uint16_t* tables, uint32_t* status, cudaStream_t st)
{
  if (nblocks == 0) return;
  dim3 grid((nblocks + ENC_WARPS - 1) / ENC_WARPS), block(ENC_WARPS * 32);
  OJB_LAUNCH(ht_encode_kernel, grid, block, 0, st, blocks, nblocks, coef, slots, results, tables, status);
}

} // namespace ojb
