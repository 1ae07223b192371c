// ojb_params.cpp -- see ojb_params.h for the reference file:line each piece mirrors.
#include "ojb_params.h"
#include <cmath>
#include <cstring>
#include <algorithm>

namespace ojb {

#include "ojb_gain_tables.inc"

namespace {

inline float energy_gain_l(uint32_t d, bool rev) { return rev ? kEnergy_5x3_l[d] : kEnergy_9x7_l[d]; }
inline float energy_gain_h(uint32_t d, bool rev) { return rev ? kEnergy_5x3_h[d] : kEnergy_9x7_h[d]; }
inline float bibo_gain_l(uint32_t d, bool rev) { return rev ? kBibo_5x3_l[d] : kBibo_9x7_l[d]; }
inline float bibo_gain_h(uint32_t d, bool rev) { return rev ? kBibo_5x3_h[d] : kBibo_9x7_h[d]; }

// Qfactor model (ojph_params.cpp:604-726): visual weight tables by chroma format and
// component type, reference step from the quality factor.
const float* visual_weights(uint32_t sx, uint32_t sy, int ctype, bool& ok) {
  ok = true;
  int fmt;  // 1: 420, 2: 422, 3: 444
  if (sx == 2 && sy == 2) fmt = 1;
  else if (sx == 2 && sy == 1) fmt = 2;
  else if (sx == 1 && sy == 1) fmt = 3;
  else { ok = false; return kVw_y; }
  if (ctype == 0) return kVw_y;
  if (ctype == 1) return fmt == 1 ? kVw_cb420 : (fmt == 2 ? kVw_cb422 : kVw_cb444);
  return fmt == 1 ? kVw_cr420 : (fmt == 2 ? kVw_cr422 : kVw_cr444);
}
inline float vw_weight(const float* v, uint32_t level, uint32_t band) {
  if (band == 0) return v[18];
  level = std::min(level, 6u);
  return v[(level - 1) * 3 + (3 - band)];
}
inline float vw_gain(int ctype) {
  if (ctype == 0) return 1.0f;
  if (ctype == 1) return 1.8051f / 1.7321f;
  return 1.5734f / 1.7321f;
}
float vw_delta_ref(uint32_t qfactor, uint32_t bit_depth, float& power) {
  constexpr uint8_t t0 = 65, t1 = 97;
  constexpr float alpha_t0 = 0.04f, alpha_t1 = 0.10f;
  constexpr float m_t0 = 2.0f * (1.0f - t0 / 100.0f);
  constexpr float m_t1 = 2.0f * (1.0f - t1 / 100.0f);
  float m_q = qfactor < 50 ? 50.0f / (float)qfactor : 2.0f * (1.0f - (float)qfactor / 100.0f);
  float alpha_q;
  if (qfactor <= t0) { power = 1.0f; alpha_q = alpha_t0; }
  else if (qfactor < t1) {
    power = std::log(m_q) - std::log(m_t1);
    power /= std::log(m_t0) - std::log(m_t1);
    alpha_q = alpha_t1 * std::pow(alpha_t0 / alpha_t1, power);
  } else { power = 0.0f; alpha_q = alpha_t1; }
  const float eps = std::sqrt(0.5f) * std::ldexp(1.0f, -(int)bit_depth);
  return alpha_q * m_q + eps;
}

// exponent/mantissa packing of an irreversible step (encode_SPqcd, ojph_params.cpp:1602)
uint16_t pack_step(float delta) {
  int exp = 0;
  while (delta < 1.0f) { exp++; delta *= 2.0f; }
  int mantissa = (int)round(delta * (float)(1 << 11)) - (1 << 11);
  mantissa = mantissa < (1 << 11) ? mantissa : 0x7FF;
  return (uint16_t)((exp << 11) | mantissa);
}

} // namespace

//------------------------------------------------------------------------------------------
// QuantSet
//------------------------------------------------------------------------------------------
void QuantSet::set_rev_quant(uint32_t nd, uint32_t bd, bool color) {
  uint32_t B = bd + (color ? 1u : 0u);          // one extra bit after the RCT
  int s = 0;
  double bl = bibo_gain_l(nd, true);
  uint32_t X = (uint32_t)ceil(log(bl * bl) / M_LN2);
  uint8_t tmp[97];
  tmp[s++] = (uint8_t)(B + X);
  uint32_t max_BX = B + X;
  for (uint32_t d = nd; d > 0; --d) {
    double l = bibo_gain_l(d, true), h = bibo_gain_h(d - 1, true);
    X = (uint32_t)ceil(log(h * l) / M_LN2);
    tmp[s++] = (uint8_t)(B + X); max_BX = std::max(max_BX, B + X);
    tmp[s++] = (uint8_t)(B + X);
    X = (uint32_t)ceil(log(h * h) / M_LN2);
    tmp[s++] = (uint8_t)(B + X); max_BX = std::max(max_BX, B + X);
  }
  if (max_BX > 38)
    fail(0x00050151, "The specified combination of bit_depth, colour transform, and type of "
         "wavelet transform requires more than 38 bits; it requires %d bits.", max_BX);
  int guard = std::max(1, (int)max_BX - 31);
  Sqcd = (uint8_t)(guard << 5);
  for (int i = 0; i < s; ++i) SP[i] = (uint16_t)(uint8_t)((uint8_t)(tmp[i] - guard) << 3);
}

void QuantSet::set_irrev_quant(uint32_t nd) {
  Sqcd = (uint8_t)((1 << 5) | 0x2);             // one guard bit, scalar expounded
  float g_c = 1.0f, delta_ref = base_delta, power = 1.0f;
  const float* weights = kVw_no_weights;
  if (qfactor != 0) {
    bool ok;
    weights = visual_weights(sx, sy, ctype, ok);
    if (!ok)
      fail(0x00050161, "Qfactor can only be used on components with 4:4:4, 4:2:2 or 4:2:0 sampling");
    if (ctype == 0 && sx != 1 && sy != 1)
      fail(0x00050162, "Qfactor can only be used for a Y or luminance component when it is "
           "not downsampled.");
    g_c = vw_gain(ctype);
    delta_ref = vw_delta_ref(qfactor, bit_depth, power);
  }
  uint32_t b = 0;
  float gl = energy_gain_l(nd, false);
  float w_b = std::pow(vw_weight(weights, nd, 0), power);
  SP[b++] = pack_step(delta_ref / (gl * gl * g_c * w_b));
  for (uint32_t d = nd; d > 0; --d) {
    float l = energy_gain_l(d, false), h = energy_gain_h(d - 1, false);
    w_b = std::pow(vw_weight(weights, d, 1), power);
    SP[b++] = pack_step(delta_ref / (h * l * g_c * w_b));
    w_b = std::pow(vw_weight(weights, d, 2), power);
    SP[b++] = pack_step(delta_ref / (l * h * g_c * w_b));
    w_b = std::pow(vw_weight(weights, d, 3), power);
    SP[b++] = pack_step(delta_ref / (h * h * g_c * w_b));
  }
}

uint32_t QuantSet::kmax(uint32_t res, uint32_t band) const {
  uint32_t idx = res ? (res - 1) * 3 + band : 0;
  if (idx >= num_subbands) idx = num_subbands - 1;
  uint32_t bits = 0;
  if ((Sqcd & 0x1F) == 0) { bits = (uint32_t)(SP[idx] & 0xFF) >> 3; bits = bits ? bits - 1 : 0; }
  else bits = (uint32_t)(SP[idx] >> 11) - 1;
  return bits + guard_bits();
}

uint32_t QuantSet::largest_kmax() const {
  uint32_t bits = 0;
  for (uint32_t i = 0; i < num_subbands; ++i) {
    uint32_t t;
    if ((Sqcd & 0x1F) == 0) { t = (uint32_t)(SP[i] & 0xFF) >> 3; t = t ? t - 1 : 0; }
    else t = (uint32_t)(SP[i] >> 11) - 1;
    bits = std::max(bits, t);
  }
  return bits + guard_bits();
}

float QuantSet::irrev_delta(uint32_t res, uint32_t band) const {
  static const float arr[4] = { 1.0f, 2.0f, 2.0f, 4.0f };
  if ((Sqcd & 0x1F) != 2)
    fail(0x00050101, "irreversible transform with reversible (no quantization) step sizes");
  uint32_t idx = res ? (res - 1) * 3 + band : 0;
  if (idx >= num_subbands) idx = num_subbands - 1;
  int eps = SP[idx] >> 11;
  float mantissa = (float)((SP[idx] & 0x7FF) | 0x800) * arr[band];
  mantissa /= (float)(1 << 11);
  mantissa /= (float)(1u << eps);
  return mantissa;
}

//------------------------------------------------------------------------------------------
// Params
//------------------------------------------------------------------------------------------
const QuantSet& Params::quant_for(uint32_t c) const {
  for (const QuantSet& q : qcc) if (q.comp_idx == c) return q;
  return qcd;
}
QuantSet* Params::find_qcc(uint32_t c) {
  for (QuantSet& q : qcc) if (q.comp_idx == c) return &q;
  return nullptr;
}
QuantSet& Params::add_qcc(uint32_t c) {
  qcc.emplace_back();
  QuantSet& q = qcc.back();
  q.is_qcc = true; q.comp_idx = (uint16_t)c;
  return q;
}

uint32_t Params::precision(uint32_t c) const {
  uint32_t p = 0;
  if (color_transform() && c < 3) {
    for (uint32_t i = 0; i < 3; ++i) p = std::max(p, quant_for(i).largest_kmax());
  } else
    p = quant_for(c).largest_kmax();
  return p + 2;   // + sign bit + one spare bit (ojph_params.cpp:1700-1706)
}

bool Params::needs_wide() const {
  bool wide = false;
  for (uint32_t c = 0; c < num_comps(); ++c)
    if (precision(c) > 32) {
      wide = true;
      bool rev = reversible(c);
      if (color_transform() && c < 3) rev = reversible(0) && reversible(1) && reversible(2);
      if (!rev)
        fail(0x000B0001, "component %u needs %u-bit coefficients and is irreversible; the 64-bit path is built for "
             "reversible components only", c, precision(c));
    }
  return wide;
}

void Params::set_block_dims(uint32_t w, uint32_t h) {
  uint32_t lw = w ? ilog2(w) : 0, lh = h ? ilog2(h) : 0;
  if (w == 0 || w != (1u << lw) || h == 0 || h != (1u << lh) || lw < 2 || lh < 2 || lw + lh > 12)
    fail(0x00050011, "incorrect code block dimensions");
  cb_w_exp = (uint8_t)(lw - 2); cb_h_exp = (uint8_t)(lh - 2);
}

void Params::set_precincts(int n, const uint32_t* w, const uint32_t* h) {
  if (n == 0 || w == nullptr || h == nullptr) { Scod &= 0xFE; return; }
  Scod |= 1;
  for (int i = 0; i <= num_decomps; ++i) {
    uint32_t pw = w[i < n ? i : n - 1], ph = h[i < n ? i : n - 1];
    if (pw == 0 || ph == 0) fail(0x00050021, "precinct width or height cannot be 0");
    uint32_t PPx = ilog2(pw), PPy = ilog2(ph);
    if (pw != (1u << PPx) || ph != (1u << PPy))
      fail(0x00050022, "precinct width and height should be a power of 2");
    if (PPx > 15 || PPy > 15) fail(0x00050023, "precinct size is too large");
    if (i > 0 && (PPx == 0 || PPy == 0)) fail(0x00050024, "precinct size is too small");
    precinct_size[i] = (uint8_t)(PPx | (PPy << 4));
  }
}

static void make_quant_steps(QuantSet& q, uint32_t comp, const Params& p) {
  if (q.is_init) fail(0x00040001, "Quantization step sizes already initialized.");
  q.is_init = true;
  q.num_decomps = p.decomps(comp);              // the component's COC when it has one (:1432-1460)
  q.bit_depth = p.comps[comp].bit_depth;
  q.is_signed = p.comps[comp].is_signed;
  q.is_color_trans = p.color_transform();
  q.wavelet = p.wavelet_of(comp);
  q.sx = p.comps[comp].dx; q.sy = p.comps[comp].dy;
  q.num_subbands = 1 + 3 * q.num_decomps;
  if (q.wavelet == DWT_REV53)
    q.set_rev_quant(q.num_decomps, q.bit_depth, comp < 3 ? q.is_color_trans : false);
  else {
    if (q.base_delta == -1.0f) {
      uint32_t t = std::min(16u, q.bit_depth);
      q.base_delta = 1.0f / (float)(1 << t);
    }
    q.set_irrev_quant(q.num_decomps);
  }
}

void Params::finalize_for_encode() {
  // tile size 0x0 => one tile covering the image (ojph_codestream_local.cpp:562-570)
  if (XTsiz == 0 && YTsiz == 0) { XTsiz = Xsiz + XOsiz; YTsiz = Ysiz + YOsiz; }
  // SIZ validity (ojph_params_local.h:234-248)
  if (Xsiz == 0 || Ysiz == 0 || XTsiz == 0 || YTsiz == 0)
    fail(0x00040001, "Image extent and/or tile size cannot be zero");
  if (XTOsiz > XOsiz || YTOsiz > YOsiz)
    fail(0x00040002, "Tile offset has to be smaller than the image offset");
  if (XTsiz + XTOsiz <= XOsiz || YTsiz + YTOsiz <= YOsiz)
    fail(0x00040003, "The top left tile must intersect with the image");
  if (Xsiz <= XOsiz || Ysiz <= YOsiz)
    fail(0x00040004, "The image extent must be larger than the image offset");
  uint32_t nc = num_comps();
  if (nc == 0 || nc > 16384) fail(0x00040005, "wrong number of components");
  // COD validity (ojph_params_local.h:450-513)
  if (mc_trans == 1 && nc < 3)
    fail(0x00040011, "color transform can only be employed when the image has 3 or more "
         "color components");
  if (mc_trans == 1) {
    for (uint32_t i = 1; i < 3; ++i) {
      if (comps[i].dx != comps[0].dx || comps[i].dy != comps[0].dy)
        fail(0x00040012, "when color transform is used, the first 3 colour components must "
             "have the same downsampling factor.");
      if (comps[i].bit_depth != comps[0].bit_depth)
        fail(0x00040014, "when color transform is used, the first 3 colour components must "
             "have the same bit depth.");
      if (comps[i].is_signed != comps[0].is_signed)
        fail(0x00040015, "when color transform is used, the first 3 colour components must "
             "have the same signedness (signed or unsigned).");
    }
  }
  if (prog_order == PO_RPCL || prog_order == PO_PCRL)
    for (uint32_t i = 0; i < nc; ++i)
      if ((comps[i].dx & (comps[i].dx - 1)) || (comps[i].dy & (comps[i].dy - 1)))
        fail(0x00040013, "For RPCL and PCRL progression orders,component downsampling "
             "factors have to be powers of 2");

  // QCD / QCC (param_qcd::check_validity, ojph_params.cpp:1359-1431)
  for (QuantSet& q : qcc) q.enabled = q.comp_idx < nc;
  uint32_t qcd_comp = 0;
  // the first component that uses COD (no COC) and has no QCC (ojph_params.cpp:1367-1377)
  for (uint32_t c = 0; c < nc; ++c) if (find_coc(c) == nullptr && find_qcc(c) == nullptr) { qcd_comp = c; break; }
  if (qcd.qfactor != 0) {
    for (uint32_t i = 0; i < nc; ++i) {
      if (find_qcc(i) == nullptr) {
        QuantSet& q = add_qcc(i);
        q.qfactor = qcd.qfactor;
        q.ctype = (nc < 3) ? 0 : (int)(i < 3u ? i : 0u);
      }
    }
  }
  make_quant_steps(qcd, qcd_comp, *this);
  for (uint32_t c = 0; c < nc; ++c) {
    QuantSet* q = find_qcc(c);
    if (q == nullptr) {
      bool needed = qcd.num_decomps != decomps(c) || qcd.bit_depth != comps[c].bit_depth ||
                    qcd.is_signed != comps[c].is_signed ||
                    qcd.is_color_trans != color_transform() || qcd.wavelet != wavelet_of(c);
      if (!needed) continue;
      QuantSet& nq = add_qcc(c);
      nq.base_delta = qcd.base_delta;
      q = &nq;
    }
    make_quant_steps(*q, c, *this);
  }

  // CAP (ojph_params_local.h:982-998); MAGB over QCD and every QCC (ojph_params.cpp:1615)
  if (wavelet == DWT_REV53) Ccap0 &= 0xFFDF; else Ccap0 |= 0x0020;
  Ccap0 &= 0xFFE0;
  uint32_t B = 0;
  auto magb = [&](const QuantSet& q) {
    uint32_t nd = (q.num_subbands - 1) / 3;
    for (uint32_t i = 0; i < q.num_subbands; ++i) {
      uint32_t t;
      if ((q.Sqcd & 0x1F) == 0) t = ((uint32_t)(q.SP[i] & 0xFF) >> 3) + q.guard_bits() - 1u;
      else { uint32_t nb = nd - (i ? (i - 1) / 3 : 0); t = (uint32_t)(q.SP[i] >> 11) + q.guard_bits() - nb; }
      B = std::max(B, t);
    }
  };
  magb(qcd);
  for (const QuantSet& q : qcc) magb(q);
  uint32_t Bp = (B <= 8) ? 0 : (B < 28 ? B - 8 : 13 + (B >> 2));
  Ccap0 = (uint16_t)(Ccap0 | (uint16_t)Bp);

  (void)needs_wide();        // 64-bit coefficients: reversible components only (refuses the rest)

  nlt_check_validity();
  if (profile) check_profile();

  // tile-part division rules (ojph_codestream_local.cpp:583-621)
  if ((prog_order == PO_LRCP || prog_order == PO_RLCP) && tilepart_div == TP_COMP)
    tilepart_div |= TP_RES;
  if (prog_order == PO_RPCL && (tilepart_div & TP_COMP)) tilepart_div &= ~(uint32_t)TP_COMP;
  if (prog_order == PO_PCRL && tilepart_div != 0) tilepart_div = 0;
  if (prog_order == PO_CPRL && (tilepart_div & TP_RES)) tilepart_div &= ~(uint32_t)TP_RES;
  if (planar == -1) planar = color_transform() ? 1 : 0;
  else if (planar == 1 && color_transform())
    fail(0x00030021, "the planar interface option cannot be used when colour transform is "
         "employed");
}

//------------------------------------------------------------------------------------------
// NLT: type 3 (two's complement <-> sign-magnitude mapping of signed samples) is the only
// non-linearity the reference implements (ojph_params.cpp:2176-2190)
//------------------------------------------------------------------------------------------
void Params::set_nonlinear_transform(uint32_t comp, uint32_t type) {
  if (type != 0 && type != 3)
    fail(0x00050171, "Nonliearities other than type 0 (No Nonlinearity) or type  3 (Binary Binary Complement to "
         "Sign Magnitude Conversion) are not supported yet");
  NltEntry* e = nullptr;
  if (comp == 0xFFFF) e = &nlt_all;
  else {
    for (NltEntry& x : nlt) if (x.comp_idx == comp) e = &x;
    if (e == nullptr) { nlt.push_back(NltEntry()); e = &nlt.back(); e->comp_idx = (uint16_t)comp; }
  }
  e->Tnlt = (uint8_t)type; e->enabled = true;
}

void Params::nlt_check_validity() {          // param_nlt::check_validity, ojph_params.cpp:2087-2173
  if (!nlt_any()) return;
  const uint32_t nc = num_comps();
  auto own = [&](uint32_t c) -> NltEntry* { for (NltEntry& x : nlt) if (x.comp_idx == c) return &x; return nullptr; };
  auto bd_of = [&](uint32_t c) { return (uint8_t)((comps[c].bit_depth - 1) | (comps[c].is_signed ? 0x80 : 0)); };
  if (nlt_all.enabled && nlt_all.Tnlt == 0) nlt_all.enabled = false;
  if (nlt_all.enabled && nlt_all.Tnlt == 3) {
    bool all_same = true;
    uint32_t bit_depth = 0; bool is_signed = false;
    for (uint32_t c = 0; c < nc; ++c) {
      NltEntry* e = own(c);
      if (e == nullptr || !e->enabled) {
        if (bit_depth != 0) all_same = all_same && bit_depth == comps[c].bit_depth && is_signed == comps[c].is_signed;
        else { bit_depth = comps[c].bit_depth; is_signed = comps[c].is_signed; }
      } else e->BDnlt = bd_of(c);
    }
    if (all_same && bit_depth != 0) nlt_all.BDnlt = (uint8_t)((bit_depth - 1) | (is_signed ? 0x80 : 0));
    else if (!all_same) {
      nlt_all.enabled = false;
      for (uint32_t c = 0; c < nc; ++c) {
        NltEntry* e = own(c);
        if (e == nullptr || !e->enabled) {
          if (e == nullptr) { nlt.push_back(NltEntry()); e = &nlt.back(); e->comp_idx = (uint16_t)c; }
          e->enabled = true; e->Tnlt = 3; e->BDnlt = bd_of(c);
        }
      }
    }
  } else {
    for (uint32_t c = 0; c < nc; ++c) { NltEntry* e = own(c); if (e && e->enabled) e->BDnlt = bd_of(c); }
  }
  for (NltEntry& e : nlt) if (e.enabled && e.comp_idx >= nc) e.enabled = false;   // trim_non_existing_components
  if (nlt_any()) Rsiz |= 0x8000 | 0x0200;    // RSIZ_EXT_FLAG | RSIZ_NLT_FLAG
}

uint32_t Params::nlt_type(uint32_t c) const {  // get_nonlinear_transform + the check of ojph_tile.cpp:292-300
  const NltEntry* e = nullptr;
  for (const NltEntry& x : nlt) if (x.comp_idx == c && x.enabled) e = &x;
  if (e == nullptr && nlt_all.enabled) e = &nlt_all;
  if (e == nullptr) return 0;
  uint32_t bd = (uint32_t)(e->BDnlt & 0x7F) + 1; if (bd > 38) bd = 38;
  const bool is = (e->BDnlt & 0x80) != 0;
  if (bd != comps[c].bit_depth || is != comps[c].is_signed)
    fail(0x000300A1, "Mismatch between Ssiz (bit_depth = %d, is_signed = %s) from SIZ marker segment, and BDnlt "
         "(bit_depth = %d, is_signed = %s) from NLT marker segment, for component %d", comps[c].bit_depth,
         comps[c].is_signed ? "True" : "False", bd, is ? "True" : "False", c);
  return e->Tnlt == 3 ? 3u : 0u;
}

//------------------------------------------------------------------------------------------
// IMF / BROADCAST profiles: parameter rules only -- but they also switch TLM on and force tile-part
// division by components, so they change the bytes (check_imf_validity :292-455,
// check_broadcast_validity :458-535)
//------------------------------------------------------------------------------------------
void Params::check_profile() {
  const bool imf = profile == 1;
  const char* nm = imf ? "IMF" : "broadcast";
  const uint32_t base = imf ? 0x000300C0u : 0x000300B0u;       // the reference's error code blocks
  const uint32_t nc = num_comps();
  const bool rev = reversible();
  // (the reference's frame-size tests for the 2K / 4K / 8K IMF levels are written `flag &= true` and never
  // reject anything, :302-330; the levels only differ in the decomposition ceilings below)
  bool lvl[3] = { true, true, true };
  if (XOsiz || YOsiz) fail(base + (imf ? 3 : 1), "For %s profile, image offset (XOsiz, YOsiz) has to be 0.", nm);
  if (XTOsiz || YTOsiz) fail(base + (imf ? 4 : 2), "For %s profile, tile offset (XTOsiz, YTOsiz) has to be 0.", nm);
  if (nc > (imf ? 3u : 4u))
    fail(base + (imf ? 5 : 3), "For %s profile, the number of components has to be less  or equal to %d", nm, imf ? 3 : 4);
  bool plain = true, x422 = true;                                // 4:4:4, or 4:2:2 on components 1 and 2
  for (uint32_t c = 0; c < nc; ++c) {
    if (comps[c].dy != 1) plain = x422 = false;
    if (comps[c].dx != 1) plain = false;
    if (comps[c].dx != ((c == 1 || c == 2) ? 2u : 1u)) x422 = false;
  }
  if (!plain && !x422)
    fail(base + (imf ? 6 : 4), "For %s profile, either no component downsampling is used, or the x-dimension of the 2nd "
         "and 3rd components is downsampled by 2.", nm);
  for (uint32_t c = 0; c < nc; ++c)
    if (comps[c].bit_depth < 8 || comps[c].bit_depth > (imf ? 16u : 12u) || comps[c].is_signed)
      fail(base + (imf ? 7 : 5), "For %s profile, compnent bit_depth has to be between 8 and %d bits inclusively, and the "
           "samples must be unsigned", nm, imf ? 16 : 12);
  if (imf) {
    if (log_cb_w() != 5 || log_cb_h() != 5)
      fail(0x000300C8, "For IMF profile, codeblock dimensions are restricted. Use \"-block_size {32,32}\" at the commandline");
  } else {
    if (num_decomps == 0 || num_decomps > 5)
      fail(0x000300B6, "For broadcast profile, number of decompositions has to be between1 and 5 inclusively.");
    if (log_cb_w() < 5 || log_cb_w() > 7)
      fail(0x000300B7, "For broadcast profile, codeblock dimensions are restricted such that codeblock width has to be "
           "either 32, 64, or 128.");
    if (log_cb_h() < 5 || log_cb_h() > 7)
      fail(0x000300B8, "For broadcast profile, codeblock dimensions are restricted such that codeblock height has to be "
           "either 32, 64, or 128.");
  }
  // precincts: 128x128 at the coarsest resolution, 256x256 above (the reference's loop keeps only the test of
  // the LAST resolution when there is more than one -- reproduced)
  bool pz = log_pp_w(0) == 7 && log_pp_h(0) == 7;
  for (uint32_t r = 1; r <= num_decomps; ++r) pz = log_pp_w(r) == 8 && log_pp_h(r) == 8;
  if (!pz) fail(base + 9, "For %s profile, precinct sizes are restricted. Use \"-precincts {128,128},{256,256}\" at the commandline", nm);
  if (prog_order != PO_CPRL)
    fail(base + 10, "For %s profile, the CPRL progression order must be used. Use \"-prog_order CPRL\".", nm);
  const uint32_t ntiles = div_ceil(Xsiz, XTsiz) * div_ceil(Ysiz, YTsiz);
  if (imf) {
    for (int i = 0; i < 3; ++i) lvl[i] = num_decomps <= 5u + (uint32_t)i;
    if (num_decomps == 0 || !(lvl[0] || lvl[1] || lvl[2]))
      fail(0x000300CB, "Number of decompositions does not match the IMF profile dictated by wavelet reversibility and "
           "image dimensions.");
    if (ntiles > 1) {
      if (!rev) fail(0x000300CC, "Lossy IMF profile must have one tile.");
      // square tiles of 1024 / 2048 / 4096, and enough tile width for the decomposition count
      const uint32_t tw = XTsiz, th = YTsiz;
      auto deep = [&](int top) {        // top: how many (width, levels) pairs the level allows
        static const uint32_t w[4] = { 1024, 2048, 4096, 8192 };
        for (int i = 0; i <= top; ++i) if (tw >= w[i] && num_decomps <= 4u + (uint32_t)i) return true;
        return false;
      };
      auto square = [&](int top) { for (int i = 0; i <= top; ++i) if (tw == (1024u << i) && th == (1024u << i)) return true; return false; };
      const bool ok2 = lvl[0] && square(0) && deep(1), ok4 = lvl[1] && square(1) && deep(2), ok8 = lvl[2] && square(2) && deep(3);
      if (!ok2 && !ok4 && !ok8)
        fail(0x000300CD, "Number of decompositions does not match the IMF profile dictated by wavelet reversibility and "
             "image dimensions and tiles.");
    }
  } else if (ntiles != 1 && ntiles != 4)
    fail(0x000300BB, "The broadcast profile can only have 1 or 4 tiles");
  need_tlm = true;
  tilepart_div = TP_COMP;                   // (the reference warns when it has to correct this)
}

static void write_quant(std::vector<uint8_t>& o, const QuantSet& q, uint32_t nc) {
  int irrev = q.Sqcd & 0x1F;
  uint32_t L = q.is_qcc ? (4 + (nc < 257 ? 0 : 1)) : 3;
  L += (irrev == 0 ? 1 : 2) * q.num_subbands;
  put_u16(o, q.is_qcc ? M_QCC : M_QCD);
  put_u16(o, L);
  if (q.is_qcc) { if (nc < 257) put_u8(o, q.comp_idx); else put_u16(o, q.comp_idx); }
  put_u8(o, q.Sqcd);
  for (uint32_t i = 0; i < q.num_subbands; ++i)
    if (irrev == 0) put_u8(o, q.SP[i] & 0xFF); else put_u16(o, q.SP[i]);
}

void Params::write_main_header(std::vector<uint8_t>& o, const char* const* comments,
                               const uint32_t* comment_lens, uint32_t n_comments) const {
  uint32_t nc = num_comps();
  put_u16(o, M_SOC);
  // SIZ
  put_u16(o, M_SIZ); put_u16(o, 38 + 3 * nc); put_u16(o, Rsiz);
  put_u32(o, Xsiz); put_u32(o, Ysiz); put_u32(o, XOsiz); put_u32(o, YOsiz);
  put_u32(o, XTsiz); put_u32(o, YTsiz); put_u32(o, XTOsiz); put_u32(o, YTOsiz);
  put_u16(o, nc);
  for (uint32_t c = 0; c < nc; ++c) {
    put_u8(o, (uint32_t)(comps[c].bit_depth - 1) + (comps[c].is_signed ? 0x80u : 0u));
    put_u8(o, comps[c].dx); put_u8(o, comps[c].dy);
  }
  // CAP
  put_u16(o, M_CAP); put_u16(o, 8); put_u32(o, Pcap); put_u16(o, Ccap0);
  // COD
  put_u16(o, M_COD);
  put_u16(o, 12 + ((Scod & 1) ? 1 + num_decomps : 0));
  put_u8(o, Scod); put_u8(o, prog_order); put_u16(o, num_layers); put_u8(o, mc_trans);
  put_u8(o, num_decomps); put_u8(o, cb_w_exp); put_u8(o, cb_h_exp); put_u8(o, block_style);
  put_u8(o, wavelet);
  if (Scod & 1) for (int i = 0; i <= num_decomps; ++i) put_u8(o, precinct_size[i]);
  // COC, in creation order (param_cod::write_coc / internal_write_coc, ojph_params.cpp:1081-1143)
  for (const CodStyle& c : coc) {
    if (c.comp_idx >= nc) continue;
    put_u16(o, M_COC);
    put_u16(o, (nc < 257 ? 9u : 10u) + ((c.Scoc & 1) ? 1u + c.num_decomps : 0u));
    if (nc < 257) put_u8(o, c.comp_idx); else put_u16(o, c.comp_idx);
    put_u8(o, c.Scoc); put_u8(o, c.num_decomps); put_u8(o, c.cb_w_exp); put_u8(o, c.cb_h_exp);
    put_u8(o, c.block_style); put_u8(o, c.wavelet);
    if (c.Scoc & 1) for (int i = 0; i <= c.num_decomps; ++i) put_u8(o, c.precinct_size[i]);
  }
  // QCD, QCC
  write_quant(o, qcd, nc);
  for (const QuantSet& q : qcc) if (q.enabled) write_quant(o, q, nc);
  // NLT: the default entry, then the per-component ones in creation order (param_nlt::write, :2210-2235)
  auto put_nlt = [&](const NltEntry& e) {
    if (!e.enabled) return;
    put_u16(o, M_NLT); put_u16(o, 6); put_u16(o, e.comp_idx); put_u8(o, e.BDnlt); put_u8(o, e.Tnlt);
  };
  put_nlt(nlt_all);
  for (const NltEntry& e : nlt) put_nlt(e);
  // COM: library signature, needed for byte-identical output
  // (ojph_codestream_local.cpp:668-686)
  static const char sig[] = "OpenJPH Ver 0.31.0.";
  put_u16(o, M_COM); put_u16(o, (uint32_t)strlen(sig) + 4); put_u16(o, 1);
  o.insert(o.end(), sig, sig + strlen(sig));
  for (uint32_t i = 0; i < n_comments; ++i) {
    put_u16(o, M_COM); put_u16(o, comment_lens[i] + 4); put_u16(o, 1);
    o.insert(o.end(), comments[i], comments[i] + comment_lens[i]);
  }
  for (const Comment& c : this->comments) {  // write_headers(file, comments, num_comments), :688-706
    put_u16(o, M_COM); put_u16(o, (uint32_t)c.data.size() + 4); put_u16(o, c.Rcom);
    o.insert(o.end(), c.data.begin(), c.data.end());
  }
}

//------------------------------------------------------------------------------------------
// main-header parsing
//------------------------------------------------------------------------------------------
namespace {
struct Reader {
  const uint8_t* d; size_t n, pos;
  bool has(size_t k) const { return pos + k <= n; }
  uint32_t u8() { if (!has(1)) fail(0x00030051, "File ended before finding a tile segment"); return d[pos++]; }
  uint32_t u16() { uint32_t a = u8(); return (a << 8) | u8(); }
  uint32_t u32() { uint32_t a = u16(); return (a << 16) | u16(); }
};

void read_quant(Reader& r, QuantSet& q, bool is_qcc, uint32_t nc) {
  uint32_t L = r.u16();
  uint32_t hdr = 3;
  q.is_qcc = is_qcc;
  if (is_qcc) {
    if (nc < 257) { q.comp_idx = (uint16_t)r.u8(); hdr = 4; }
    else { q.comp_idx = (uint16_t)r.u16(); hdr = 5; }
  }
  q.Sqcd = (uint8_t)r.u8();
  int style = q.Sqcd & 0x1F;
  if (style == 0) {
    q.num_subbands = L - hdr;
    if (q.num_subbands == 0 || q.num_subbands > 97)
      fail(0x00050083, "wrong Lqcd value of %d in QCD marker", L);
    for (uint32_t i = 0; i < q.num_subbands; ++i) q.SP[i] = (uint16_t)r.u8();
  } else if (style == 2) {
    q.num_subbands = (L - hdr) / 2;
    if (q.num_subbands == 0 || q.num_subbands > 97 || L != hdr + 2 * q.num_subbands)
      fail(0x00050086, "wrong Lqcd value of %d in QCD marker", L);
    for (uint32_t i = 0; i < q.num_subbands; ++i) q.SP[i] = (uint16_t)r.u16();
  } else if (style == 1)
    fail(0x00050089, "Scalar derived quantization is not supported yet in QCD marker");
  else
    fail(0x00050088, "wrong Sqcd value in QCD marker");
  q.is_init = true;
}
} // namespace

size_t Params::read_main_header(const uint8_t* data, size_t len) {
  Reader r{ data, len, 0 };
  // find SOC then SIZ (find_marker, ojph_codestream_local.cpp:706-730: after a 0xFF the next byte is consumed
  // whether it matches or not; read_headers ignores a failed search, so a stream without SOC / SIZ ends up
  // in param_siz::read at the end of the data, :855-857)
  auto find = [&](uint16_t m) {
    while (r.pos < r.n) {
      if (r.d[r.pos++] != 0xFF) continue;
      if (r.pos >= r.n) return false;
      if (r.d[r.pos++] == (m & 0xFF)) return true;
    }
    return false;
  };
  if (!find(M_SOC) || !find(M_SIZ) || !r.has(2)) fail(0x00050041, "error reading SIZ marker");
  // SIZ
  {
    uint32_t L = r.u16();
    int nc = ((int)L - 38) / 3;
    if ((int)L != 38 + 3 * nc) fail(0x00050042, "error in SIZ marker length");
    Rsiz = (uint16_t)r.u16();
    if ((Rsiz & 0x4000) == 0) fail(0x00050044, "Rsiz bit 14 is not set (this is not a JPH file)");
    Xsiz = r.u32(); Ysiz = r.u32(); XOsiz = r.u32(); YOsiz = r.u32();
    XTsiz = r.u32(); YTsiz = r.u32(); XTOsiz = r.u32(); YTOsiz = r.u32();
    uint32_t C = r.u16();
    if ((int)C != nc) fail(0x0005004E, "Csiz does not match the SIZ marker size");
    if (C == 0) fail(0x0005004F, "Wrong Csiz value of 0 in SIZ marker segment");
    comps.resize(C);
    for (uint32_t c = 0; c < C; ++c) {
      uint32_t s = r.u8(), xr = r.u8(), yr = r.u8();
      if ((s & 0x7F) > 37) fail(0x00050054, "Wrong SIZ-SSiz value of %d", s);
      if (xr == 0) fail(0x00050055, "Wrong SIZ-XRsiz value of %d", xr);
      if (yr == 0) fail(0x00050056, "Wrong SIZ-YRsiz value of %d", yr);
      comps[c].bit_depth = (uint8_t)((s & 0x7F) + 1);
      comps[c].is_signed = (s & 0x80) != 0;
      comps[c].dx = (uint8_t)xr; comps[c].dy = (uint8_t)yr;
    }
    if (Xsiz == 0 || Ysiz == 0 || XTsiz == 0 || YTsiz == 0)
      fail(0x00040001, "Image extent and/or tile size cannot be zero");
    if (XTOsiz > XOsiz || YTOsiz > YOsiz)
      fail(0x00040002, "Tile offset has to be smaller than the image offset");
    if (XTsiz + XTOsiz <= XOsiz || YTsiz + YTOsiz <= YOsiz)
      fail(0x00040003, "The top left tile must intersect with the image");
    if (Xsiz <= XOsiz || Ysiz <= YOsiz)
      fail(0x00040004, "The image extent must be larger than the image offset");
    if (Rsiz & 0x00A0)
      fail(0x000B0002, "codestreams that need ATK/DFS (Part 2 wavelet structures) are not "
           "supported by the GPU path");
  }
  int received = 0;
  qcc.clear(); coc.clear(); nlt.clear(); nlt_all = NltEntry();
  for (;;) {
    // scan to the next 0xFF xx marker of interest (the reference skips unknown bytes too)
    if (r.pos + 1 >= r.n) fail(0x00030051, "File ended before finding a tile segment");
    if (r.d[r.pos] != 0xFF) { ++r.pos; continue; }
    uint16_t m = (uint16_t)(0xFF00 | r.d[r.pos + 1]);
    bool known = m == M_CAP || m == M_PRF || m == M_CPF || m == M_COD || m == M_COC || m == M_QCD ||
                 m == M_QCC || m == M_RGN || m == M_POC || m == M_PPM || m == M_TLM || m == M_PLM ||
                 m == M_CRG || m == M_COM || m == M_DFS || m == M_ATK || m == M_NLT || m == M_SOT;
    if (!known) { ++r.pos; continue; }
    if (m == M_SOT) break;
    r.pos += 2;
    if (m == M_CAP) {
      uint32_t L = r.u16();
      Pcap = r.u32();
      if (Pcap & 0xFFFDFFFF) fail(0x00050063, "error Pcap in CAP has options that are not supported");
      if ((Pcap & 0x00020000) == 0)
        fail(0x00050064, "error Pcap should have its 15th MSB set, Pcap^15.  This is not a JPH file");
      Ccap0 = (uint16_t)r.u16();
      if (L != 8) fail(0x00050066, "error in CAP marker length");
    } else if (m == M_COD) {
      uint32_t L = r.u16();
      Scod = (uint8_t)r.u8(); prog_order = (uint8_t)r.u8(); num_layers = (uint16_t)r.u16();
      mc_trans = (uint8_t)r.u8(); num_decomps = (uint8_t)r.u8(); cb_w_exp = (uint8_t)r.u8();
      cb_h_exp = (uint8_t)r.u8(); block_style = (uint8_t)r.u8(); wavelet = (uint8_t)r.u8();
      if (num_decomps > 32 || cb_w_exp > 8 || cb_h_exp > 8 || cb_w_exp + cb_h_exp > 8 ||
          (block_style & 0x40) != 0x40 || (block_style & 0xB7) != 0x00)
        fail(0x0005007D, "wrong settings in a COD-SPcod parameter");
      if (wavelet > 1)
        fail(0x000B0003, "arbitrary transformation kernels (ATK) are not supported by the GPU path");
      if (Scod & 1)
        for (int i = 0; i <= num_decomps; ++i) {
          precinct_size[i] = (uint8_t)r.u8();
          if (i && ((precinct_size[i] & 0xF) == 0 || (precinct_size[i] >> 4) == 0))
            fail(0x0005007F, "Precinct width or height for resolutions other than the coarsest "
                 "must be larger than 1");
        }
      if (L != 12u + ((Scod & 1) ? 1u + num_decomps : 0u)) fail(0x0005007C, "error in COD segment length");
      if (num_layers != 1)
        fail(0x00030053, "The current implementation supports 1 quality layer only.  This "
             "codestream has %d quality layers", num_layers);
      if (prog_order > 4) fail(0x0005007D, "wrong settings in a COD-SPcod parameter");
      received |= 1;
    } else if (m == M_QCD) {
      read_quant(r, qcd, false, num_comps());
      received |= 2;
    } else if (m == M_QCC) {
      QuantSet q;
      read_quant(r, q, true, num_comps());
      if (q.comp_idx >= num_comps())
        fail(0x00030054, "The codestream carries a QCC marker segment for a component indexed "
             "by %d, which is more than the allowed index number, since the codestream has %d "
             "components", q.comp_idx, num_comps());
      if (find_qcc(q.comp_idx))
        fail(0x00030055, "The codestream has two QCC marker segments for one component of "
             "index %d", q.comp_idx);
      qcc.push_back(q);
    } else if (m == M_COC) {                    // param_cod::read_coc, ojph_params.cpp:1207-1290
      uint32_t L = r.u16();
      const uint32_t nc = num_comps();
      CodStyle c;
      c.comp_idx = (uint16_t)(nc < 257 ? r.u8() : r.u16());
      c.Scoc = (uint8_t)r.u8(); c.num_decomps = (uint8_t)r.u8(); c.cb_w_exp = (uint8_t)r.u8();
      c.cb_h_exp = (uint8_t)r.u8(); c.block_style = (uint8_t)r.u8(); c.wavelet = (uint8_t)r.u8();
      if (c.comp_idx >= nc)
        fail(0x00030056, "The codestream carries a COC marker segment for a component indexed by %d, which is "
             "more than the allowed index number, since the codestream has %d components", c.comp_idx, nc);
      if (find_coc(c.comp_idx))
        fail(0x00030057, "The codestream has two COC marker segments for one component of index %d", c.comp_idx);
      if (c.num_decomps > 32 || c.cb_w_exp > 8 || c.cb_h_exp > 8 || c.cb_w_exp + c.cb_h_exp > 8 ||
          (c.block_style & 0x40) != 0x40 || (c.block_style & 0xB7) != 0x00)
        fail(0x0005012D, "wrong settings in a COC-SPcoc parameter");
      if (c.wavelet > 1)
        fail(0x000B0003, "arbitrary transformation kernels (ATK) are not supported by the GPU path");
      if (c.Scoc & 1)
        for (int i = 0; i <= c.num_decomps; ++i) {
          c.precinct_size[i] = (uint8_t)r.u8();
          if (i && ((c.precinct_size[i] & 0xF) == 0 || (c.precinct_size[i] >> 4) == 0))
            fail(0x0005012E, "Precinct width or height for resolutions other than the coarsest must be larger than 1");
        }
      if (L != (nc < 257 ? 9u : 10u) + ((c.Scoc & 1) ? 1u + c.num_decomps : 0u))
        fail(0x0005012F, "error in COC segment length");
      coc.push_back(c);
    } else if (m == M_NLT) {                // param_nlt::read, ojph_params.cpp:2238-2266
      if (!r.has(6)) fail(0x00050141, "error reading NLT marker segment");
      uint32_t L = r.u16(), comp = r.u16(), bd = r.u8(), t = r.u8();
      if (L != 6 || (t != 3 && t != 0)) fail(0x00050145, "Unsupported NLT type %d\n", t);
      NltEntry* e = nullptr;
      if (comp == 0xFFFF) e = &nlt_all;
      else {
        for (NltEntry& x : nlt) if (x.comp_idx == comp) e = &x;
        if (e == nullptr) { nlt.push_back(NltEntry()); e = &nlt.back(); }
      }
      e->enabled = true; e->comp_idx = (uint16_t)comp; e->BDnlt = (uint8_t)bd; e->Tnlt = (uint8_t)t;
    } else if (m == M_DFS || m == M_ATK)
      fail(0x000B0005, "DFS/ATK marker segments are not supported by the GPU path");
    else {  // PRF CPF RGN POC PPM TLM PLM CRG COM: skipped
      uint32_t L = r.u16();
      if (L < 2 || !r.has(L - 2)) fail(0x00030041, "error reading marker");
      r.pos += L - 2;
    }
  }
  if (received != 3) fail(0x00030052, "markers error, COD and QCD are required");
  return r.pos;
}

} // namespace ojb
