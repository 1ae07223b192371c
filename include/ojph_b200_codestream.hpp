// ojph_b200_codestream.hpp -- header-only C++ facade over the C-ABI (ojph_b200.h) with the class and
// method names of OpenJPH's public interface, so that code written against ojph::codestream
// (src/core/openjph/ojph_codestream.h:88-383), ojph::param_siz / param_cod / param_qcd
// (src/core/openjph/ojph_params.h) -- e.g. ojph_compress.cpp:1165-1203, ojph_expand.cpp:224-421 --
// switches to the B200 path by replacing `ojph::codestream` with `ojph::b200::codestream`.
//
// The including translation unit provides OpenJPH's own public headers first (they are not part of
// this repository):  ojph_base.h (ui8/ui32/si32, point, size), ojph_mem.h (line_buf), ojph_file.h
// (outfile_base, infile_base).  Only what the reference's apps call is mirrored; Part-2 items the
// hot path does not cover (DFS/ATK) raise the
// same kind of std::runtime_error the reference raises for invalid settings.
#pragma once
#include "ojph_b200.h"
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace ojph { namespace b200 {

inline void raise(const char* what) {           // OJPH_ERROR: message on stderr + std::runtime_error
  fprintf(stderr, "%s\n", what);
  throw std::runtime_error("ojph error");
}
inline void check(int rc) { if (rc != 0) raise(ojb_last_error()); }

struct state {                                   // what write_headers / read_headers consume
  ojb_params p;
  ojb_frame_info info;                           // decode side, after read_headers
  ojb_decoder* dec = nullptr;                    // read side: for the per-component getters
  bool reading = false;
  int planar = -1;
  uint32_t nlt_calls = 0, q_calls = 0;
  state() { ojb_params_default(&p); memset(&info, 0, sizeof(info)); }
};

inline ojb_coding_style state_style(const state* s, ui32 c) {
  ojb_coding_style cs; memset(&cs, 0, sizeof(cs));
  if (s->dec == nullptr) raise("ojph error: coding-style getters need read_headers first");
  check(ojb_dec_get_coding_style(s->dec, c, &cs));
  return cs;
}

class param_siz {                                // ojph_params.h:55-101
  state* s;
public:
  explicit param_siz(state* st) : s(st) {}
  void set_image_extent(point extent) { s->p.width = extent.x; s->p.height = extent.y; }
  void set_tile_size(size ts) { s->p.tile_w = ts.w; s->p.tile_h = ts.h; }
  void set_image_offset(point o) { s->p.off_x = o.x; s->p.off_y = o.y; }
  void set_tile_offset(point o) { s->p.tile_off_x = o.x; s->p.tile_off_y = o.y; }
  void set_num_components(ui32 n) {
    if (n == 0 || n > 16) raise("ojph error 0x00040005: this build supports 1..16 components");
    s->p.num_comps = n;
  }
  void set_component(ui32 c, const point& downsampling, ui32 bit_depth, bool is_signed) {
    if (c >= s->p.num_comps) raise("ojph error 0x00040002: component number is larger than the number of components");
    s->p.dx[c] = downsampling.x; s->p.dy[c] = downsampling.y;
    s->p.bit_depth[c] = bit_depth; s->p.is_signed[c] = is_signed ? 1u : 0u;
  }
  ojb_coding_style style(ui32 c) const { return state_style(s, c); }
  point get_image_extent() const { return s->reading ? point(s->info.width, s->info.height) : point(s->p.width, s->p.height); }
  point get_image_offset() const { return s->reading ? point(s->info.off_x, s->info.off_y) : point(s->p.off_x, s->p.off_y); }
  ui32 get_num_components() const { return s->reading ? s->info.num_comps : s->p.num_comps; }
  ui32 get_bit_depth(ui32 c) const { return s->reading ? s->info.bit_depth[c] : s->p.bit_depth[c]; }
  bool is_signed(ui32 c) const { return (s->reading ? s->info.is_signed[c] : s->p.is_signed[c]) != 0; }
  point get_downsampling(ui32 c) const { return s->reading ? point(s->info.dx[c], s->info.dy[c]) : point(s->p.dx[c], s->p.dy[c]); }
  size get_tile_size() const { if (!s->reading) return size(s->p.tile_w, s->p.tile_h); ojb_coding_style cs = style(0); return size(cs.tile_w, cs.tile_h); }
  point get_tile_offset() const { if (!s->reading) return point(s->p.tile_off_x, s->p.tile_off_y); ojb_coding_style cs = style(0); return point(cs.tile_off_x, cs.tile_off_y); }
  ui32 get_recon_width(ui32 c) const { return s->info.comp_w[c]; }
  ui32 get_recon_height(ui32 c) const { return s->info.comp_h[c]; }
};

class param_cod {                                // ojph_params.h:103-160
  state* s;
  static ui32 lg(ui32 v) { ui32 n = 0; while ((1u << n) < v) ++n; return n; }
  void coc(ui32 c) { if (c >= 16) raise("ojph error: per-component coding styles are limited to 16 components"); s->p.coc_present[c] = 1; }
public:
  explicit param_cod(state* st) : s(st) {}
  void set_num_decomposition(ui32 n) { s->p.num_decomps = n; }
  void set_block_dims(ui32 w, ui32 h) { s->p.block_w = w; s->p.block_h = h; }
  void set_precinct_size(int num_levels, size* precinct_size) {
    if (num_levels < 0 || num_levels > 33) raise("ojph error: too many precinct sizes");
    s->p.num_precincts = (uint32_t)num_levels;
    for (int i = 0; i < num_levels; ++i) { s->p.precinct_w[i] = precinct_size[i].w; s->p.precinct_h[i] = precinct_size[i].h; }
  }
  void set_progression_order(const char* name) {
    static const char* names[5] = { "LRCP", "RLCP", "RPCL", "PCRL", "CPRL" };
    for (uint32_t i = 0; i < 5; ++i) if (strcmp(name, names[i]) == 0) { s->p.prog_order = i; return; }
    raise("ojph error 0x00050031: unknown progression order");
  }
  void set_color_transform(bool on) { s->p.color_transform = on ? 1u : 0u; }
  void set_reversible(bool on) { s->p.reversible = on ? 1u : 0u; }
  // per-component coding styles (COC; ojph_params.cpp:255-281): the component's style starts from the
  // library defaults, as in the reference
  void set_num_decomposition(ui32 comp_idx, ui32 n) { coc(comp_idx); s->p.coc_num_decomps[comp_idx] = n; }
  void set_block_dims(ui32 comp_idx, ui32 w, ui32 h) { coc(comp_idx); s->p.coc_block_w[comp_idx] = w; s->p.coc_block_h[comp_idx] = h; }
  void set_reversible(ui32 comp_idx, bool on) { coc(comp_idx); s->p.coc_reversible[comp_idx] = on ? 1u : 0u; }
  void set_precinct_size(ui32 comp_idx, int num_levels, size* precinct_size) {
    coc(comp_idx);
    if (num_levels < 0 || num_levels > 33) raise("ojph error: too many precinct sizes");
    s->p.coc_num_precincts[comp_idx] = (uint32_t)num_levels;
    for (int i = 0; i < num_levels; ++i) { s->p.coc_precinct_w[comp_idx][i] = precinct_size[i].w; s->p.coc_precinct_h[comp_idx][i] = precinct_size[i].h; }
  }
  // read-side getters (after read_headers): the component's COC when it has one, else the COD
  size get_block_dims(ui32 c = 0) const { ojb_coding_style t = state_style(s, c); return size(t.block_w, t.block_h); }
  size get_log_block_dims(ui32 c = 0) const { ojb_coding_style t = state_style(s, c); return size(lg(t.block_w), lg(t.block_h)); }
  size get_precinct_size(ui32 level_num) const { ojb_coding_style t = state_style(s, 0); return size(t.precinct_w[level_num], t.precinct_h[level_num]); }
  size get_precinct_size(ui32 c, ui32 level_num) const { ojb_coding_style t = state_style(s, c); return size(t.precinct_w[level_num], t.precinct_h[level_num]); }
  size get_log_precinct_size(ui32 level_num) const { size p = get_precinct_size(level_num); return size(lg(p.w), lg(p.h)); }
  size get_log_precinct_size(ui32 c, ui32 level_num) const { size p = get_precinct_size(c, level_num); return size(lg(p.w), lg(p.h)); }
  int get_progression_order() const { return s->reading ? (int)state_style(s, 0).prog_order : (int)s->p.prog_order; }
  const char* get_progression_order_as_string() const {
    static const char* names[5] = { "LRCP", "RLCP", "RPCL", "PCRL", "CPRL" };
    return names[get_progression_order() % 5];
  }
  int get_num_layers() const { return s->reading ? (int)state_style(s, 0).num_layers : 1; }
  bool packets_may_use_sop() const { return s->reading && state_style(s, 0).may_use_sop != 0; }
  bool packets_use_eph() const { return s->reading && state_style(s, 0).use_eph != 0; }
  bool get_block_vertical_causality(ui32 c = 0) const { return s->reading && state_style(s, c).vertical_causality != 0; }
  ui32 get_num_decompositions(ui32 c) const { return s->reading ? state_style(s, c).num_decomps : (s->p.coc_present[c] ? s->p.coc_num_decomps[c] : s->p.num_decomps); }
  bool is_reversible(ui32 c) const { return s->reading ? state_style(s, c).reversible != 0 : (s->p.coc_present[c] ? s->p.coc_reversible[c] != 0 : s->p.reversible != 0); }
  ui32 get_num_decompositions() const { return s->reading ? s->info.num_decomps : s->p.num_decomps; }
  bool is_reversible() const { return (s->reading ? s->info.reversible : s->p.reversible) != 0; }
  bool is_using_color_transform() const { return (s->reading ? s->info.color_transform : s->p.color_transform) != 0; }
};

class param_qcd {                                // ojph_params.h:186-250
  state* s;
public:
  explicit param_qcd(state* st) : s(st) {}
  enum comp_type : ui8 { OJPH_COMP_Y = 0, OJPH_COMP_CB = 1, OJPH_COMP_CR = 2 };
  void set_irrev_quant(float delta) { s->p.qstep = delta; }
  void set_qfactor(ui8 qfactor) { s->p.qfactor = qfactor; }
  // per-component (QCC) forms, ojph_params.h:229-243; the call order is recorded (see ojb_params)
  void set_irrev_quant(ui32 comp_idx, float delta) {
    if (comp_idx >= 16) return;
    s->p.qcc_calls[comp_idx] |= 1u; s->p.qcc_qstep[comp_idx] = delta; s->p.qcc_qstep_seq[comp_idx] = ++s->q_calls;
  }
  void set_qfactor(ui32 comp_idx, comp_type ctype, ui8 qfactor) {
    if (comp_idx >= 16) return;
    s->p.qcc_calls[comp_idx] |= 2u; s->p.qcc_ctype[comp_idx] = ctype; s->p.qcc_qfactor[comp_idx] = qfactor;
    s->p.qcc_qfactor_seq[comp_idx] = ++s->q_calls;
  }
};

class param_nlt {                                // ojph_params.h:299-342
  state* s;
public:
  enum special_comp_num : ui16 { ALL_COMPS = 65535 };
  enum nonlinearity : ui8 { OJPH_NLT_NO_NLT = 0, OJPH_NLT_GAMMA_STYLE_NLT = 1, OJPH_NLT_LUT_STYLE_NLT = 2,
                            OJPH_NLT_BINARY_COMPLEMENT_NLT = 3, OJPH_NLT_UNDEFINED = 255 };
  explicit param_nlt(state* st) : s(st) {}
  void set_nonlinear_transform(ui32 comp_num, ui8 nl_type) {
    if (nl_type != OJPH_NLT_NO_NLT && nl_type != OJPH_NLT_BINARY_COMPLEMENT_NLT)
      raise("ojph error 0x00050171: Nonliearities other than type 0 (No Nonlinearity) or type  3 (Binary Binary "
            "Complement to Sign Magnitude Conversion) are not supported yet");
    if (comp_num == ALL_COMPS) { s->p.nlt_all = 1u + nl_type; return; }
    if (comp_num >= 16) return;                  // entries for non-existing components are dropped anyway
    if (s->p.nlt_comp[comp_num] == 0) s->p.nlt_seq[comp_num] = s->nlt_calls++;
    s->p.nlt_comp[comp_num] = 1u + nl_type;
  }
  // decode side, after read_headers: the transform in force for the component
  bool get_nonlinear_transform(ui32 comp_num, ui8& bit_depth, bool& is_signed, ui8& nl_type) const {
    if (!s->reading || comp_num >= s->info.num_comps) return false;
    bit_depth = (ui8)s->info.bit_depth[comp_num]; is_signed = s->info.is_signed[comp_num] != 0;
    nl_type = (ui8)s->info.nlt_type[comp_num];
    return nl_type != 0;
  }
};

class comment_exchange {                         // ojph_params.h:345-358
  friend class codestream;
  const char* data = nullptr;
  ui16 len = 0, Rcom = 0;
public:
  void set_string(const char* str) {
    size_t t = strlen(str);
    if (t > 65531) raise("ojph error 0x000500C1: COM marker string length cannot be larger than 65531");
    data = str; len = (ui16)t; Rcom = 1;
  }
  void set_data(const char* d, ui16 l) {
    if (l > 65531) raise("ojph error 0x000500C2: COM marker string length cannot be larger than 65531");
    data = d; len = l; Rcom = 0;
  }
};

class codestream {                               // ojph_codestream.h:88-383
  state st;
  ojb_encoder* enc = nullptr;
  ojb_decoder* dec = nullptr;
  outfile_base* out = nullptr;
  std::vector<ui8> j2c;                          // decode: the whole stream (the GPU path is frame based)
  line_buf lines[16];                            // one per component: callers keep the pointers of a row's components
  std::vector<si32> stage[16];                   // pulled rows are handed out from 64-byte aligned, padded storage, like
                                                 // the reference's own line buffers (its file writers use SIMD loads)
  si32* staged(ui32 c, const si32* row, ui32 w) {
    std::vector<si32>& v = stage[c & 15];
    if (v.size() < (size_t)w + 48) v.assign((size_t)w + 48, 0);
    si32* a = reinterpret_cast<si32*>((reinterpret_cast<size_t>(v.data()) + 63) & ~(size_t)63);
    memcpy(a, row, (size_t)w * sizeof(si32));
    return a;
  }
  bool resilient = false;
public:
  codestream() {}
  ~codestream() { close(); }
  codestream(const codestream&) = delete;
  codestream& operator=(const codestream&) = delete;

  param_siz access_siz() { return param_siz(&st); }
  param_cod access_cod() { return param_cod(&st); }
  param_qcd access_qcd() { return param_qcd(&st); }
  param_nlt access_nlt() { return param_nlt(&st); }

  // ---- write side
  void set_planar(bool planar) { st.planar = planar ? 1 : 0; }
  // (the library defaults when set_planar was not called: ojph_codestream_local.cpp:623 on the write side,
  // :879 on the read side)
  bool is_planar() const {
    if (st.planar >= 0) return st.planar != 0;
    return st.reading ? st.info.color_transform == 0 : st.p.color_transform != 0;
  }
  void set_profile(const char* name) {           // ojph_codestream_local.cpp:1124-1133; rules applied at write_headers
    if (name != nullptr && strcmp(name, "IMF") == 0) st.p.profile = 1;
    else if (name != nullptr && strcmp(name, "BROADCAST") == 0) st.p.profile = 2;
    else raise("ojph error 0x000300A1: unkownn or unsupported profile");
  }
  void set_tilepart_divisions(bool at_resolutions, bool at_components) {
    st.p.tilepart_div = (at_resolutions ? 1u : 0u) | (at_components ? 2u : 0u);
  }
  bool is_tilepart_division_at_resolutions() { return (st.p.tilepart_div & 1u) != 0; }
  bool is_tilepart_division_at_components() { return (st.p.tilepart_div & 2u) != 0; }
  void request_tlm_marker(bool needed) { st.p.tlm = needed ? 1u : 0u; }
  bool is_tlm_requested() { return st.p.tlm != 0; }
  // codestream::restart(): back to a freshly constructed object, device buffers kept for the next frame
  void restart() {
    if (out) { out->close(); out = nullptr; }
    ojb_encoder* e = enc; ojb_decoder* d = dec;
    st = state(); enc = e; dec = d; j2c.clear(); resilient = false;
  }
  void write_headers(outfile_base* file, const comment_exchange* comments = nullptr, ui32 num_comments = 0) {
    st.p.planar = st.planar;
    if (enc == nullptr) enc = ojb_enc_create();
    if (enc == nullptr) raise(ojb_last_error());
    std::vector<ojb_comment> com(num_comments);
    for (ui32 i = 0; i < num_comments; ++i) { com[i].data = comments[i].data; com[i].len = comments[i].len; com[i].rcom = comments[i].Rcom; }
    check(ojb_enc_set_comments(enc, com.data(), num_comments));
    check(ojb_enc_configure(enc, &st.p, OJB_I32));
    out = file;
  }
  // first call with NULL; returns the line to fill and the component it belongs to; NULL after the last line
  line_buf* exchange(line_buf* filled, ui32& next_component) {
    si32* p = ojb_enc_exchange(enc, filled ? filled->i32 : nullptr, &next_component);
    if (p == nullptr) return nullptr;
    line_buf& line = lines[next_component & 15];
    line.i32 = p; line.size = comp_width(next_component); line.pre_size = 0;
    line.flags = line_buf::LFT_32BIT | line_buf::LFT_INTEGER;
    return &line;
  }
  void flush() {
    uint64_t cap = 1u << 20;
    for (ui32 c = 0; c < st.p.num_comps; ++c) cap += (uint64_t)comp_width(c) * comp_height(c) * 4;
    std::vector<ui8> buf(cap);
    uint64_t n = 0;
    check(ojb_enc_flush(enc, buf.data(), buf.size(), &n));
    if (out->write(buf.data(), (size_t)n) != (size_t)n) raise("ojph error: could not write the codestream");
  }

  // ---- read side
  void enable_resilience() { resilient = true; }
  void read_headers(infile_base* file) {
    ui8 tmp[1 << 16];
    size_t k;
    j2c.clear();
    while ((k = file->read(tmp, sizeof(tmp))) > 0) j2c.insert(j2c.end(), tmp, tmp + k);
    if (dec == nullptr) dec = ojb_dec_create();
    if (dec == nullptr) raise(ojb_last_error());
    if (resilient) ojb_dec_enable_resilience(dec);
    check(ojb_dec_read_headers(dec, j2c.data(), j2c.size(), OJB_I32, &st.info));
    st.reading = true; st.dec = dec;
  }
  void restrict_input_resolution(ui32 skipped_res_for_data, ui32 skipped_res_for_recon) {
    check(ojb_dec_restrict_input_resolution(dec, skipped_res_for_data, skipped_res_for_recon, &st.info));
  }
  void create() {
    if (st.planar >= 0) ojb_dec_set_planar(dec, st.planar);
    check(ojb_dec_begin_pull(dec));
  }
  line_buf* pull(ui32& comp_num) {
    const si32* p = ojb_dec_pull(dec, &comp_num);
    if (p == nullptr) return nullptr;
    line_buf& line = lines[comp_num & 15];                      // (ppm_out keeps the three pointers until the row is complete)
    line.i32 = staged(comp_num, p, st.info.comp_w[comp_num]); line.size = st.info.comp_w[comp_num]; line.pre_size = 0;
    line.flags = line_buf::LFT_32BIT | line_buf::LFT_INTEGER;
    return &line;
  }

  void close() {
    if (enc) { ojb_enc_destroy(enc); enc = nullptr; }
    if (dec) { ojb_dec_destroy(dec); dec = nullptr; st.dec = nullptr; }
    if (out) { out->close(); out = nullptr; }
  }

private:
  static ui32 cdiv(ui32 a, ui32 b) { return (a + b - 1) / b; }
  ui32 comp_width(ui32 c) const { return cdiv(st.p.width, st.p.dx[c]) - cdiv(st.p.off_x, st.p.dx[c]); }
  ui32 comp_height(ui32 c) const { return cdiv(st.p.height, st.p.dy[c]) - cdiv(st.p.off_y, st.p.dy[c]); }
};

}} // namespace ojph::b200
