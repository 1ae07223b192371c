// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked by the product).
//
// A thin extern "C" harness over the UNMODIFIED reference library
// (aous72/OpenJPH v0.31.0, compiled from /root/reference by oracle/Makefile into
// oracle/_ref/libopenjph_ref.so).  It exposes, with plain pointers and sizes:
//   * whole-image encode / decode through the reference's public API
//     (ojph::codestream + mem_outfile / mem_infile, src/core/openjph/ojph_codestream.h:88-383,
//      ojph_file.h:126,308) -- the same call sequence as ojph_compress.cpp:1165-1203 and
//      ojph_expand.cpp:389-421;
//   * the reference's exported kernels for kernel-level parity:
//     ojph_encode_codeblock32 (ojph_block_encoder.cpp:542), ojph_decode_codeblock32
//     (ojph_block_decoder32.cpp:742), the dispatched (SIMD) variants through
//     codeblock_fun (ojph_codeblock_fun.h:93-120), the generic lifting / colour / quantise
//     line functions (ojph_transform.cpp:209-855, ojph_colour.cpp:238-571,
//     ojph_codestream_gen.cpp:59-168).
// Every entry point catches the reference's exceptions and returns a negative int.
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
#include <stdexcept>

#include "ojph_params.h"
#define private public   // test harness only: reach param_atk::init_rev53/init_irv97
#include "ojph_params_local.h"
#undef private

#include "ojph_arch.h"
#include "ojph_file.h"
#include "ojph_mem.h"
#include "ojph_params.h"
#include "ojph_codestream.h"
#include "ojph_message.h"
#include "ojph_block_encoder.h"
#include "ojph_block_decoder.h"
#include "ojph_codeblock_fun.h"
#include "ojph_transform.h"
#include "ojph_colour.h"
#include "ojph_colour_local.h"
#include "ojph_img_io.h"       // src/apps/common: the apps' file readers / writers (N2 oracle)

using namespace ojph;

namespace ojph { namespace local {
  // generic (scalar) kernels -- exported from the library, declared in *_local.h
  void gen_rev_vert_step(const lifting_step*, const line_buf*, const line_buf*,
                         const line_buf*, ui32, bool);
  void gen_rev_horz_ana(const param_atk*, const line_buf*, const line_buf*,
                        const line_buf*, ui32, bool);
  void gen_rev_horz_syn(const param_atk*, const line_buf*, const line_buf*,
                        const line_buf*, ui32, bool);
  void gen_irv_vert_step(const lifting_step*, const line_buf*, const line_buf*,
                         const line_buf*, ui32, bool);
  void gen_irv_vert_times_K(float, const line_buf*, ui32);
  void gen_irv_horz_ana(const param_atk*, const line_buf*, const line_buf*,
                        const line_buf*, ui32, bool);
  void gen_irv_horz_syn(const param_atk*, const line_buf*, const line_buf*,
                        const line_buf*, ui32, bool);
  void gen_rct_forward(const line_buf*, const line_buf*, const line_buf*,
                       line_buf*, line_buf*, line_buf*, ui32);
  void gen_rct_backward(const line_buf*, const line_buf*, const line_buf*,
                        line_buf*, line_buf*, line_buf*, ui32);
  void gen_ict_forward(const float*, const float*, const float*,
                       float*, float*, float*, ui32);
  void gen_ict_backward(const float*, const float*, const float*,
                        float*, float*, float*, ui32);
  void gen_irv_convert_to_float(const line_buf*, ui32, line_buf*, ui32, bool, ui32);
  void gen_irv_convert_to_integer(const line_buf*, line_buf*, ui32, ui32, bool, ui32);
  void gen_rev_tx_to_cb32(const void*, ui32*, ui32, float, ui32, ui32*);
  void gen_irv_tx_to_cb32(const void*, ui32*, ui32, float, ui32, ui32*);
  void gen_rev_tx_from_cb32(const ui32*, void*, ui32, float, ui32);
  void gen_irv_tx_from_cb32(const ui32*, void*, ui32, float, ui32);
}}

extern "C" {

struct ojr_params {
  uint32_t width, height;          // image extent (Xsiz, Ysiz)
  uint32_t off_x, off_y;           // image offset (XOsiz, YOsiz)
  uint32_t tile_w, tile_h;         // 0,0 => one tile
  uint32_t tile_off_x, tile_off_y;
  uint32_t num_comps;
  uint32_t bit_depth[16];
  uint32_t is_signed[16];
  uint32_t dx[16], dy[16];         // component sub-sampling
  uint32_t num_decomps;
  uint32_t block_w, block_h;       // e.g. 64,64
  uint32_t num_precincts;          // 0 => default 2^15
  uint32_t precinct_w[33], precinct_h[33];
  uint32_t reversible;
  uint32_t color_transform;
  uint32_t prog_order;             // 0 LRCP 1 RLCP 2 RPCL 3 PCRL 4 CPRL
  float    qstep;                  // <=0 => library default
  uint32_t qfactor;                // 0 => unset
  uint32_t tlm;                    // request TLM marker
  uint32_t tilepart_div;           // 0 none, 1 resolutions, 2 components, 3 both
  int32_t  planar;                 // -1 => library default
  // per-component coding styles (COC), same meaning as in include/ojph_b200.h
  uint32_t coc_present[16], coc_reversible[16], coc_num_decomps[16], coc_block_w[16], coc_block_h[16];
  uint32_t coc_num_precincts[16], coc_precinct_w[16][33], coc_precinct_h[16][33];
  // NLT: 0 = not set, else 1 + type; nlt_seq = order of the per-component calls
  uint32_t nlt_all, nlt_comp[16], nlt_seq[16];
  uint32_t profile;                // 0 none, 1 IMF, 2 BROADCAST
  // per-component quantisation calls (see include/ojph_b200.h)
  uint32_t qcc_calls[16]; float qcc_qstep[16]; uint32_t qcc_qstep_seq[16];
  uint32_t qcc_qfactor[16], qcc_ctype[16], qcc_qfactor_seq[16];
};

static char g_err[512] = "";
const char* ojr_last_error() { return g_err; }
// the reference reports errors through a message object (code, file, line, text) and then throws
// std::runtime_error("ojph error"): keep the code and the text of the last one for the error-parity tests
static uint32_t g_err_code = 0;
static char g_err_text[512] = "";
class capture_error : public ojph::message_error {
public:
  virtual void operator()(int error_code, const char* file_name, int line_num, const char* fmt, ...) {
    (void)file_name; (void)line_num;
    g_err_code = (uint32_t)error_code;
    va_list args; va_start(args, fmt); vsnprintf(g_err_text, sizeof(g_err_text), fmt, args); va_end(args);
    throw std::runtime_error("ojph error");
  }
};
static capture_error g_capture;
static ojph::message_error g_plain;          // the library's own behaviour (prints, then throws)
extern "C" void ojr_capture_errors(int on) { ojph::configure_error(on ? (ojph::message_error*)&g_capture : &g_plain); g_err_code = 0; g_err_text[0] = 0; }
extern "C" uint32_t ojr_last_error_code() { return g_err_code; }
extern "C" const char* ojr_last_error_text() { return g_err_text; }
int ojr_cpu_ext_level() { return ojph::get_cpu_ext_level(); }

static const char* po_name(uint32_t po) {
  static const char* n[] = { "LRCP", "RLCP", "RPCL", "PCRL", "CPRL" };
  return n[po < 5 ? po : 2];
}

// planes[c] : int32 samples, row stride = comp width, as the caller would hand to
// ojph::codestream::exchange (unsigned samples un-shifted).
// extra COM segments for the next ojr_encode calls (write_headers(file, comments, num_comments))
static std::vector<std::string> g_comment_data;
static std::vector<int> g_comment_text;
void ojr_set_comments(const char* const* data, const uint16_t* lens, const uint16_t* is_text, uint32_t n) {
  g_comment_data.clear(); g_comment_text.clear();
  for (uint32_t i = 0; i < n; ++i) { g_comment_data.emplace_back(data[i], data[i] + lens[i]); g_comment_text.push_back(is_text[i]); }
}

int ojr_encode(const ojr_params* p, const int32_t* const* planes,
               uint8_t* out, uint64_t out_cap, uint64_t* out_len)
{
  try {
    ojph::codestream cs;
    param_siz siz = cs.access_siz();
    siz.set_image_extent(point(p->width, p->height));
    siz.set_num_components(p->num_comps);
    for (ui32 c = 0; c < p->num_comps; ++c)
      siz.set_component(c, point(p->dx[c], p->dy[c]), p->bit_depth[c],
                        p->is_signed[c] != 0);
    siz.set_image_offset(point(p->off_x, p->off_y));
    siz.set_tile_size(size(p->tile_w, p->tile_h));
    siz.set_tile_offset(point(p->tile_off_x, p->tile_off_y));
    param_cod cod = cs.access_cod();
    cod.set_num_decomposition(p->num_decomps);
    cod.set_block_dims(p->block_w, p->block_h);
    if (p->num_precincts) {
      size ps[33];
      for (ui32 i = 0; i < p->num_precincts && i < 33; ++i)
        ps[i] = size(p->precinct_w[i], p->precinct_h[i]);
      cod.set_precinct_size((int)p->num_precincts, ps);
    }
    cod.set_progression_order(po_name(p->prog_order));
    cod.set_color_transform(p->color_transform != 0);
    cod.set_reversible(p->reversible != 0);
    for (uint32_t c = 0; c < p->num_comps; ++c)
      if (p->coc_present[c]) {
        cod.set_num_decomposition(c, p->coc_num_decomps[c]);
        cod.set_block_dims(c, p->coc_block_w[c], p->coc_block_h[c]);
        cod.set_reversible(c, p->coc_reversible[c] != 0);
        if (p->coc_num_precincts[c]) {
          size ps[33];
          for (uint32_t i = 0; i < p->coc_num_precincts[c] && i < 33; ++i) ps[i] = size(p->coc_precinct_w[c][i], p->coc_precinct_h[c][i]);
          cod.set_precinct_size(c, (int)p->coc_num_precincts[c], ps);
        }
      }
    if (p->nlt_all) cs.access_nlt().set_nonlinear_transform(param_nlt::ALL_COMPS, (ui8)(p->nlt_all - 1));
    for (uint32_t k = 0; k < 16; ++k)            // per-component calls in their recorded order
      for (uint32_t c = 0; c < 16; ++c)
        if (p->nlt_comp[c] && p->nlt_seq[c] == k) cs.access_nlt().set_nonlinear_transform(c, (ui8)(p->nlt_comp[c] - 1));
    if (!p->reversible) {
      if (p->qstep > 0.0f) cs.access_qcd().set_irrev_quant(p->qstep);
      if (p->qfactor) cs.access_qcd().set_qfactor((ui8)p->qfactor);
    }
    for (uint32_t k = 1; k <= 32; ++k)              // the per-component calls in their recorded order
      for (uint32_t c = 0; c < 16; ++c) {
        if ((p->qcc_calls[c] & 1) && p->qcc_qstep_seq[c] == k) cs.access_qcd().set_irrev_quant(c, p->qcc_qstep[c]);
        if ((p->qcc_calls[c] & 2) && p->qcc_qfactor_seq[c] == k)
          cs.access_qcd().set_qfactor(c, ojph::param_qcd::ui8_2_comp_type((ui8)p->qcc_ctype[c]), (ui8)p->qcc_qfactor[c]);
      }
    // the apps always set planar explicitly (ojph_compress.cpp:763,868,1021); the
    // library default (planar = colour_transform, ojph_codestream_local.cpp:623) is unusable
    cs.set_planar(p->planar > 0);
    if (p->profile) cs.set_profile(p->profile == 1 ? "IMF" : "BROADCAST");
    if (p->tlm) cs.request_tlm_marker(true);
    if (p->tilepart_div) cs.set_tilepart_divisions((p->tilepart_div & 1) != 0,
                                                   (p->tilepart_div & 2) != 0);

    mem_outfile f;
    f.open(1 << 20);
    std::vector<comment_exchange> com(g_comment_data.size());
    for (size_t i = 0; i < com.size(); ++i) {
      if (g_comment_text[i]) com[i].set_string(g_comment_data[i].c_str());
      else com[i].set_data(g_comment_data[i].data(), (ui16)g_comment_data[i].size());
    }
    cs.write_headers(&f, com.empty() ? NULL : com.data(), (ui32)com.size());

    ui32 nc = p->num_comps;
    std::vector<ui32> cw(nc), ch(nc), row(nc, 0);
    for (ui32 c = 0; c < nc; ++c) {
      cw[c] = ojph_div_ceil(p->width, p->dx[c]) - ojph_div_ceil(p->off_x, p->dx[c]);
      ch[c] = ojph_div_ceil(p->height, p->dy[c]) - ojph_div_ceil(p->off_y, p->dy[c]);
    }
    ui32 next_comp;
    line_buf* cur = cs.exchange(NULL, next_comp);
    while (cur) {
      ui32 c = next_comp;
      memcpy(cur->i32, planes[c] + (size_t)row[c] * cw[c], (size_t)cw[c] * 4);
      row[c]++;
      cur = cs.exchange(cur, next_comp);
    }
    cs.flush();
    size_t n = (size_t)f.tell();
    *out_len = n;
    int rc = 0;
    if (n > out_cap) rc = -2; else memcpy(out, f.get_data(), n);
    cs.close();
    return rc;
  } catch (const std::exception& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.what()); return -1;
  } catch (const char* s) {
    snprintf(g_err, sizeof(g_err), "%s", s); return -1;
  }
}

struct ojr_info {
  uint32_t width, height, off_x, off_y, num_comps;
  uint32_t bit_depth[16], is_signed[16], dx[16], dy[16];
  uint32_t comp_w[16], comp_h[16];
  uint32_t num_decomps, reversible, color_transform;
};

static void fill_info(ojph::codestream& cs, ojr_info* info) {
  param_siz siz = cs.access_siz();
  param_cod cod = cs.access_cod();
  info->width = siz.get_image_extent().x; info->height = siz.get_image_extent().y;
  info->off_x = siz.get_image_offset().x; info->off_y = siz.get_image_offset().y;
  info->num_comps = siz.get_num_components();
  for (ui32 c = 0; c < info->num_comps && c < 16; ++c) {
    info->bit_depth[c] = siz.get_bit_depth(c);
    info->is_signed[c] = siz.is_signed(c);
    info->dx[c] = siz.get_downsampling(c).x; info->dy[c] = siz.get_downsampling(c).y;
    info->comp_w[c] = siz.get_recon_width(c); info->comp_h[c] = siz.get_recon_height(c);
  }
  info->num_decomps = cod.get_num_decompositions();
  info->reversible = cod.is_reversible();
  info->color_transform = cod.is_using_color_transform();
}

int ojr_read_info(const uint8_t* j2c, uint64_t len, ojr_info* info)
{
  try {
    ojph::codestream cs; mem_infile f; f.open(j2c, (size_t)len);
    cs.read_headers(&f); fill_info(cs, info); cs.close(); return 0;
  } catch (const std::exception& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.what()); return -1;
  } catch (const char* s) { snprintf(g_err, sizeof(g_err), "%s", s); return -1; }
}

// planes[c] : caller-allocated int32[comp_w*comp_h]
static int decode_impl(const uint8_t* j2c, uint64_t len, int32_t* const* planes, int resilient,
                       uint32_t skip_read, uint32_t skip_recon, ojr_info* info_out);

int ojr_decode(const uint8_t* j2c, uint64_t len, int32_t* const* planes, int resilient)
{ return decode_impl(j2c, len, planes, resilient, 0, 0, NULL); }

// restrict_input_resolution(skip_read, skip_recon); planes == NULL: only report the reconstruction sizes
int ojr_decode_restricted(const uint8_t* j2c, uint64_t len, int32_t* const* planes, int resilient,
                          uint32_t skip_read, uint32_t skip_recon, ojr_info* info)
{ return decode_impl(j2c, len, planes, resilient, skip_read, skip_recon, info); }

static int decode_impl(const uint8_t* j2c, uint64_t len, int32_t* const* planes, int resilient,
                       uint32_t skip_read, uint32_t skip_recon, ojr_info* info_out)
{
  try {
    ojph::codestream cs; mem_infile f; f.open(j2c, (size_t)len);
    if (resilient) cs.enable_resilience();
    cs.read_headers(&f);
    if (skip_read || skip_recon) cs.restrict_input_resolution(skip_read, skip_recon);
    ojr_info info; fill_info(cs, &info);
    if (info_out) *info_out = info;
    if (planes == NULL) { cs.close(); return 0; }
    // keep the library's own choice (read_headers: planar = !colour_transform,
    // ojph_codestream_local.cpp:879): interleaved pulls stall on sub-sampled components
    cs.create();
    ui32 nc = info.num_comps;
    std::vector<ui32> row(nc, 0);
    // interleaved pull: rows x comps (sub-sampled comps have fewer rows)
    ui32 maxh = 0; for (ui32 c = 0; c < nc; ++c) maxh = ojph_max(maxh, info.comp_h[c]);
    ui64 total = 0; for (ui32 c = 0; c < nc; ++c) total += info.comp_h[c];
    for (ui64 i = 0; i < total; ) {
      ui32 c;
      line_buf* l = cs.pull(c);
      if (l == NULL) break;
      if (row[c] < info.comp_h[c]) {
        memcpy(planes[c] + (size_t)row[c] * info.comp_w[c], l->i32,
               (size_t)info.comp_w[c] * 4);
        row[c]++; ++i;
      }
    }
    cs.close();
    return 0;
  } catch (const std::exception& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.what()); return -1;
  } catch (const char* s) { snprintf(g_err, sizeof(g_err), "%s", s); return -1; }
}

//---------------------------------------------------------------------------------------
// block coder kernels
//---------------------------------------------------------------------------------------
// variant: 0 = generic scalar (ojph_encode_codeblock32), 1 = what codeblock_fun dispatches
int ojr_encode_block32(uint32_t* buf, uint32_t missing_msbs, uint32_t w, uint32_t h,
                       uint32_t stride, uint8_t* out, uint32_t out_cap,
                       uint32_t* out_len, int variant)
{
  try {
    static mem_elastic_allocator* elastic = new mem_elastic_allocator(1 << 20);
    local::initialize_block_encoder_tables();
    coded_lists* coded = NULL;
    ui32 lengths[2] = { 0, 0 };
    if (variant == 0)
      local::ojph_encode_codeblock32(buf, missing_msbs, 1, w, h, stride, lengths,
                                     elastic, coded);
    else {
      local::codeblock_fun fun; fun.init(true);
      fun.encode_cb32(buf, missing_msbs, 1, w, h, stride, lengths, elastic, coded);
    }
    *out_len = lengths[0];
    int rc = 0;
    if (lengths[0] > out_cap) rc = -2; else memcpy(out, coded->buf, lengths[0]);
    elastic->restart();
    return rc;
  } catch (const std::exception& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.what()); return -1;
  }
}

// coded must have >= 8 readable bytes before and 16 zero bytes after (harness copies)
int ojr_decode_block32(const uint8_t* coded, uint32_t* out, uint32_t missing_msbs,
                       uint32_t num_passes, uint32_t len1, uint32_t len2,
                       uint32_t w, uint32_t h, uint32_t stride, int causal, int variant)
{
  try {
    std::vector<ui8> tmp(8 + (size_t)len1 + len2 + 64, 0);
    memcpy(tmp.data() + 8, coded, (size_t)len1 + len2);
    bool ok;
    if (variant == 0)
      ok = local::ojph_decode_codeblock32(tmp.data() + 8, out, missing_msbs, num_passes,
                                          len1, len2, w, h, stride, causal != 0);
    else {
      local::codeblock_fun fun; fun.init(true);
      ok = fun.decode_cb32(tmp.data() + 8, out, missing_msbs, num_passes, len1, len2,
                           w, h, stride, causal != 0);
    }
    return ok ? 0 : 1;
  } catch (const std::exception& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.what()); return -1;
  }
}

//---------------------------------------------------------------------------------------
// line kernels (generic scalar variants), exposed on raw pointers
//---------------------------------------------------------------------------------------
static void wrap_i32(line_buf& l, int32_t* p, size_t n) {
  l.size = n; l.pre_size = 0; l.flags = line_buf::LFT_32BIT | line_buf::LFT_INTEGER; l.i32 = p;
}
static void wrap_f32(line_buf& l, float* p, size_t n) {
  l.size = n; l.pre_size = 0; l.flags = line_buf::LFT_32BIT; l.f32 = p;
}
static local::param_atk* atk53() {
  static local::param_atk a; static bool init = false;
  if (!init) { a.init_rev53(); init = true; } return &a;
}
static local::param_atk* atk97() {
  static local::param_atk a; static bool init = false;
  if (!init) { a.init_irv97(); init = true; } return &a;
}

// one lifting step over lines; step index s as param_atk::get_step(s)
void ojr_rev_vert_step(int s, const int32_t* sig, const int32_t* other, int32_t* aug,
                       uint32_t repeat, int synthesis)
{
  line_buf a, b, c; wrap_i32(a, (int32_t*)sig, repeat); wrap_i32(b, (int32_t*)other, repeat);
  wrap_i32(c, aug, repeat);
  local::gen_rev_vert_step(atk53()->get_step((ui32)s), &a, &b, &c, repeat, synthesis != 0);
}
void ojr_irv_vert_step(int s, const float* sig, const float* other, float* aug,
                       uint32_t repeat, int synthesis)
{
  line_buf a, b, c; wrap_f32(a, (float*)sig, repeat); wrap_f32(b, (float*)other, repeat);
  wrap_f32(c, aug, repeat);
  local::gen_irv_vert_step(atk97()->get_step((ui32)s), &a, &b, &c, repeat, synthesis != 0);
}
float ojr_irv_K() { return atk97()->get_K(); }
float ojr_irv_step(int s) { return atk97()->get_step((ui32)s)->irv.Aatk; }

// horizontal analysis / synthesis of one line; ldst/hdst/src need 1 sample of slack
// before and after (the reference writes lp[-1], lp[w]); harness copies into padded bufs.
void ojr_rev_horz_ana(int32_t* ldst, int32_t* hdst, const int32_t* src, uint32_t width,
                      int even)
{
  std::vector<int32_t> L(width + 4), H(width + 4), S(src, src + width);
  line_buf l, h, s; wrap_i32(l, L.data() + 1, width + 1); wrap_i32(h, H.data() + 1, width + 1);
  wrap_i32(s, S.data(), width);
  local::gen_rev_horz_ana(atk53(), &l, &h, &s, width, even != 0);
  ui32 lw = (width + (even ? 1 : 0)) >> 1, hw = (width + (even ? 0 : 1)) >> 1;
  memcpy(ldst, L.data() + 1, lw * 4); memcpy(hdst, H.data() + 1, hw * 4);
}
void ojr_rev_horz_syn(int32_t* dst, const int32_t* lsrc, const int32_t* hsrc,
                      uint32_t width, int even)
{
  ui32 lw = (width + (even ? 1 : 0)) >> 1, hw = (width + (even ? 0 : 1)) >> 1;
  std::vector<int32_t> L(width + 4), H(width + 4), D(width + 4);
  memcpy(L.data() + 1, lsrc, lw * 4); memcpy(H.data() + 1, hsrc, hw * 4);
  line_buf l, h, d; wrap_i32(l, L.data() + 1, width + 1); wrap_i32(h, H.data() + 1, width + 1);
  wrap_i32(d, D.data(), width);
  local::gen_rev_horz_syn(atk53(), &d, &l, &h, width, even != 0);
  memcpy(dst, D.data(), width * 4);
}
void ojr_irv_horz_ana(float* ldst, float* hdst, const float* src, uint32_t width, int even)
{
  std::vector<float> L(width + 4), H(width + 4), S(src, src + width);
  line_buf l, h, s; wrap_f32(l, L.data() + 1, width + 1); wrap_f32(h, H.data() + 1, width + 1);
  wrap_f32(s, S.data(), width);
  local::gen_irv_horz_ana(atk97(), &l, &h, &s, width, even != 0);
  ui32 lw = (width + (even ? 1 : 0)) >> 1, hw = (width + (even ? 0 : 1)) >> 1;
  memcpy(ldst, L.data() + 1, lw * 4); memcpy(hdst, H.data() + 1, hw * 4);
}
void ojr_irv_horz_syn(float* dst, const float* lsrc, const float* hsrc, uint32_t width,
                      int even)
{
  ui32 lw = (width + (even ? 1 : 0)) >> 1, hw = (width + (even ? 0 : 1)) >> 1;
  std::vector<float> L(width + 4), H(width + 4), D(width + 4);
  memcpy(L.data() + 1, lsrc, lw * 4); memcpy(H.data() + 1, hsrc, hw * 4);
  line_buf l, h, d; wrap_f32(l, L.data() + 1, width + 1); wrap_f32(h, H.data() + 1, width + 1);
  wrap_f32(d, D.data(), width);
  local::gen_irv_horz_syn(atk97(), &d, &l, &h, width, even != 0);
  memcpy(dst, D.data(), width * 4);
}
void ojr_irv_vert_times_K(float K, float* aug, uint32_t repeat)
{ line_buf a; wrap_f32(a, aug, repeat); local::gen_irv_vert_times_K(K, &a, repeat); }

void ojr_rct_forward(const int32_t* r, const int32_t* g, const int32_t* b,
                     int32_t* y, int32_t* cb, int32_t* cr, uint32_t n)
{
  line_buf lr, lg, lb, ly, lcb, lcr;
  wrap_i32(lr, (int32_t*)r, n); wrap_i32(lg, (int32_t*)g, n); wrap_i32(lb, (int32_t*)b, n);
  wrap_i32(ly, y, n); wrap_i32(lcb, cb, n); wrap_i32(lcr, cr, n);
  local::gen_rct_forward(&lr, &lg, &lb, &ly, &lcb, &lcr, n);
}
void ojr_rct_backward(const int32_t* y, const int32_t* cb, const int32_t* cr,
                      int32_t* r, int32_t* g, int32_t* b, uint32_t n)
{
  line_buf lr, lg, lb, ly, lcb, lcr;
  wrap_i32(lr, r, n); wrap_i32(lg, g, n); wrap_i32(lb, b, n);
  wrap_i32(ly, (int32_t*)y, n); wrap_i32(lcb, (int32_t*)cb, n); wrap_i32(lcr, (int32_t*)cr, n);
  local::gen_rct_backward(&ly, &lcb, &lcr, &lr, &lg, &lb, n);
}
void ojr_ict_forward(const float* r, const float* g, const float* b,
                     float* y, float* cb, float* cr, uint32_t n)
{ local::gen_ict_forward(r, g, b, y, cb, cr, n); }
void ojr_ict_backward(const float* y, const float* cb, const float* cr,
                      float* r, float* g, float* b, uint32_t n)
{ local::gen_ict_backward(y, cb, cr, r, g, b, n); }
void ojr_irv_convert_to_float(const int32_t* src, float* dst, uint32_t bit_depth,
                              int is_signed, uint32_t n)
{
  line_buf s, d; wrap_i32(s, (int32_t*)src, n); wrap_f32(d, dst, n);
  local::gen_irv_convert_to_float(&s, 0, &d, bit_depth, is_signed != 0, n);
}
void ojr_irv_convert_to_integer(const float* src, int32_t* dst, uint32_t bit_depth,
                                int is_signed, uint32_t n)
{
  line_buf s, d; wrap_f32(s, (float*)src, n); wrap_i32(d, dst, n);
  local::gen_irv_convert_to_integer(&s, &d, 0, bit_depth, is_signed != 0, n);
}
// NLT type 3 variants (pin the oracle's restatement)
void ojr_rev_convert_nlt_type3(const int32_t* src, int32_t* dst, int64_t shift, uint32_t n)
{
  line_buf s, d; wrap_i32(s, (int32_t*)src, n); wrap_i32(d, dst, n);
  local::gen_rev_convert_nlt_type3(&s, 0, &d, 0, shift, n);
}
void ojr_irv_convert_to_float_nlt_type3(const int32_t* src, float* dst, uint32_t bit_depth, int is_signed, uint32_t n)
{
  line_buf s, d; wrap_i32(s, (int32_t*)src, n); wrap_f32(d, dst, n);
  local::gen_irv_convert_to_float_nlt_type3(&s, 0, &d, bit_depth, is_signed != 0, n);
}
void ojr_irv_convert_to_integer_nlt_type3(const float* src, int32_t* dst, uint32_t bit_depth, int is_signed, uint32_t n)
{
  line_buf s, d; wrap_f32(s, (float*)src, n); wrap_i32(d, dst, n);
  local::gen_irv_convert_to_integer_nlt_type3(&s, &d, 0, bit_depth, is_signed != 0, n);
}
void ojr_rev_tx_to_cb32(const int32_t* sp, uint32_t* dp, uint32_t K_max, uint32_t n,
                        uint32_t* max_val)
{ local::gen_rev_tx_to_cb32(sp, dp, K_max, 0.0f, n, max_val); }
void ojr_irv_tx_to_cb32(const float* sp, uint32_t* dp, float delta_inv, uint32_t n,
                        uint32_t* max_val)
{ local::gen_irv_tx_to_cb32(sp, dp, 0, delta_inv, n, max_val); }
void ojr_rev_tx_from_cb32(const uint32_t* sp, int32_t* dp, uint32_t K_max, uint32_t n)
{ local::gen_rev_tx_from_cb32(sp, dp, K_max, 0.0f, n); }
void ojr_irv_tx_from_cb32(const uint32_t* sp, float* dp, float delta, uint32_t n)
{ local::gen_irv_tx_from_cb32(sp, dp, 0, delta, n); }

} // extern "C"


// ---- the apps' image readers / writers (src/apps/others/ojph_img_io.cpp) ---------------------------
// kind 0: .pgm/.ppm through ppm_in / ppm_out (rows, components interleaved, as ojph_compress / ojph_expand
// call them); kind 1: .yuv through yuv_in / yuv_out (one component after the other).
extern "C" int ojr_read_image(const char* path, int kind, uint32_t w, uint32_t h, uint32_t nc, uint32_t bit_depth,
                              const uint32_t* dx, const uint32_t* dy, int32_t* const* planes)
{
  try {
    std::vector<si32> row(w + 64);
    line_buf line; line.wrap(row.data(), w, 0);
    if (kind == 0) {
      ppm_in in; in.open(path);
      if (in.get_width() != w || in.get_height() != h || in.get_num_components() != nc) throw std::runtime_error("pnm header mismatch");
      for (uint32_t y = 0; y < h; ++y)
        for (uint32_t c = 0; c < nc; ++c) { in.read(&line, c); memcpy(planes[c] + (size_t)y * w, line.i32, (size_t)w * 4); }
      in.close();
    } else if (kind == 2) {
      dpx_in in; in.open(path);
      if (in.get_size().w != w || in.get_size().h != h || in.get_num_components() != nc || in.get_bit_depth(0) != bit_depth)
        throw std::runtime_error("dpx header mismatch");
      for (uint32_t y = 0; y < h; ++y)
        for (uint32_t c = 0; c < nc; ++c) { in.read(&line, c); memcpy(planes[c] + (size_t)y * w, line.i32, (size_t)w * 4); }
      in.close();
    } else {
      yuv_in in;
      ui32 bd = bit_depth; in.set_bit_depth(1, &bd);
      std::vector<point> sub(nc);
      for (uint32_t c = 0; c < nc; ++c) sub[c] = point(dx[c], dy[c]);
      in.set_img_props(size(w, h), nc, nc, sub.data());
      in.open(path);
      for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t cw = (w + dx[c] - 1) / dx[c], ch = (h + dy[c] - 1) / dy[c];
        for (uint32_t y = 0; y < ch; ++y) { in.read(&line, c); memcpy(planes[c] + (size_t)y * cw, line.i32, (size_t)cw * 4); }
      }
      in.close();
    }
    return 0;
  } catch (const std::exception& e) { snprintf(g_err, sizeof(g_err), "%s", e.what()); return 1; }
}

extern "C" int ojr_write_image(const char* path, int kind, uint32_t nc, uint32_t bit_depth, const uint32_t* comp_w,
                               const uint32_t* comp_h, const int32_t* const* planes)
{
  try {
    uint32_t wmax = 0;
    for (uint32_t c = 0; c < nc; ++c) wmax = comp_w[c] > wmax ? comp_w[c] : wmax;
    std::vector<std::vector<si32>> rows(nc, std::vector<si32>(wmax + 64));
    std::vector<line_buf> lines(nc);
    for (uint32_t c = 0; c < nc; ++c) lines[c].wrap(rows[c].data(), comp_w[c], 0);
    std::string name(path);
    if (kind == 0) {
      ppm_out out; out.configure(comp_w[0], comp_h[0], nc, bit_depth); out.open(&name[0]);
      for (uint32_t y = 0; y < comp_h[0]; ++y)
        for (uint32_t c = 0; c < nc; ++c) {
          memcpy(rows[c].data(), planes[c] + (size_t)y * comp_w[c], (size_t)comp_w[c] * 4);
          out.write(&lines[c], c);
        }
      out.close();
    } else {
      yuv_out out; std::vector<ui32> cw(comp_w, comp_w + nc);
      out.configure(bit_depth, nc, cw.data()); out.open(&name[0]);
      for (uint32_t c = 0; c < nc; ++c)
        for (uint32_t y = 0; y < comp_h[c]; ++y) {
          memcpy(rows[c].data(), planes[c] + (size_t)y * comp_w[c], (size_t)comp_w[c] * 4);
          out.write(&lines[c], c);
        }
      out.close();
    }
    return 0;
  } catch (const std::exception& e) { snprintf(g_err, sizeof(g_err), "%s", e.what()); return 1; }
}
