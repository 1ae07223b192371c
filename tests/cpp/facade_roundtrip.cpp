// Test-only: the same application code (the call sequence of ojph_compress.cpp:1165-1203 and
// ojph_expand.cpp:224-421) instantiated once with OpenJPH's own ojph::codestream and once with the
// facade ojph::b200::codestream; the codestreams must be byte-identical and the pulled lines equal.
// Built by tests/test_cpp_facade.py against the reference's public headers (CPU tier, here only).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ojph_arch.h"
#include "ojph_base.h"
#include "ojph_mem.h"
#include "ojph_file.h"
#include "ojph_params.h"
#include "ojph_codestream.h"
#include "ojph_b200_codestream.hpp"

using namespace ojph;

// extras: signed samples with NLT type 3 for all components (ojph_compress.cpp:859-866) and two COM segments
template <typename CS, typename COM, typename NLT>
static std::vector<ui8> compress(const std::vector<std::vector<si32>>& planes, ui32 w, ui32 h, ui32 depth, bool planar,
                                 bool extras = false)
{
  CS cs;
  auto siz = cs.access_siz();
  siz.set_image_extent(point(w, h));
  siz.set_num_components((ui32)planes.size());
  for (ui32 c = 0; c < planes.size(); ++c) siz.set_component(c, point(1, 1), depth, extras);
  siz.set_image_offset(point(0, 0));
  siz.set_tile_size(size(0, 0));
  siz.set_tile_offset(point(0, 0));
  auto cod = cs.access_cod();
  cod.set_num_decomposition(4);
  cod.set_block_dims(64, 64);
  cod.set_progression_order("RPCL");
  cod.set_color_transform(planes.size() == 3 && !planar);
  cod.set_reversible(true);
  cs.set_planar(planar);
  cs.request_tlm_marker(true);
  mem_outfile f;
  f.open();
  COM com[2];
  if (extras) {
    cs.access_nlt().set_nonlinear_transform(NLT::ALL_COMPS, NLT::OJPH_NLT_BINARY_COMPLEMENT_NLT);
    com[0].set_string("facade round trip");
    static const char raw[5] = { 1, 0, 2, (char)0xFF, 3 };
    com[1].set_data(raw, 5);
    cs.write_headers(&f, com, 2);
  } else
    cs.write_headers(&f);
  ui32 next = 0;
  std::vector<ui32> row(planes.size(), 0);
  line_buf* line = cs.exchange(NULL, next);
  while (line) {
    const si32* src = planes[next].data() + (size_t)row[next] * w;
    for (ui32 x = 0; x < w; ++x) line->i32[x] = src[x];
    row[next]++;
    line = cs.exchange(line, next);
  }
  cs.flush();
  std::vector<ui8> out(f.get_data(), f.get_data() + f.tell());
  cs.close();
  return out;
}

// what the last expand() saw without ever calling set_planar: is_planar(), is_using_color_transform(), then the
// component index of every pulled line -- the reference's defaults (read_headers: planar = !colour transform,
// ojph_codestream_local.cpp:879) decide the order a caller's `if (is_planar())` loop relies on
static std::vector<si32> g_trace;

template <typename CS>
static std::vector<std::vector<si32>> expand(const std::vector<ui8>& j2c, ui32 skip = 0)
{
  CS cs;
  mem_infile f;
  f.open(j2c.data(), j2c.size());
  cs.read_headers(&f);
  if (skip) cs.restrict_input_resolution(skip, skip);          // ojph_expand -skip_res
  {   // the read-side getters an application may print (ojph_expand.cpp); both classes must agree
    auto cod = cs.access_cod();
    auto sz = cs.access_siz();
    printf("  cod: %u levels, blocks %ux%u, precinct(0) %ux%u, order %s, layers %d, sop %d eph %d causal %d, tile %ux%u\n",
           cod.get_num_decompositions(), cod.get_block_dims().w, cod.get_block_dims().h,
           cod.get_precinct_size(0).w, cod.get_precinct_size(0).h, cod.get_progression_order_as_string(),
           cod.get_num_layers(), (int)cod.packets_may_use_sop(), (int)cod.packets_use_eph(),
           (int)cod.get_block_vertical_causality(), sz.get_tile_size().w, sz.get_tile_size().h);
  }
  auto siz = cs.access_siz();
  const ui32 nc = siz.get_num_components();
  std::vector<std::vector<si32>> planes(nc);
  std::vector<ui32> row(nc, 0);
  cs.create();
  g_trace.clear();
  g_trace.push_back(cs.is_planar() ? 1 : 0);
  g_trace.push_back(cs.access_cod().is_using_color_transform() ? 1 : 0);
  ui32 total = 0;
  for (ui32 c = 0; c < nc; ++c) { planes[c].resize((size_t)siz.get_recon_width(c) * siz.get_recon_height(c)); total += siz.get_recon_height(c); }
  for (ui32 i = 0; i < total; ++i) {
    ui32 c;
    line_buf* line = cs.pull(c);
    if (!line) { fprintf(stderr, "pull ended early\n"); exit(2); }
    g_trace.push_back((si32)c);
    const ui32 w = siz.get_recon_width(c);
    for (ui32 x = 0; x < w; ++x) planes[c][(size_t)row[c] * w + x] = line->i32[x];
    row[c]++;
  }
  cs.close();
  return planes;
}

int main()
{
  const ui32 w = 203, h = 131, depth = 10;
  int fails = 0;
  for (int planar = 0; planar < 2; ++planar) {
    std::vector<std::vector<si32>> planes(3, std::vector<si32>((size_t)w * h));
    unsigned s = 12345u + (unsigned)planar;
    for (auto& p : planes)
      for (size_t i = 0; i < p.size(); ++i) { s = s * 1664525u + 1013904223u; p[i] = (si32)(((i % w) * 3 + (i / w) * 5 + (s >> 27)) & ((1u << depth) - 1)); }
    std::vector<ui8> a = compress<ojph::codestream, ojph::comment_exchange, ojph::param_nlt>(planes, w, h, depth, planar != 0);
    std::vector<ui8> b = compress<ojph::b200::codestream, ojph::b200::comment_exchange, ojph::b200::param_nlt>(planes, w, h, depth, planar != 0);
    const bool same = a == b;
    std::vector<std::vector<si32>> ra = expand<ojph::codestream>(a);
    const std::vector<si32> ta = g_trace;
    std::vector<std::vector<si32>> rb = expand<ojph::b200::codestream>(a);
    const bool defaults = ta == g_trace && ta[0] == planar && ta[1] == 1 - planar;
    if (!defaults) { printf("read-side defaults differ: planar %d/%d colour transform %d/%d\n", ta[0], g_trace[0], ta[1], g_trace[1]); ++fails; }
    const bool lossless = ra == planes && rb == planes;
    printf("planar=%d codestream %zu bytes identical=%d lossless=%d\n", planar, a.size(), (int)same, (int)lossless);
    if (!same || !lossless) ++fails;
  }
  {   // signed samples + NLT type 3 + comments; then a reduced-resolution expand
    std::vector<std::vector<si32>> planes(3, std::vector<si32>((size_t)w * h));
    unsigned s = 777u;
    for (auto& p : planes)
      for (size_t i = 0; i < p.size(); ++i) {
        s = s * 1664525u + 1013904223u;
        p[i] = (si32)(((i % w) * 3 + (i / w) * 5 + (s >> 27)) & ((1u << depth) - 1)) - (si32)(1u << (depth - 1));
      }
    std::vector<ui8> a = compress<ojph::codestream, ojph::comment_exchange, ojph::param_nlt>(planes, w, h, depth, false, true);
    std::vector<ui8> b = compress<ojph::b200::codestream, ojph::b200::comment_exchange, ojph::b200::param_nlt>(planes, w, h, depth, false, true);
    const bool same = a == b;
    const bool lossless = expand<ojph::codestream>(a) == planes && expand<ojph::b200::codestream>(a) == planes;
    const bool reduced = expand<ojph::codestream>(a, 2) == expand<ojph::b200::codestream>(a, 2);
    printf("nlt+comments codestream %zu bytes identical=%d lossless=%d reduced=%d\n", a.size(), (int)same, (int)lossless, (int)reduced);
    if (!same || !lossless || !reduced) ++fails;
  }
  return fails ? 1 : 0;
}
