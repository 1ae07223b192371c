// ojb_layout.cpp -- geometry, packet sequencing, packet headers (see ojb_layout.h).
#include "ojb_layout.h"
#include <cstdlib>
#include <algorithm>
#include <climits>
#include <cstring>

namespace ojb {

//------------------------------------------------------------------------------------------
// geometry
//------------------------------------------------------------------------------------------
Rect BandGeom::block_rect(uint32_t bx, uint32_t by) const {
  // code-block grid is anchored at multiples of the nominal size in band coordinates
  // (ojph_subband.cpp:186-203, :313-333)
  uint32_t nw = 1u << xcb, nh = 1u << ycb;
  uint32_t gx = (rect.x0 >> xcb) << xcb, gy = (rect.y0 >> ycb) << ycb;
  uint32_t x0 = std::max(rect.x0, gx + bx * nw), x1 = std::min(rect.x1(), gx + (bx + 1) * nw);
  uint32_t y0 = std::max(rect.y0, gy + by * nh), y1 = std::min(rect.y1(), gy + (by + 1) * nh);
  Rect r; r.x0 = x0; r.y0 = y0; r.w = x1 - x0; r.h = y1 - y0;
  return r;
}

static void build_band(const Params& p, uint32_t comp, ResGeom& rg, uint32_t b, const Rect& br,
                       uint32_t& block_counter, uint64_t& arena) {
  BandGeom& bg = rg.bands[b];
  bg.rect = br; bg.band_num = b;
  bg.K_max = p.band_kmax(comp, rg.res_num, b);
  if (!p.reversible(comp)) {
    float d = p.band_delta(comp, rg.res_num, b);
    d /= (float)(1u << (31 - std::min(bg.K_max, 31u)));
    bg.delta = d; bg.delta_inv = 1.0f / d;
  }
  bg.empty = br.empty();
  if (bg.empty) return;
  bg.xcb = std::min(p.log_cb_w(comp), rg.log_ppx - rg.hsplit);      // subband::finalize_alloc, ojph_subband.cpp:70-74
  bg.ycb = std::min(p.log_cb_h(comp), rg.log_ppy - rg.vsplit);
  bg.nbw = ((br.x1() + (1u << bg.xcb) - 1) >> bg.xcb) - (br.x0 >> bg.xcb);
  bg.nbh = ((br.y1() + (1u << bg.ycb) - 1) >> bg.ycb) - (br.y0 >> bg.ycb);
  bg.block_base = block_counter;
  block_counter += bg.nbw * bg.nbh;
  bg.plane_pad_x = br.x0 & 3u;
  bg.plane_stride = (bg.plane_pad_x + br.w + 15u) & ~15u;     // 64-byte rows
  bg.plane_off = arena;
  arena += (uint64_t)bg.plane_stride * br.h + 16;            // slack for vector tails
  arena = (arena + 31) & ~(uint64_t)31;
}

static void build_precincts(const Params& p, const TileGeom& tile, const TileCompGeom& tc,
                            ResGeom& rg) {
  const Rect& rr = rg.rect;
  rg.npw = rg.nph = 0;
  if (rr.empty()) return;
  rg.npw = ((rr.x1() + (1u << rg.log_ppx) - 1) >> rg.log_ppx) - (rr.x0 >> rg.log_ppx);
  rg.nph = ((rr.y1() + (1u << rg.log_ppy) - 1) >> rg.log_ppy) - (rr.y0 >> rg.log_ppy);
  rg.precincts.assign((size_t)rg.npw * rg.nph, PrecinctGeom());
  uint32_t xlb = (rr.x0 >> rg.log_ppx) << rg.log_ppx, ylb = (rr.y0 >> rg.log_ppy) << rg.log_ppy;
  // resolution down-sampling relative to the canvas, including component sub-sampling
  // (tile_comp::finalize_alloc passes comp_downsamp as the initial res_downsamp)
  uint64_t dsx = (uint64_t)p.comps[tc.comp].dx * rg.dsx;
  uint64_t dsy = (uint64_t)p.comps[tc.comp].dy * rg.dsy;
  for (uint32_t y = 0; y < rg.nph; ++y)
    for (uint32_t x = 0; x < rg.npw; ++x) {
      PrecinctGeom& pc = rg.precincts[(size_t)y * rg.npw + x];
      uint32_t tx = (uint32_t)(dsx * (xlb + (x << rg.log_ppx)));
      uint32_t ty = (uint32_t)(dsy * (ylb + (y << rg.log_ppy)));
      pc.img_x = std::max(tx, tile.rect.x0);
      pc.img_y = std::max(ty, tile.rect.y0);
    }
  // code-block index rectangles (subband::get_cb_indices)
  const uint32_t xshift = rg.hsplit, yshift = rg.vsplit;
  for (uint32_t b = (rg.res_num ? 1 : 0); b < (rg.res_num ? 4u : 1u); ++b) {
    const BandGeom& bg = rg.bands[b];
    if (bg.empty) continue;
    uint32_t coly = 0;
    for (uint32_t y = 0; y < rg.nph; ++y) {
      uint32_t pcy0 = std::max(rr.y0, ylb + (y << rg.log_ppy));
      uint32_t pcy1 = std::min(rr.y1(), ylb + ((y + 1) << rg.log_ppy));
      pcy0 = (pcy0 - (b >> 1) + (1u << yshift) - 1) >> yshift;
      pcy1 = (pcy1 - (b >> 1) + (1u << yshift) - 1) >> yshift;
      uint32_t yb = ((pcy1 + (1u << bg.ycb) - 1) >> bg.ycb) - (pcy0 >> bg.ycb);
      uint32_t colx = 0;
      for (uint32_t x = 0; x < rg.npw; ++x) {
        uint32_t pcx0 = std::max(rr.x0, xlb + (x << rg.log_ppx));
        uint32_t pcx1 = std::min(rr.x1(), xlb + ((x + 1) << rg.log_ppx));
        pcx0 = (pcx0 - (b & 1) + (1u << xshift) - 1) >> xshift;
        pcx1 = (pcx1 - (b & 1) + (1u << xshift) - 1) >> xshift;
        uint32_t xb = ((pcx1 + (1u << bg.xcb) - 1) >> bg.xcb) - (pcx0 >> bg.xcb);
        Rect& r = rg.precincts[(size_t)y * rg.npw + x].cb_idx[b];
        r.x0 = colx; r.y0 = coly; r.w = xb; r.h = yb;
        colx += xb;
      }
      coly += yb;
    }
  }
}

void Layout::build(const Params& params) {
  p = &params;
  const Params& P = params;
  uint32_t nc = P.num_comps();
  ntw = div_ceil(P.Xsiz - P.XTOsiz, P.XTsiz);
  nth = div_ceil(P.Ysiz - P.YTOsiz, P.YTsiz);
  if ((uint64_t)ntw * nth > 65535) fail(0x00030011, "the number of tiles cannot exceed 65535");
  if ((uint64_t)ntw * nth == 0) fail(0x00030012, "the number of tiles cannot be 0");
  tiles.assign((size_t)ntw * nth, TileGeom());
  num_blocks = 0;
  uint64_t arena = 0;
  for (uint32_t ty = 0; ty < nth; ++ty)
    for (uint32_t tx = 0; tx < ntw; ++tx) {
      TileGeom& t = tiles[(size_t)ty * ntw + tx];
      t.idx = ty * ntw + tx;
      uint32_t x0 = P.XTOsiz + tx * P.XTsiz, x1 = x0 + P.XTsiz;
      uint32_t y0 = P.YTOsiz + ty * P.YTsiz, y1 = y0 + P.YTsiz;
      t.rect.x0 = std::max(x0, P.XOsiz); t.rect.w = std::min(x1, P.Xsiz) - t.rect.x0;
      t.rect.y0 = std::max(y0, P.YOsiz); t.rect.h = std::min(y1, P.Ysiz) - t.rect.y0;
      t.comps.assign(nc, TileCompGeom());
      for (uint32_t c = 0; c < nc; ++c) {
        TileCompGeom& tc = t.comps[c];
        tc.comp = c;
        uint32_t dx = P.comps[c].dx, dy = P.comps[c].dy;
        tc.rect.x0 = div_ceil(t.rect.x0, dx); tc.rect.w = div_ceil(t.rect.x1(), dx) - tc.rect.x0;
        tc.rect.y0 = div_ceil(t.rect.y0, dy); tc.rect.h = div_ceil(t.rect.y1(), dy) - tc.rect.y0;
        const uint32_t D = P.decomps(c);           // the component's own coding style (COC) if it has one
        tc.res.assign(D + 1, ResGeom());
        Rect rr = tc.rect;
        uint32_t dsx = 1, dsy = 1;
        for (int r = (int)D; r >= 0; --r) {
          ResGeom& rg = tc.res[r];
          rg.rect = rr; rg.res_num = (uint32_t)r;
          rg.log_ppx = P.log_pp_w(c, (uint32_t)r); rg.log_ppy = P.log_pp_h(c, (uint32_t)r);
          // sample plane of this resolution (input of the level-r analysis / output of synthesis)
          rg.plane_stride = (rr.w + 15u) & ~15u;
          rg.plane_off = arena;
          if (r != (int)D || true) {   // top level planes are used by the decoder as well
            arena += (uint64_t)rg.plane_stride * rr.h + 16;
            arena = (arena + 31) & ~(uint64_t)31;
          }
          uint32_t trx0 = rr.x0, trx1 = rr.x1(), try0 = rr.y0, try1 = rr.y1();
          rg.dsx = dsx; rg.dsy = dsy;
          if (r > 0) {
            // how level D - r + 1 splits this resolution: both ways, or -- with a DFS marker segment -- one way or not
            // at all (resolution::finalize_alloc, ojph_resolution.cpp:264-396)
            const DfsType ds = P.dwt_type(c, D - (uint32_t)r + 1);
            rg.hsplit = (ds == DFS_BIDIR || ds == DFS_HORZ) ? 1u : 0u;
            rg.vsplit = (ds == DFS_BIDIR || ds == DFS_VERT) ? 1u : 0u;
            for (uint32_t i = 1; i < 4; ++i) {
              if (((i & 1) && !rg.hsplit) || ((i >> 1) && !rg.vsplit)) continue;
              Rect br = rr;
              if (rg.hsplit) { br.x0 = (trx0 - (i & 1) + 1) >> 1; br.w = ((trx1 - (i & 1) + 1) >> 1) - br.x0; }
              if (rg.vsplit) { br.y0 = (try0 - (i >> 1) + 1) >> 1; br.h = ((try1 - (i >> 1) + 1) >> 1) - br.y0; }
              build_band(P, c, rg, i, br, num_blocks, arena);
            }
            Rect ll = rr;
            if (rg.hsplit) { ll.x0 = (trx0 + 1) >> 1; ll.w = ((trx1 + 1) >> 1) - ll.x0; dsx *= 2; }
            if (rg.vsplit) { ll.y0 = (try0 + 1) >> 1; ll.h = ((try1 + 1) >> 1) - ll.y0; dsy *= 2; }
            rr = ll;
          } else
            build_band(P, c, rg, 0, rr, num_blocks, arena);
        }
        for (uint32_t r = 0; r <= D; ++r) build_precincts(P, t, tc, tc.res[r]);
      }
    }
  coef_words = arena;
}

//------------------------------------------------------------------------------------------
// packet sequencing
//------------------------------------------------------------------------------------------
void Layout::packet_sequence(uint32_t tile, std::vector<PacketRef>& seq,
                             std::vector<uint32_t>& tp_first, std::vector<uint32_t>* tp_index,
                             uint32_t* tp_total) const {
  const TileGeom& t = tiles[tile];
  const Params& P = *p;
  uint32_t nc = P.num_comps(), D = P.max_decomps();
  seq.clear(); tp_first.clear();
  // per (comp,res) cursor over precincts in raster order; a component with fewer decompositions
  // simply has no precincts at the higher resolution numbers (tile_comp::get_top_left_precinct,
  // ojph_tile_comp.cpp:132-145)
  std::vector<uint32_t> cur((size_t)nc * (D + 1), 0);
  auto npre = [&](uint32_t c, uint32_t r) { return r < t.comps[c].res.size() ? (uint32_t)t.comps[c].res[r].precincts.size() : 0u; };
  auto emit_all = [&](uint32_t c, uint32_t r) {
    for (uint32_t i = 0; i < npre(c, r); ++i) seq.push_back(PacketRef{ tile, c, r, i });
  };
  auto emit_one = [&](uint32_t c, uint32_t r) {
    seq.push_back(PacketRef{ tile, c, r, cur[(size_t)c * (D + 1) + r]++ });
  };
  auto top_left = [&](uint32_t c, uint32_t r, uint32_t& x, uint32_t& y) {
    uint32_t i = cur[(size_t)c * (D + 1) + r];
    if (i >= npre(c, r)) return false;
    x = t.comps[c].res[r].precincts[i].img_x; y = t.comps[c].res[r].precincts[i].img_y;
    return true;
  };
  uint32_t div = P.tilepart_div;
  if (div == TP_NONE) tp_first.push_back(0);
  std::vector<uint32_t> idx;                 // TPsot of every tile-part when it is not simply its rank
  uint32_t total = 0;
  if (P.prog_order == PO_LRCP || P.prog_order == PO_RLCP) {
    for (uint32_t r = 0; r <= D; ++r) {
      if (div == TP_RES) tp_first.push_back((uint32_t)seq.size());
      for (uint32_t c = 0; c < nc; ++c) {
        if (div & TP_COMP) {
          // one tile-part per EXISTING (resolution, component), numbered c + r * nc out of nc * (D + 1)
          // -- the numbering has gaps when components have different decomposition counts (tile::flush,
          // ojph_tile.cpp:634-652)
          if (r >= t.comps[c].res.size()) continue;
          tp_first.push_back((uint32_t)seq.size());
          idx.push_back(c + r * nc); total = nc * (D + 1);
        }
        emit_all(c, r);
      }
    }
  } else if (P.prog_order == PO_RPCL) {
    for (uint32_t r = 0; r <= D; ++r) {
      if (div == TP_RES) tp_first.push_back((uint32_t)seq.size());
      for (;;) {
        bool found = false; uint32_t best_c = 0, bx = INT_MAX, by = INT_MAX, x, y;
        for (uint32_t c = 0; c < nc; ++c) {
          if (!top_left(c, r, x, y)) continue;
          found = true;
          if (y < by || (y == by && x < bx)) { bx = x; by = y; best_c = c; }
        }
        if (!found) break;
        emit_one(best_c, r);
      }
    }
  } else if (P.prog_order == PO_PCRL) {
    for (;;) {
      bool found = false; uint32_t bc = 0, br = 0, bx = INT_MAX, by = INT_MAX, x, y;
      for (uint32_t c = 0; c < nc; ++c)
        for (uint32_t r = 0; r <= D; ++r) {
          if (!top_left(c, r, x, y)) continue;
          found = true;
          if (y < by || (y == by && x < bx) || (y == by && x == bx && c < bc) ||
              (y == by && x == bx && c == bc && r < br)) { bx = x; by = y; bc = c; br = r; }
        }
      if (!found) break;
      emit_one(bc, br);
    }
  } else {  // CPRL
    for (uint32_t c = 0; c < nc; ++c) {
      if (div == TP_COMP) tp_first.push_back((uint32_t)seq.size());
      for (;;) {
        bool found = false; uint32_t br = 0, bx = INT_MAX, by = INT_MAX, x, y;
        for (uint32_t r = 0; r <= D; ++r) {
          if (!top_left(c, r, x, y)) continue;
          found = true;
          if (y < by || (y == by && x < bx)) { bx = x; by = y; br = r; }
        }
        if (!found) break;
        emit_one(c, br);
      }
    }
  }
  if (tp_first.empty()) tp_first.push_back(0);
  if (tp_index) {
    tp_index->clear();
    for (uint32_t i = 0; i < tp_first.size(); ++i) tp_index->push_back(idx.size() == tp_first.size() ? idx[i] : i);
  }
  if (tp_total) *tp_total = (idx.size() == tp_first.size() && total) ? total : (uint32_t)tp_first.size();
}

//------------------------------------------------------------------------------------------
// packet headers
//------------------------------------------------------------------------------------------
namespace {

inline uint32_t log2ceil(uint32_t x) {
  uint32_t t = 31u - (uint32_t)__builtin_clz(x);
  return t + ((x & (x - 1)) ? 1u : 0u);
}

// Tag tree with the reference's storage behaviour: level i is a linear array of
// 4^(num_levels-1-i) entries pre-filled with init; an entry is addressed as
// x + y * ceil(w / 2^i).  When a level's width is odd, the "right child" of the last
// column aliases the first entry of the next row -- the reference computes its minima
// that way (ojph_precinct.cpp:57-87,142-165), so byte-identical headers need the same.
// Flat storage, reused across bands (one instance per thread).
struct TagTree {
  uint32_t w = 0, h = 0, nl = 0;
  uint32_t off[18], wl[18];
  std::vector<uint8_t> buf;
  void init(uint32_t nlev, uint32_t ww, uint32_t hh, uint8_t v) {
    w = ww; h = hh; nl = nlev;
    uint32_t o = 0;
    for (uint32_t i = 0; i < nl; ++i) { off[i] = o; wl[i] = (w + (1u << i) - 1) >> i; o += 1u << ((nl - 1 - i) << 1); }
    off[nl] = o; wl[nl] = 1; o += 1;
    if (buf.size() < o) buf.resize(o);
    memset(buf.data(), v, o - 1);
    buf[o - 1] = 0;
  }
  uint8_t& at(uint32_t x, uint32_t y, uint32_t l) { return buf[off[l] + x + (size_t)y * wl[l]]; }
};

struct BitWriter {   // MSB first; a byte after 0xFF carries 7 bits (ojph_bitbuffer_write.h:85-143)
  std::vector<uint8_t>& o; int avail = 8; uint32_t tmp = 0;
  explicit BitWriter(std::vector<uint8_t>& out) : o(out) {}
  inline void bits(uint32_t v, int n) {      // the n low bits of v, MSB first
    while (n > 0) {
      int k = n < avail ? n : avail;
      uint32_t chunk = (n >= 32 && k == 32) ? v : ((v >> (n - k)) & ((1u << k) - 1u));
      avail -= k; n -= k;
      tmp |= chunk << avail;
      if (avail == 0) { o.push_back((uint8_t)tmp); avail = (tmp != 0xFF) ? 8 : 7; tmp = 0; }
    }
  }
  inline void bit(uint32_t b) { bits(b & 1u, 1); }
  inline void zeros(int n) { while (n > 0) { int k = n < 24 ? n : 24; bits(0, k); n -= k; } }
  void finish() { if (avail < 8) o.push_back((uint8_t)tmp); }
};

struct BitReader {   // ojph_bitbuffer_read.h:73-130
  // `left` counts the bytes the tile-part SAYS it still has (Psot); `end` is where the buffer really
  // ends: running out of buffer while `left` > 0 is the reference's failed file read (bb_read, :80-86)
  const uint8_t* d; size_t& pos; uint32_t& left; size_t end; HostMirror* hm; uint32_t tmp = 0; int avail = 0; bool unstuff = false;
  BitReader(const uint8_t* data, size_t& p, uint32_t& l, size_t e, HostMirror* m) : d(data), pos(p), left(l), end(e), hm(m) {}
  inline bool fill() {
    if (left > 0) {
      if (pos >= end) throw Error(0x00030092, "error reading from file");
      if (hm && !hm->present[pos >> HostMirror::PAGE_SHIFT]) hm->need(pos);
      uint8_t t = d[pos++]; tmp = t; avail = 8 - (unstuff ? 1 : 0); unstuff = (t == 0xFF); --left; return true;
    }
    tmp = 0; avail = 8 - (unstuff ? 1 : 0); unstuff = false; return false;
  }
  inline bool bit(uint32_t& b) { bool r = true; if (avail == 0) r = fill(); b = (tmp >> --avail) & 1u; return r; }
  inline bool bits(int n, uint32_t& v) {
    v = 0; bool r = true;
    while (n) {
      if (avail == 0) r = fill();
      int t = std::min(avail, n);
      v <<= t; avail -= t; n -= t;
      v |= (tmp >> avail) & ((1u << t) - 1);
    }
    return r;
  }
  bool finish() { bool r = true; if (unstuff) r = fill(); tmp = 0; avail = 0; return r; }
};

struct Trees { TagTree inc, incf, mm, mmf; };
static thread_local Trees t_trees;

} // namespace

uint32_t write_packet_header(const ResGeom& res, const PrecinctGeom& pc,
                             const CodedBlock* blocks, std::vector<uint8_t>& out) {
  size_t start = out.size();
  BitWriter bw(out);
  bool coded = false;
  int skipped_bands = 0;
  uint32_t body = 0;
  for (uint32_t s = 0; s < 4; ++s) {
    const BandGeom& bg = res.bands[s];
    if (bg.empty) continue;
    const Rect& ci = pc.cb_idx[s];
    if (ci.w == 0 || ci.h == 0) continue;
    uint32_t nl = 1 + std::max(log2ceil(ci.w), log2ceil(ci.h));
    TagTree &inc = t_trees.inc, &incf = t_trees.incf, &mm = t_trees.mm, &mmf = t_trees.mmf;
    inc.init(nl, ci.w, ci.h, 255); incf.init(nl, ci.w, ci.h, 0);
    mm.init(nl, ci.w, ci.h, 255); mmf.init(nl, ci.w, ci.h, 0);
    const CodedBlock* base = blocks + bg.block_base;
    for (uint32_t y = 0; y < ci.h; ++y)
      for (uint32_t x = 0; x < ci.w; ++x) {
        const CodedBlock& cb = base[(size_t)(ci.y0 + y) * bg.nbw + ci.x0 + x];
        inc.at(x, y, 0) = cb.num_passes == 0 ? 1 : 0;
        mm.at(x, y, 0) = cb.missing_msbs;
      }
    for (uint32_t l = 1; l < nl; ++l) {
      uint32_t hh = (ci.h + (1u << l) - 1) >> l, ww = (ci.w + (1u << l) - 1) >> l;
      for (uint32_t y = 0; y < hh; ++y)
        for (uint32_t x = 0; x < ww; ++x) {
          uint8_t a = std::min(inc.at(x << 1, y << 1, l - 1), inc.at((x << 1) + 1, y << 1, l - 1));
          uint8_t b = std::min(inc.at(x << 1, (y << 1) + 1, l - 1), inc.at((x << 1) + 1, (y << 1) + 1, l - 1));
          inc.at(x, y, l) = std::min(a, b); incf.at(x, y, l) = 0;
          a = std::min(mm.at(x << 1, y << 1, l - 1), mm.at((x << 1) + 1, y << 1, l - 1));
          b = std::min(mm.at(x << 1, (y << 1) + 1, l - 1), mm.at((x << 1) + 1, (y << 1) + 1, l - 1));
          mm.at(x, y, l) = std::min(a, b); mmf.at(x, y, l) = 0;
        }
    }
    if (inc.at(0, 0, nl - 1) != 0) {          // nothing included in this band
      if (coded) bw.bit(0); else ++skipped_bands;
      continue;
    }
    if (!coded) {
      coded = true;
      bw.bit(1);                               // non-empty packet
      for (int i = 0; i < skipped_bands; ++i) bw.bit(0);
      skipped_bands = 0;
    }
    for (uint32_t y = 0; y < ci.h; ++y)
      for (uint32_t x = 0; x < ci.w; ++x) {
        const CodedBlock& cb = base[(size_t)(ci.y0 + y) * bg.nbw + ci.x0 + x];
        // inclusion: levels along the leaf-to-root path are sent top-down, so the unsent ones are
        // a prefix [0, u); a sent ancestor with a positive value means "nothing included below"
        {
          uint32_t u = 0;
          while (u < nl && incf.at(x >> u, y >> u, u) == 0) ++u;
          if (!(u < nl && inc.at(x >> u, y >> u, u) > 0))
            for (uint32_t cl = u; cl > 0; --cl) {
              uint32_t lm = cl - 1;
              uint32_t v = inc.at(x >> lm, y >> lm, lm);
              bw.bit(1 - (v - inc.at(x >> cl, y >> cl, cl)));
              incf.at(x >> lm, y >> lm, lm) = 1;
              if (v > 0) break;
            }
        }
        if (cb.num_passes == 0) continue;
        {   // missing msbs
          uint32_t u = 0;
          while (u < nl && mmf.at(x >> u, y >> u, u) == 0) ++u;
          for (uint32_t cl = u; cl > 0; --cl) {
            uint32_t lm = cl - 1;
            int nz = (int)mm.at(x >> lm, y >> lm, lm) - (int)mm.at(x >> cl, y >> cl, cl);
            bw.zeros(nz);
            bw.bit(1);
            mmf.at(x >> lm, y >> lm, lm) = 1;
          }
        }
        if (cb.num_passes == 3) bw.bits(12, 4);
        else if (cb.num_passes == 2) bw.bits(2, 2);
        else bw.bits(0, 1);
        int bits1 = 32 - (cb.pass_len[0] ? __builtin_clz(cb.pass_len[0]) : 32);
        int extra = cb.num_passes > 2 ? 1 : 0;
        int bits2 = 0;
        if (cb.num_passes > 1) bits2 = 32 - (cb.pass_len[1] ? __builtin_clz(cb.pass_len[1]) : 32);
        int nb = std::max(std::max(bits1, bits2 - extra) - 3, 0);
        bw.bits(0xFFFFFFFEu, nb + 1);
        bw.bits(cb.pass_len[0], nb + 3);
        if (cb.num_passes > 1) bw.bits(cb.pass_len[1], nb + 3 + extra);
        body += cb.pass_len[0] + cb.pass_len[1];
      }
  }
  if (coded) bw.finish();
  else { out.resize(start); out.push_back(0); }   // empty packet: one zero byte
  return body;
}

namespace {
// The same header grammar as the loop in parse_packet, read through a 64-bit window of de-stuffed bits instead of
// one bit at a time (the host parse is the longest host phase of a decode: 49 152 blocks per headline frame).  It
// is an optimistic pass: anything out of the ordinary -- an empty packet, the byte budget or the buffer running
// out, a value the slow loop would refuse -- makes it return false WITHOUT having moved pos / data_left, and the
// byte-at-a-time loop then re-reads the packet and reports whatever there is to report, exactly as before.
struct BitWindow {
  const uint8_t* d; size_t pos, start, limit; HostMirror* hm;
  uint64_t acc = 0; int n = 0;             // n valid bits at the top of acc; the rest of acc is zero
  BitWindow(const uint8_t* data, size_t p, size_t lim, HostMirror* m) : d(data), pos(p), start(p), limit(lim), hm(m) {}
  inline int bits_of(size_t i) const { return (i > start && d[i - 1] == 0xFF) ? 7 : 8; }   // a byte after 0xFF carries 7
  inline void refill() {
    // the usual case in one step: the next eight bytes hold no 0xFF and neither does the byte before them, so each
    // carries eight bits -- as many whole bytes as fit are appended at once (with a device-resident codestream only
    // when their page has already been fetched: a page is never fetched for bytes that may not be read)
    if (n <= 56 && pos + 8 <= limit && (!hm || (hm->present[pos >> HostMirror::PAGE_SHIFT] && hm->present[(pos + 7) >> HostMirror::PAGE_SHIFT]))) {
      uint64_t v; memcpy(&v, d + pos, 8);
      const bool prev_ff = pos > start && d[pos - 1] == 0xFF;
      if (!prev_ff && ((((v & 0x7F7F7F7F7F7F7F7Full) + 0x0101010101010101ull) & v & 0x8080808080808080ull) == 0)) {
        v = __builtin_bswap64(v);
        const int take = (64 - n) >> 3;              // 1..8 whole bytes
        const uint64_t top = take == 8 ? v : (v & (~0ull << (64 - 8 * take)));
        acc |= n ? (top >> n) : top;
        n += 8 * take; pos += (size_t)take;
        return;
      }
    }
    while (n <= 56 && pos < limit) {
      if (hm && !hm->present[pos >> HostMirror::PAGE_SHIFT]) hm->need(pos);
      const int k = bits_of(pos);
      acc |= (uint64_t)(d[pos] & (k == 7 ? 0x7F : 0xFF)) << (64 - n - k);
      n += k; ++pos;
    }
  }
  inline bool need(int k) { if (n < k) refill(); return n >= k; }
  inline uint32_t take(int k) { uint32_t v = (uint32_t)(acc >> (64 - k)); acc <<= k; n -= k; return v; }   // 1 <= k <= 32
  // length of the run of `bit` ahead, consumed together with the bit that ends it; -1 when the data runs out
  inline int run(uint32_t bit) {
    int total = 0;
    for (;;) {
      if (n == 0 && !need(1)) return -1;
      const uint64_t a = bit ? ~acc : acc;
      int z = a ? __builtin_clzll(a) : 64;
      if (z >= n) { total += n; acc = 0; n = 0; continue; }      // the whole window is the run: go on
      total += z; acc <<= (z + 1); n -= z + 1;
      return total;
    }
  }
  // bytes whose bits were (at least partly) consumed: give back the look-ahead
  inline size_t consumed_end() { size_t p = pos; int m = n; while (p > start && m >= bits_of(p - 1)) { m -= bits_of(p - 1); --p; } return p; }
};

// a tag tree of the fast parser: one byte per node, 0xFF = nothing received for it yet (the values it can hold are 0 / 1
// for inclusion and at most K_max for missing MSBs).  Same node addressing as TagTree.
struct FastTree {
  uint32_t nl = 0, off[18], wl[18];
  std::vector<uint8_t> buf;
  void init(uint32_t nlev, uint32_t w) {
    nl = nlev;
    uint32_t o = 0;
    for (uint32_t i = 0; i < nl; ++i) { off[i] = o; wl[i] = (w + (1u << i) - 1) >> i; o += 1u << ((nl - 1 - i) << 1); }
    if (buf.size() < o) buf.resize(o);
    memset(buf.data(), 0xFF, o);
  }
};
static thread_local FastTree t_finc, t_fmm;

bool parse_header_fast(const ResGeom& res, const PrecinctGeom& pc, CodedBlock* blocks, const uint8_t* data,
                       size_t& pos, uint32_t& data_left, size_t data_end, HostMirror* mirror, bool& unstuff, uint64_t& body_total) {
  if (data_left == 0 || pos >= data_end) return false;
  uint64_t body = 0;                      // bytes of the bodies read so far: a block's data_off is set RELATIVE to the first body
  BitWindow w(data, pos, std::min(data_end, pos + (size_t)data_left), mirror);
  bool first_band = true;
  for (uint32_t s = 0; s < 4; ++s) {
    const BandGeom& bg = res.bands[s];
    if (bg.empty) continue;
    const Rect& ci = pc.cb_idx[s];
    if (ci.w == 0 || ci.h == 0) continue;
    if (first_band) {
      if (!w.need(1) || w.take(1) == 0) return false;             // empty packets go the slow way
      first_band = false;
    }
    const uint32_t nl = 1 + std::max(log2ceil(ci.w), log2ceil(ci.h));
    if (nl > 16) return false;
    FastTree &ti = t_finc, &tm = t_fmm;
    ti.init(nl, ci.w); tm.init(nl, ci.w);
    uint8_t* I = ti.buf.data(); uint8_t* M = tm.buf.data();
    CodedBlock* base = blocks + bg.block_base;
    const uint32_t kmax = bg.K_max;
    for (uint32_t y = 0; y < ci.h; ++y) {
      uint32_t rb[18];                                            // first node of this row's ancestors, per level
      for (uint32_t l = 0; l < nl; ++l) rb[l] = ti.off[l] + (y >> l) * ti.wl[l];
      // levels 0 .. u0-1 of a leaf's path cannot have been reached before it: the leaf is the raster-first one below
      // those nodes (x and y are multiples of 2^l); the search for the lowest level already received starts there
      const uint32_t uy = y ? (uint32_t)__builtin_ctz(y) : 31u;
      CodedBlock* row = base + (size_t)(ci.y0 + y) * bg.nbw + ci.x0;
      for (uint32_t x = 0; x < ci.w; ++x) {
        const uint32_t u0 = std::min(std::min(x ? (uint32_t)__builtin_ctz(x) : 31u, uy) + 1u, nl);
        if (w.n < 40) w.refill();
        // inclusion: the bits of the levels below the lowest one already received, top-down, up to the first 0
        {
          uint32_t u = u0;
          while (u < nl && I[rb[u] + (x >> u)] == 0xFF) ++u;
          if (u < nl && I[rb[u] + (x >> u)] == 1) continue;       // an ancestor said: nothing included below
          if (w.n < (int)u) return false;                         // (u <= 16 bits; only at the very end of the data)
          const uint32_t v = (uint32_t)(w.acc >> 48) | (0xFFFFu >> u);     // the u bits ahead, ones behind them
          const uint32_t ones = (uint32_t)__builtin_clz(~(v << 16) | 1u);   // leading ones among them, 0 .. u
          if (ones >= u) {
            for (uint32_t l = 0; l < u; ++l) I[rb[l] + (x >> l)] = 0;
            if (u) { w.acc <<= u; w.n -= (int)u; }
          } else {
            for (uint32_t l = u - ones; l < u; ++l) I[rb[l] + (x >> l)] = 0;
            I[rb[u - ones - 1] + (x >> (u - ones - 1))] = 1;
            w.acc <<= (ones + 1); w.n -= (int)(ones + 1);
            continue;
          }
        }
        uint32_t mmsbs;
        {
          uint32_t u = u0;
          while (u < nl && M[rb[u] + (x >> u)] == 0xFF) ++u;
          mmsbs = u < nl ? M[rb[u] + (x >> u)] : 0u;
          for (uint32_t lp = u; lp > 0; --lp) {
            const uint32_t l = lp - 1;
            const int z = w.run(0);
            if (z < 0 || z > 254) return false;
            mmsbs += (uint32_t)z;
            if (mmsbs > 254) return false;
            M[rb[l] + (x >> l)] = (uint8_t)mmsbs;
          }
        }
        if (mmsbs > kmax) return false;
        if (!w.need(1)) return false;
        if (w.take(1)) return false;                               // more than one coding pass: the general loop
        const int ones = w.run(1);
        if (ones < 0 || ones > 28) return false;
        const int nb = 3 + ones;                                   // Lblock; no placeholder passes here
        if (!w.need(nb)) return false;
        const uint32_t len = w.take(nb);
        if (len < 2 || len >= 65535) return false;
        CodedBlock& cb = row[x];
        cb.missing_msbs = (uint8_t)mmsbs;
        cb.num_passes = 1;
        cb.pass_len[0] = len; cb.pass_len[1] = 0;
        cb.data_off = body; body += len;
      }
    }
  }
  body_total = body;
  if (first_band) return false;                                    // no band with blocks: one bit to read, slow way
  const size_t end = w.consumed_end();
  if (end == pos) return false;
  data_left -= (uint32_t)(end - pos);
  pos = end;
  unstuff = data[end - 1] == 0xFF;
  return true;
}
} // namespace

void parse_packet(const Params& P, const ResGeom& res, const PrecinctGeom& pc,
                  CodedBlock* blocks, const uint8_t* data, size_t& pos, uint32_t& data_left, size_t data_end,
                  HostMirror* mirror) {
  BitReader br(data, pos, data_left, data_end, mirror);
  if (P.uses_sop() && data_left >= 2) {            // optional SOP marker segment
    if (pos + 2 > data_end) throw Error(0x00030092, "error reading from file");
    if (mirror) mirror->need(pos, 6);
    if (data[pos] == 0xFF && data[pos + 1] == 0x91) {
      pos += 2; data_left -= 2;
      if (data_left < 4) throw Error(0x00030092, "precinct truncated early");
      if (pos + 4 > data_end) throw Error(0x00030092, "error reading from file");
      uint32_t L = ((uint32_t)data[pos] << 8) | data[pos + 1];
      if (L != 4) throw Error(0x00030092, "something is wrong with SOP length");
      pos += 4; data_left -= 4;
    }
  }
  bool empty_packet = true;
  // A header that fails to parse delivers no body: the blocks of this packet whose lengths were already
  // read must not reach the block decoder (the reference decodes a block only once its body has been
  // attached, ojph_codeblock.cpp:192-193) -- resilient mode swallows the exception and carries on.
  struct DropOnError {
    const ResGeom& res; const PrecinctGeom& pc; CodedBlock* blocks; bool armed = true;
    ~DropOnError() {
      if (!armed) return;
      for (uint32_t s = 0; s < 4; ++s) {
        const BandGeom& bg = res.bands[s];
        if (bg.empty) continue;
        const Rect& ci = pc.cb_idx[s];
        for (uint32_t y = 0; y < ci.h; ++y)
          for (uint32_t x = 0; x < ci.w; ++x) {
            CodedBlock& cb = blocks[bg.block_base + (size_t)(ci.y0 + y) * bg.nbw + ci.x0 + x];
            if (cb.data_off == 0) { cb.pass_len[0] = cb.pass_len[1] = 0; cb.num_passes = 0; }
          }
      }
    }
  } drop{res, pc, blocks};
  bool fast_unstuff = false;
  // OJB_PARSE_SLOW=1: the byte-wise loop only (tests compare the two readers)
  const char* slow_only = getenv("OJB_PARSE_SLOW");
  uint64_t fast_body = 0;
  const bool fast_done = !(slow_only && *slow_only == '1') &&
                         parse_header_fast(res, pc, blocks, data, pos, data_left, data_end, mirror, fast_unstuff, fast_body);
  if (!fast_done && !(slow_only && *slow_only == '1'))
    for (uint32_t s = 0; s < 4; ++s) {            // the optimistic pass left relative offsets behind: not delivered
      const BandGeom& bg = res.bands[s];
      if (bg.empty) continue;
      const Rect& ci = pc.cb_idx[s];
      for (uint32_t y = 0; y < ci.h; ++y)
        for (uint32_t x = 0; x < ci.w; ++x) blocks[bg.block_base + (size_t)(ci.y0 + y) * bg.nbw + ci.x0 + x].data_off = 0;
    }
  if (fast_done) { br.unstuff = fast_unstuff; empty_packet = false; }
  for (uint32_t s = 0; s < 4 && !fast_done; ++s) {
    const BandGeom& bg = res.bands[s];
    if (bg.empty) continue;
    const Rect& ci = pc.cb_idx[s];
    if (ci.w == 0 || ci.h == 0) continue;
    uint32_t bit;
    if (empty_packet) {
      br.bit(bit);
      if (bit == 0) {                              // empty packet
        br.finish();
        if (P.uses_eph() && data_left >= 2) {
          if (pos + 2 > data_end) throw Error(0x00030092, "error reading from file");
          pos += 2; data_left -= 2;
        }
        drop.armed = false;
        return;
      }
      empty_packet = false;
    }
    uint32_t nl = 1 + std::max(log2ceil(ci.w), log2ceil(ci.h));
    TagTree &inc = t_trees.inc, &incf = t_trees.incf, &mm = t_trees.mm, &mmf = t_trees.mmf;
    inc.init(nl, ci.w, ci.h, 0); incf.init(nl, ci.w, ci.h, 0);
    mm.init(nl, ci.w, ci.h, 0); mmf.init(nl, ci.w, ci.h, 0);
    CodedBlock* base = blocks + bg.block_base;
    for (uint32_t y = 0; y < ci.h; ++y)
      for (uint32_t x = 0; x < ci.w; ++x) {
        CodedBlock& cb = base[(size_t)(ci.y0 + y) * bg.nbw + ci.x0 + x];
        bool empty_cb = false;
        {
          // received levels form a suffix of the leaf-to-root path; start at the lowest received one
          uint32_t u = 0;
          while (u < nl && incf.at(x >> u, y >> u, u) == 0) ++u;
          if (u < nl && inc.at(x >> u, y >> u, u) == 1) empty_cb = true;
          for (uint32_t cl = u; cl > 0 && !empty_cb; --cl) {
            uint32_t l = cl - 1;
            if (!br.bit(bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p1"); }
            empty_cb = (bit == 0);
            inc.at(x >> l, y >> l, l) = (uint8_t)(1 - bit);
            incf.at(x >> l, y >> l, l) = 1;
          }
        }
        if (empty_cb) continue;
        uint32_t mmsbs = 0;
        {
          uint32_t u = 0;
          while (u < nl && mmf.at(x >> u, y >> u, u) == 0) ++u;
          mmsbs = mm.at(x >> u, y >> u, u);       // lowest received level (or the root's 0)
          for (uint32_t lp = u; lp > 0; --lp) {
            uint32_t l = lp - 1;
            bit = 0;
            while (bit == 0) {
              if (!br.bit(bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p2"); }
              mmsbs += 1 - bit;
            }
            mm.at(x >> l, y >> l, l) = (uint8_t)mmsbs;
            mmf.at(x >> l, y >> l, l) = 1;
          }
        }
        if (mmsbs > bg.K_max)
          throw Error(0x00030092, "error in parsing a tile header; missing msbs are larger or "
                      "equal to Kmax. The most likely cause is a corruption in the bitstream.");
        cb.missing_msbs = (uint8_t)mmsbs;
        uint32_t np = 1;
        if (!br.bit(bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p3"); }
        if (bit) {
          np = 2;
          if (!br.bit(bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p4"); }
          if (bit) {
            if (!br.bits(2, bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p5"); }
            np = 3 + bit;
            if (bit == 3) {
              if (!br.bits(5, bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p6"); }
              np = 6 + bit;
              if (bit == 31) {
                if (!br.bits(7, bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p7"); }
                np = 37 + bit;
              }
            }
          }
        }
        // placeholder passes: every group of 3 adds one missing msb (ojph_precinct.cpp:466-480)
        uint32_t phld = (np - 1) / 3;
        cb.missing_msbs = (uint8_t)(cb.missing_msbs + phld);
        phld *= 3;
        cb.num_passes = (uint8_t)(np - phld);
        cb.pass_len[0] = cb.pass_len[1] = 0;
        int Lblock = 3;
        bit = 1;
        while (bit) {
          if (!br.bit(bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p8"); }
          Lblock += (int)bit;
        }
        int nb = Lblock + 31 - __builtin_clz(phld + 1);
        if (!br.bits(nb, bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p9"); }
        if (bit < 2)
          throw Error(0x00030092, "The cleanup segment of an HT codeblock cannot contain less than 2 bytes");
        if (bit >= 65535)
          throw Error(0x00030092, "The cleanup segment of an HT codeblock must contain less than 65535 bytes");
        cb.pass_len[0] = bit;
        if (cb.num_passes > 1) {
          nb = Lblock + (cb.num_passes > 2 ? 1 : 0);
          if (!br.bits(nb, bit)) { data_left = 0; throw Error(0x00030092, "error reading from file p10"); }
          if (bit >= 2047)
            throw Error(0x00030092, "The refinement segment (SigProp and MagRep passes) of an HT "
                        "codeblock must contain less than 2047 bytes");
          cb.pass_len[1] = bit;
        }
      }
  }
  if (empty_packet) { uint32_t bit = 0; br.bit(bit); }
  br.finish();
  if (P.uses_eph() && data_left >= 2) {
    if (pos + 2 > data_end) throw Error(0x00030092, "error reading from file");
    if (mirror) mirror->need(pos, 2);
    if (!(data[pos] == 0xFF && data[pos + 1] == 0x92))
      throw Error(0x00030092, "should find EPH, but found something else");
    pos += 2; data_left -= 2;
  }
  // code-block bodies follow in band / raster order.  A body the buffer cannot deliver is dropped and
  // the rest of this packet with it (bb_read_chunk, ojph_bitbuffer_read.h:134-150; ojph_precinct.cpp:
  // 536-573), but the tile-part's byte budget stays what Psot says: the next packet header then fails
  // to read, which is how a truncation is detected outside resilient mode.
  drop.armed = false;
  if (fast_done && fast_body <= data_left && pos < data_end && fast_body <= data_end - pos) {
    // every body is there: the offsets were laid out while the header was read, relative to the first body
    for (uint32_t s = 0; s < 4; ++s) {
      const BandGeom& bg = res.bands[s];
      if (bg.empty) continue;
      const Rect& ci = pc.cb_idx[s];
      for (uint32_t y = 0; y < ci.h; ++y) {
        CodedBlock* row = blocks + bg.block_base + (size_t)(ci.y0 + y) * bg.nbw + ci.x0;
        for (uint32_t x = 0; x < ci.w; ++x) if (row[x].pass_len[0]) row[x].data_off += pos;
      }
    }
    pos += (size_t)fast_body; data_left -= (uint32_t)fast_body;
    return;
  }
  if (fast_done)            // a body the buffer cannot deliver: the general rule below decides block by block
    for (uint32_t s = 0; s < 4; ++s) {
      const BandGeom& bg = res.bands[s];
      if (bg.empty) continue;
      const Rect& ci = pc.cb_idx[s];
      for (uint32_t y = 0; y < ci.h; ++y)
        for (uint32_t x = 0; x < ci.w; ++x) blocks[bg.block_base + (size_t)(ci.y0 + y) * bg.nbw + ci.x0 + x].data_off = 0;
    }
  bool body_ok = true;
  for (uint32_t s = 0; s < 4; ++s) {
    const BandGeom& bg = res.bands[s];
    if (bg.empty) continue;
    const Rect& ci = pc.cb_idx[s];
    CodedBlock* base = blocks + bg.block_base;
    for (uint32_t y = 0; y < ci.h; ++y)
      for (uint32_t x = 0; x < ci.w; ++x) {
        CodedBlock& cb = base[(size_t)(ci.y0 + y) * bg.nbw + ci.x0 + x];
        uint32_t nbytes = cb.pass_len[0] + cb.pass_len[1];
        if (data_left && body_ok) {
          if (nbytes) {
            const uint32_t want = std::min(nbytes, data_left);
            const size_t have = pos < data_end ? data_end - pos : 0;
            const uint32_t got = (uint32_t)std::min<size_t>(want, have);
            if (got == nbytes) cb.data_off = pos;
            else { cb.pass_len[0] = cb.pass_len[1] = 0; body_ok = false; }       // truncated block: not decoded
            pos += got; data_left -= got;
          }
        } else
          cb.pass_len[0] = cb.pass_len[1] = 0;
      }
  }
}

} // namespace ojb
