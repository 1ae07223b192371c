// ojb_layout.h -- canvas geometry (tile / tile-component / resolution / sub-band / precinct /
// code-block rectangles), packet sequencing and packet headers.
//
// The GPU path replaces the reference's line-push engine with whole-plane kernels but must
// reproduce exactly the same rectangles and the same packet bytes:
//   tiles           src/core/codestream/ojph_codestream_local.cpp:113-170
//   tile-components src/core/codestream/ojph_tile.cpp:110-140,266-295
//   resolutions / sub-bands   ojph_resolution.cpp:104-124,299-337
//   code-block grid           ojph_subband.cpp:133-203, precinct index rects :224-276
//   precinct grid             ojph_resolution.cpp:400-439 (img_point), :444-458 (tag-tree levels)
//   packet header write       ojph_precinct.cpp:94-278, parse :328-573
//   packet order              ojph_tile.cpp:584-772 (write), :777-936 (parse)
#pragma once
#include <algorithm>
#include "ojb_params.h"

namespace ojb {

// what the packet layer knows about one code-block (coded_cb_header,
// src/core/codestream/ojph_codeblock.h:115-126)
struct CodedBlock {
  uint32_t pass_len[2] = {0, 0};   // cleanup length, SPP+MRP length
  uint8_t num_passes = 0;          // 0 => not included
  uint8_t missing_msbs = 0;
  uint64_t data_off = 0;           // decode: byte offset of the block's bytes in the codestream
};

struct BandGeom {
  Rect rect;                       // in sub-band coordinates
  bool empty = true;
  uint32_t band_num = 0;           // 0 LL, 1 HL, 2 LH, 3 HH
  uint32_t K_max = 0;
  float delta = 0.f, delta_inv = 0.f;    // irreversible step (already / 2^(31-K_max))
  uint32_t xcb = 0, ycb = 0;       // log2 nominal block size after precinct restriction
  uint32_t nbw = 0, nbh = 0;       // code-blocks across / down
  uint32_t block_base = 0;         // index of block (0,0) in the per-image block array
  // device plane of 32-bit words holding this band (sign-magnitude); origin at
  // (rect.x0 & ~3, rect.y0) so that every code-block row starts 16-byte aligned
  uint64_t plane_off = 0;          // word offset in the coefficient arena
  uint32_t plane_stride = 0;       // words
  uint32_t plane_pad_x = 0;        // rect.x0 & 3

  Rect block_rect(uint32_t bx, uint32_t by) const;   // in band coordinates
};

struct PrecinctGeom {
  Rect cb_idx[4];                  // code-block index rectangle per band
  uint32_t img_x = 0, img_y = 0;   // reference point on the canvas (RPCL/PCRL/CPRL ordering)
};

struct ResGeom {
  Rect rect;                       // resolution rectangle (tile-component coordinates / 2^(D-r))
  uint32_t res_num = 0;
  BandGeom bands[4];               // r == 0: bands[0]; r > 0: bands[1..3] (HL only / LH only when the level splits one way)
  uint32_t hsplit = 0, vsplit = 0; // 1: this resolution is split horizontally / vertically into resolution r - 1 and its bands
                                   // (resolution::transform_flags, ojph_resolution.cpp:290-297); both 0 at r == 0 and for a
                                   // DFS level without transform
  uint32_t dsx = 1, dsy = 1;       // down-sampling of this resolution relative to the tile-component (res_downsamp / comp_downsamp)
  uint32_t log_ppx = 15, log_ppy = 15;
  uint32_t npw = 0, nph = 0;       // precincts across / down
  std::vector<PrecinctGeom> precincts;
  // device plane holding this resolution's samples (LL of the level above), int32 or float;
  // the top resolution (r == D) with no colour/level-shift fusion reads the input image
  uint64_t plane_off = 0;
  uint32_t plane_stride = 0;
};

struct TileCompGeom {
  Rect rect;
  uint32_t comp = 0;
  std::vector<ResGeom> res;        // res[0] coarsest ... res[D]
};

struct TileGeom {
  Rect rect;
  uint32_t idx = 0;
  std::vector<TileCompGeom> comps;
};

struct PacketRef { uint32_t tile, comp, res, precinct; };

struct Layout {
  const Params* p = nullptr;
  uint32_t ntw = 0, nth = 0;
  std::vector<TileGeom> tiles;
  uint32_t num_blocks = 0;
  uint64_t coef_words = 0;         // size of the coefficient arena (32-bit words)

  void build(const Params& params);
  // packets of one tile in progression order; tile-part boundaries in tp_first
  // (index into the sequence where each tile-part starts)
  void packet_sequence(uint32_t tile, std::vector<PacketRef>& seq, std::vector<uint32_t>& tp_first,
                       std::vector<uint32_t>* tp_index = nullptr, uint32_t* tp_total = nullptr) const;
  const ResGeom& res_of(const PacketRef& pk) const
  { return tiles[pk.tile].comps[pk.comp].res[pk.res]; }
};

// packet header coder ------------------------------------------------------------------
// Appends the header of one packet (one precinct of one resolution of one tile-component,
// single quality layer) to out and returns the number of body bytes that follow it.
uint32_t write_packet_header(const ResGeom& res, const PrecinctGeom& pc,
                             const CodedBlock* blocks, std::vector<uint8_t>& out);

// Parses one packet header starting at data[pos]; fills blocks[] (lengths, passes, missing
// msbs, data_off) and advances pos past header and body.  data_left is the number of bytes
// left in the tile-part according to its SOT (Psot), data_end the size of the buffer: running out of
// buffer with data_left > 0 throws like the reference's failed file read.  Throws Error on malformed input.
// Host view of a codestream that lives in device memory: `data` is a host buffer of the same size of which only
// the 32 KB pages marked present have been fetched.  The parsers touch marker segments and packet HEADERS only
// (never code-block bodies), so a device-resident decode moves kilobytes, not the codestream, to the host.
struct HostMirror {
  enum : unsigned { PAGE_SHIFT = 15 };
  std::vector<uint8_t> present; size_t len = 0;
  virtual void fetch(size_t first_page, size_t npages) = 0;
  virtual ~HostMirror() {}
  inline void need(size_t pos, size_t n = 1) {
    if (pos >= len) return;
    size_t a = pos >> PAGE_SHIFT, b = (std::min(pos + n, len) - 1) >> PAGE_SHIFT;
    for (size_t pg = a; pg <= b; ++pg) if (!present[pg]) { fetch(pg, b - pg + 1); break; }
  }
};
void parse_packet(const Params& p, const ResGeom& res, const PrecinctGeom& pc,
                  CodedBlock* blocks, const uint8_t* data, size_t& pos, uint32_t& data_left, size_t data_end,
                  HostMirror* mirror = nullptr);

} // namespace ojb
