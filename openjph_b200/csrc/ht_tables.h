// ht_tables.h -- HT block-coder lookup tables, derived at start-up from the ITU-T T.814
// Annex C CxtVLC rows (ht_cxtvlc_rows.inc) and the U-VLC definition (T.814 Table 3).
//
// Reference behaviour mirrored (not copied):
//   encoder tables  : src/core/coding/ojph_block_encoder.cpp:61-66,76-193  (index
//                     (c_q<<8)|(rho<<4)|emb -> (cwd<<8)|(len<<4)|e_k)
//   encoder U-VLC   : ojph_block_encoder.cpp:196-256
//   decoder tables  : src/core/coding/ojph_block_common.cpp:124-190 (index (c_q<<7)|7 bits ->
//                     (e_k<<12)|(e_1<<8)|(rho<<4)|(u_off<<3)|len), U-VLC :196-337
#pragma once
#include <cstdint>

namespace ojb {

struct HtTables {
  // encoder: CxtVLC for the initial quad-row [0] and the other rows [1]
  uint16_t enc_vlc[2][2048];
  // encoder U-VLC, u in 0..32 : pre | pre_len<<3 | suf<<6 | suf_len<<11  (ext unused: 32-bit path)
  uint16_t enc_uvlc[33];
  // decoder: CxtVLC
  uint16_t dec_vlc[2][1024];
  // decoder: U-VLC for a quad pair. [0]: initial row, index = mode(3 bits)<<6 | 6 bits of VLC
  //          (mode = u_off0 + 2*u_off1 (+1 if both and MEL event)), [1]: other rows (256)
  //          entry = total_prefix | total_suffix<<3 | u0_suffix_len<<7 | u0_prefix<<10 | u1_prefix<<13
  uint16_t dec_uvlc0[320];
  uint16_t dec_uvlc1[256];
};

const HtTables& ht_tables();   // built once, thread-safe

} // namespace ojb
