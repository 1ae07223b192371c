// ojb_codec.h -- encoder / decoder objects: the B200 counterparts of the reference's
// local::codestream (src/core/codestream/ojph_codestream_local.h:63-169).  exchange()/pull()
// only move lines in and out of a pinned frame buffer; all GPU work happens in flush() (encode)
// and create() (decode) -- SURVEY §8(b).
#pragma once
#include "ojb_layout.h"
#include "ojb_kernels.h"
#include <memory>
#include <cstdlib>

namespace ojb {

void cuda_check(cudaError_t e, const char* what);

struct DeviceBuf {
  void* p = nullptr; size_t cap = 0;
  ~DeviceBuf();
  void reserve(size_t n);
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
struct PinnedBuf {
  void* p = nullptr; size_t cap = 0;
  ~PinnedBuf();
  void reserve(size_t n);
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

struct Timings { float h2d_ms = 0, kernels_ms = 0, host_ms = 0, d2h_ms = 0; };

// sample container of the image buffers handed to / returned by the frame calls
enum SampleType : uint32_t { ST_U8 = 0, ST_U16 = 1, ST_I32 = 2 };

class CodecBase {
public:
  CodecBase();
  virtual ~CodecBase();
  Params params;
  Layout layout;
  cudaStream_t stream = nullptr;
  SideStream side;                     // second stream + fork / join events (see ojb_kernels.h)
  uint32_t last_launches = 0;          // kernels launched by the last frame call
  uint32_t host_threads = 4;           // host threads used for packet headers
  // stage timing of the last frame call (CUDA events on `stream`, ms)
  enum { EV_MAX = 12 };
  cudaEvent_t ev[EV_MAX] = {nullptr};
  float stage_ms[EV_MAX] = {0};
  double host_ms = 0;                  // host time between the two device phases
  bool capturing = false;              // inside run_frame's stream capture: stage marks become event-record NODES
#ifdef OJB_EMU_BUILD
  void mark(int i) { cudaEventRecord(ev[i], stream); }
#else
  void mark(int i) { if (capturing) cudaEventRecordWithFlags(ev[i], stream, cudaEventRecordExternal); else cudaEventRecord(ev[i], stream); }
#endif
  // The device work of a frame call as a CUDA graph.  Small frames are bound by the launch rate of the process
  // (about 45 launches per encode + decode), not by the SMs: the resident forms, whose launches and arguments repeat
  // frame after frame, run eagerly the first time a given argument set is seen, are captured the second time and are
  // one cudaGraphLaunch from then on.  `key` holds everything the enqueued work depends on.  OJB_NO_GRAPHS=1 disables it.
  struct FrameGraph { std::vector<uint64_t> key, seen; cudaGraphExec_t exec = nullptr; uint32_t launches = 0; bool off = false; };
  FrameGraph fgraph;
  template <typename F> void run_frame(const std::vector<uint64_t>& key, F&& enqueue);
  void drop_graph();
  void collect(int n) { for (int i = 0; i + 1 < n; ++i) cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]); }
  void upload_tables();
  DeviceBuf d_tables_enc, d_tables_dec;
  // image planes on the device: component c at byte offset img_off[c], tight rows
  DeviceBuf d_image;
  std::vector<uint64_t> img_off;
  std::vector<uint32_t> img_w, img_h;
  uint32_t img_type = ST_I32;
  uint32_t max_block_w = 64;            // widest nominal code-block of the current geometry
  // decoder: resolutions dropped from the top (codestream::restrict_input_resolution,
  // ojph_codestream_local.cpp:883-900): skip_recon levels are not reconstructed (smaller output),
  // skip_read >= skip_recon levels are not decoded (their bands read as zero)
  uint32_t skip_read = 0, skip_recon = 0;
  size_t img_bytes = 0;
  void plan_image(uint32_t sample_type);
  DeviceBuf d_coef;                    // coefficient arena (32-bit words; 64-bit elements in wide mode)
  bool wide = false;                   // some component needs more than 32 bits: every plane holds 64-bit elements, general kernels only
  // DWT jobs per level (index 0: full resolution level D, ...), uploaded once
  // jobs of one level are grouped by the kernel variant that runs them
  struct JobGroup { std::vector<DwtJob> jobs; uint32_t ctas = 0, ncomp = 1; bool first = false, stream = false, reversible = true; size_t dev_off = 0; };
  std::vector<std::vector<JobGroup>> jobs;
  DeviceBuf d_jobs;
  void build_dwt_jobs(bool forward);
  // one image over several GPUs (ojb_shard.cpp): this object works on the tiles whose mask byte is set (empty =
  // all tiles).  The geometry, arenas and block tables stay those of the whole image, so Isot, canvas coordinates
  // and byte offsets are the single-encoder ones; DWT jobs, code-blocks and packets of the other tiles are skipped.
  std::vector<uint8_t> tile_mask;
  bool tile_wanted(uint32_t t) const { return tile_mask.empty() || (t < tile_mask.size() && tile_mask[t] != 0); }
  // one image over several GPUs by ROW REGIONS (ojb_shard.cpp, partition "regions"): a tile-component is
  // cut into `world` horizontal slabs.  Encoder: rank g codes the code-blocks whose first row falls into slab g and runs,
  // at every decomposition level, only the row chunks of the streaming DWT kernels those blocks depend on -- the input
  // rows it needs are its slab plus a halo of about 2 (2^D - 1) rows (5/3) or 4 (2^D - 1) rows (9/7) each side, the
  // exact footprint of one lifting step reaching +-1 (ojph_transform.cpp:376-390), so no coefficient crosses a rank
  // boundary in mid-pipeline.  Decoder: rank g delivers the image rows of slab g, runs the synthesis chunks that
  // produce them and decodes every code-block those chunks read (its own and a halo of its neighbours').  The
  // geometry, arenas and block tables stay those of the whole image, as with a tile mask.  A tile-component whose
  // levels do not all run on the streaming kernels (degenerate sizes, DFS / ATK, 64-bit path) keeps whole-plane
  // transforms on every rank; only its block coding is shared.
  struct RegionSpec { uint32_t rank = 0, world = 0; bool on() const { return world > 1; } } region;
  struct RegionWin { uint32_t chunk0 = 0, nchunks = 0; bool full = true; };   // row chunks of one level's job; full: all of them
  std::vector<std::vector<std::vector<RegionWin>>> region_win;   // [tile][comp][resolution being split / rebuilt]
  std::vector<uint8_t> region_owner;   // encoder: the rank that codes block b
  std::vector<uint8_t> region_blocks;  // the blocks this object works on (encoder: the ones it owns; decoder: all its rows depend on)
  struct RowSpan { uint32_t lo = 0, hi = 0; };       // rows [lo, hi) in absolute tile-component coordinates
  std::vector<std::vector<RowSpan>> region_rows;     // [tile][comp]: encoder: image rows this rank reads; decoder: rows it delivers
  static RowSpan region_slab(const Rect& tc, uint32_t g, uint32_t world);
  void plan_region(bool forward);
  bool no_stream_dwt = false;          // force the general shared-memory DWT kernels (tests)
  bool no_fast_blocks = getenv("OJB_NO_FAST_BLOCKS") != nullptr;   // force the general block-coder kernels (tests, A/B)
};

class Encoder : public CodecBase {
public:
  // finalises parameters, builds geometry and device arenas, produces the main header
  void configure(const Params& p, uint32_t sample_type);
  // encode one frame whose component planes are on the host (pinned or pageable) or on the
  // device; planes[c] points at sample (0,0) of component c, stride in samples
  size_t encode(const void* const* planes, const uint32_t* strides, bool planes_on_device,
                uint8_t* out, size_t out_cap, bool out_on_device);
  uint32_t status_flags = 0;
  // row regions (ojb_shard.cpp): the frame call split at the point where code-block bytes change hands
  enum Phase : uint32_t { PHASE_ALL = 0, PHASE_FRONT = 1, PHASE_BACK = 2 };
  uint32_t phase = PHASE_ALL;
  // line-based front end (ojph::codestream::exchange): lines go to a pinned frame
  PinnedBuf h_frame;
  std::vector<uint32_t> line_cur;
  uint32_t cur_comp = 0;
  bool lines_done = false;
  int32_t* exchange(const int32_t* line_written, uint32_t& next_comp);
  std::vector<int32_t> line_buf;
  std::vector<uint8_t> main_header;
  std::vector<EncBlock> h_blocks;
  DeviceBuf d_blocks, d_results, d_status, d_slots, d_dst, d_hdr, d_pieces, d_out;
  PinnedBuf h_results, h_dst, h_hdr, h_pieces, h_status;
  std::vector<CodedBlock> coded;
  // with a tile mask the output is this object's tile-parts only (no main header, no EOC); where each one sits:
  struct TilePartOut { uint32_t tile; uint64_t offset; uint32_t bytes; };
  std::vector<TilePartOut> last_tileparts;
  size_t slot_bytes = 0;
  uint32_t num_fast_blocks = 0;        // blocks flagged ENC_FLAG_FAST
  bool few_blocks_warp = false;        // sharded image, few blocks on this rank: the warp-per-block encoder (see configure)
  // packet headers, tile-part markers and the byte layout on the device (pkt_headers.cu): no host round trip between
  // the block coder and the finished codestream.  OJB_HOST_HEADERS=1 keeps the host writer (ojb_layout.cpp), which also
  // serves the odd configuration with an empty tile-part.  With a tile mask (ojb_shard.cpp) the output is this rank's
  // tile-parts only and their positions come back in a few bytes
  bool device_headers = false;
  HdrPlanDev hplan;
  std::vector<uint8_t> fixed_blob;     // main header [+ TLM with blank lengths]
  DeviceBuf d_fixed, d_hplan[32];
  PinnedBuf h_total, h_tpout;
  std::vector<uint32_t> h_tp_tile;
  size_t hscr_bytes = 0;
  void build_header_plan();
  size_t out_cap_dev = 0;
};

struct FrameInfo {
  uint32_t width, height, off_x, off_y, num_comps;
  uint32_t bit_depth[16], is_signed[16], dx[16], dy[16], comp_w[16], comp_h[16];
  uint32_t num_decomps, reversible, color_transform, num_tiles;
  uint32_t nlt_type[16];
};

// block-coder variants, identical results.  Encoder: one thread per code-block by default (fewest
// instructions, best with several frames in flight), OJB_BLOCK_ENCODER=warp selects one warp per
// block.  Decoder: the single-pass thread-per-block kernel by default, OJB_BLOCK_DECODER=twostep selects
// step 1 (thread per block) + step 2 (warp per block).
bool serial_block_encoder();
bool serial_block_decoder();

class Decoder : public CodecBase {
public:
  bool resilient = false;
  // parse main header; (re)builds geometry when parameters changed
  void read_headers(const uint8_t* j2c, size_t len, uint32_t sample_type);
  // the same for a codestream that is in DEVICE memory only (complete, with >= 32 readable bytes after its end):
  // the marker segments and packet headers the host parsers touch are fetched page by page into a pinned mirror
  void read_headers_device(const uint8_t* dev_j2c, size_t len, uint32_t sample_type);
  struct Mirror : HostMirror {
    Decoder* owner = nullptr; const uint8_t* dev = nullptr; PinnedBuf host; size_t fetched_bytes = 0;
    std::vector<size_t> last_pages;     // pages the previous frame's parse needed
    void fetch(size_t first_page, size_t npages) override;
  } mirror;
  bool mirrored = false;                // j2c points into mirror.host
  // after read_headers, before decode: rebuilds the output planes and the synthesis schedule
  void restrict_resolution(uint32_t skipped_res_for_read, uint32_t skipped_res_for_recon);
  void setup_geometry(uint32_t sample_type);
  std::vector<uint8_t> block_res;       // per block: how many resolutions lie above its own (D_c - r)
  std::vector<uint8_t> block_wanted;    // with a tile mask: the block belongs to one of this object's tiles
  void info(FrameInfo& fi) const;
  // decode into planes (host or device); returns number of code-blocks that failed to decode
  uint32_t decode(void* const* planes, const uint32_t* strides, bool planes_on_device);
  uint32_t failed_blocks = 0;
  const uint8_t* j2c = nullptr; size_t j2c_len = 0; size_t first_sot = 0;
  const uint8_t* dev_cs = nullptr;     // optional: the same bytes already resident in HBM (+32 bytes slack)
  std::vector<uint8_t> header_sig;     // bytes of the main header the geometry was built for
  std::vector<CodedBlock> coded;
  std::vector<DecBlock> h_dec_proto;   // geometry part of DecBlock, per block
  // what the per-frame record loop of decode() needs of it: quad-record words (scratch), static flags, K_max, and
  // whether the block's shape qualifies for ht_decode_fast_kernel (dec_block_is_fast minus the frame's fields)
  struct DecStatic { uint32_t quad_words = 0; uint8_t flags = 0, K_max = 0, fast_shape = 0, pad = 0; };
  std::vector<DecStatic> block_static;
  DeviceBuf d_cs, d_dec, d_proto, d_scratch, d_bstatus;
  PinnedBuf h_dyn, h_scr, h_bstatus;
  bool any_rev_blocks = false, any_irv_blocks = false;
  void parse_tiles();
};


template <typename F> void CodecBase::run_frame(const std::vector<uint64_t>& key, F&& enqueue) {
#ifdef OJB_EMU_BUILD
  (void)key; enqueue();
#else
  static const bool disabled = [] { const char* e = getenv("OJB_NO_GRAPHS"); return e && *e && *e != '0'; }();
  FrameGraph& g = fgraph;
  if (disabled || g.off) { enqueue(); return; }
  if (g.exec && g.key == key) {
    if (cudaGraphLaunch(g.exec, stream) == cudaSuccess) { last_launches += g.launches; return; }
    cudaGetLastError(); drop_graph(); g.off = true; enqueue(); return;
  }
  if (g.seen != key) { g.seen = key; enqueue(); return; }       // first sight of this argument set: eager
  drop_graph();
  if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); g.off = true; enqueue(); return; }
  const uint32_t l0 = last_launches;
  cudaGraph_t graph = nullptr;
  capturing = true;
  try { enqueue(); } catch (...) { capturing = false; cudaStreamEndCapture(stream, &graph); if (graph) cudaGraphDestroy(graph); cudaGetLastError(); g.off = true; throw; }
  capturing = false;
  g.launches = last_launches - l0;
  cudaError_t rc = cudaStreamEndCapture(stream, &graph);
  if (rc == cudaSuccess) rc = cudaGraphInstantiate(&g.exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (rc == cudaSuccess) rc = cudaGraphLaunch(g.exec, stream);
  if (rc != cudaSuccess) {          // whatever the capture could not take: this object stays eager
    cudaGetLastError(); drop_graph(); g.off = true; last_launches = l0; enqueue(); return;
  }
  g.key = key;
#endif
}

} // namespace ojb
