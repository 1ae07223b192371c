// raster.cu -- sample layouts of the files either side of the codec (SURVEY §8 f, N2).
//
// The reference's apps turn file payloads into si32 lines and back on the CPU, one line at a time:
// .pgm/.ppm = interleaved components, 8-bit or 16-bit big-endian (ppm_in::read,
// src/apps/others/ojph_img_io.cpp:338-374; ppm_out + gen_cvrt_32b{1,3}c_to_{8ub,16ub_be},
// :99-235, :539-556); .yuv/.raw = component planes one after the other, little-endian (yuv_in::read
// :1350-1378, yuv_out::write :1477-1516).  The planar layout is what the DWT kernels read and write
// directly (their stores clamp to [0, 2^B - 1] as yuv_out does); for the interleaved layout these two
// kernels move a whole frame between the file's pixel order and the codec's component planes in HBM,
// so the payload crosses PCIe exactly once in the form the file has.
#include "ojb_device.h"
#include "ojb_kernels.h"

namespace ojb {

namespace {

template <typename T> __device__ __forceinline__ uint32_t load_be(const T* p);
template <> __device__ __forceinline__ uint32_t load_be<uint8_t>(const uint8_t* p) { return *p; }
template <> __device__ __forceinline__ uint32_t load_be<uint16_t>(const uint16_t* p) { const uint32_t v = *p; return ((v & 0xFFu) << 8) | (v >> 8); }

// interleaved pixels (nc samples each, big-endian) -> component planes (native order)
template <typename T, int NC>
__global__ void __launch_bounds__(256)
raster_unpack_kernel(const T* __restrict__ src, uint8_t* __restrict__ image, RasterPlanes pl, uint32_t width, uint32_t height)
{
  // a thread converts four consecutive pixels of one row (4 * NC samples in, NC vectors of 4 out)
  const uint32_t groups = (width + 3) >> 2;
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = (uint32_t)(idx / groups);
  if (y >= height) return;
  const uint32_t x = (uint32_t)(idx % groups) << 2;
  const T* s = src + ((size_t)y * width + x) * NC;
  const uint32_t n = min(4u, width - x);
  uint32_t v[NC][4];
  #pragma unroll
  for (int i = 0; i < 4; ++i)
    #pragma unroll
    for (int c = 0; c < NC; ++c) v[c][i] = ((uint32_t)i < n) ? load_be<T>(s + i * NC + c) : 0u;
  #pragma unroll
  for (int c = 0; c < NC; ++c) {
    T* d = reinterpret_cast<T*>(image + pl.off[c]) + (size_t)y * pl.stride[c] + x;
    const bool vec = n == 4 && (((size_t)d) & (4 * sizeof(T) - 1)) == 0;
    if (vec) {
      if (sizeof(T) == 2) *reinterpret_cast<uint2*>(d) = make_uint2(v[c][0] | (v[c][1] << 16), v[c][2] | (v[c][3] << 16));
      else *reinterpret_cast<uint32_t*>(d) = v[c][0] | (v[c][1] << 8) | (v[c][2] << 16) | (v[c][3] << 24);
    } else {
      for (uint32_t i = 0; i < n; ++i) d[i] = (T)v[c][i];
    }
  }
}

// component planes -> interleaved big-endian pixels.  The planes already hold values clamped to
// [0, 2^B - 1] (the DWT store does what the gen_cvrt_* converters do), so this is a pure permutation.
template <typename T, int NC>
__global__ void __launch_bounds__(256)
raster_pack_kernel(T* __restrict__ dst, const uint8_t* __restrict__ image, RasterPlanes pl, uint32_t width, uint32_t height)
{
  const uint32_t groups = (width + 3) >> 2;
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = (uint32_t)(idx / groups);
  if (y >= height) return;
  const uint32_t x = (uint32_t)(idx % groups) << 2;
  const uint32_t n = min(4u, width - x);
  uint32_t v[NC][4];
  #pragma unroll
  for (int c = 0; c < NC; ++c) {
    const T* s = reinterpret_cast<const T*>(image + pl.off[c]) + (size_t)y * pl.stride[c] + x;
    #pragma unroll
    for (int i = 0; i < 4; ++i) v[c][i] = ((uint32_t)i < n) ? (uint32_t)s[i] : 0u;
  }
  T* d = dst + ((size_t)y * width + x) * NC;
  for (uint32_t i = 0; i < n; ++i)
    #pragma unroll
    for (int c = 0; c < NC; ++c) {
      const uint32_t t = v[c][i];
      d[i * NC + c] = sizeof(T) == 2 ? (T)(((t & 0xFFu) << 8) | (t >> 8)) : (T)t;
    }
}

// .dpx image data (dpx_in::read, src/apps/others/ojph_img_io.cpp:2064-2150): 10-bit RGB, one 32-bit word per
// pixel (R in bits 31..22, G 21..12, B 11..2; "packing 1"), or 16-bit RGB samples with rows padded to whole
// 32-bit words; the file may be written in either byte order (magic number SDPX / XPDS)
template <bool SWAP>
__global__ void __launch_bounds__(256)
raster_unpack_dpx10_kernel(const uint32_t* __restrict__ src, uint8_t* __restrict__ image, RasterPlanes pl, uint32_t width, uint32_t height)
{
  const uint32_t groups = (width + 3) >> 2;
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = (uint32_t)(idx / groups);
  if (y >= height) return;
  const uint32_t x = (uint32_t)(idx % groups) << 2;
  const uint32_t n = min(4u, width - x);
  const uint32_t* s = src + (size_t)y * width + x;
  uint32_t v[3][4];
  #pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t w = ((uint32_t)i < n) ? s[i] : 0u;
    if (SWAP) w = __byte_perm(w, 0, 0x0123);
    v[0][i] = w >> 22; v[1][i] = (w >> 12) & 0x3FFu; v[2][i] = (w >> 2) & 0x3FFu;
  }
  #pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint16_t* d = reinterpret_cast<uint16_t*>(image + pl.off[c]) + (size_t)y * pl.stride[c] + x;
    if (n == 4 && (((size_t)d) & 7) == 0) *reinterpret_cast<uint2*>(d) = make_uint2(v[c][0] | (v[c][1] << 16), v[c][2] | (v[c][3] << 16));
    else for (uint32_t i = 0; i < n; ++i) d[i] = (uint16_t)v[c][i];
  }
}
template <bool SWAP>
__global__ void __launch_bounds__(256)
raster_unpack_dpx16_kernel(const uint16_t* __restrict__ src, uint8_t* __restrict__ image, RasterPlanes pl, uint32_t width,
                           uint32_t height, uint32_t row_samples)      // row_samples: 3 * width rounded up to even
{
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = (uint32_t)(idx / width), x = (uint32_t)(idx % width);
  if (y >= height) return;
  const uint16_t* s = src + (size_t)y * row_samples + 3u * x;
  #pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t t = s[c];
    if (SWAP) t = ((t & 0xFFu) << 8) | (t >> 8);
    reinterpret_cast<uint16_t*>(image + pl.off[c])[(size_t)y * pl.stride[c] + x] = (uint16_t)t;
  }
}

} // namespace

void launch_raster_unpack_dpx(const void* src, void* image, const RasterPlanes& pl, uint32_t bit_depth, bool swap,
                              uint32_t width, uint32_t height, cudaStream_t st)
{
  if (width == 0 || height == 0) return;
  uint8_t* img = static_cast<uint8_t*>(image);
  dim3 block(256);
  if (bit_depth == 10) {
    const unsigned long long threads = (unsigned long long)((width + 3) / 4) * height;
    dim3 grid((unsigned)((threads + 255) / 256));
    auto k = swap ? raster_unpack_dpx10_kernel<true> : raster_unpack_dpx10_kernel<false>;
    OJB_LAUNCH(k, grid, block, 0, st, static_cast<const uint32_t*>(src), img, pl, width, height);
  } else {
    const unsigned long long threads = (unsigned long long)width * height;
    dim3 grid((unsigned)((threads + 255) / 256));
    auto k = swap ? raster_unpack_dpx16_kernel<true> : raster_unpack_dpx16_kernel<false>;
    OJB_LAUNCH(k, grid, block, 0, st, static_cast<const uint16_t*>(src), img, pl, width, height, (3u * width + 1u) & ~1u);
  }
}

void launch_raster_unpack(const void* src, void* image, const RasterPlanes& pl, uint32_t ncomp, uint32_t bytes_per_sample,
                          uint32_t width, uint32_t height, cudaStream_t st)
{
  if (width == 0 || height == 0) return;
  const unsigned long long threads = (unsigned long long)((width + 3) / 4) * height;
  dim3 block(256), grid((unsigned)((threads + 255) / 256));
  uint8_t* img = static_cast<uint8_t*>(image);
  if (bytes_per_sample == 1) {
    auto k = ncomp == 1 ? raster_unpack_kernel<uint8_t, 1> : raster_unpack_kernel<uint8_t, 3>;
    OJB_LAUNCH(k, grid, block, 0, st, static_cast<const uint8_t*>(src), img, pl, width, height);
  } else {
    auto k = ncomp == 1 ? raster_unpack_kernel<uint16_t, 1> : raster_unpack_kernel<uint16_t, 3>;
    OJB_LAUNCH(k, grid, block, 0, st, static_cast<const uint16_t*>(src), img, pl, width, height);
  }
}

void launch_raster_pack(void* dst, const void* image, const RasterPlanes& pl, uint32_t ncomp, uint32_t bytes_per_sample,
                        uint32_t width, uint32_t height, cudaStream_t st)
{
  if (width == 0 || height == 0) return;
  const unsigned long long threads = (unsigned long long)((width + 3) / 4) * height;
  dim3 block(256), grid((unsigned)((threads + 255) / 256));
  const uint8_t* img = static_cast<const uint8_t*>(image);
  if (bytes_per_sample == 1) {
    auto k = ncomp == 1 ? raster_pack_kernel<uint8_t, 1> : raster_pack_kernel<uint8_t, 3>;
    OJB_LAUNCH(k, grid, block, 0, st, static_cast<uint8_t*>(dst), img, pl, width, height);
  } else {
    auto k = ncomp == 1 ? raster_pack_kernel<uint16_t, 1> : raster_pack_kernel<uint16_t, 3>;
    OJB_LAUNCH(k, grid, block, 0, st, static_cast<uint16_t*>(dst), img, pl, width, height);
  }
}

} // namespace ojb
