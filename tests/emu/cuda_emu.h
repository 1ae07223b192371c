// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal SIMT emulator that lets the product's .cu sources be compiled with g++ and run
// on the CPU so that kernel LOGIC (warp shuffles, ballots, shared-memory staging, block
// barriers) can be checked against the oracle in the `-m "not gpu"` test tier, where no GPU
// exists.  Every CUDA thread is a ucontext fiber; all fibers of a thread block run
// round-robin on one OS thread and yield at warp/block collectives; blocks are spread over a
// small pool of OS threads.  It is built only into tests/emu/libojph_b200_emu.so, is never
// loaded by the product (openjph_b200/_lib.py loads the nvcc-built library and fails loudly
// without it), and nothing timed by bench.py goes through it.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <functional>
#include <algorithm>

#define OJB_EMU 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct ushort2 { unsigned short x, y; };
struct ushort4 { unsigned short x, y, z, w; };
struct uchar2 { unsigned char x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

namespace ojb_emu {
struct Fiber;
struct ThreadCtx { uint3 tid, bid; dim3 bdim, gdim; };
ThreadCtx& ctx();
void* dyn_smem();
unsigned long long collective(unsigned mask, unsigned long long v, unsigned long long* all32);
void block_barrier();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}

#define threadIdx (ojb_emu::ctx().tid)
#define blockIdx (ojb_emu::ctx().bid)
#define blockDim (ojb_emu::ctx().bdim)
#define gridDim (ojb_emu::ctx().gdim)
#define warpSize 32

// ---- warp collectives -------------------------------------------------------------------
static inline unsigned emu_lane() {
  auto& c = ojb_emu::ctx();
  return (c.tid.x + c.tid.y * c.bdim.x + c.tid.z * c.bdim.x * c.bdim.y) & 31u;
}
template <typename T> static inline unsigned long long emu_pack(T v) {
  unsigned long long u = 0; memcpy(&u, &v, sizeof(T)); return u;
}
template <typename T> static inline T emu_unpack(unsigned long long u) {
  T v; memcpy(&v, &u, sizeof(T)); return v;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  unsigned long long all[32];
  ojb_emu::collective(mask, emu_pack(v), all);
  unsigned lane = emu_lane();
  unsigned base = lane & ~(unsigned)(width - 1);
  return emu_unpack<T>(all[base + ((unsigned)src & (unsigned)(width - 1))]);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
  unsigned long long all[32];
  ojb_emu::collective(mask, emu_pack(v), all);
  unsigned lane = emu_lane();
  unsigned base = lane & ~(unsigned)(width - 1);
  return (lane - base >= d) ? emu_unpack<T>(all[lane - d]) : v;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
  unsigned long long all[32];
  ojb_emu::collective(mask, emu_pack(v), all);
  unsigned lane = emu_lane();
  unsigned base = lane & ~(unsigned)(width - 1);
  return (lane - base + d < (unsigned)width) ? emu_unpack<T>(all[lane + d]) : v;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) {
  unsigned long long all[32];
  ojb_emu::collective(mask, emu_pack(v), all);
  (void)width;
  return emu_unpack<T>(all[emu_lane() ^ (unsigned)m]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned long long all[32];
  ojb_emu::collective(mask, pred ? 1ull : 0ull, all);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r |= (unsigned)(all[i] & 1ull) << i;
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) {
  unsigned long long all[32]; ojb_emu::collective(mask, 0, all);
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  unsigned long long all[32]; ojb_emu::collective(mask, v, all);
  unsigned r = 0; for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r |= (unsigned)all[i];
  return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  unsigned long long all[32]; ojb_emu::collective(mask, v, all);
  unsigned r = 0; for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r += (unsigned)all[i];
  return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  unsigned long long all[32]; ojb_emu::collective(mask, v, all);
  unsigned r = 0; for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r = std::max(r, (unsigned)all[i]);
  return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
  unsigned long long all[32]; ojb_emu::collective(mask, v, all);
  unsigned r = 0xFFFFFFFFu; for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r = std::min(r, (unsigned)all[i]);
  return r;
}
static inline unsigned __activemask() { return 0xFFFFFFFFu; }
static inline void __syncthreads() { ojb_emu::block_barrier(); }
namespace ojb_emu { void named_barrier(unsigned id, unsigned count, bool wait); }
static inline void __threadfence() { __sync_synchronize(); }
static inline void __threadfence_block() { __sync_synchronize(); }

// ---- scalar intrinsics ------------------------------------------------------------------
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  return __builtin_bswap32(v);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
  unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)(v >> (s & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
  unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32);
}
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) {
  unsigned long long v = ((unsigned long long)hi << 32) | lo; s = s > 32 ? 32 : s;
  return s == 32 ? hi : (unsigned)(v >> s);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  unsigned long long v = ((unsigned long long)b << 32) | a; unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
}
static inline int __float2int_rn(float f) { return (int)nearbyintf(f); }
static inline int __float2int_rz(float f) { return (int)f; }
static inline float __int2float_rn(int i) { return (float)i; }
static inline float __uint2float_rn(unsigned i) { return (float)i; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
using std::min; using std::max;
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

// ---- atomics ----------------------------------------------------------------------------
static inline unsigned atomicOr(unsigned* a, unsigned v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAnd(unsigned* a, unsigned v) { return __atomic_fetch_and(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v)
{ return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicMax(unsigned* a, unsigned v) {
  unsigned o = *a; while (o < v && !__atomic_compare_exchange_n(a, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o;
}
static inline unsigned atomicExch(unsigned* a, unsigned v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned* a, unsigned c, unsigned v)
{ __atomic_compare_exchange_n(a, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return c; }

// ---- runtime API subset -----------------------------------------------------------------
typedef int cudaError_t;
typedef struct emuStream_st* cudaStream_t;
typedef struct emuEvent_st* cudaEvent_t;
typedef struct emuGraphExec_st* cudaGraphExec_t;   // never instantiated: the emulator runs every frame eagerly
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostRegisterDefault = 0,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; char name[256]; size_t totalGlobalMem; int major, minor; };
cudaError_t cudaMalloc(void** p, size_t n);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t n);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaHostRegister(void* p, size_t n, unsigned flags);
cudaError_t cudaHostUnregister(void* p);
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st = 0);
cudaError_t cudaMemset(void* d, int v, size_t n);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = 0);
cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind k, cudaStream_t st = 0);
cudaError_t cudaStreamCreate(cudaStream_t* s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned f);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned f = 0);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned f);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n,
                                                                   size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice)
{ memcpy((char*)&sym + off, src, n); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyToSymbolAsync(T& sym, const void* src, size_t n,
                                                                        size_t off, cudaMemcpyKind, cudaStream_t)
{ memcpy((char*)&sym + off, src, n); return cudaSuccess; }
