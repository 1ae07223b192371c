// ojb_async.cuh -- cp.async (LDGSTS) wrappers: a thread requests global bytes into shared memory
// without holding registers while they are in flight; cp.async.wait_group orders a thread's own
// copies, so a per-thread private FIFO needs no other synchronisation.
#pragma once
#include "ojb_device.h"
#include <cstring>

namespace ojb {

#ifdef OJB_EMU_BUILD
template <int N> __device__ __forceinline__ void cp_async(void* dst, const void* src) { memcpy(dst, src, N); }
__device__ __forceinline__ void cp_commit() {}
template <int N> __device__ __forceinline__ void cp_wait() {}
#else
template <int N> __device__ __forceinline__ void cp_async(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(d), "l"(src), "n"(N) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#endif

} // namespace ojb
