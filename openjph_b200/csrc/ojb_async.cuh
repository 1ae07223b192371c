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

// ---- named barriers: producer warps `arrive`, consumer warps `sync` (or the reverse); `count` = all threads that
// take part either way (a multiple of 32)
#ifdef OJB_EMU_BUILD
__device__ __forceinline__ void named_bar_sync(unsigned id, unsigned count) { ojb_emu::named_barrier(id, count, true); }
__device__ __forceinline__ void named_bar_arrive(unsigned id, unsigned count) { ojb_emu::named_barrier(id, count, false); }
#else
// (the barrier number is an immediate so that ptxas reserves only the barriers that are used: a register operand
// makes it reserve all 16, and the SM's barrier pool then caps the CTAs per SM)
template <unsigned ID> __device__ __forceinline__ void named_bar_sync_i(unsigned count) { asm volatile("bar.sync %0, %1;" :: "n"(ID), "r"(count) : "memory"); }
template <unsigned ID> __device__ __forceinline__ void named_bar_arrive_i(unsigned count) { asm volatile("bar.arrive %0, %1;" :: "n"(ID), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_sync(unsigned id, unsigned count) {
  switch (id) { case 1: named_bar_sync_i<1>(count); break; case 2: named_bar_sync_i<2>(count); break;
                case 3: named_bar_sync_i<3>(count); break; default: named_bar_sync_i<4>(count); break; }
}
__device__ __forceinline__ void named_bar_arrive(unsigned id, unsigned count) {
  switch (id) { case 1: named_bar_arrive_i<1>(count); break; case 2: named_bar_arrive_i<2>(count); break;
                case 3: named_bar_arrive_i<3>(count); break; default: named_bar_arrive_i<4>(count); break; }
}
#endif

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier: one elected lane requests a whole contiguous run of
// global bytes into shared memory; the copy engine completes a transaction count on an mbarrier every reader waits
// on.  Source, destination and size must be multiples of 16 bytes.
#ifdef OJB_EMU_BUILD
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned) { *b = 0; }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long*, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(unsigned long long*, unsigned) {}
__device__ __forceinline__ void fence_proxy_async() {}
#else
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(b);
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra WAIT_%=;\n}"
               :: "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

} // namespace ojb
