// 1-D bulk copies of the TMA unit (cp.async.bulk, SASS UBLKCP) and the mbarrier that tracks them (sm_90+; written for
// sm_100a). One thread issues a copy of a whole tile between global and shared memory; the data path does not pass
// through registers or the LSU instruction stream, so the other warps keep issuing while the tile is in flight.
// Rules (PTX ISA "cp.async.bulk"): both addresses 16-byte aligned, size a multiple of 16 bytes; a load completes on an
// mbarrier (complete_tx counts the bytes); a store completes in a bulk async-group of the issuing thread. Shared memory
// that was written by ordinary stores must be made visible to the async proxy (fence.proxy.async) before a bulk store
// reads it, and shared memory that ordinary loads have read must be fenced the same way before a bulk load overwrites it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cbgpu {

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barrier visible to the async proxy (the copy engine signals it)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one arrival that also announces `bytes` of copy traffic still to come
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// announces further bytes without arriving (several copies on one phase)
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// global -> shared, completion on the barrier
__device__ __forceinline__ void bulk_load(void *dst_shared, const void *src_global, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_shared)),
               "l"(src_global), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// shared -> global, completion in the issuing thread's bulk group
__device__ __forceinline__ void bulk_store(void *dst_global, const void *src_shared, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(smem_addr(src_shared)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed store has been read (it may be overwritten); the writes may still be in flight
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

} // namespace cbgpu
