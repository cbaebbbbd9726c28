// decompress_b200/csrc/tma.cuh — 1-D bulk asynchronous copies (TMA, `cp.async.bulk`) and the
// mbarrier plumbing around them, as inline PTX for sm_100a.  SASS: UBLKCP / SYNCS.
//
// Used to stage compressed input chunks global -> shared (completion signalled on an
// mbarrier) and to drain finished output shared -> global (bulk_group completion).  Both
// directions need 16-byte aligned addresses and sizes.
#pragma once
#include "simt.cuh"

namespace dcmp {

#ifdef DCMP_EMU
// CPU debugging build: synchronous copies with the same call shape.
__device__ inline void mbar_init(uint64_t *bar, int) { *bar = 0; }
__device__ inline void mbar_expect_tx(uint64_t *, uint32_t) {}
__device__ inline void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *) { memcpy(smem_dst, gsrc, bytes); }
__device__ inline void mbar_wait(uint64_t *, uint32_t) {}
__device__ inline void fence_proxy_async() {}
__device__ inline void fence_mbar_init() {}
__device__ inline void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) { memcpy(gdst, smem_src, bytes); }
__device__ inline void tma_store_commit() {}
__device__ inline void tma_store_wait_read() {}
__device__ inline void tma_store_wait_all() {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the mbarrier initialisation visible to the async proxy
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a bulk store)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif

} // namespace dcmp
