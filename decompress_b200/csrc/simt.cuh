// decompress_b200/csrc/simt.cuh — the few macros that differ between the real CUDA build and
// the CPU debugging build of the same kernel sources (tests/simt_emu, -DDCMP_EMU; test-only).
#pragma once
#include <stdint.h>
#ifdef DCMP_EMU
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#define DCMP_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define DCMP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
// inside a polling loop on a shared-memory flag: let the other warps (emulator: the other fibers) run
#ifdef DCMP_EMU
#define DCMP_SPIN() simt::yield_to_sched()
#else
#ifndef DCMP_SPIN_NS
#define DCMP_SPIN_NS 32
#endif
#define DCMP_SPIN() __nanosleep(DCMP_SPIN_NS)
#endif

namespace dcmp {
// Relaxed, CTA-scope loads and stores for the few places where threads of a CTA read and write the same 8/16-bit
// shared-memory cell without a barrier in between ON PURPOSE (any value a reader can see is a correct one: the jump
// table of lzo_parallel.cu, the marks of inflate_parallel.cu's orbit, the write-back of resolved symbols in
// symbol_drain.cuh).  As plain accesses these were data races — undefined behaviour the compiler may cache or tear;
// as morally-strong relaxed operations of the PTX memory model they are defined, and compute-sanitizer's racecheck
// does not flag them.
#ifdef DCMP_EMU
__device__ __forceinline__ uint16_t ld_relaxed_u16(const uint16_t *p) { return *p; }
__device__ __forceinline__ void st_relaxed_u16(uint16_t *p, uint16_t v) { *p = v; }
__device__ __forceinline__ uint8_t ld_relaxed_u8(const uint8_t *p) { return *p; }
__device__ __forceinline__ void st_relaxed_u8(uint8_t *p, uint8_t v) { *p = v; }
#else
__device__ __forceinline__ uint16_t ld_relaxed_u16(const uint16_t *p) {
  uint16_t v;
  asm volatile("ld.relaxed.cta.shared.u16 %0, [%1];" : "=h"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u16(uint16_t *p, uint16_t v) {
  asm volatile("st.relaxed.cta.shared.u16 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "h"(v) : "memory");
}
__device__ __forceinline__ uint8_t ld_relaxed_u8(const uint8_t *p) {
  uint16_t v;
  asm volatile("ld.relaxed.cta.shared.u8 %0, [%1];" : "=h"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return (uint8_t)v;
}
__device__ __forceinline__ void st_relaxed_u8(uint8_t *p, uint8_t v) {
  asm volatile("st.relaxed.cta.shared.u8 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "h"((uint16_t)v) : "memory");
}
#endif
// The same load for a hot loop: not `volatile`, no memory clobber — the compiler may schedule it like a plain load (its
// address always depends on data read after the barrier that precedes the loop, so it cannot move above it).
__device__ __forceinline__ uint32_t ld_relaxed_u16_fast(const uint16_t *p) {
#ifdef DCMP_EMU
  return *p;
#else
  uint32_t v;
  asm("ld.relaxed.cta.shared.u16 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
#endif
}
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }
} // namespace dcmp
