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
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }
} // namespace dcmp
