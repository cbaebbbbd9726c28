// tests/simt_emu/simt_emu.h — a tiny SIMT emulator for DEBUGGING kernel logic on a CPU.
//
// TEST INFRASTRUCTURE ONLY.  The build container has nvcc but no GPU, so kernel bugs would
// otherwise cost a GPU round-trip each.  This header lets the *same* kernel source
// (decompress_b200/csrc/*.cu, compiled as C++ with -DDCMP_EMU) run one CTA at a time on
// cooperative fibers: every CUDA thread is a ucontext fiber, __syncthreads / __syncwarp /
// warp collectives are scheduling points.  It is never linked into libdcmp_b200.so, is not
// a fallback, and nothing in decompress_b200/ can reach it; tests/test_emu_*.py build it into
// tests/_build/ to check kernel logic before it goes to the B200.
//
// Fidelity: any interleaving it produces is legal under CUDA's independent-thread-scheduling
// model (threads only ever switch at barriers/collectives), so code that needs lock-step
// warps fails here.  It does not model timing, bank conflicts or memory-ordering hazards —
// compute-sanitizer on the GPU box covers those.
#pragma once
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <functional>
#include <string>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__ static const
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

namespace simt {

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

enum Wait { RUN = 0, WAIT_CTA, WAIT_WARP, DONE };

struct Fiber {
  ucontext_t ctx;
  char *stack = nullptr;
  int state = RUN;
  unsigned warp_mask = 0;
  unsigned long long warp_gen = 0;
};

struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int nthreads = 0;
  int cur = -1;
  // per-warp exchange slots for collectives
  std::vector<unsigned long long> slot;      // [nthreads]
  std::vector<unsigned> warp_arrived;        // [nwarps] lanes currently waiting at a warp barrier
  std::function<void()> body;
};

extern Block *g_blk;
extern dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
extern unsigned char *g_dyn_smem;
extern long g_switches;

inline int tid() { return g_blk->cur; }
inline int lane() { return g_blk->cur & 31; }
inline int warp() { return g_blk->cur >> 5; }

inline void yield_to_sched() {
  Block *b = g_blk;
  int me = b->cur;
  g_switches++;
  swapcontext(&b->fibers[me].ctx, &b->sched);
}

inline unsigned live_mask_of_warp(int w) {
  Block *b = g_blk;
  unsigned m = 0;
  for (int l = 0; l < 32; l++) {
    int t = w * 32 + l;
    if (t < b->nthreads && b->fibers[t].state != DONE) m |= 1u << l;
  }
  return m;
}

inline void cta_barrier() {
  g_blk->fibers[g_blk->cur].state = WAIT_CTA;
  yield_to_sched();
}

inline void warp_barrier(unsigned mask) {
  Fiber &f = g_blk->fibers[g_blk->cur];
  f.state = WAIT_WARP;
  f.warp_mask = mask;
  yield_to_sched();
}

void run_block(Block &b);
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);

// ---- collectives --------------------------------------------------------------------------
template <class T> inline T exchange(unsigned mask, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "");
  Block *b = g_blk;
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b->slot[b->cur] = raw;
  warp_barrier(mask);
  unsigned long long got = b->slot[(b->cur & ~31) + (src_lane & 31)];
  if (!((mask >> (src_lane & 31)) & 1)) got = raw; // undefined in CUDA; keep own value
  warp_barrier(mask);
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}

inline unsigned ballot(unsigned mask, int pred) {
  Block *b = g_blk;
  b->slot[b->cur] = pred ? 1 : 0;
  warp_barrier(mask);
  unsigned r = 0;
  int base = b->cur & ~31;
  for (int l = 0; l < 32; l++)
    if (((mask >> l) & 1) && base + l < b->nthreads && b->slot[base + l]) r |= 1u << l;
  warp_barrier(mask);
  return r;
}

template <class T, class F> inline T reduce(unsigned mask, T v, F f) {
  Block *b = g_blk;
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b->slot[b->cur] = raw;
  warp_barrier(mask);
  int base = b->cur & ~31;
  bool first = true;
  T acc = v;
  for (int l = 0; l < 32; l++)
    if (((mask >> l) & 1) && base + l < b->nthreads) {
      T o;
      memcpy(&o, &b->slot[base + l], sizeof(T));
      acc = first ? o : f(acc, o);
      first = false;
    }
  warp_barrier(mask);
  return acc;
}

} // namespace simt

// ---- CUDA surface -----------------------------------------------------------------------------
using simt::dim3;
#define threadIdx simt::g_threadIdx
#define blockIdx simt::g_blockIdx
#define blockDim simt::g_blockDim
#define gridDim simt::g_gridDim
typedef void *cudaStream_t;

inline void __syncthreads() { simt::cta_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_barrier(mask); }
inline int __syncthreads_or(int p) {
  static int acc;
  simt::cta_barrier();
  if (simt::tid() == 0) acc = 0;
  simt::cta_barrier();
  if (p) acc = 1;
  simt::cta_barrier();
  return acc;
}
template <class T> inline T __shfl_sync(unsigned m, T v, int src, int w = 32) { return simt::exchange(m, v, src); }
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x, int w = 32) { return simt::exchange(m, v, simt::lane() ^ x); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int w = 32) {
  int s = simt::lane() - (int)d;
  return simt::exchange(m, v, s < 0 ? simt::lane() : s);
}
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int w = 32) {
  int s = simt::lane() + (int)d;
  return simt::exchange(m, v, s > 31 ? simt::lane() : s);
}
inline unsigned __ballot_sync(unsigned m, int p) { return simt::ballot(m, p); }
inline int __any_sync(unsigned m, int p) { return simt::ballot(m, p) != 0; }
inline int __all_sync(unsigned m, int p) { return simt::ballot(m, !p) == 0; }
inline unsigned __activemask() { return simt::live_mask_of_warp(simt::warp()); }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return simt::reduce(m, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return simt::reduce(m, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return simt::reduce(m, v, [](unsigned a, unsigned b) { return a + b; }); }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return simt::reduce(m, v, [](unsigned a, unsigned b) { return a | b; }); }
inline unsigned __match_any_sync(unsigned m, unsigned v) {
  simt::Block *b = simt::g_blk;
  b->slot[b->cur] = v;
  simt::warp_barrier(m);
  unsigned r = 0;
  int base = b->cur & ~31;
  for (int l = 0; l < 32; l++)
    if (((m >> l) & 1) && base + l < b->nthreads && (unsigned)b->slot[base + l] == v) r |= 1u << l;
  simt::warp_barrier(m);
  return r;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __brev(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  return __builtin_bswap32(v);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
  unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (unsigned)(v >> (s & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
  unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (unsigned)((v << (s & 31)) >> 32);
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  unsigned long long v = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
  return c;
}
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
inline void __threadfence_block() {}
inline void __threadfence() {}
inline void __nanosleep(unsigned) {}
inline void __trap() { fprintf(stderr, "simt: __trap()\n"); abort(); }

template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
inline unsigned atomicAnd(unsigned *p, unsigned v) { unsigned o = *p; *p = o & v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <class T> inline T min(T a, T b) { return a < b ? a : b; }
template <class T> inline T max(T a, T b) { return a > b ? a : b; }

#define DCMP_DYN_SMEM(name) unsigned char *name = simt::g_dyn_smem
#define DCMP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  simt::launch(simt::dim3(grid), simt::dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })

struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
