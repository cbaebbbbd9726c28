// decompress_b200/csrc/symbol_drain.cuh — the last stage of both parallel decoders (inflate_parallel.cu,
// lzo_parallel.cu): a buffer of 16-bit symbols in shared memory — a byte, or a marker 256 + (index of the
// symbol that holds the byte) left by a back-reference whose source was not known yet — is resolved by
// pointer chasing and drained to HBM with aligned 128-bit stores; Adler-32 (zlib streams) is folded over
// the same registers.  Replaces WInf.blit/fill/update/tail (lib/de.ml:436-508) and Lzo's `copy`
// (lib/lzo.ml:194-216).  SM: the kernel's shared-memory struct (sym, adl_s, adl_b, adler_a, adler_b,
// out_pos; THREADS = CTA size).
#pragma once
#include "simt.cuh"
#include "../../include/dcmp.h"

#ifndef DCMP_HIST_GLOBAL
#define DCMP_HIST_GLOBAL 1                // 1: read the history from the output buffer in HBM/L2 instead of a shared ring
#endif
#ifndef STAT
#define STAT(x)
#endif
#ifndef DCMP_CHASE_GUARD
#define DCMP_CHASE_GUARD 0                // phase E: step over resolved pairs in the later chase rounds (measured: slower, 131.7 vs 139.8 GB/s)
#endif
#ifndef DCMP_E_WRITEBACK
#define DCMP_E_WRITEBACK 1                // phase E: resolved chunks are written back (path compression)
#endif
#ifndef DCMP_E_RELAXED
#define DCMP_E_RELAXED 0                  // phase E chase loads: 0 plain, 1 relaxed (schedulable), 2 relaxed volatile
#endif

namespace dcmp {

constexpr int HIST = 32768;               // history ring (the sliding window, de.ml:439-440) in bytes
constexpr int HIST_MASK = HIST - 1;

// One 16-symbol chunk (symbol indices [i0, i0 + 16), packed into four words) leaves for HBM — one aligned 128-bit
// store, byte-wise at the ragged ends of the segment [bias, E) — and is folded into the Adler-32 partial sums.
template <class SM>
__device__ __forceinline__ void emit_chunk16(SM &S, const int fmt, uint8_t *const G0, const uint32_t g0_low,
                                             const uint32_t bias, const uint32_t E, const uint32_t i0, uint32_t (&wv)[4],
                                             unsigned long long &s_sum, long long &b_sum) {
          const bool full = i0 >= bias && i0 + 16u <= E;
          if (full) {
            uint4 q = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            *(uint4 *)(G0 + i0) = q;
#if !DCMP_HIST_GLOBAL
            *(uint4 *)(S.ring + ((g0_low + i0) & HIST_MASK)) = q;
#endif
          } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              uint32_t m = 0;
#pragma unroll
              for (int j = 0; j < 4; j++) {
                uint32_t idx = i0 + 4u * k + j;
                if (idx >= bias && idx < E) {
                  uint8_t b = (uint8_t)(wv[k] >> (8 * j));
                  G0[idx] = b;
#if !DCMP_HIST_GLOBAL
                  S.ring[(g0_low + idx) & HIST_MASK] = b;
#endif
                  m |= 0xffu << (8 * j);
                }
              }
              wv[k] &= m;
            }
          }
          if (fmt == DCMP_FMT_ZLIB) {
            uint32_t sb = 0, bb = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              sb = __dp4a(wv[k], 0x01010101u, sb);
              // weight of byte j of word k: distance to the chunk end = 16 - (4k + j)
              bb = __dp4a(wv[k], (uint32_t)(((16 - 4 * k) & 0xff) | (((15 - 4 * k) & 0xff) << 8) |
                                            (((14 - 4 * k) & 0xff) << 16) | (((13 - 4 * k) & 0xff) << 24)), bb);
            }
            // bytes of the round after this chunk's end (negative by < 16 for a ragged last chunk)
            long long after = (long long)E - (long long)(i0 + 16u);
            s_sum += sb;
            b_sum += (long long)bb + (long long)sb * after;
          }
}

// Close a segment of seg_out bytes that started at stream position out0: reduce the Adler-32 partial sums over the
// CTA, append them to the stream's running (A, B) and advance the output cursor.  All threads call.
template <class SM>
__device__ __forceinline__ void segment_finish(SM &S, const int fmt, const uint32_t out0, const uint32_t round_out,
                                               unsigned long long s_sum, long long b_sum, const int tid, const int lane,
                                               const int warp) {
        if (fmt == DCMP_FMT_ZLIB) {
          for (int o = 16; o > 0; o >>= 1) {
            s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o);
            b_sum += __shfl_xor_sync(0xffffffffu, b_sum, o);
          }
          if (lane == 0) { S.adl_s[warp] = s_sum; S.adl_b[warp] = b_sum; }
        }
        __syncthreads();
        if (tid == 0) {
          if (fmt == DCMP_FMT_ZLIB && round_out) {
            unsigned long long ts = 0;
            long long tb = 0;
            for (int w = 0; w < SM::THREADS / 32; w++) { ts += S.adl_s[w]; tb += S.adl_b[w]; }
            const unsigned long long M = 65521ull;
            // (A,B) after appending a segment of n bytes with byte sum s and suffix-weighted sum b:
            // B' = B + n*A + b ; A' = A + s
            unsigned long long a = S.adler_a, b = S.adler_b;
            b = (b + (round_out % M) * a + (unsigned long long)tb % M) % M;
            a = (a + ts % M) % M;
            S.adler_a = (uint32_t)a;
            S.adler_b = (uint32_t)b;
          }
          S.out_pos = out0 + round_out;
          STAT(g_stats.out_bytes += round_out;)
        }
        __syncthreads();
}

// ---- E: resolve markers, drain to HBM (and to the history ring), fold Adler-32.  All threads call.
__device__ __forceinline__ uint32_t sym_load(const uint16_t *p) {
#if DCMP_E_RELAXED == 2
  return ld_relaxed_u16(p);
#elif DCMP_E_RELAXED == 1
  return ld_relaxed_u16_fast(p);
#else
  return *p;
#endif
}

template <class SM>
__device__ __forceinline__ void phase_e_drain(SM &S, const int fmt, uint8_t *const G0, const uint32_t g0_low,
                                              const uint32_t bias, const uint32_t E, const uint32_t out0,
                                              const uint32_t round_out, const int tid, const int lane, const int warp) {
        const uint32_t nchunks = (E + 15u) >> 4;
        unsigned long long s_sum = 0;
        long long b_sum = 0;
        for (uint32_t c = (uint32_t)tid; c < nchunks; c += SM::THREADS) {
          const uint32_t i0 = c << 4;
          uint4 qa = *(const uint4 *)(S.sym + i0);
          uint4 qb = *(const uint4 *)(S.sym + i0 + 8);
          uint32_t pr[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
          uint32_t wv[4];
          const bool any_marker = ((qa.x | qa.y | qa.z | qa.w | qb.x | qb.y | qb.z | qb.w) & 0xff00ff00u) != 0u;
          if (any_marker) {
            // chase in rounds: every unresolved symbol of the chunk hops once per round, so up to 16 independent
            // shared-memory loads are in flight instead of one dependent chain per symbol
            bool pending = true;
            bool first = true;
            while (pending) {
              uint32_t acc = 0;
#pragma unroll
              for (int q = 0; q < 8; q++) {
                // (after the first hop most pairs are resolved: later rounds step over them with one test)
                if (!DCMP_CHASE_GUARD || first || (pr[q] & 0xff00ff00u) != 0u) {
                  uint32_t lo = pr[q] & 0xffffu, hi = pr[q] >> 16;
                  if (lo >= 256u) { lo = sym_load(S.sym + (lo - 256u)); STAT(g_stats.chase_hops++;) }
                  if (hi >= 256u) { hi = sym_load(S.sym + (hi - 256u)); STAT(g_stats.chase_hops++;) }
                  pr[q] = lo | (hi << 16);
                  acc |= pr[q];
                }
              }
              first = false;
              pending = (acc & 0xff00ff00u) != 0u;
            }
            // path compression: later chases that land here stop after one hop.  A reader racing with the write-back sees
            // the old marker or the final literal, never a mix (a 16-bit cell is written whole) — both lead to the same byte
#if DCMP_E_WRITEBACK == 1
            *(uint4 *)(S.sym + i0) = make_uint4(pr[0], pr[1], pr[2], pr[3]);
            *(uint4 *)(S.sym + i0 + 8) = make_uint4(pr[4], pr[5], pr[6], pr[7]);
#elif DCMP_E_WRITEBACK == 2
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const uint32_t was = q < 4 ? (q == 0 ? qa.x : q == 1 ? qa.y : q == 2 ? qa.z : qa.w) : (q == 4 ? qb.x : q == 5 ? qb.y : q == 6 ? qb.z : qb.w);
              if ((was & 0xffffu) >= 256u) st_relaxed_u16(S.sym + i0 + 2 * q, (uint16_t)pr[q]);
              if ((was >> 16) >= 256u) st_relaxed_u16(S.sym + i0 + 2 * q + 1, (uint16_t)(pr[q] >> 16));
            }
#endif
          }
#pragma unroll
          for (int k = 0; k < 4; k++) wv[k] = __byte_perm(pr[2 * k], pr[2 * k + 1], 0x6420);
          emit_chunk16(S, fmt, G0, g0_low, bias, E, i0, wv, s_sum, b_sum);
        }
        segment_finish(S, fmt, out0, round_out, s_sum, b_sum, tid, lane, warp);
}

// A run of n raw bytes (a stored DEFLATE payload, de.ml:837-865 / 1613-1627; a long LZO1X literal run,
// lzo.ml:188-192) goes straight from the input to the output with 128-bit loads and stores.  The bytes
// pass through registers once — re-aligned with funnel shifts, Adler-32 folded on the way — and never
// touch shared memory.  Only 16-byte granules that hold a byte of [in_lo, in_hi) are read.  All threads call.
template <class SM>
__device__ __forceinline__ void raw_copy(SM &S, const int fmt, const uint8_t *sp, const uint32_t n, uint8_t *const dp,
                                         const uint32_t out0, const uint8_t *in_lo, const uint8_t *in_hi, const int tid,
                                         const int lane, const int warp) {
  constexpr int NT = SM::THREADS;
  const uint32_t bias = (uint32_t)((uintptr_t)dp & 15);
  uint8_t *const G0 = dp - bias;
  const uint32_t g0_low = (uint32_t)((uintptr_t)G0 & HIST_MASK);
  const uint32_t E = bias + n;
  const uint32_t nchunks = (E + 15u) >> 4;
  // symbol index i <-> source address SA + i
  const uintptr_t SA = (uintptr_t)sp - bias;
  const uint32_t a = (uint32_t)(SA & 15);
  const uintptr_t AB = SA - a;
  const uintptr_t lo_ok = (uintptr_t)in_lo & ~(uintptr_t)15, hi_ok = (uintptr_t)in_hi;
  const uint32_t q = a >> 2, r8 = 8u * (a & 3u);
  unsigned long long s_sum = 0;
  long long b_sum = 0;
  constexpr int U = 4; // chunks in flight per thread
  for (uint32_t c0 = 0; c0 < nchunks; c0 += U * NT) {
    uint4 L0[U], L1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t c = c0 + (uint32_t)u * NT + (uint32_t)tid;
      const uintptr_t g = AB + 16ull * c;
      L0[u] = make_uint4(0, 0, 0, 0);
      L1[u] = make_uint4(0, 0, 0, 0);
      if (c < nchunks) {
        if (g >= lo_ok && g < hi_ok) L0[u] = __ldcs(reinterpret_cast<const uint4 *>(g));
        if (a != 0u && g + 16 >= lo_ok && g + 16 < hi_ok) L1[u] = __ldcs(reinterpret_cast<const uint4 *>(g + 16));
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t c = c0 + (uint32_t)u * NT + (uint32_t)tid;
      if (c >= nchunks) continue;
      const uint32_t w[8] = {L0[u].x, L0[u].y, L0[u].z, L0[u].w, L1[u].x, L1[u].y, L1[u].z, L1[u].w};
      uint32_t x[5];
#pragma unroll
      for (int k = 0; k < 5; k++) x[k] = q == 0 ? w[k] : q == 1 ? w[k + 1] : q == 2 ? w[k + 2] : w[(k + 3) & 7];
      uint32_t wv[4];
#pragma unroll
      for (int k = 0; k < 4; k++) wv[k] = __funnelshift_r(x[k], x[k + 1], r8);
      emit_chunk16(S, fmt, G0, g0_low, bias, E, c << 4, wv, s_sum, b_sum);
    }
  }
  segment_finish(S, fmt, out0, n, s_sum, b_sum, tid, lane, warp);
}

} // namespace dcmp
