// decompress_b200/csrc/lzo_decode.cu — LZO1X decode, one warp per block (BASELINE config C5).
//
// Replaces Lzo.uncompress / Lzo.uncompress_with_buffer (lib/lzo.ml:395-414): the opcode
// interpreter `run`/`fiber` (lzo.ml:246-393), `count` (218-234), `transmit` (188-192) and
// `copy` (194-197, byte-replicating as in uncompress_with_buffer, 206-216).  The reference
// interprets one byte at a time through a free-monad DSL; here the warp parses the opcode
// stream uniformly from a shared-memory cache of the input (filled by coalesced loads) and
// spreads the two byte movers — literal runs input->output and back-references output->output
// — over its 32 lanes.  The sliding window (LZO1X distances reach 49151) is a 64 KiB ring in
// shared memory: a back-reference costs a shared-memory round trip instead of an L2 one, and
// every produced byte is stored once to the ring and once (fire and forget) to HBM.  Blocks
// are independent, so a batch shards one block per single-warp CTA, three CTAs per SM.
//
// Error vocabulary (include/dcmp.h): DCMP_LZO_MALFORMED / _INVALID_ARGUMENT / _INVALID_DICTIONARY.
// mode 0 = Lzo.uncompress (Out_of_bound -> Invalid_argument, lzo.ml:395-403)
// mode 1 = Lzo.uncompress_with_buffer (Out_of_bound -> Malformed, distance beyond the produced
//          output -> Invalid_dictionary, lzo.ml:405-414; the reference's output Buffer.t is unbounded, so a
//          full dst is reported as DCMP_UNEXPECTED_END_OF_OUTPUT: grow it and call again)
#include <type_traits>
#include "simt.cuh"
#include "../../include/dcmp.h"

namespace dcmp {

constexpr int LZO_WARPS = 1;    // one warp per CTA: 64 KiB ring + input cache each, 3 CTAs per SM
constexpr int LZO_CACHE = 2048; // input bytes cached per warp
constexpr int LZO_RING = 65536;
constexpr int LZO_RING_MASK = LZO_RING - 1;
enum : int { LZ_OK = 0, LZ_MALFORMED = 1, LZ_OOB = 2, LZ_DICT = 3, LZ_OUTFULL = 4 };

template <typename P> // P: int for blocks below 2 GiB (the common case: half the instructions), int64_t otherwise
struct LzoWarp {
  const uint8_t *in;
  P i_pos, i_len;
  uint8_t *out;
  P o_pos, o_len;
  P cbase; // first input position held in the cache (-1: empty)
  uint8_t *cache;
  uint8_t *ring; // last 64 KiB of output, indexed by (o_pos & LZO_RING_MASK)
  int lane, mode;
  // The next 8 input bytes, in registers: an opcode, its operand bytes and up to three trailing literals are
  // picked out of it by shifts, so the common instruction costs ONE shared-memory round trip on its critical path.
  uint64_t W;
  P wbase;
  // A back-reference's byte is loaded from the ring when the copy is decoded but stored only when the next
  // copy (or the end) needs the ring to be current: the load latency overlaps the parse of the next instruction.
  uint32_t pend_pos; // output position of this lane's pending byte
  uint32_t pend_b;
  bool pend;

  __device__ __forceinline__ void flush_pending() {
    if (pend) {
      ring[pend_pos & LZO_RING_MASK] = (uint8_t)pend_b;
      out[pend_pos] = (uint8_t)pend_b;
      pend = false;
    }
  }
  // (re)load the window at i_pos (bytes past the input read as zero: the cache is zero-padded)
  __device__ __forceinline__ void load_window() {
    typedef typename std::make_unsigned<P>::type U;
    if ((U)(i_pos - cbase) > (U)(LZO_CACHE - 16)) { // keep 12 readable bytes behind i_pos
      __syncwarp();
      cbase = i_pos & ~(P)3;
      for (int k = lane; k < LZO_CACHE; k += 32) cache[k] = cbase + k < i_len ? in[cbase + k] : 0;
      __syncwarp();
    }
    const uint32_t o = (uint32_t)(i_pos - cbase);
    const uint32_t *c32 = reinterpret_cast<const uint32_t *>(cache);
    const uint32_t a = c32[o >> 2], b = c32[(o >> 2) + 1], c = c32[(o >> 2) + 2];
    const uint32_t sh = 8u * (o & 3u);
    W = (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
    wbase = i_pos;
  }
  // byte at input position p: from the register window when it is there, else through the cache
  __device__ __forceinline__ int bat(P p) {
    typedef typename std::make_unsigned<P>::type U;
    const U d = (U)(p - wbase);
    if (d < (U)8) return (int)((W >> (8u * (uint32_t)d)) & 0xffu);
    return byte_at(p);
  }

  // byte at input position p (p < i_len), through the cache
  __device__ __forceinline__ int byte_at(P p) {
    if ((typename std::make_unsigned<P>::type)(p - cbase) >= (typename std::make_unsigned<P>::type)LZO_CACHE) {
      __syncwarp();
      cbase = p;
      for (int k = lane; k < LZO_CACHE; k += 32) cache[k] = cbase + k < i_len ? in[cbase + k] : 0;
      __syncwarp();
    }
    return cache[p - cbase];
  }
  // lzo.ml:188-192 (+ blit bounds check lzo.ml:81-90)
  __device__ __forceinline__ int transmit(P len) {
    if (len < 0 || i_pos > i_len - len) return LZ_OOB;
    if (o_pos > o_len - len) return mode == 1 ? LZ_OUTFULL : LZ_OOB;
    typedef typename std::make_unsigned<P>::type U;
    if ((U)(i_pos - wbase) + (U)len <= (U)8) { // all of them are in the register window (trailing literals)
      if (lane < (int)len) {
        const uint8_t b = (uint8_t)(W >> (8u * (uint32_t)(i_pos - wbase + lane)));
        ring[(uint32_t)(o_pos + lane) & LZO_RING_MASK] = b;
        out[o_pos + lane] = b;
      }
      i_pos += len;
      o_pos += len;
      return LZ_OK;
    }
    for (P k = lane; k < len; k += 32) {
      const P p = i_pos + k;
      const uint8_t b = (typename std::make_unsigned<P>::type)(p - cbase) < (typename std::make_unsigned<P>::type)LZO_CACHE ? cache[p - cbase] : in[p];
      ring[(uint32_t)(o_pos + k) & LZO_RING_MASK] = b;
      out[o_pos + k] = b;
    }
    i_pos += len;
    o_pos += len;
    return LZ_OK;
  }
  // lzo.ml:194-197 / 206-216: out[o_pos + k] = out[o_pos - off + k], byte-replicating
  __device__ __forceinline__ int copy(P off, P len) {
    const P s = o_pos - off;
    if (mode == 1) { if (o_pos < off) return LZ_DICT; }
    else if (s < 0) return LZ_OOB;
    if (o_pos > o_len - len) return mode == 1 ? LZ_OUTFULL : LZ_OOB;
    flush_pending();
    __syncwarp(); // bytes stored by other lanes before this point are visible
    if (len <= 32) {
      if (lane < (int)len) {
        const P k = off >= len ? (P)lane : (P)lane % off;
        pend_b = ring[(uint32_t)(s + k) & LZO_RING_MASK];
        pend_pos = (uint32_t)(o_pos + lane);
        pend = true;
      }
      o_pos += len;
      return LZ_OK;
    }
    if (off >= len) {
      for (P k = lane; k < len; k += 32) {
        const uint8_t b = ring[(uint32_t)(s + k) & LZO_RING_MASK];
        ring[(uint32_t)(o_pos + k) & LZO_RING_MASK] = b;
        out[o_pos + k] = b;
      }
    } else if (len <= (P)LZO_RING - off) { // the copy overlaps itself: byte k repeats byte k mod off of the first period
      for (P k = lane; k < len; k += 32) {
        const uint8_t b = ring[(uint32_t)(s + k % off) & LZO_RING_MASK];
        ring[(uint32_t)(o_pos + k) & LZO_RING_MASK] = b;
        out[o_pos + k] = b;
      }
    } else {
      // ... and is so long that its own ring writes wrap round onto the slots of that period: the period is read
      // from the output instead (final since flush_pending + __syncwarp above; never written by this loop)
      for (P k = lane; k < len; k += 32) {
        const uint8_t b = out[s + k % off];
        ring[(uint32_t)(o_pos + k) & LZO_RING_MASK] = b;
        out[o_pos + k] = b;
      }
    }
    o_pos += len;
    return LZ_OK;
  }
  // lzo.ml:218-234
  __device__ int count(P *res_out) {
    P res = 0, idx = i_pos;
    while (idx < i_len && byte_at(idx) == 0) { idx++; res++; }
    if (idx < i_len) { *res_out = res * 255 + byte_at(idx); i_pos = idx + 1; return LZ_OK; }
    return LZ_MALFORMED;
  }
  // Copy (sub, state) in run, lzo.ml:286-290: len + 2 bytes, then state & 3 literals
  __device__ __forceinline__ int do_copy(P off, P len, int st, int *state) {
    if (i_pos >= i_len) return LZ_MALFORMED; // guard of `go`, lzo.ml:267-268
    *state = st;
    int r = copy(off, len + 2);
    if (r) return r;
    return transmit(st & 3);
  }
};

#define LZ_NEED_BYTE() do { if (w.i_pos >= w.i_len) return LZ_MALFORMED; } while (0)
// Peek/Junk (Short _): guard lzo.ml:267-268, then get_int16 bounds lzo.ml:133-135
#define LZ_NEED_SHORT() do { if (w.i_pos >= w.i_len) return LZ_MALFORMED; if (w.i_pos > w.i_len - 2) return LZ_OOB; } while (0)

// lzo.ml:315-368 (main opcode loop) preceded by lzo.ml:372-393 (first byte)
template <typename P>
__device__ int lzo_run(LzoWarp<P> &w) {
  int state = 0;
  P cnt = 0;
  int r;
  w.load_window();
  LZ_NEED_BYTE();
  {
    int chr = w.bat(w.i_pos);
    if (chr == 16) return LZ_MALFORMED; // "No dictionary at offset 0 available"
    if (chr > 17) {
      w.i_pos++;
      if (w.i_pos >= w.i_len) return LZ_MALFORMED; // guard of `go`
      state = 0;
      r = w.transmit(chr - 17);
      if (r) return r;
    }
  }
  for (;;) {
    w.load_window();
    LZ_NEED_BYTE();
    const int chr = w.bat(w.i_pos++);
    const int st = (int)(state & 3); // -1 land 3 = 3
    if (chr <= 15 && st == 0) {
      P len;
      if (chr == 0) {
        LZ_NEED_BYTE();
        r = w.count(&cnt);
        if (r) return r;
        len = 3 + 15 + cnt;
      } else {
        len = chr + 3;
      }
      if (w.i_pos >= w.i_len) return LZ_MALFORMED; // guard of `go`
      state = -1;
      r = w.transmit(len);
      if (r) return r;
    } else if (chr <= 15) {
      // st in {1,2,3}; the `-1` arm of lzo.ml:332-336 is unreachable
      const int d = chr >> 2;
      LZ_NEED_BYTE();
      const int h = w.bat(w.i_pos++);
      r = w.do_copy((P)((h << 2) + d + 1), 0, chr & 3, &state);
      if (r) return r;
    } else if (chr <= 31) {
      P len = chr & 7;
      if (len == 0) { LZ_NEED_BYTE(); r = w.count(&cnt); if (r) return r; len = 7 + cnt; }
      LZ_NEED_SHORT();
      const int s = w.bat(w.i_pos) | (w.bat(w.i_pos + 1) << 8);
      w.i_pos += 2;
      const P off = 16384 + (((chr & 8) >> 3) << 14) + (s >> 2);
      if (off == 16384) return LZ_OK; // end of stream
      r = w.do_copy(off, len, s & 0xff, &state);
      if (r) return r;
    } else if (chr <= 63) {
      P len = chr & 31;
      if (len == 0) { LZ_NEED_BYTE(); r = w.count(&cnt); if (r) return r; len = 31 + cnt; }
      LZ_NEED_SHORT();
      const int s = w.bat(w.i_pos) | (w.bat(w.i_pos + 1) << 8);
      w.i_pos += 2;
      r = w.do_copy((P)((s >> 2) + 1), len, s & 0xff, &state);
      if (r) return r;
    } else {
      const int len = (chr >> 5) - 1, d = (chr >> 2) & 7;
      LZ_NEED_BYTE();
      const int h = w.bat(w.i_pos++);
      r = w.do_copy((P)((h << 3) + d + 1), (P)len, chr, &state);
      if (r) return r;
    }
  }
}

__global__ void __launch_bounds__(LZO_WARPS * 32)
lzo_decode_kernel(int mode, int64_t nblocks, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                  uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                  int64_t *produced, int32_t *status, int only_status, unsigned long long *redo_count) {
  DCMP_DYN_SMEM(smem_raw);
  uint8_t(*cache)[LZO_CACHE] = reinterpret_cast<uint8_t(*)[LZO_CACHE]>(smem_raw + (size_t)LZO_WARPS * LZO_RING);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * LZO_WARPS + warp;
  if (k >= nblocks) return;
  // second stage of the dispatcher: only the blocks the CTA-per-block kernel (lzo_parallel.cu) gave up
  if (only_status >= 0 && status[k] != only_status) return;
  if (only_status >= 0 && redo_count && lane == 0) atomicAdd(redo_count, 1ull); // dcmp_engine_stats: blocks the fast kernel gave up
  const int64_t n_in = src_len[k], n_out = dst_cap[k];
  int r;
  int64_t cons, prod;
  if (n_in < (1ll << 30) && n_out < (1ll << 30)) {
    LzoWarp<int> w;
    w.in = src_base + src_off[k]; w.i_pos = 0; w.i_len = (int)n_in;
    w.out = dst_base + dst_off[k]; w.o_pos = 0; w.o_len = (int)n_out;
    w.cbase = -LZO_CACHE - 1; w.cache = cache[warp]; w.ring = smem_raw + (size_t)warp * LZO_RING; w.lane = lane; w.mode = mode; w.pend = false; w.W = 0; w.wbase = 0;
    r = lzo_run(w);
    w.flush_pending();
    cons = w.i_pos; prod = w.o_pos;
  } else {
    LzoWarp<int64_t> w;
    w.in = src_base + src_off[k]; w.i_pos = 0; w.i_len = n_in;
    w.out = dst_base + dst_off[k]; w.o_pos = 0; w.o_len = n_out;
    w.cbase = -(int64_t)LZO_CACHE - 1; w.cache = cache[warp]; w.ring = smem_raw + (size_t)warp * LZO_RING; w.lane = lane; w.mode = mode; w.pend = false; w.W = 0; w.wbase = 0;
    r = lzo_run(w);
    w.flush_pending();
    cons = w.i_pos; prod = w.o_pos;
  }
  __syncwarp();
  if (lane == 0) {
    int st = DCMP_OK;
    if (r == LZ_MALFORMED) st = DCMP_LZO_MALFORMED;
    else if (r == LZ_OOB) st = mode == 1 ? DCMP_LZO_MALFORMED : DCMP_LZO_INVALID_ARGUMENT;
    else if (r == LZ_DICT) st = DCMP_LZO_INVALID_DICTIONARY;
    else if (r == LZ_OUTFULL) st = DCMP_UNEXPECTED_END_OF_OUTPUT; // mode 1 only: the reference's Buffer.t is unbounded, grow and retry
    consumed[k] = cons;
    produced[k] = prod;
    status[k] = st;
  }
}

#ifndef DCMP_EMU
int configure_lzo_decode() { // per device (see configure_inflate_parallel)
  return (int)cudaFuncSetAttribute(lzo_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LZO_WARPS * (LZO_RING + LZO_CACHE));
}
void launch_lzo_decode(int mode, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                       uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                       int64_t *produced, int32_t *status, int only_status, unsigned long long *redo_count,
                       cudaStream_t stream) {
  if (n <= 0) return;
  const int64_t blocks = (n + LZO_WARPS - 1) / LZO_WARPS;
  const size_t smem = (size_t)LZO_WARPS * (LZO_RING + LZO_CACHE);
  lzo_decode_kernel<<<(unsigned)blocks, LZO_WARPS * 32, smem, stream>>>(mode, n, src_base, src_off, src_len, dst_base,
                                                                    dst_off, dst_cap, consumed, produced, status, only_status, redo_count);
}
#endif

} // namespace dcmp
