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
#include "simt.cuh"
#include "../../include/dcmp.h"

namespace dcmp {

constexpr int LZO_WARPS = 1;    // one warp per CTA: 64 KiB ring + input cache each, 3 CTAs per SM
constexpr int LZO_CACHE = 2048; // input bytes cached per warp
constexpr int LZO_RING = 65536;
constexpr int LZO_RING_MASK = LZO_RING - 1;
enum : int { LZ_OK = 0, LZ_MALFORMED = 1, LZ_OOB = 2, LZ_DICT = 3, LZ_OUTFULL = 4 };

struct LzoWarp {
  const uint8_t *in;
  int64_t i_pos, i_len;
  uint8_t *out;
  int64_t o_pos, o_len;
  int64_t cbase; // first input position held in the cache (-1: empty)
  uint8_t *cache;
  uint8_t *ring; // last 64 KiB of output, indexed by (o_pos & LZO_RING_MASK)
  int lane, mode;

  // byte at input position p (p < i_len), through the cache
  __device__ __forceinline__ int byte_at(int64_t p) {
    if (p < cbase || p >= cbase + LZO_CACHE) {
      __syncwarp();
      cbase = p;
      for (int k = lane; k < LZO_CACHE; k += 32) cache[k] = cbase + k < i_len ? in[cbase + k] : 0;
      __syncwarp();
    }
    return cache[p - cbase];
  }
  // lzo.ml:188-192 (+ blit bounds check lzo.ml:81-90)
  __device__ int transmit(int64_t len) {
    if (len < 0 || i_pos > i_len - len) return LZ_OOB;
    if (o_pos > o_len - len) return mode == 1 ? LZ_OUTFULL : LZ_OOB;
    for (int64_t k = lane; k < len; k += 32) {
      const int64_t p = i_pos + k;
      const uint8_t b = (p >= cbase && p < cbase + LZO_CACHE) ? cache[p - cbase] : in[p];
      ring[(o_pos + k) & LZO_RING_MASK] = b;
      out[o_pos + k] = b;
    }
    i_pos += len;
    o_pos += len;
    return LZ_OK;
  }
  // lzo.ml:194-197 / 206-216: out[o_pos + k] = out[o_pos - off + k], byte-replicating
  __device__ int copy(int64_t off, int64_t len) {
    const int64_t s = o_pos - off;
    if (mode == 1) { if (o_pos < off) return LZ_DICT; }
    else if (s < 0) return LZ_OOB;
    if (o_pos > o_len - len) return mode == 1 ? LZ_OUTFULL : LZ_OOB;
    __syncwarp(); // bytes stored by other lanes before this point are visible
    for (int64_t k = lane; k < len; k += 32) {
      const uint8_t b = ring[(s + (off >= len ? k : k % off)) & LZO_RING_MASK];
      ring[(o_pos + k) & LZO_RING_MASK] = b;
      out[o_pos + k] = b;
    }
    o_pos += len;
    return LZ_OK;
  }
  // lzo.ml:218-234
  __device__ int count(int64_t *res_out) {
    int64_t res = 0, idx = i_pos;
    while (idx < i_len && byte_at(idx) == 0) { idx++; res++; }
    if (idx < i_len) { *res_out = res * 255 + byte_at(idx); i_pos = idx + 1; return LZ_OK; }
    return LZ_MALFORMED;
  }
  // Copy (sub, state) in run, lzo.ml:286-290: len + 2 bytes, then state & 3 literals
  __device__ int do_copy(int64_t off, int64_t len, int64_t st, int64_t *state) {
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
__device__ int lzo_run(LzoWarp &w) {
  int64_t state = 0, cnt = 0;
  int r;
  LZ_NEED_BYTE();
  {
    int chr = w.byte_at(w.i_pos);
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
    LZ_NEED_BYTE();
    const int chr = w.byte_at(w.i_pos++);
    const int st = (int)(state & 3); // -1 land 3 = 3
    if (chr <= 15 && st == 0) {
      int64_t len;
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
      const int h = w.byte_at(w.i_pos++);
      r = w.do_copy(((int64_t)h << 2) + d + 1, 0, chr & 3, &state);
      if (r) return r;
    } else if (chr <= 31) {
      int64_t len = chr & 7;
      if (len == 0) { LZ_NEED_BYTE(); r = w.count(&cnt); if (r) return r; len = 7 + cnt; }
      LZ_NEED_SHORT();
      const int64_t s = w.byte_at(w.i_pos) | (w.byte_at(w.i_pos + 1) << 8);
      w.i_pos += 2;
      const int64_t off = 16384 + ((int64_t)((chr & 8) >> 3) << 14) + (s >> 2);
      if (off == 16384) return LZ_OK; // end of stream
      r = w.do_copy(off, len, s & 0xff, &state);
      if (r) return r;
    } else if (chr <= 63) {
      int64_t len = chr & 31;
      if (len == 0) { LZ_NEED_BYTE(); r = w.count(&cnt); if (r) return r; len = 31 + cnt; }
      LZ_NEED_SHORT();
      const int64_t s = w.byte_at(w.i_pos) | (w.byte_at(w.i_pos + 1) << 8);
      w.i_pos += 2;
      r = w.do_copy((s >> 2) + 1, len, s & 0xff, &state);
      if (r) return r;
    } else {
      const int len = (chr >> 5) - 1, d = (chr >> 2) & 7;
      LZ_NEED_BYTE();
      const int h = w.byte_at(w.i_pos++);
      r = w.do_copy(((int64_t)h << 3) + d + 1, len, chr, &state);
      if (r) return r;
    }
  }
}

__global__ void __launch_bounds__(LZO_WARPS * 32)
lzo_decode_kernel(int mode, int64_t nblocks, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                  uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                  int64_t *produced, int32_t *status) {
  DCMP_DYN_SMEM(smem_raw);
  uint8_t(*cache)[LZO_CACHE] = reinterpret_cast<uint8_t(*)[LZO_CACHE]>(smem_raw + (size_t)LZO_WARPS * LZO_RING);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * LZO_WARPS + warp;
  if (k >= nblocks) return;
  LzoWarp w;
  w.in = src_base + src_off[k];
  w.i_pos = 0;
  w.i_len = src_len[k];
  w.out = dst_base + dst_off[k];
  w.o_pos = 0;
  w.o_len = dst_cap[k];
  w.cbase = -(int64_t)LZO_CACHE - 1;
  w.cache = cache[warp];
  w.ring = smem_raw + (size_t)warp * LZO_RING;
  w.lane = lane;
  w.mode = mode;
  int r = lzo_run(w);
  __syncwarp();
  if (lane == 0) {
    int st = DCMP_OK;
    if (r == LZ_MALFORMED) st = DCMP_LZO_MALFORMED;
    else if (r == LZ_OOB) st = mode == 1 ? DCMP_LZO_MALFORMED : DCMP_LZO_INVALID_ARGUMENT;
    else if (r == LZ_DICT) st = DCMP_LZO_INVALID_DICTIONARY;
    else if (r == LZ_OUTFULL) st = DCMP_UNEXPECTED_END_OF_OUTPUT; // mode 1 only: the reference's Buffer.t is unbounded, grow and retry
    consumed[k] = w.i_pos;
    produced[k] = w.o_pos;
    status[k] = st;
  }
}

#ifndef DCMP_EMU
void launch_lzo_decode(int mode, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                       uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                       int64_t *produced, int32_t *status, cudaStream_t stream) {
  if (n <= 0) return;
  const int64_t blocks = (n + LZO_WARPS - 1) / LZO_WARPS;
  const size_t smem = (size_t)LZO_WARPS * (LZO_RING + LZO_CACHE);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(lzo_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  lzo_decode_kernel<<<(unsigned)blocks, LZO_WARPS * 32, smem, stream>>>(mode, n, src_base, src_off, src_len, dst_base,
                                                                    dst_off, dst_cap, consumed, produced, status);
}
#endif

} // namespace dcmp
