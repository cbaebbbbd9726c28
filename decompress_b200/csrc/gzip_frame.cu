// decompress_b200/csrc/gzip_frame.cu — gzip member framing on the device, so that a batch of members can be
// verified without the inflated bytes ever visiting the host.
//
// Replaces, for one member per unit, Gz.Inf's header / trailer handling (lib/gz.ml:463-492 header, 358-461
// FEXTRA / FNAME / FCOMMENT / FHCRC, 344-356 CRC-32 + ISIZE) and the Checkseum.Crc32 folds of lib/gz.ml:503-513.
// once inflate runs at tens of GB/s a host CRC-32 (1-3 GB/s per core) is the end-to-end
// bottleneck, and CRC-32 is parallel-reducible (crc(A||B) = crc(A) * x^(8|B|) + crc(B) over GF(2)).
//   gzip_header_kernel   one thread per member: parse the header exactly as gz.ml does (only the magic is checked;
//                        XLEN is read big-endian, gz.ml:455; FHCRC covers the 10 fixed bytes + FNAME + FCOMMENT),
//                        hand the DEFLATE body to the raw inflate kernels through rebased offset tables;
//   gzip_trailer_kernel  one warp per member: every lane folds a contiguous slice of the inflated bytes, eight
//                        bytes per step through slicing-by-8 tables in shared memory, the 32 partial CRCs are combined with carry-less multiplications
//                        by x^(8*len) mod P, then CRC-32 and ISIZE are compared (the checksum first, gz.ml:352-354).
#include "simt.cuh"
#include "../../include/dcmp.h"

namespace dcmp {

constexpr uint32_t CRC_POLY = 0xedb88320u;

// (a * b) mod P over GF(2), reflected representation (bit 31 = x^0), as zlib's multmodp
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#pragma unroll 4
  for (int i = 31; i >= 0; i--) {
    if ((a >> i) & 1u) p ^= b;
    b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
  }
  return p;
}
// x^(8 * nbytes) mod P by square-and-multiply
__device__ uint32_t gf2_xpow8n(uint64_t nbytes) {
  uint32_t sq = 0x00800000u; // x^8: bit (31 - 8)
  uint32_t p = 0x80000000u;  // x^0
  while (nbytes) {
    if (nbytes & 1ull) p = gf2_mulmod(sq, p);
    sq = gf2_mulmod(sq, sq);
    nbytes >>= 1;
  }
  return p;
}

__device__ uint32_t crc_bytes(const uint32_t *T, uint32_t c, const uint8_t *p, int64_t n) {
  for (int64_t i = 0; i < n; i++) c = T[(c ^ p[i]) & 0xffu] ^ (c >> 8);
  return c;
}

// Slicing-by-8: T8[k][b] = CRC of byte b followed by k zero bytes, so eight bytes cost eight independent
// table reads (one shared-memory round trip on the dependency chain) instead of eight dependent ones.
__device__ __forceinline__ void crc_build_slices(uint32_t (*T8)[256], int tid, int nthreads) {
  for (int i = tid; i < 256; i += nthreads) {
    uint32_t c = (uint32_t)i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? CRC_POLY ^ (c >> 1) : c >> 1;
    T8[0][i] = c;
  }
  __syncthreads();
  for (int i = tid; i < 256; i += nthreads) {
    uint32_t c = T8[0][i];
    for (int k = 1; k < 8; k++) {
      c = T8[0][c & 0xffu] ^ (c >> 8);
      T8[k][i] = c;
    }
  }
  __syncthreads();
}
__device__ uint32_t crc_bytes8(const uint32_t (*T8)[256], uint32_t c, const uint8_t *p, int64_t n) {
  int64_t i = 0;
  while (i < n && ((uintptr_t)(p + i) & 7)) { c = T8[0][(c ^ p[i]) & 0xffu] ^ (c >> 8); i++; }
  for (; i + 8 <= n; i += 8) {
    const uint2 w = *reinterpret_cast<const uint2 *>(p + i);
    const uint32_t a = w.x ^ c, b = w.y;
    c = T8[7][a & 0xffu] ^ T8[6][(a >> 8) & 0xffu] ^ T8[5][(a >> 16) & 0xffu] ^ T8[4][a >> 24] ^
        T8[3][b & 0xffu] ^ T8[2][(b >> 8) & 0xffu] ^ T8[1][(b >> 16) & 0xffu] ^ T8[0][b >> 24];
  }
  for (; i < n; i++) c = T8[0][(c ^ p[i]) & 0xffu] ^ (c >> 8);
  return c;
}

// one thread per member
__global__ void gzip_header_kernel(int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                                   int64_t *body_off, int64_t *body_len, int32_t *pre_status, int rfc_xlen) {
  __shared__ uint32_t T[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t c = (uint32_t)i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? CRC_POLY ^ (c >> 1) : c >> 1;
    T[i] = c;
  }
  __syncthreads();
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint8_t *p = src_base + src_off[k];
  const int64_t len = src_len[k];
  int st = DCMP_OK;
  int64_t i = 10;
  if (len < 10) st = DCMP_UNEXPECTED_END_OF_INPUT;
  else if (p[0] != 0x1f || p[1] != 0x8b) st = DCMP_INVALID_GZIP_HEADER;
  else {
    const int flg = p[3];
    if (flg & 4) {
      if (len - i < 2) st = DCMP_UNEXPECTED_END_OF_INPUT;
      else {
        // XLEN: big-endian in the reference (gz.ml:455); little-endian in RFC1952 and in every file other tools write
        // (BGZF ...): DCMP_GZIP_RFC_XLEN picks the latter
        const int64_t xlen = rfc_xlen ? ((int64_t)p[i + 1] << 8) | p[i] : ((int64_t)p[i] << 8) | p[i + 1];
        i += 2;
        if (len - i < xlen) st = DCMP_UNEXPECTED_END_OF_INPUT;
        else i += xlen;
      }
    }
    uint32_t crc = ~crc_bytes(T, 0xffffffffu, p, 10);
    for (int bit = 8; bit <= 16 && st == DCMP_OK; bit <<= 1) { // FNAME, then FCOMMENT: zero-terminated
      if (!(flg & bit)) continue;
      int64_t j = i;
      while (j < len && p[j] != 0) j++;
      if (j >= len) { st = DCMP_UNEXPECTED_END_OF_INPUT; break; }
      crc = ~crc_bytes(T, ~crc, p + i, j + 1 - i);
      i = j + 1;
    }
    if (st == DCMP_OK && (flg & 2)) {
      if (len - i < 2) st = DCMP_UNEXPECTED_END_OF_INPUT;
      else {
        const uint32_t have = ((uint32_t)p[i] << 8) | p[i + 1];
        if ((crc >> 16) != have) st = DCMP_INVALID_GZIP_HEADER_CHECKSUM;
        i += 2;
      }
    }
  }
  pre_status[k] = st;
  body_off[k] = src_off[k] + (st == DCMP_OK ? i : 0);
  body_len[k] = st == DCMP_OK ? len - i : 0;
}

// one warp per member
__global__ void __launch_bounds__(128)
gzip_trailer_kernel(int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *body_off, const int64_t *body_len,
                    const int32_t *pre_status, const uint8_t *dst_base, const int64_t *dst_off, int64_t *consumed,
                    int64_t *produced, int32_t *status) {
  __shared__ uint32_t T8[8][256];
  crc_build_slices(T8, threadIdx.x, blockDim.x);
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= n) return;
  const int pre = pre_status[k];
  int st = status[k];
  if (pre != DCMP_OK) {
    if (lane == 0) { status[k] = pre; consumed[k] = 0; produced[k] = 0; }
    return;
  }
  if (st != DCMP_OK) return; // the inflate error stands (its consumed / produced too)
  const int64_t prod = produced[k], cons = consumed[k];
  // CRC-32 of dst[0..prod): contiguous slices, raw remainders (no pre/post conditioning), combined left to right
  const uint8_t *o = dst_base + dst_off[k];
  const int64_t per = (prod + 31) / 32;
  const int64_t lo = per * lane < prod ? per * lane : prod, hi = lo + per < prod ? lo + per : prod;
  uint32_t c = crc_bytes8(T8, lane == 0 ? 0xffffffffu : 0u, o + lo, hi - lo);
  // tree combine: at step s, lane l (l % 2^(s+1) == 0) absorbs lane l + 2^s, whose span is known from lo/hi
  for (int s = 1; s < 32; s <<= 1) {
    const uint32_t right = __shfl_down_sync(0xffffffffu, c, s);
    const int64_t right_lo = __shfl_down_sync(0xffffffffu, lo, s);
    int64_t span_end = per * (lane + 2 * s) < prod ? per * (lane + 2 * s) : prod; // end of the absorbed lanes' range
    if (lane + s < 32 && (lane & (2 * s - 1)) == 0) {
      const int64_t rlen = span_end - right_lo;
      c = rlen > 0 ? gf2_mulmod(gf2_xpow8n((uint64_t)rlen), c) ^ right : c;
    }
  }
  if (lane == 0) {
    const uint32_t crc = ~c;
    const int64_t blen = body_len[k];
    if (blen - cons < 8) st = DCMP_UNEXPECTED_END_OF_INPUT;
    else {
      const uint8_t *t = src_base + body_off[k] + cons;
      const uint32_t want = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
      const uint32_t isize = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
      if (crc != want) st = DCMP_INVALID_CHECKSUM; // lib/gz.ml:352-354: the checksum is compared first
      else if ((uint32_t)prod != isize) st = DCMP_INVALID_INPUT_SIZE;
      else consumed[k] = cons + (body_off[k] - src_off[k]) + 8;
    }
    status[k] = st;
  }
}


// Gz.Def's framing around a DEFLATE body the deflate kernels wrote at dst + 10 (lib/gz.ml:794-813 header: magic, CM = 8,
// FLG = 0, MTIME = 0 — big-endian in the reference, all zero here —, XFL = 2 for level 9 and above, OS; lib/gz.ml:725-733
// trailer: CRC-32 and ISIZE of the input, little-endian).  One warp per stream; the CRC-32 of the input is folded
// like gzip_trailer_kernel folds the output.
__global__ void __launch_bounds__(128)
gzip_wrap_kernel(int64_t n, int level, int os, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                 uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *written, int32_t *status) {
  __shared__ uint32_t T8[8][256];
  crc_build_slices(T8, threadIdx.x, blockDim.x);
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= n) return;
  if (status[k] != DCMP_OK) return;
  if (dst_cap[k] < 18) { if (lane == 0) { status[k] = DCMP_UNEXPECTED_END_OF_OUTPUT; written[k] = 0; } return; }
  const uint8_t *p = src_base + src_off[k];
  const int64_t len = src_len[k];
  const int64_t per = (len + 31) / 32;
  const int64_t lo = per * lane < len ? per * lane : len, hi = lo + per < len ? lo + per : len;
  uint32_t c = crc_bytes8(T8, lane == 0 ? 0xffffffffu : 0u, p + lo, hi - lo);
  for (int s = 1; s < 32; s <<= 1) {
    const uint32_t right = __shfl_down_sync(0xffffffffu, c, s);
    const int64_t right_lo = __shfl_down_sync(0xffffffffu, lo, s);
    const int64_t span_end = per * (lane + 2 * s) < len ? per * (lane + 2 * s) : len;
    if (lane + s < 32 && (lane & (2 * s - 1)) == 0) {
      const int64_t rlen = span_end - right_lo;
      c = rlen > 0 ? gf2_mulmod(gf2_xpow8n((uint64_t)rlen), c) ^ right : c;
    }
  }
  if (lane == 0) {
    const uint32_t crc = ~c, isize = (uint32_t)len;
    uint8_t *o = dst_base + dst_off[k];
    const int64_t body = written[k];
    const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, (uint8_t)(level >= 9 ? 2 : 0), (uint8_t)os};
    for (int i = 0; i < 10; i++) o[i] = hdr[i];
    uint8_t *t = o + 10 + body;
    for (int i = 0; i < 4; i++) { t[i] = (uint8_t)(crc >> (8 * i)); t[4 + i] = (uint8_t)(isize >> (8 * i)); }
    written[k] = body + 18;
  }
}
#ifndef DCMP_EMU
void launch_gzip_header(int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len, int64_t *body_off,
                        int64_t *body_len, int32_t *pre_status, int rfc_xlen, cudaStream_t stream) {
  gzip_header_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, src_base, src_off, src_len, body_off, body_len, pre_status, rfc_xlen);
}
void launch_gzip_trailer(int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *body_off, const int64_t *body_len,
                         const int32_t *pre_status, const uint8_t *dst_base, const int64_t *dst_off, int64_t *consumed,
                         int64_t *produced, int32_t *status, cudaStream_t stream) {
  gzip_trailer_kernel<<<(unsigned)((n + 3) / 4), 128, 0, stream>>>(n, src_base, src_off, body_off, body_len, pre_status, dst_base,
                                                                  dst_off, consumed, produced, status);
}
void launch_gzip_wrap(int64_t n, int level, int os, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                      uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *written, int32_t *status,
                      cudaStream_t stream) {
  gzip_wrap_kernel<<<(unsigned)((n + 3) / 4), 128, 0, stream>>>(n, level, os, src_base, src_off, src_len, dst_base, dst_off, dst_cap,
                                                               written, status);
}
#endif

} // namespace dcmp
