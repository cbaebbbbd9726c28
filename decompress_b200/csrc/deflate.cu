// decompress_b200/csrc/deflate.cu — batched DEFLATE compression (BASELINE config C3).
//
// Replaces, for the batched / one-shot case, the compress side of lib/de.ml: the match finder
// De.Lz77 (de.ml:4013-4515; hash of 4 bytes * 0x9e3779b1, de.ml:4067-4071), the frequency
// counters (de.ml:2330-2351), the Huffman builder De.T (de.ml:1828-2192: lengths limited to
// 15 / 7 bits, canonical codes emitted bit-reversed), the dynamic block header
// (Def.encode_dynamic_header / code-length RLE with 16/17/18, de.ml:2100-2192, 2583-2633) and the
// LSB-first bit writer (Def.write / c_bits, de.ml:2529-2897).  The reference is a serial,
// resumable state machine; lazy matching in particular is sequential by definition, so the
// parse here is a different (greedy, warp-parallel) algorithm that produces a different but
// valid DEFLATE stream — parity for this path is "the reference's Inf and system zlib both
// inflate it back byte-identically" (SURVEY.md §8c), which tests/test_gpu_deflate.py checks.
//
// Layout.  Every input stream is cut into independent 64 KiB chunks; one warp compresses one
// chunk into one DEFLATE block (dynamic Huffman, or stored when that is smaller):
//   1 tokenise   32 positions per step: every lane hashes its 4 bytes, takes the two most recent
//                earlier positions with the same hash (two-way buckets in a shared-memory table,
//                __match_any_sync for the ones inside the step) and extends both matches four
//                bytes at a time (level 1: one candidate); from level 4 on a match yields to a
//                longer one starting at the next position (one step of lazy evaluation); the
//                parse of the 32 positions is the orbit of the carry-in position under
//                next(p) = p + (match ? len : 1), found by pointer doubling with shuffles;
//                tokens go to an L2-resident scratch, symbol frequencies to shared memory;
//   2 build      lengths of the literal/length, distance and code-length codes (sort by warp
//                bitonic network, two-queue Huffman, zlib's overflow repair), canonical codes;
//   3 pack       32 tokens per step: bit widths -> exclusive scan -> atomicOr into a
//                shared-memory bit window that is drained to HBM with coalesced 32-bit stores.
// A chunk that is not the last of its stream ends with an empty stored block, which byte-aligns
// it (the classic sync-flush marker 00 00 ff ff), so chunk outputs concatenate; a second kernel
// gathers them to their final place, adds the zlib header / Adler-32 trailer if asked to.
#include "simt.cuh"
#include "../../include/dcmp.h"

namespace dcmp {

constexpr int DF_CHUNK = 65536;
#ifndef DCMP_DF_WARPS
#define DCMP_DF_WARPS 4
#endif
constexpr int DF_WARPS = DCMP_DF_WARPS;
constexpr int DF_HASH_BITS_FAST = 11;     // levels 1..6: 8 KiB of hash table per warp, four CTAs per SM
constexpr int DF_HASH_BITS_BEST = 12;     // levels 7..12: 16 KiB, two CTAs per SM, ~2 % smaller output
constexpr int DF_SLOT = DF_CHUNK + 64;   // bytes of scratch output per chunk (stored worst case: len + 15)
constexpr int DF_WIN = 256;              // words of the output bit window (ring)
constexpr int DF_MIN_MATCH = 4;          // 3-byte matches only when close (zlib's TOO_FAR idea, de.ml:4102)
constexpr int DF_TOO_FAR = 4096;

struct DfChunk {       // built on the host by engine.cu
  int64_t src_off;     // chunk start, relative to src_base
  int32_t len;         // 0..DF_CHUNK
  int32_t stream;      // owning stream
  int32_t flags;       // bit0: last chunk of its stream
  int32_t hist;        // bytes of the same stream in front of the chunk (at most 32768): the decoder's window reaches them
};
struct DfChunkOut {
  uint32_t bytes;      // compressed bytes in the chunk's slot
  uint32_t sum_a;      // Adler-32 partials of the chunk's input: sum of bytes (mod 65521)
  uint32_t sum_b;      // ... sum of (len - j) * byte_j (mod 65521)
  uint32_t pad;
};

constexpr int DF_HEAP = 2 * 286 + 1;   // zlib's HEAP_SIZE (de.ml:233)
struct DfTree {            // De.T.make's state (de.ml:2013-2026); only alive after tokenising, so it overlays the hash table
  uint32_t freq[DF_HEAP + 3];   // leaves [0, length), then the internal nodes
  uint16_t heap[DF_HEAP + 3];
  uint16_t dad[DF_HEAP + 3];
  uint16_t depth[DF_HEAP + 3];  // depth of a subtree while the tree is built (up to 285), then the bit length of every node
};
template <int HB>
struct DfSmemT {
  static constexpr int HASH_BITS = HB;
  // The hash table is dead once the chunk is tokenised; everything that is built afterwards lives in its space: the
  // Huffman construction scratch, then (over the scratch) the output bit window, and behind them the codes.
  union {
    uint32_t head[1 << HB];   // two most recent positions per hash: [15:0] newest, [31:16] the one before
    struct {
      union {
        DfTree tree;
        uint32_t win[DF_WIN]; // output bit window: cleared and used only after the last tree is built
      };
      uint16_t code_l[288];
      uint16_t code_d[32];
      uint16_t code_bl[19];
      uint8_t len_l[288];
      uint8_t len_d[32];
      uint8_t len_bl[19];
      uint8_t rle_sym[320];
      uint8_t rle_extra[320];
    };
  };
  uint32_t freq_l[288];
  uint32_t freq_d[32];
  uint32_t freq_bl[19];
  uint32_t extra_bits; // total extra bits of the tokens (for the stored/dynamic decision)
  int n_rle;
  int hlit, hdist, hclen;
};
static_assert(sizeof(DfTree) + 2 * (288 + 32 + 19) + (288 + 32 + 19) + 640 + 16 <= sizeof(uint32_t) * (1 << DF_HASH_BITS_FAST), "what follows tokenising overlays the hash table");
static_assert(sizeof(DfTree) <= sizeof(uint32_t) * (1 << DF_HASH_BITS_FAST), "tree scratch overlays the hash table");

__device__ __constant__ uint8_t c_bl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15}; // de.ml:237

__device__ __forceinline__ uint32_t df_len_sym(uint32_t len, uint32_t *ebits, uint32_t *eval) { // len 3..258
  if (len == 258) { *ebits = 0; *eval = 0; return 285; }
  const uint32_t x = len - 3;
  const uint32_t e = x < 8 ? 0u : (uint32_t)(31 - __clz((int)x)) - 2u;
  *ebits = e;
  *eval = x & ((1u << e) - 1u);
  return 257u + 4u * e + (x >> e);
}
__device__ __forceinline__ uint32_t df_dist_sym(uint32_t dist, uint32_t *ebits, uint32_t *eval) { // dist 1..32768
  const uint32_t D = dist - 1;
  if (D < 4) { *ebits = 0; *eval = 0; return D; }
  const uint32_t n = (uint32_t)(31 - __clz((int)D));
  *ebits = n - 1;
  *eval = D & ((1u << (n - 1)) - 1u);
  return 2u * n + ((D >> (n - 1)) & 1u);
}

// ---- match extension ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t df_load32(const uint8_t *p) { // unaligned little-endian load from two aligned words
  const uintptr_t a = (uintptr_t)p;
  const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
  return __funnelshift_r(w[0], w[1], 8u * (uint32_t)(a & 3));
}
// length of the common prefix of d[a..] and d[b..] (a < b), at most maxlen; `safe` = bytes from b that may be
// read with the 8-byte-footprint word loads
__device__ __forceinline__ unsigned long long df_load64(const uint8_t *p) { // the same from two aligned 64-bit words
  const uintptr_t a = (uintptr_t)p;
  const uint2 *w = reinterpret_cast<const uint2 *>(a & ~(uintptr_t)7);
  const uint2 w0 = w[0], w1 = w[1];
  const uint32_t sh = 8u * (uint32_t)(a & 3);
  const bool hi = (a & 4) != 0; // start in the upper half of the first word
  const uint32_t x0 = hi ? w0.y : w0.x, x1 = hi ? w1.x : w0.y, x2 = hi ? w1.y : w1.x;
  return (unsigned long long)__funnelshift_r(x0, x1, sh) | ((unsigned long long)__funnelshift_r(x1, x2, sh) << 32);
}
__device__ __forceinline__ int df_match_len(const uint8_t *d, int a, int b, int maxlen, int safe, int l = 0) { // l: bytes already known equal
  // eight bytes per step while the 16-byte footprint of the loads stays inside the chunk: most matches end in the first
  const int wmax8 = maxlen < safe - 15 ? maxlen : safe - 15;
  while (l + 8 <= wmax8) {
    const unsigned long long x = df_load64(d + a + l) ^ df_load64(d + b + l);
    if (x) return l + ((__ffsll((long long)x) - 1) >> 3);
    l += 8;
  }
  const int wmax = maxlen < safe - 7 ? maxlen : safe - 7;
  while (l + 4 <= wmax) {
    const uint32_t x = df_load32(d + a + l) ^ df_load32(d + b + l);
    if (x) return l + ((__ffs((int)x) - 1) >> 3);
    l += 4;
  }
  while (l < maxlen && d[a + l] == d[b + l]) l++;
  return l;
}

// ---- output bit window -----------------------------------------------------------------------
struct DfBits {
  uint32_t *win;       // shared ring of DF_WIN words
  uint32_t *out;       // global, word aligned
  uint32_t bitpos;     // bits written so far
  uint32_t flushed;    // words already stored to global
  __device__ __forceinline__ void put(uint64_t v, uint32_t n, uint32_t at) { // n <= 57 bits of v at bit `at`
    if (n == 0) return;
    const uint32_t w = at >> 5, sh = at & 31;
    const uint64_t lo = v << sh;
    atomicOr(&win[w & (DF_WIN - 1)], (uint32_t)lo);
    if (sh + n > 32) atomicOr(&win[(w + 1) & (DF_WIN - 1)], (uint32_t)(lo >> 32));
    if (sh + n > 64) atomicOr(&win[(w + 2) & (DF_WIN - 1)], (uint32_t)(v >> (64 - sh)));
  }
  // all lanes: store the words that are complete (or everything, at the end)
  __device__ __forceinline__ void drain(int lane, bool all) {
    __syncwarp();
    const uint32_t upto = all ? (bitpos + 31) >> 5 : bitpos >> 5;
    for (uint32_t i = flushed + lane; i < upto; i += 32) {
      out[i] = win[i & (DF_WIN - 1)];
      win[i & (DF_WIN - 1)] = 0;
    }
    flushed = upto;
    __syncwarp();
  }
};

// ---- Huffman code construction ------------------------------------------------------------------
// freq[0..n) -> len[0..n) (<= max_bits) and bit-reversed canonical codes: De.T.make (de.ml:2013-2068), i.e. zlib's
// build_tree / gen_bitlen / gen_codes step for step — heap ordered by (frequency, depth), at least two codes (de.ml:1862-1874),
// overflow repair (de.ml:1978-2009) — so the code lengths are the REFERENCE's, not just valid ones (test/test.ml:1169-1237:
// "the same tree as zlib").  The merge loop is a chain of dependent heap moves: one lane runs it (a few hundred
// microseconds per 64 KiB chunk at most, ~2 % of a chunk's time; the other lanes of the SM's warps fill the gaps).
// Returns max_code (the highest symbol with a code), the same in every lane.
template <class SM>
__device__ int df_build_code(SM &S, const uint32_t *freq, int n, int max_bits, uint8_t *len, uint16_t *code, int lane) {
  DfTree &T = S.tree;
  __syncwarp();
  for (int i = lane; i < n; i += 32) { T.freq[i] = freq[i]; len[i] = 0; code[i] = 0; }
  __syncwarp();
  int max_code = -1;
  if (lane == 0) {
    auto smaller = [&](int a, int b) { return T.freq[a] < T.freq[b] || (T.freq[a] == T.freq[b] && T.depth[a] <= T.depth[b]); };
    int heap_len = 0, heap_max = DF_HEAP;
    auto down = [&](int k) { // pqdownheap, de.ml:1879-1898
      const int v = T.heap[k];
      int j = k << 1;
      while (j <= heap_len) {
        if (j < heap_len && smaller(T.heap[j + 1], T.heap[j])) j++;
        if (smaller(v, T.heap[j])) break;
        T.heap[k] = T.heap[j];
        k = j;
        j <<= 1;
      }
      T.heap[k] = (uint16_t)v;
    };
    for (int s = 0; s < n; s++) {
      T.depth[s] = 0; // (a symbol that is not used keeps length 0, de.ml:1849-1853)
      if (T.freq[s] != 0) { T.heap[++heap_len] = (uint16_t)s; max_code = s; }
    }
    while (heap_len < 2) { // the pkzip rule
      const int node = max_code < 2 ? ++max_code : 0;
      T.freq[node] = 1;
      T.heap[++heap_len] = (uint16_t)node;
      T.depth[node] = 0;
    }
    for (int k = heap_len / 2; k >= 1; k--) down(k);
    int node = n;
    do {
      const int a = T.heap[1];
      T.heap[1] = T.heap[heap_len--];
      down(1);
      const int b = T.heap[1];
      T.heap[--heap_max] = (uint16_t)a;
      T.heap[--heap_max] = (uint16_t)b;
      T.freq[node] = T.freq[a] + T.freq[b];
      T.depth[node] = (uint16_t)((T.depth[a] >= T.depth[b] ? T.depth[a] : T.depth[b]) + 1);
      T.dad[a] = (uint16_t)node;
      T.dad[b] = (uint16_t)node;
      T.heap[1] = (uint16_t)node++;
      down(1);
    } while (heap_len >= 2);
    T.heap[--heap_max] = T.heap[1];
    // generate_lengths (de.ml:1952-2009); depth[] becomes the bit length of every node
    int bl_count[16];
    for (int bq = 0; bq < 16; bq++) bl_count[bq] = 0;
    T.depth[T.heap[heap_max]] = 0;
    int overflow = 0;
    for (int h = heap_max + 1; h < DF_HEAP; h++) {
      const int m = T.heap[h];
      int bits = T.depth[T.dad[m]] + 1;
      if (bits > max_bits) { bits = max_bits; overflow++; }
      T.depth[m] = (uint16_t)bits;
      if (m <= max_code) bl_count[bits]++;
    }
    if (overflow != 0) {
      do {
        int bits = max_bits - 1;
        while (bl_count[bits] == 0) bits--;
        bl_count[bits]--;
        bl_count[bits + 1] += 2;
        bl_count[max_bits]--;
        overflow -= 2;
      } while (overflow > 0);
      int h = DF_HEAP;
      for (int bits = max_bits; bits >= 1; bits--) {
        int c = bl_count[bits];
        while (c != 0) {
          const int m = T.heap[--h];
          if (m <= max_code) { T.depth[m] = (uint16_t)bits; c--; }
        }
      }
    }
    // generate_codes (de.ml:1926-1950): canonical, bit-reversed
    uint32_t next_code[16];
    uint32_t c = 0;
    next_code[0] = 0;
    for (int bits = 1; bits <= max_bits; bits++) { c = (c + (uint32_t)bl_count[bits - 1]) << 1; next_code[bits] = c; }
    for (int s = 0; s <= max_code; s++) {
      const int l = T.depth[s];
      len[s] = (uint8_t)l;
      code[s] = l ? (uint16_t)(__brev(next_code[l]++) >> (32 - l)) : 0;
    }
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, max_code, 0);
}

// Adler-32 partials of p[0..n): sum of bytes and sum of (n - j) * p[j], both mod 65521.
__device__ void df_adler_partials(const uint8_t *p, int n, int lane, uint32_t *sum_a, uint32_t *sum_b) {
  unsigned long long a = 0, b = 0;
  for (int j = lane; j < n; j += 32) {
    const uint32_t v = p[j];
    a += v;
    b += (unsigned long long)v * (uint32_t)(n - j);
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  *sum_a = (uint32_t)(a % 65521ull);
  *sum_b = (uint32_t)(b % 65521ull);
}

// ---- a Dynamic block from the tokens' frequencies ---------------------------------------------------
// 2 build: the three codes (De.T.make each: Def.dynamic_of_frequencies, de.ml:2387-2407) and the header's RLE symbols;
// returns the size of the block in bits.  S.freq_l / S.freq_d / S.extra_bits describe the tokens.
template <class SM>
__device__ uint32_t df_build_block(SM &S, bool marker, int lane) {
  const int l_max = df_build_code(S, S.freq_l, 286, 15, S.len_l, S.code_l, lane);
  const int d_max = df_build_code(S, S.freq_d, 30, 15, S.len_d, S.code_d, lane);
  if (lane == 0) {
    // code-length RLE: T.scan + T.symbols (de.ml:2070-2190, zlib's scan_tree / send_tree), the literal/length tree and
    // the distance tree one after the other — a run never crosses from one into the other
    for (int i = 0; i < 19; i++) S.freq_bl[i] = 0;
    int nr = 0;
    for (int which = 0; which < 2; which++) {
      const uint8_t *tl = which ? S.len_d : S.len_l;
      const int max_code = which ? d_max : l_max;
      int prevlen = -1, nextlen = tl[0], count = 0, max_count = 7, min_count = 4;
      if (nextlen == 0) { max_count = 138; min_count = 3; }
      for (int n = 0; n <= max_code; n++) {
        const int curlen = nextlen;
        nextlen = n + 1 <= max_code ? tl[n + 1] : 0xffff; // (the guard of de.ml:2087)
        count++;
        if (count < max_count && curlen == nextlen) continue;
        if (count < min_count) {
          S.freq_bl[curlen] += (uint32_t)count;
          do { S.rle_sym[nr] = (uint8_t)curlen; S.rle_extra[nr++] = 0; } while (--count != 0);
        } else if (curlen != 0) {
          if (curlen != prevlen) { S.rle_sym[nr] = (uint8_t)curlen; S.rle_extra[nr++] = 0; S.freq_bl[curlen]++; count--; }
          S.rle_sym[nr] = 16; S.rle_extra[nr++] = (uint8_t)(count - 3); S.freq_bl[16]++;
        } else if (count <= 10) {
          S.rle_sym[nr] = 17; S.rle_extra[nr++] = (uint8_t)(count - 3); S.freq_bl[17]++;
        } else {
          S.rle_sym[nr] = 18; S.rle_extra[nr++] = (uint8_t)(count - 11); S.freq_bl[18]++;
        }
        count = 0;
        prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
      }
    }
    S.n_rle = nr;
    S.hlit = l_max + 1;
    S.hdist = d_max + 1;
  }
  __syncwarp();
  const int n_rle = S.n_rle;
  __syncwarp();
  df_build_code(S, S.freq_bl, 19, 7, S.len_bl, S.code_bl, lane);
  // cost of the dynamic block vs stored blocks
  uint32_t bits_dyn = 0;
  for (int s = lane; s < 286; s += 32) bits_dyn += S.freq_l[s] * S.len_l[s];
  if (lane < 30) bits_dyn += S.freq_d[lane] * S.len_d[lane];
  if (lane < 19) bits_dyn += S.freq_bl[lane] * S.len_bl[lane];
  bits_dyn = __reduce_add_sync(0xffffffffu, bits_dyn);
  int hclen = 19;
  while (hclen > 4 && S.len_bl[c_bl_order[hclen - 1]] == 0) hclen--;
  uint32_t rle_extra_bits = 0;
  for (int i = lane; i < n_rle; i += 32) rle_extra_bits += S.rle_sym[i] == 16 ? 2u : S.rle_sym[i] == 17 ? 3u : S.rle_sym[i] == 18 ? 7u : 0u;
  rle_extra_bits = __reduce_add_sync(0xffffffffu, rle_extra_bits);
  bits_dyn += 3 + 14 + 3 * (uint32_t)hclen + rle_extra_bits + S.extra_bits + (marker ? 3u + 7u + 32u : 0u);
  if (lane == 0) S.hclen = hclen;
  __syncwarp();
  return bits_dyn;
}

// 3 pack: header (encode_dynamic_header / zigzag / huffman, de.ml:2583-2624), the tokens (Def.write, de.ml:2657-2897), the end
// of block and — `marker` — an empty stored block that byte-aligns the output; returns the bytes written to out32.
template <class SM>
__device__ uint32_t df_pack_block(SM &S, const uint32_t *tok, int ntok, bool last, bool marker, uint32_t *out32, int lane,
                                  bool fixed = false) {
  const int n_rle = S.n_rle, hlit = S.hlit, hdist = S.hdist, hclen = S.hclen;
  if (fixed) {
    // the Fixed codes (de.ml:821-833, RFC1951 3.2.6) in place of the block's own: no tree in the header
    __syncwarp();
    for (int s2 = lane; s2 < 288; s2 += 32) {
      const uint32_t l = s2 < 144 ? 8u : s2 < 256 ? 9u : s2 < 280 ? 7u : 8u;
      const uint32_t c = s2 < 144 ? 0x30u + (uint32_t)s2 : s2 < 256 ? 0x190u + (uint32_t)(s2 - 144) : s2 < 280 ? (uint32_t)(s2 - 256) : 0xc0u + (uint32_t)(s2 - 280);
      S.len_l[s2] = (uint8_t)l;
      S.code_l[s2] = (uint16_t)(__brev(c) >> (32 - l));
    }
    S.len_d[lane] = 5;
    S.code_d[lane] = (uint16_t)(__brev((uint32_t)lane) >> 27);
    __syncwarp();
  }
  for (int i = lane; i < DF_WIN; i += 32) S.win[i] = 0;
  __syncwarp();
  DfBits bw{S.win, out32, 0u, 0u};
  if (lane == 0 && fixed) {
    bw.put(last ? 1u : 0u, 1, 0);
    bw.put(1u, 2, 1);
    S.n_rle = 3;
  }
  if (lane == 0 && !fixed) {
    uint32_t at = 0;
    auto put = [&](uint32_t v, uint32_t n) { bw.put(v, n, at); at += n; };
    put(last ? 1u : 0u, 1);
    put(2u, 2);
    put((uint32_t)(hlit - 257), 5);
    put((uint32_t)(hdist - 1), 5);
    put((uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) put(S.len_bl[c_bl_order[i]], 3);
    for (int i = 0; i < n_rle; i++) {
      const int s = S.rle_sym[i];
      put(S.code_bl[s], S.len_bl[s]);
      if (s == 16) put(S.rle_extra[i], 2);
      else if (s == 17) put(S.rle_extra[i], 3);
      else if (s == 18) put(S.rle_extra[i], 7);
    }
    S.n_rle = (int)at;
  }
  __syncwarp();
  bw.bitpos = (uint32_t)S.n_rle;
  bw.drain(lane, false);
  for (int g = 0; g < ntok; g += 32) {
    const int i = g + lane;
    uint64_t v = 0;
    uint32_t nb = 0;
    if (i < ntok) {
      const uint32_t t = tok[i];
      if ((t >> 23) & 1u) {
        uint32_t eb, ev, eb2, ev2;
        const uint32_t ls = df_len_sym((t & 0xffu) + 3u, &eb, &ev), ds = df_dist_sym(((t >> 8) & 0x7fffu) + 1u, &eb2, &ev2);
        v = S.code_l[ls];
        nb = S.len_l[ls];
        v |= (uint64_t)ev << nb; nb += eb;
        v |= (uint64_t)S.code_d[ds] << nb; nb += S.len_d[ds];
        v |= (uint64_t)ev2 << nb; nb += eb2;
      } else {
        v = S.code_l[t];
        nb = S.len_l[t];
      }
    }
    uint32_t incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    bw.put(v, nb, bw.bitpos + incl - nb);
    bw.bitpos += __shfl_sync(0xffffffffu, incl, 31);
    bw.drain(lane, false);
  }
  if (lane == 0) {
    uint32_t at = bw.bitpos;
    bw.put(S.code_l[256], S.len_l[256], at); at += S.len_l[256];
    if (marker) {
      // empty stored block: byte-aligns the chunk so that chunk outputs concatenate
      at += 3;                       // BFINAL=0, BTYPE=00
      at = (at + 7u) & ~7u;
      bw.put(0xffff0000ull, 32, at); // LEN=0, NLEN=0xffff
      at += 32;
    }
    S.n_rle = (int)at;
  }
  __syncwarp();
  bw.bitpos = (uint32_t)S.n_rle;
  bw.drain(lane, true);
  return (bw.bitpos + 7u) >> 3;
}

// ---- one chunk ----------------------------------------------------------------------------------
template <class SM>
__device__ void df_compress_chunk(SM &S, const uint8_t *d, int len, int hist, bool last, int level, uint32_t *tok,
                                  uint8_t *slot, DfChunkOut *res, bool want_adler, int lane) {
  uint32_t sum_a = 0, sum_b = 0;
  if (want_adler) df_adler_partials(d, len, lane, &sum_a, &sum_b);
  uint32_t *out32 = reinterpret_cast<uint32_t *>(slot);
  // ---------------------------------------------------------------- 1 tokenise
  int ntok = 0;
  bool stored = level == 0;
  if (!stored) {
    for (int i = lane; i < (1 << SM::HASH_BITS); i += 32) S.head[i] = 0;
    for (int i = lane; i < 288; i += 32) S.freq_l[i] = 0;
    if (lane < 32) S.freq_d[lane] = 0;
    if (lane == 0) S.extra_bits = 0;
    __syncwarp();
    // Positions live in the table as 16 bits (p mod 65536): a candidate is the nearest earlier position with those low
    // bits (an alias 64 KiB further back is beyond the window anyway; a wrong one fails the byte comparison).
    auto pos_of = [](int p_cur, uint32_t c16) { return p_cur - (int)(((uint32_t)p_cur - c16) & 0xffffu); };
    // The best levels first enter the `hist` bytes in front of the chunk into the table (no tokens): chunks are
    // encoded independently, but the decoder's 32 KiB window does not know that.
    for (int base = -((hist + 31) & ~31); base < 0; base += 32) {
      const int p = base + lane;
      uint32_t h = 0xffffffffu;
      if (p >= -hist) {
        const uint32_t v = (uint32_t)d[p] | ((uint32_t)d[p + 1] << 8) | ((uint32_t)d[p + 2] << 16) | ((uint32_t)d[p + 3] << 24);
        h = (v * 0x9e3779b1u) >> (32 - SM::HASH_BITS);
      }
      const uint32_t same = __match_any_sync(0xffffffffu, h);
      if (h != 0xffffffffu && (same >> lane) == 1u) {
        const uint32_t lower = same & ((1u << lane) - 1u);
        const uint32_t prev = lower ? (uint32_t)(base + (31 - __clz((int)lower))) : (S.head[h] & 0xffffu);
        S.head[h] = ((prev & 0xffffu) << 16) | ((uint32_t)p & 0xffffu);
      }
      __syncwarp();
    }
    int carry = 0; // first position of the step not covered by the previous token (may be >= 32)
    uint32_t my_extra = 0;
    for (int base = 0; base < len; base += 32) {
      if (carry >= 32) { carry -= 32; continue; } // the whole step lies inside the previous match
      const int p = base + lane;
      const int maxlen = min(258, len - p);
      int mlen = 0, mdist = 0;
      uint32_t h = 0xffffffffu; // lanes that cannot start a match do not take part in the hash table
      if (maxlen >= DF_MIN_MATCH) {
        const uint32_t v = (uint32_t)d[p] | ((uint32_t)d[p + 1] << 8) | ((uint32_t)d[p + 2] << 16) | ((uint32_t)d[p + 3] << 24);
        h = (v * 0x9e3779b1u) >> (32 - SM::HASH_BITS);
      }
      const uint32_t same = __match_any_sync(0xffffffffu, h);
      int cand0 = -1, cand1 = -1; // the two most recent earlier positions with this hash
      if (h != 0xffffffffu) {
        const uint32_t lower = same & ((1u << lane) - 1u);
        const uint32_t hd = S.head[h];
        if (lower) {
          const int l0 = 31 - __clz((int)lower);
          cand0 = base + l0;
          const uint32_t lower2 = lower & ~(1u << l0);
          cand1 = lower2 ? base + (31 - __clz((int)lower2)) : pos_of(p, hd & 0xffffu);
        } else {
          cand0 = pos_of(p, hd & 0xffffu);
          cand1 = pos_of(p, hd >> 16);
        }
      }
      __syncwarp();
      if (h != 0xffffffffu && (same >> lane) == 1u) { // the last lane of a group records the group's two newest positions
        const uint32_t lower = same & ((1u << lane) - 1u);
        const uint32_t prev = lower ? (uint32_t)(base + (31 - __clz((int)lower))) : (S.head[h] & 0xffffu);
        S.head[h] = ((prev & 0xffffu) << 16) | ((uint32_t)p & 0xffffu);
      }
      if (cand0 >= -hist && cand0 < p && p - cand0 <= 32768) {
        const bool two = level >= 2 && cand1 >= -hist && cand1 < cand0 && p - cand1 <= 32768;
        int l, dist = p - cand0;
        if (maxlen >= 8 && len - p >= 16) {
          // the first eight bytes of both candidates in one round trip: most matches end there
          const unsigned long long P8 = df_load64(d + p), A8 = df_load64(d + cand0) ^ P8;
          const unsigned long long B8 = two ? df_load64(d + cand1) ^ P8 : 1ull;
          l = A8 ? (__ffsll((long long)A8) - 1) >> 3 : df_match_len(d, cand0, p, maxlen, len - p, 8);
          if (two && l < maxlen) {
            const int l1 = B8 ? (__ffsll((long long)B8) - 1) >> 3 : df_match_len(d, cand1, p, maxlen, len - p, 8);
            if (l1 > l) { l = l1; dist = p - cand1; }
          }
        } else {
          l = df_match_len(d, cand0, p, maxlen, len - p);
          if (two && l < maxlen) {
            const int l1 = df_match_len(d, cand1, p, maxlen, len - p);
            if (l1 > l) { l = l1; dist = p - cand1; }
          }
        }
        if (l >= DF_MIN_MATCH || (l == 3 && dist < DF_TOO_FAR)) { mlen = l; mdist = dist; }
      }
      if (level >= 4) {
        // lazy evaluation, one step (the idea of de.ml:4351-4410): a match loses to a longer one starting at the
        // next position; the byte in front of it goes out as a literal
        const int nlen = __shfl_down_sync(0xffffffffu, mlen, 1);
        if (lane < 31 && mlen && nlen > mlen) mlen = 0;
      }
      // greedy parse of the 32 positions: orbit of `carry` under next()
      const int nxt = lane + (mlen ? mlen : 1);
      int jump = nxt > 32 ? 32 : nxt;
      uint32_t marked = 1u << carry;
#pragma unroll
      for (int r = 0; r < 5; r++) {
        const bool me = (marked >> lane) & 1u;
        marked |= __reduce_or_sync(0xffffffffu, (me && jump < 32) ? (1u << jump) : 0u);
        const int j2 = __shfl_sync(0xffffffffu, jump, jump & 31);
        if (jump < 32) jump = j2;
      }
      // positions beyond the input do not exist
      if (base + 32 > len) marked &= (len - base) >= 32 ? 0xffffffffu : ((1u << (len - base)) - 1u);
      const bool start = (marked >> lane) & 1u;
      // the token that leaves the step tells where the next one starts
      const uint32_t leaving = __ballot_sync(0xffffffffu, start && nxt >= 32);
      carry = leaving ? __shfl_sync(0xffffffffu, nxt, __ffs((int)leaving) - 1) - 32 : 0;
      if (start) {
        uint32_t t;
        if (mlen) {
          uint32_t eb, ev, eb2, ev2;
          const uint32_t ls = df_len_sym((uint32_t)mlen, &eb, &ev), ds = df_dist_sym((uint32_t)mdist, &eb2, &ev2);
          atomicAdd(&S.freq_l[ls], 1u);
          atomicAdd(&S.freq_d[ds], 1u);
          my_extra += eb + eb2;
          t = (1u << 23) | ((uint32_t)(mdist - 1) << 8) | (uint32_t)(mlen - 3);
        } else {
          t = d[p];
          atomicAdd(&S.freq_l[t], 1u);
        }
        tok[ntok + __popc(marked & ((1u << lane) - 1u))] = t;
      }
      ntok += __popc(marked);
    }
    my_extra = __reduce_add_sync(0xffffffffu, my_extra);
    if (lane == 0) { S.freq_l[256] = 1; S.extra_bits = my_extra; }
    __syncwarp();
    // ---------------------------------------------------------------- 2 build
    const uint32_t bits_dyn = df_build_block(S, !last, lane);
    // the same tokens under the Fixed codes (Def.block_of_frequencies weighs the two as well, de.ml:2411-2449): small inputs
    // do not pay for a tree in the header
    uint32_t bits_fixed = 0;
    for (int s2 = lane; s2 < 286; s2 += 32) bits_fixed += S.freq_l[s2] * (s2 < 144 ? 8u : s2 < 256 ? 9u : s2 < 280 ? 7u : 8u);
    if (lane < 30) bits_fixed += S.freq_d[lane] * 5u;
    bits_fixed = __reduce_add_sync(0xffffffffu, bits_fixed) + 3u + S.extra_bits + (last ? 0u : 3u + 7u + 32u);
    const bool fixed = bits_fixed < bits_dyn;
    const uint32_t bits_stored = 8u * ((uint32_t)len + 5u * (uint32_t)((len + 65534) / 65535));
    if ((fixed ? bits_fixed : bits_dyn) >= bits_stored) stored = true;
    if (!stored) {
      // ---------------------------------------------------------------- 3 pack
      const uint32_t nbytes = df_pack_block(S, tok, ntok, last, !last, out32, lane, fixed);
      if (lane == 0) { res->bytes = nbytes; res->sum_a = sum_a; res->sum_b = sum_b; }
      return;
    }
  }
  // ---------------------------------------------------------------- stored blocks (de.ml:2566-2581)
  {
    uint32_t o = 0;
    int pos = 0;
    do {
      const int n = min(65535, len - pos);
      const bool fin = last && pos + n == len;
      if (lane == 0) {
        slot[o] = fin ? 1 : 0;
        slot[o + 1] = (uint8_t)n; slot[o + 2] = (uint8_t)(n >> 8);
        slot[o + 3] = (uint8_t)~n; slot[o + 4] = (uint8_t)(~n >> 8);
      }
      for (int k = lane; k < n; k += 32) slot[o + 5 + k] = d[pos + k];
      o += 5 + (uint32_t)n;
      pos += n;
    } while (pos < len);
    if (lane == 0) { res->bytes = o; res->sum_a = sum_a; res->sum_b = sum_b; }
  }
}

// ---- De.Def.encode over a Queue, De.T.make: the command-level half of the reference's encoder on the device -----------
// One Dynamic block from a list of Queue commands (de.ml:2245-2266: literal = byte, end of block = 256, copy =
// 0x2000000 | (len - 3) << 16 | (off - 1)): frequencies counted as De.Lz77 counts them (succ_literal / succ_length /
// succ_distance, de.ml:2333-2351; the end of block starts at 1), Def.dynamic_of_frequencies, the header, the codes.
// The bytes are the reference's (test/test.ml:507-531).  One warp.
__global__ void __launch_bounds__(32)
deflate_queue_kernel(const int32_t *cmds, int64_t n, int last, uint32_t *tok, uint32_t *slot, uint8_t *dst, int64_t cap,
                     int64_t *written, int32_t *status) {
  DCMP_DYN_SMEM(smem_raw);
  typedef DfSmemT<DF_HASH_BITS_FAST> DfSmem;
  DfSmem &S = *reinterpret_cast<DfSmem *>(smem_raw);
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < 288; i += 32) S.freq_l[i] = 0;
  S.freq_d[lane] = 0;
  __syncwarp();
  uint32_t my_extra = 0;
  int ntok = 0;
  for (int64_t base = 0; base < n; base += 32) {
    const int64_t k = base + lane;
    const int32_t c = k < n ? cmds[k] : 256;
    const bool is_tok = k < n && c != 256;
    const uint32_t m = __ballot_sync(0xffffffffu, is_tok);
    if (is_tok) {
      uint32_t t;
      if (c & 0x2000000) {
        const uint32_t len = (uint32_t)((c >> 16) & 0x1ff) + 3u, dist = (uint32_t)(c & 0xffff) + 1u;
        uint32_t eb, ev, eb2, ev2;
        atomicAdd(&S.freq_l[df_len_sym(len, &eb, &ev)], 1u);
        atomicAdd(&S.freq_d[df_dist_sym(dist, &eb2, &ev2)], 1u);
        my_extra += eb + eb2;
        t = (1u << 23) | ((dist - 1u) << 8) | (len - 3u);
      } else {
        t = (uint32_t)c & 0xffu;
        atomicAdd(&S.freq_l[t], 1u);
      }
      tok[ntok + __popc(m & ((1u << lane) - 1u))] = t;
    }
    ntok += __popc(m);
  }
  my_extra = __reduce_add_sync(0xffffffffu, my_extra);
  if (lane == 0) { S.freq_l[256] = 1; S.extra_bits = my_extra; }
  __syncwarp();
  df_build_block(S, false, lane);
  const uint32_t nbytes = df_pack_block(S, tok, ntok, last != 0, false, slot, lane);
  __syncwarp();
  if ((int64_t)nbytes > cap) {
    if (lane == 0) { *written = 0; *status = DCMP_UNEXPECTED_END_OF_OUTPUT; }
    return;
  }
  const uint8_t *sb = reinterpret_cast<const uint8_t *>(slot);
  for (uint32_t i = lane; i < nbytes; i += 32) dst[i] = sb[i];
  if (lane == 0) { *written = nbytes; *status = DCMP_OK; }
}

// De.T.make ~length ?max_length freqs (de.ml:2013-2068) for one alphabet: lengths and bit-reversed codes per symbol, max_code.
__global__ void __launch_bounds__(32)
deflate_tmake_kernel(const int32_t *freqs, int length, int max_bits, int32_t *lens_out, int32_t *codes_out, int32_t *max_code_out) {
  DCMP_DYN_SMEM(smem_raw);
  typedef DfSmemT<DF_HASH_BITS_FAST> DfSmem;
  DfSmem &S = *reinterpret_cast<DfSmem *>(smem_raw);
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < 288; i += 32) S.freq_l[i] = i < length ? (uint32_t)freqs[i] : 0u;
  __syncwarp();
  const int mc = df_build_code(S, S.freq_l, length, max_bits, S.len_l, S.code_l, lane);
  for (int i = lane; i < length; i += 32) { lens_out[i] = S.len_l[i]; codes_out[i] = S.code_l[i]; }
  if (lane == 0) *max_code_out = mc;
}

template <int HB>
__global__ void __launch_bounds__(DF_WARPS * 32)
deflate_chunk_kernel(int level, int want_adler, int64_t nchunks, const uint8_t *src_base, const DfChunk *chunks,
                     DfChunkOut *outs, uint8_t *slots, uint32_t *tok_scratch, unsigned long long *work_counter) {
  DCMP_DYN_SMEM(smem_raw);
  typedef DfSmemT<HB> DfSmem;
  DfSmem *SS = reinterpret_cast<DfSmem *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DfSmem &S = SS[warp];
  uint32_t *tok = tok_scratch + ((size_t)blockIdx.x * DF_WARPS + warp) * DF_CHUNK;
  for (;;) {
    unsigned long long c = 0;
    if (lane == 0) c = atomicAdd(work_counter, 1ull);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= (unsigned long long)nchunks) break;
    const DfChunk ch = chunks[c];
    df_compress_chunk(S, src_base + ch.src_off, ch.len, HB == DF_HASH_BITS_BEST ? ch.hist : 0, (ch.flags & 1) != 0, level, tok, slots + (size_t)c * DF_SLOT,
                      &outs[c], want_adler != 0, lane);
    __syncwarp();
  }
}

// One warp per stream: where every chunk goes (exclusive scan of the chunk sizes), the Adler-32 of the whole
// input from the per-chunk partial sums, the framing, the result.
// first_chunk[k] .. first_chunk[k+1] are stream k's chunks.
__global__ void deflate_plan_kernel(int fmt, int level, int64_t nstreams, const int64_t *first_chunk, const DfChunkOut *outs,
                                    const DfChunk *chunks, uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                                    int64_t *chunk_dst, int64_t *written, int32_t *status) {
  const int lane = threadIdx.x & 31;
  const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= nstreams) return;
  const int64_t c0 = first_chunk[k], c1 = first_chunk[k + 1];
  const int64_t cap = dst_cap[k];
  int64_t total = fmt == DCMP_FMT_ZLIB ? 2 : 0;
  const unsigned long long M = 65521ull;
  unsigned long long A = 1, B = 0; // Adler-32 state after the chunks seen so far
  for (int64_t base = c0; base < c1; base += 32) {
    const int64_t c = base + lane;
    const bool have = c < c1;
    // a chunk as an Adler segment (a, b, n): appended to a state (A, B) it gives (A + a, B + n*A + b)
    unsigned long long a = 0, bq = 0, n = 0;
    int64_t bytes = 0;
    if (have) {
      bytes = outs[c].bytes;
      a = outs[c].sum_a % M;
      bq = outs[c].sum_b % M;
      n = (unsigned long long)chunks[c].len % M;
    }
    int64_t incl = bytes;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (have) chunk_dst[c] = total + incl - bytes;
    total += __shfl_sync(0xffffffffu, incl, 31);
    // ordered reduction of the 32 segments: (a1, b1, n1) then (a2, b2, n2) = (a1 + a2, b1 + b2 + n2*a1, n1 + n2)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long a2 = __shfl_down_sync(0xffffffffu, a, o), b2 = __shfl_down_sync(0xffffffffu, bq, o),
                               n2 = __shfl_down_sync(0xffffffffu, n, o);
      if (lane + o < 32) {
        bq = (bq + b2 + (n2 * a) % M) % M;
        a = (a + a2) % M;
        n = (n + n2) % M;
      }
    }
    a = __shfl_sync(0xffffffffu, a, 0);
    bq = __shfl_sync(0xffffffffu, bq, 0);
    n = __shfl_sync(0xffffffffu, n, 0);
    B = (B + (n * A) % M + bq) % M;
    A = (A + a) % M;
  }
  if (fmt == DCMP_FMT_ZLIB) total += 4;
  // De.Def.Ns.deflate needs _end_padding = 8 bytes of dst (de.ml:3998-4001) and reports a dst that is too
  // small as Unexpected_end_of_output
  if (total > cap) {
    for (int64_t c = c0 + lane; c < c1; c += 32) chunk_dst[c] = -1;
    if (lane == 0) {
      written[k] = 0;
      status[k] = DCMP_UNEXPECTED_END_OF_OUTPUT;
    }
    return;
  }
  if (lane != 0) return;
  uint8_t *o = dst_base + dst_off[k];
  if (fmt == DCMP_FMT_ZLIB) {
    // Zl.Def.Ns.header (lib/zl.ml:600-608)
    uint32_t hdr = (8u + (7u << 4)) << 8;
    const uint32_t fl = level <= 1 ? 0u : level <= 5 ? 1u : level == 6 ? 2u : 3u;
    hdr |= fl << 6;
    hdr += 31u - (hdr % 31u);
    o[0] = (uint8_t)(hdr >> 8);
    o[1] = (uint8_t)hdr;
    const uint32_t ad = (uint32_t)((B << 16) | A);
    o[total - 4] = (uint8_t)(ad >> 24); o[total - 3] = (uint8_t)(ad >> 16); o[total - 2] = (uint8_t)(ad >> 8); o[total - 1] = (uint8_t)ad;
  }
  written[k] = total;
  status[k] = DCMP_OK;
}

// One CTA per chunk: copy the chunk's slot to its place in the stream's output.
__global__ void __launch_bounds__(256)
deflate_gather_kernel(const DfChunk *chunks, const DfChunkOut *outs, const uint8_t *slots, const int64_t *chunk_dst,
                      uint8_t *dst_base, const int64_t *dst_off) {
  const int64_t c = blockIdx.x;
  const int64_t at = chunk_dst[c];
  if (at < 0) return;
  const uint8_t *s = slots + (size_t)c * DF_SLOT;
  uint8_t *o = dst_base + dst_off[chunks[c].stream] + at;
  const uint32_t n = outs[c].bytes;
  // head to 16-byte alignment of the destination, 16-byte body, tail
  uint32_t headn = (uint32_t)((16 - ((uintptr_t)o & 15)) & 15);
  if (headn > n) headn = n;
  for (uint32_t i = threadIdx.x; i < headn; i += blockDim.x) o[i] = s[i];
  const uint32_t body = (n - headn) & ~15u;
  for (uint32_t i = threadIdx.x * 16; i < body; i += blockDim.x * 16) {
    uint8_t tmp[16];
#pragma unroll
    for (int b = 0; b < 16; b++) tmp[b] = s[headn + i + b];
    *reinterpret_cast<uint4 *>(o + headn + i) = *reinterpret_cast<uint4 *>(tmp);
  }
  for (uint32_t i = headn + body + threadIdx.x; i < n; i += blockDim.x) o[i] = s[i];
}

inline int deflate_hash_bits(int level) { return level >= 7 ? DF_HASH_BITS_BEST : DF_HASH_BITS_FAST; }
size_t deflate_smem_bytes(int level) {
  return (deflate_hash_bits(level) == DF_HASH_BITS_BEST ? sizeof(DfSmemT<DF_HASH_BITS_BEST>) : sizeof(DfSmemT<DF_HASH_BITS_FAST>)) * DF_WARPS + 128;
}
size_t deflate_tok_bytes_per_cta() { return (size_t)DF_WARPS * DF_CHUNK * sizeof(uint32_t); }
#ifndef DCMP_EMU
// resident CTAs of the chunk kernel per SM at its smallest footprint (the persistent grid and its scratch are sized with it)
int deflate_ctas_per_sm() {
  int n = 0;
  const size_t smem = deflate_smem_bytes(1);
  cudaFuncSetAttribute(deflate_chunk_kernel<DF_HASH_BITS_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, deflate_chunk_kernel<DF_HASH_BITS_FAST>, DF_WARPS * 32, smem) != cudaSuccess || n < 1) n = 2;
  if (getenv("DCMP_DEF_CTAS")) n = atoi(getenv("DCMP_DEF_CTAS")) > 0 ? atoi(getenv("DCMP_DEF_CTAS")) : n; // tuning knob
  return n;
}
#endif
size_t deflate_chunk_desc_bytes() { return sizeof(DfChunk); }
size_t deflate_chunk_out_bytes() { return sizeof(DfChunkOut); }
size_t deflate_slot_bytes() { return DF_SLOT; }
int deflate_chunk_size() { return DF_CHUNK; }

#ifndef DCMP_EMU
int configure_deflate() { // per device (see configure_inflate_parallel)
  cudaError_t e = cudaFuncSetAttribute(deflate_chunk_kernel<DF_HASH_BITS_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)deflate_smem_bytes(1));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(deflate_chunk_kernel<DF_HASH_BITS_BEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)deflate_smem_bytes(9));
  return (int)e;
}
void launch_deflate_queue(const int32_t *cmds, int64_t n, int last, uint32_t *tok, uint32_t *slot, uint8_t *dst, int64_t cap, int64_t *written,
                          int32_t *status, cudaStream_t stream) {
  deflate_queue_kernel<<<1, 32, sizeof(DfSmemT<DF_HASH_BITS_FAST>) + 128, stream>>>(cmds, n, last, tok, slot, dst, cap, written, status);
}
void launch_deflate_tmake(const int32_t *freqs, int length, int max_bits, int32_t *lens, int32_t *codes, int32_t *max_code, cudaStream_t stream) {
  deflate_tmake_kernel<<<1, 32, sizeof(DfSmemT<DF_HASH_BITS_FAST>) + 128, stream>>>(freqs, length, max_bits, lens, codes, max_code);
}
int launch_deflate(int level, int fmt, int64_t nstreams, int64_t nchunks, const uint8_t *src_base, const void *chunks,
                   const int64_t *first_chunk, void *outs, uint8_t *slots, int64_t *chunk_dst, uint32_t *tok_scratch,
                   unsigned long long *work_counter, int ctas, uint8_t *dst_base, const int64_t *dst_off,
                   const int64_t *dst_cap, int64_t *written, int32_t *status, cudaStream_t stream) {
  const size_t smem = deflate_smem_bytes(level);
  if (nchunks > 0) {
    int64_t want = (nchunks + DF_WARPS - 1) / DF_WARPS;
    if (ctas > want) ctas = (int)want;
    if (deflate_hash_bits(level) == DF_HASH_BITS_BEST)
      deflate_chunk_kernel<DF_HASH_BITS_BEST><<<(unsigned)ctas, DF_WARPS * 32, smem, stream>>>(
          level, fmt == DCMP_FMT_ZLIB ? 1 : 0, nchunks, src_base, (const DfChunk *)chunks, (DfChunkOut *)outs, slots, tok_scratch, work_counter);
    else
      deflate_chunk_kernel<DF_HASH_BITS_FAST><<<(unsigned)ctas, DF_WARPS * 32, smem, stream>>>(
          level, fmt == DCMP_FMT_ZLIB ? 1 : 0, nchunks, src_base, (const DfChunk *)chunks, (DfChunkOut *)outs, slots, tok_scratch, work_counter);
  }
  deflate_plan_kernel<<<(unsigned)((nstreams + 3) / 4), 128, 0, stream>>>(fmt, level, nstreams, first_chunk, (const DfChunkOut *)outs,
                                                                              (const DfChunk *)chunks, dst_base, dst_off, dst_cap,
                                                                              chunk_dst, written, status);
  if (nchunks > 0)
    deflate_gather_kernel<<<(unsigned)nchunks, 256, 0, stream>>>((const DfChunk *)chunks, (const DfChunkOut *)outs, slots, chunk_dst,
                                                                 dst_base, dst_off);
  return 0;
}
#endif

} // namespace dcmp
