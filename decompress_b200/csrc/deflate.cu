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
// Layout.  Every input stream is cut into 128 KiB chunks, one DEFLATE block each (Dynamic, Fixed or Stored, whichever is
// smallest).  Two kernels share a batch:
//   deflate_cta_kernel    one CTA of 512 (1024) threads per run of consecutive chunks: input staged in shared memory by TMA,
//                         the 32 KiB window and its hash buckets carried from chunk to chunk, per-level chain / nice /
//                         lazy (de.ml:4030-4049), block-wide histogram, parallel Huffman build, all warps packing
//                         bits at scanned offsets — described where it is defined, further down;
//   deflate_chunk_kernel  one warp per chunk, for the small single-chunk streams of a batch (<= DF_SMALL bytes: a CTA
//                         per few hundred bytes would leave the GPU idle), and the round-1 path for A/B measurements:
//   1 tokenise   32 positions per step: every lane hashes its 4 bytes, takes the two most recent
//                earlier positions with the same hash (two-way buckets in a shared-memory table,
//                __match_any_sync for the ones inside the step) and extends both matches eight
//                bytes at a time (level 1: one candidate); from level 4 on a match yields to a
//                longer one starting at the next position (one step of lazy evaluation); the
//                parse of the 32 positions is the orbit of the carry-in position under
//                next(p) = p + (match ? len : 1), found by pointer doubling with shuffles;
//                tokens go to an L2-resident scratch, symbol frequencies to shared memory;
//   2 build      De.T.make exactly (zlib's heap, its tie-breaks, gen_bitlen's overflow repair), T.scan / T.symbols;
//   3 pack       32 tokens per step: bit widths -> exclusive scan -> atomicOr into a
//                shared-memory bit window that is drained to HBM with coalesced 32-bit stores.
// A chunk that is not the last of its stream ends with an empty stored block, which byte-aligns
// it (the classic sync-flush marker 00 00 ff ff), so chunk outputs concatenate; a second kernel
// gathers them to their final place, adds the zlib header / Adler-32 trailer if asked to.
#include <type_traits>
#include "simt.cuh"
#include "tma.cuh"
#include "../../include/dcmp.h"

namespace dcmp {

constexpr int DF_CHUNK = 131072;          // input bytes per DEFLATE block (four 32 KiB halves of the ring)
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
constexpr int DF_SMALL = 4096;           // single-chunk streams up to this size stay with the warp-per-chunk kernel

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
// De.T.make's state (de.ml:2013-2026); only alive after tokenising, so it overlays the hash table.  A heap slot carries
// its ordering key — frequency and subtree depth, what zlib's `smaller` compares (de.ml:1879-1885) — next to the node,
// so a step down the heap is one round of loads, not a chain through freq[heap[j]] and depth[heap[j]].
// KEY: 32 bits hold frequencies below 2^23 (a 128 KiB chunk has at most 2^17 tokens), the command-level entry points use 64.
template <class KEY>
struct DfTreeT {
  static constexpr int SHIFT = 9;   // depth < 2^9 (at most 285)
  KEY key[DF_HEAP + 3];             // heap slots 1..heap_len: (frequency << SHIFT) | depth
  uint16_t node[DF_HEAP + 3];       // the node in a heap slot; from heap_max up: the nodes in the order they left the heap
  uint16_t dad[DF_HEAP + 3];
  uint16_t depth[DF_HEAP + 3];      // the bit length of every node (generate_lengths)
};
template <int HB, class KEY = uint32_t>
struct DfSmemT {
  static constexpr int HASH_BITS = HB;
  typedef DfTreeT<KEY> DfTree;
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
static_assert(sizeof(DfTreeT<uint32_t>) + 2 * (288 + 32 + 19) + (288 + 32 + 19) + 640 + 16 <= sizeof(uint32_t) * (1 << DF_HASH_BITS_FAST), "what follows tokenising overlays the hash table");
static_assert(sizeof(DfTreeT<uint64_t>) + 2 * (288 + 32 + 19) + (288 + 32 + 19) + 640 + 16 <= sizeof(uint32_t) * (1 << DF_HASH_BITS_BEST), "... of the command-level kernels");

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
  auto &T = S.tree;
  typedef typename std::remove_reference<decltype(T.key[0])>::type KEY;
  constexpr int SH = std::remove_reference<decltype(T)>::type::SHIFT;
  __syncwarp();
  for (int i = lane; i < n; i += 32) { len[i] = 0; code[i] = 0; T.depth[i] = 0; } // (a symbol that is not used keeps length 0, de.ml:1849-1853)
  __syncwarp();
  int max_code = -1;
  if (lane == 0) {
    int heap_len = 0, heap_max = DF_HEAP;
    // smaller(a, b), de.ml:1879-1885: freq[a] < freq[b] || (freq[a] = freq[b] && depth[a] <= depth[b])  <=>  key(a) <= key(b)
    auto down = [&](int k) { // pqdownheap, de.ml:1887-1898
      const KEY vk = T.key[k];
      const uint16_t vn = T.node[k];
      int j = k << 1;
      while (j <= heap_len) {
        KEY kj = T.key[j];
        if (j < heap_len) { const KEY kr = T.key[j + 1]; if (kr <= kj) { j++; kj = kr; } }
        if (vk <= kj) break;
        T.key[k] = kj;
        T.node[k] = T.node[j];
        k = j;
        j <<= 1;
      }
      T.key[k] = vk;
      T.node[k] = vn;
    };
    for (int s = 0; s < n; s++) {
      const uint32_t f = freq[s];
      if (f != 0) { ++heap_len; T.key[heap_len] = (KEY)f << SH; T.node[heap_len] = (uint16_t)s; max_code = s; }
    }
    while (heap_len < 2) { // the pkzip rule
      const int node = max_code < 2 ? ++max_code : 0;
      ++heap_len;
      T.key[heap_len] = (KEY)1 << SH;
      T.node[heap_len] = (uint16_t)node;
    }
    for (int k = heap_len / 2; k >= 1; k--) down(k);
    int node = n;
    do {
      const KEY ak = T.key[1];
      const int a = T.node[1];
      T.key[1] = T.key[heap_len];
      T.node[1] = T.node[heap_len];
      heap_len--;
      down(1);
      const KEY bk = T.key[1];
      const int b = T.node[1];
      T.node[--heap_max] = (uint16_t)a;
      T.node[--heap_max] = (uint16_t)b;
      const KEY da = ak & (((KEY)1 << SH) - 1), db = bk & (((KEY)1 << SH) - 1);
      T.dad[a] = (uint16_t)node;
      T.dad[b] = (uint16_t)node;
      T.key[1] = (((ak >> SH) + (bk >> SH)) << SH) | ((da >= db ? da : db) + 1);
      T.node[1] = (uint16_t)node++;
      down(1);
    } while (heap_len >= 2);
    T.node[--heap_max] = T.node[1];
    // generate_lengths (de.ml:1952-2009); depth[] is the bit length of every node
    int bl_count[16];
    for (int bq = 0; bq < 16; bq++) bl_count[bq] = 0;
    T.depth[T.node[heap_max]] = 0;
    int overflow = 0;
    for (int h = heap_max + 1; h < DF_HEAP; h++) {
      const int m = T.node[h];
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
          const int m = T.node[--h];
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

// The CTA-per-chunk compressor's builder: the same optimal, length-limited code (Huffman + zlib's overflow repair), built
// without the heap — sort, two-queue merge, depths by pointer jumping, canonical codes by ranks — so that one warp is done
// in a few microseconds instead of a few hundred.  What it does not keep is zlib's ORDER among equal frequencies (the
// heap's tie-breaks decide which of two equally frequent symbols gets the longer code): the block costs the same number
// of bits, its bytes are not the ones De.T.make would choose.  The command-level entry points (dcmp_t_make,
// dcmp_def_encode_queue) and the small-stream kernel use df_build_code, which is De.T.make to the bit.
struct DfFastTree {
  uint32_t sortbuf[512];           // (frequency << 9) | symbol, ascending; unused slots 0xffffffff
  uint32_t weight[2 * 286];        // two-queue merge: leaves [0, m), internal nodes [m, 2m - 1)
  uint32_t info[2][2 * 286];       // (distance to an ancestor << 16) | that ancestor: the parent at first, the root in the end
                                   // (two copies: a doubling round reads one and writes the other)
  uint32_t bl_count[16];
  uint32_t next_code[16];
  int m;
};
template <class SM>
__device__ int df_build_code_fast(SM &S, const uint32_t *freq, int n, int max_bits, uint8_t *len, uint16_t *code, int lane) {
  DfFastTree &T = S.ftree;
  __syncwarp();
  // keys of the used symbols, packed; the pkzip rule (de.ml:1862-1874): at least two codes
  int m = 0, max_code = -1;
  for (int base = 0; base < n; base += 32) {
    const int s2 = base + lane;
    const uint32_t f = s2 < n ? freq[s2] : 0u;
    if (s2 < n) { len[s2] = 0; code[s2] = 0; }
    const uint32_t used = __ballot_sync(0xffffffffu, f != 0u);
    if (f) T.sortbuf[m + __popc(used & ((1u << lane) - 1u))] = (f << 9) | (uint32_t)s2;
    m += __popc(used);
    if (used) max_code = base + 31 - __clz((int)used);
  }
  if (lane < 16) T.bl_count[lane] = 0;
  while (m < 2) {
    const int node = max_code < 2 ? ++max_code : 0;
    if (lane == 0) T.sortbuf[m] = (1u << 9) | (uint32_t)node;
    m++;
  }
  // bitonic sort of the keys, padded to a power of two
  int N = 32, lgN = 5;
  while (N < m) { N <<= 1; lgN++; }
  for (int i = m + lane; i < N; i += 32) T.sortbuf[i] = 0xffffffffu;
  __syncwarp();
  for (int lk = 1; lk <= lgN; lk++) {
    for (int lj = lk - 1; lj >= 0; lj--) {
      for (int t = lane; t < (N >> 1); t += 32) {
        const int i = ((t >> lj) << (lj + 1)) | (t & ((1 << lj) - 1));
        const int q = i | (1 << lj);
        const bool up = ((i >> lk) & 1) == 0;
        const uint32_t a = T.sortbuf[i], b = T.sortbuf[q];
        if ((a > b) == up) { T.sortbuf[i] = b; T.sortbuf[q] = a; }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < m; i += 32) T.weight[i] = T.sortbuf[i] >> 9;
  __syncwarp();
  if (lane == 0) {
    // two-queue Huffman: the leaves ascending, the internal nodes in the order they are made (ascending too)
    int li = 0, ii = m, nn = m;
    uint32_t wl = T.weight[0], wi = 0xffffffffu;
    while (nn < 2 * m - 1) {
      uint32_t w = 0;
#pragma unroll
      for (int k = 0; k < 2; k++) {
        if (li < m && (ii >= nn || wl <= wi)) { w += wl; T.info[0][li] = (1u << 16) | (uint32_t)nn; li++; wl = li < m ? T.weight[li] : 0xffffffffu; }
        else { w += wi; T.info[0][ii] = (1u << 16) | (uint32_t)nn; ii++; wi = ii < nn ? T.weight[ii] : 0xffffffffu; }
      }
      T.weight[nn] = w;
      if (ii == nn) wi = w; // (the queue of internal nodes was empty: the new one is its head)
      nn++;
    }
    T.info[0][2 * m - 2] = (uint32_t)(2 * m - 2); // the root: distance 0 to itself
  }
  __syncwarp();
  // depths by pointer jumping (a code over at most 2^17 + 1 tokens is at most 26 deep: six doublings reach 64)
  const int nodes = 2 * m - 1;
  for (int r = 0; r < 6; r++) {
    const uint32_t *src = T.info[r & 1];
    uint32_t *dst = T.info[(r + 1) & 1];
    for (int i = lane; i < nodes; i += 32) {
      const uint32_t a = src[i], b = src[a & 0xffffu];
      dst[i] = ((a >> 16) + (b >> 16)) << 16 | (b & 0xffffu);
    }
    __syncwarp();
  }
  for (int i = lane; i < m; i += 32) {
    uint32_t d = T.info[0][i] >> 16; // (six rounds: back in the first copy)
    if ((int)d > max_bits) d = (uint32_t)max_bits;
    atomicAdd(&T.bl_count[d], 1u);
  }
  __syncwarp();
  if (lane == 0) {
    // Codes cut to max_bits over-subscribe the code space; gen_bitlen's repair (de.ml:1978-1994) — a leaf of the deepest
    // level above max_bits moves one level down and takes one of the cut leaves as its brother — gives back one unit
    // of 2^-max_bits per round: as many rounds as the Kraft sum is above 1
    int ov = -(1 << max_bits);
    for (int bits = 1; bits <= max_bits; bits++) ov += (int)T.bl_count[bits] << (max_bits - bits);
    while (ov > 0) {
      int bits = max_bits - 1;
      while (T.bl_count[bits] == 0) bits--;
      T.bl_count[bits]--;
      T.bl_count[bits + 1] += 2;
      T.bl_count[max_bits]--;
      ov--;
    }
    uint32_t c = 0;
    T.next_code[0] = 0;
    for (int bits = 1; bits <= max_bits; bits++) { c = (c + T.bl_count[bits - 1]) << 1; T.next_code[bits] = c; }
  }
  __syncwarp();
  // the longest codes to the least frequent symbols
  for (int i = lane; i < m; i += 32) {
    int start = 0, bits = max_bits;
    while (bits > 1 && i >= start + (int)T.bl_count[bits]) { start += (int)T.bl_count[bits]; bits--; }
    len[T.sortbuf[i] & 511u] = (uint8_t)bits;
  }
  __syncwarp();
  // canonical codes, bit-reversed (de.ml:1926-1950): a symbol's rank among the symbols of its length
  for (int base = 0; base <= max_code; base += 32) {
    const int s2 = base + lane;
    const uint32_t l = s2 <= max_code ? len[s2] : 0u;
    const uint32_t same = __match_any_sync(0xffffffffu, l);
    if (l) {
      const uint32_t first = T.next_code[l];
      const uint32_t cd = first + (uint32_t)__popc(same & ((1u << lane) - 1u));
      code[s2] = (uint16_t)(__brev(cd) >> (32 - l));
    }
    __syncwarp();
    if (l && (same >> lane) == 1u) T.next_code[l] += (uint32_t)__popc(same); // (the group's highest lane)
    __syncwarp();
  }
  return max_code;
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
template <bool EXACT = true, class SM>
__device__ uint32_t df_build_block(SM &S, bool marker, int lane) {
  int l_max, d_max;
  if constexpr (EXACT) {
    l_max = df_build_code(S, S.freq_l, 286, 15, S.len_l, S.code_l, lane);
    d_max = df_build_code(S, S.freq_d, 30, 15, S.len_d, S.code_d, lane);
  } else {
    l_max = df_build_code_fast(S, S.freq_l, 286, 15, S.len_l, S.code_l, lane);
    d_max = df_build_code_fast(S, S.freq_d, 30, 15, S.len_d, S.code_d, lane);
  }
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
  if constexpr (EXACT) df_build_code(S, S.freq_bl, 19, 7, S.len_bl, S.code_bl, lane);
  else df_build_code_fast(S, S.freq_bl, 19, 7, S.len_bl, S.code_bl, lane);
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
  typedef DfSmemT<DF_HASH_BITS_BEST, uint64_t> DfSmem;
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
  typedef DfSmemT<DF_HASH_BITS_BEST, uint64_t> DfSmem;
  DfSmem &S = *reinterpret_cast<DfSmem *>(smem_raw);
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < 288; i += 32) S.freq_l[i] = i < length ? (uint32_t)freqs[i] : 0u;
  __syncwarp();
  const int mc = df_build_code(S, S.freq_l, length, max_bits, S.len_l, S.code_l, lane);
  for (int i = lane; i < length; i += 32) { lens_out[i] = S.len_l[i]; codes_out[i] = S.code_l[i]; }
  if (lane == 0) *max_code_out = mc;
}

#ifdef DCMP_EMU
// (tests: the CTA-per-chunk compressor's builder on its own)
__global__ void __launch_bounds__(32)
deflate_tmake_fast_kernel(const int32_t *freqs, int length, int max_bits, int32_t *lens_out, int32_t *codes_out, int32_t *max_code_out) {
  DCMP_DYN_SMEM(smem_raw);
  struct Sm { DfFastTree ftree; uint32_t freq[288]; uint8_t len[288]; uint16_t code[288]; };
  Sm &S = *reinterpret_cast<Sm *>(smem_raw);
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < 288; i += 32) S.freq[i] = i < length ? (uint32_t)freqs[i] : 0u;
  __syncwarp();
  const int mc = df_build_code_fast(S, S.freq, length, max_bits, S.len, S.code, lane);
  __syncwarp();
  for (int i = lane; i < length; i += 32) { lens_out[i] = S.len[i]; codes_out[i] = S.code[i]; }
  if (lane == 0) *max_code_out = mc;
}
#endif

template <int HB>
__global__ void __launch_bounds__(DF_WARPS * 32)
deflate_chunk_kernel(int level, int want_adler, int64_t nchunks, const uint8_t *src_base, const DfChunk *chunks,
                     DfChunkOut *outs, uint8_t *slots, uint32_t *tok_scratch, unsigned long long *work_counter, int small_only) {
  DCMP_DYN_SMEM(smem_raw);
  typedef DfSmemT<HB> DfSmem;
  DfSmem *SS = reinterpret_cast<DfSmem *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DfSmem &S = SS[warp];
  // (a warp that only takes the small single-chunk streams needs a token per input byte of those)
  uint32_t *tok = tok_scratch + ((size_t)blockIdx.x * DF_WARPS + warp) * (small_only ? DF_SMALL : DF_CHUNK);
  for (;;) {
    unsigned long long c = 0;
    if (lane == 0) c = atomicAdd(work_counter, 1ull);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= (unsigned long long)nchunks) break;
    const DfChunk ch = chunks[c];
    if (small_only && !(ch.len <= DF_SMALL && ch.hist == 0)) continue; // deflate_cta_kernel's
    df_compress_chunk(S, src_base + ch.src_off, ch.len, HB == DF_HASH_BITS_BEST ? ch.hist : 0, (ch.flags & 1) != 0, level, tok, slots + (size_t)c * DF_SLOT,
                      &outs[c], want_adler != 0, lane);
    __syncwarp();
  }
}

// =====================================================================================================================
// The CTA-per-chunk compressor (round 2): the kernel behind C3.  deflate_chunk_kernel above (a warp per chunk, the input
// read through L1/L2) stays for the small inputs of a batch — thousands of few-hundred-byte streams keep every warp of the
// GPU busy there — this one takes everything longer than DF_SMALL.
//
// One CTA of 512 threads (1024 with the 16-way table) owns a RUN of consecutive 128 KiB chunks of one stream and keeps De.Lz77's state on chip
// across them (de.ml:4076-4100: the window, its hash heads and chains):
//   ring   64 KiB of shared memory = the last 64 KiB of the input, indexed by (global address & 0xffff): the 32 KiB window
//          plus the 32 KiB half that is being tokenised (staged by TMA; a chunk is four halves);
//   table  2048 hash buckets of WAYS 16-bit ring positions, replaced round-robin: the `max_chain` most recent
//          positions with the same hash4 (de.ml:4067-4071).  What zlib keeps as head[] + prev[] chains; WAYS = 8 for
//          levels 1..6, 16 for 7..9, of which a level examines min(max_chain, WAYS) (level table de.ml:4030-4049).
// A step = as many consecutive positions as the CTA has threads, one per thread:
//   insert   every position enters its bucket (an atomic counter per bucket hands out the slot)          -- barrier
//   match    every position compares itself with the bucket's entries that lie before it (positions of the same
//            step included), eight bytes at a time out of the ring; longest wins, nearest on ties; the search stops
//            at `nice_length`; from level 4 on a match shorter than `max_lazy` yields to a longer one starting at
//            the next position (lazy evaluation, de.ml:4351-4410)
//   parse    which positions start a token is the orbit of the carry-in under p -> p + len(p): every warp computes
//            the transfer function of its 32 positions for ALL 32 entry points by pointer doubling (exit and token
//            count), one thread chains the sixteen of them (16 dependent shared-memory loads), every warp then marks
//            its orbit and stores its tokens at the offsets the chain gave it                           -- 2 barriers
// After the two halves: the block's three Huffman codes (df_build_block: De.T.make), Dynamic / Fixed / Stored by cost,
// and the bits, packed by all sixteen warps at once — per-warp bit totals, a scan, every warp into its own bit window
// that drains to the chunk's slot with plain 32-bit stores (an atomicOr for the two words a warp shares with its
// neighbours).  Matches do not cross a half (the bytes behind it are not in the ring yet): one match per 32 KiB is cut.
// threads of a CTA: 512 with the 8-way table (two CTAs per SM), 1024 with the 16-way table (its 146 KiB of shared memory
// leave room for one CTA per SM only: twice the warps keep the SM as busy)
__host__ __device__ constexpr int dc_threads(int ways) { return ways == 16 ? 1024 : 512; }
constexpr int DC_HB = 11;
constexpr int DC_HALF = 32768;
constexpr int DC_RING = 65536;
constexpr int DC_MIRROR = 288;           // >= 258 + 8 + 16: the longest match plus the footprint of the word loads
constexpr int DC_WIN = 64;               // words of a warp's output bit window
constexpr int DC_RUN = 2;                // consecutive chunks a CTA claims at once (the window carries over inside a run)
#ifndef DCMP_DC_EXACT_TREE
#define DCMP_DC_EXACT_TREE 0             // 1: De.T.make's heap in this kernel too (its tie-breaks), ~15 % slower on C3
#endif
constexpr uint32_t DC_EXIT_END = 1000;   // transfer function: ran into the end of the half

template <int WAYS>
struct DcSmemT {
  static constexpr int HASH_BITS = DC_HB;
  static constexpr int DC_NW = dc_threads(WAYS) / 32;
  __align__(16) uint8_t ring[DC_RING + DC_MIRROR]; // + a copy of its first bytes: a read that starts inside the ring never wraps
  __align__(16) uint16_t tab[(1 << DC_HB) * WAYS];
  uint32_t cnt[(1 << DC_HB) / 4];          // four 8-bit insertion counters per word
  union {
    DfTreeT<uint32_t> tree;                // only while the codes are built
    DfFastTree ftree;
    struct {
      uint32_t win_w[DC_NW][DC_WIN];       // pack: a bit window per warp
      uint32_t xfer[DC_NW][32];            // parse: exit | tokens << 16 per entry lane
    };
  };
  uint16_t code_l[288];
  uint16_t code_d[32];
  uint16_t code_bl[19];
  uint8_t len_l[288];
  uint8_t len_d[32];
  uint8_t len_bl[19];
  uint8_t rle_sym[320];
  uint8_t rle_extra[320];
  uint32_t freq_l[288];
  uint32_t freq_d[32];
  uint32_t freq_bl[19];
  uint32_t extra_bits;
  int n_rle;
  int hlit, hdist, hclen;
  int32_t carryq[2];                       // first position of a step that is not inside a token of the step before, and the
  uint32_t ntokq[2];                       // tokens so far; by step parity: the last warp writes the next step's while others read
  uint32_t warp_bits[DC_NW + 1];
  uint32_t hdr_bits, blk_kind, blk_bytes;
  unsigned long long sum_a, sum_b;
  long long claim;
  __align__(8) uint64_t mbar;              // completion of the bulk copies into the ring
};

// four / eight ring bytes at byte index p < DC_RING (any alignment; the mirror behind the ring takes the wrap)
__device__ __forceinline__ uint32_t rg32(const uint32_t *r32, uint32_t p) {
  const uint32_t *w = r32 + (p >> 2);
  return __funnelshift_r(w[0], w[1], p << 3);
}
__device__ __forceinline__ unsigned long long rg64(const uint32_t *r32, uint32_t p) {
  const uint32_t *w = r32 + (p >> 2);
  const uint32_t a = w[0], b = w[1], c = w[2];
  return (unsigned long long)__funnelshift_r(a, b, p << 3) | ((unsigned long long)__funnelshift_r(b, c, p << 3) << 32);
}

// Input bytes [g0, g1) of a stream into the ring, at (address & 0xffff): TMA bulk copies of the whole 16-byte granules
// inside the range (at most two copies: the ring wraps once), completion on the mbarrier; the ragged bytes at the two
// ends go one by one — EXACTLY the range is written: the ring bytes behind it are the far end of the window (distance
// 32768 from the first positions of the half), not padding.  All threads call; `phase` is the mbarrier's parity, kept by
// every thread.  Callers have a barrier behind the last reads of the bytes that are replaced.
template <class SM>
__device__ __forceinline__ void dc_load(SM &S, const uint8_t *g0, const uint8_t *g1, int tid, uint32_t &phase) {
  if (g1 <= g0) return;
  const uintptr_t a0 = (uintptr_t)g0, a1 = (uintptr_t)g1;
  uintptr_t ga = (a0 + 15) & ~(uintptr_t)15, gz = a1 & ~(uintptr_t)15; // whole granules: [ga, gz)
  if (gz < ga) { ga = a1; gz = a1; }                                    // (the range lies inside one granule)
  if (ga > a1) ga = a1;
  if ((uintptr_t)tid < ga - a0) S.ring[(uint32_t)(a0 + (uintptr_t)tid) & 0xffffu] = g0[tid];
  if (gz >= ga && (uintptr_t)tid < a1 - gz) S.ring[(uint32_t)(gz + (uintptr_t)tid) & 0xffffu] = *reinterpret_cast<const uint8_t *>(gz + (uintptr_t)tid);
  if (gz > ga) {
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(gz - ga), r0 = (uint32_t)ga & 0xffffu;
      const uint32_t first = bytes < DC_RING - r0 ? bytes : DC_RING - r0;
      mbar_expect_tx(&S.mbar, bytes);
      tma_load_1d(S.ring + r0, reinterpret_cast<const void *>(ga), first, &S.mbar);
      if (bytes > first) tma_load_1d(S.ring, reinterpret_cast<const void *>(ga + first), bytes - first, &S.mbar);
    }
    mbar_wait(&S.mbar, phase & 1u);
    phase ^= 1u;
  }
}

// A warp's share of the block's bits: its own window, drained to the slot; the first and the last word of the share
// may be shared with the neighbours (atomicOr; zeroed beforehand), the words in between are its own.
struct DcBits {
  uint32_t *win;       // DC_WIN words
  uint32_t *out;       // the slot, word aligned
  uint32_t bitpos;     // next bit (block coordinates)
  uint32_t flushed;    // words below this index are out
  uint32_t first_word, last_word;
  __device__ __forceinline__ void put(uint64_t v, uint32_t n, uint32_t at) {
    if (n == 0) return;
    const uint32_t w = at >> 5, sh = at & 31;
    const uint64_t lo = v << sh;
    atomicOr(&win[w & (DC_WIN - 1)], (uint32_t)lo);
    if (sh + n > 32) atomicOr(&win[(w + 1) & (DC_WIN - 1)], (uint32_t)(lo >> 32));
    if (sh + n > 64) atomicOr(&win[(w + 2) & (DC_WIN - 1)], (uint32_t)(v >> (64 - sh)));
  }
  // one thread on its own (the header): the complete words below bit `at` leave with plain stores
  __device__ __forceinline__ void drain_serial(uint32_t at) {
    for (; flushed < (at >> 5); flushed++) {
      out[flushed] = win[flushed & (DC_WIN - 1)];
      win[flushed & (DC_WIN - 1)] = 0;
    }
  }
  __device__ __forceinline__ void drain(int lane, bool all) {
    __syncwarp();
    const uint32_t upto = all ? (bitpos + 31) >> 5 : bitpos >> 5;
    for (uint32_t i = flushed + lane; i < upto; i += 32) {
      const uint32_t v = win[i & (DC_WIN - 1)];
      if (i == first_word || i >= last_word) { if (v) atomicOr(&out[i], v); }
      else out[i] = v;
      win[i & (DC_WIN - 1)] = 0;
    }
    flushed = upto;
    __syncwarp();
  }
};

struct DcLevel { int ways, nice, lazy; };
__device__ __forceinline__ DcLevel dc_level(int level, int WAYS) {
  // {good_length; max_lazy; nice_length; max_chain} of de.ml:4030-4049; levels 1..3 are zlib's deflate_fast (no lazy
  // evaluation); good_length shortens the chain of the LAZY search after a good match: there is no "previous match"
  // when every position is searched at once
  DcLevel c;
  const int chain = level <= 1 ? 4 : level == 2 ? 8 : level == 3 ? 32 : level == 4 ? 16 : level == 5 ? 32 : level == 6 ? 128 : level == 7 ? 256 : level == 8 ? 1024 : 4096;
  c.ways = chain < WAYS ? chain : WAYS;
  c.nice = level <= 1 ? 8 : level == 2 ? 16 : level == 3 ? 32 : level == 4 ? 16 : level == 5 ? 32 : level <= 7 ? 128 : 258;
  c.lazy = level <= 3 ? 0 : level == 4 ? 4 : level <= 6 ? 16 : level == 7 ? 32 : level == 8 ? 128 : 258;
  return c;
}

template <int WAYS>
__global__ void __launch_bounds__(dc_threads(WAYS), WAYS == 8 ? 2 : 1)
deflate_cta_kernel(int level, int want_adler, int64_t nchunks, const uint8_t *src_base, const DfChunk *chunks,
                   DfChunkOut *outs, uint8_t *slots, uint32_t *tok_scratch, unsigned long long *work_counter, int tune_ways) {
  DCMP_DYN_SMEM(smem_raw);
  typedef DcSmemT<WAYS> Smem;
  constexpr int DC_T = dc_threads(WAYS), DC_NW = DC_T / 32;
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *tok = tok_scratch + (size_t)blockIdx.x * DF_CHUNK;
  const uint32_t *r32 = reinterpret_cast<const uint32_t *>(S.ring);
  uint32_t mphase = 0; // parity of the ring's mbarrier (CTA-uniform)
  if (tid == 0) { mbar_init(&S.mbar, 1); fence_mbar_init(); }
  DcLevel L = dc_level(level, WAYS);
  if (tune_ways > 0 && tune_ways <= WAYS) L.ways = tune_ways; // (DCMP_DC_WAYS: measurements)
  for (;;) {
    __syncthreads();
    if (tid == 0) S.claim = (long long)atomicAdd(work_counter, (unsigned long long)DC_RUN);
    __syncthreads();
    const int64_t c0 = S.claim;
    if (c0 >= nchunks) break;
    bool carried = false; // ring and table hold the 32 KiB in front of the next chunk
    for (int64_t c = c0; c < c0 + DC_RUN && c < nchunks; c++) {
      const DfChunk ch = chunks[c];
      if (ch.len <= DF_SMALL && ch.hist == 0) { carried = false; continue; } // the warp-per-chunk kernel's
      const uint8_t *d = src_base + ch.src_off;
      const int len = ch.len, hist = ch.hist;
      const bool last = (ch.flags & 1) != 0;
      const uint32_t rb = (uint32_t)(uintptr_t)d & 0xffffu; // ring index of the chunk's first byte
      uint8_t *slot = slots + (size_t)c * DF_SLOT;
      uint32_t *out32 = reinterpret_cast<uint32_t *>(slot);
      __syncthreads();
      uint32_t stepno = 0; // (CTA-uniform)
      if (tid == 0) { S.carryq[0] = 0; S.ntokq[0] = 0; S.sum_a = 0; S.sum_b = 0; S.extra_bits = 0; }
      for (int i = tid; i < 288; i += DC_T) S.freq_l[i] = 0;
      if (tid < 32) S.freq_d[tid] = 0;
      bool stored = level == 0;
      uint32_t my_extra = 0;
      if (!stored) {
        if (!carried || hist == 0) {
          // a new window: empty table, and the (up to) 32 KiB in front of the chunk entered without producing tokens
          for (int i = tid; i < (1 << DC_HB) * WAYS / 8; i += DC_T) reinterpret_cast<uint4 *>(S.tab)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
          for (int i = tid; i < (1 << DC_HB) / 4; i += DC_T) S.cnt[i] = 0;
          if (hist > 0) {
            dc_load(S, d - hist, d, tid, mphase);
            __syncthreads();
            if (tid < DC_MIRROR / 16) reinterpret_cast<uint4 *>(S.ring + DC_RING)[tid] = reinterpret_cast<const uint4 *>(S.ring)[tid];
          }
          __syncthreads();
          // (the last three window positions hash bytes of the chunk, which are not loaded yet: left out)
          for (int j = -hist + tid; j < -3; j += DC_T) {
            const uint32_t p = (rb + (uint32_t)j) & 0xffffu;
            const uint32_t h = (rg32(r32, p) * 0x9e3779b1u) >> (32 - DC_HB);
            const uint32_t sh = 8u * (h & 3u);
            const uint32_t slot_i = (atomicAdd(&S.cnt[h >> 2], 1u << sh) >> sh) & (uint32_t)(WAYS - 1);
            st_relaxed_u16(&S.tab[h * WAYS + slot_i], (uint16_t)p);
          }
        }
        // ------------------------------------------------------------ 1 tokenise, half by half
        for (int h0 = 0; h0 < len; h0 += DC_HALF) {
          const int hend = min(len, h0 + DC_HALF);
          __syncthreads(); // everybody is done with the bytes this half replaces
          dc_load(S, d + h0, d + hend, tid, mphase);
          __syncthreads();
          if (tid < DC_MIRROR / 16) reinterpret_cast<uint4 *>(S.ring + DC_RING)[tid] = reinterpret_cast<const uint4 *>(S.ring)[tid];
          __syncthreads();
          if (want_adler) {
            // Adler-32 partials of the chunk (sum of bytes, sum of (len - j) * byte_j), from the ring
            unsigned long long a = 0, b = 0;
            for (int j = h0 + tid * 4; j < hend; j += DC_T * 4) {
              const uint32_t v = rg32(r32, (rb + (uint32_t)j) & 0xffffu);
#pragma unroll
              for (int k = 0; k < 4; k++)
                if (j + k < hend) { const uint32_t x = (v >> (8 * k)) & 0xffu; a += x; b += (unsigned long long)x * (uint32_t)(len - j - k); }
            }
            for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            if (lane == 0) { atomicAdd(&S.sum_a, a); atomicAdd(&S.sum_b, b % 65521ull); }
          }
          if (tid == 0) S.carryq[stepno & 1] = 0; // (a match never crosses a half)
          for (int base = h0; base < hend; base += DC_T) {
            const int j = base + tid;
            const uint32_t p = (rb + (uint32_t)j) & 0xffffu;
            const int avail = hend - j;              // bytes of this half from j on
            const bool can = avail >= DF_MIN_MATCH;
            // -- insert
            uint32_t hsh = 0;
            if (can) {
              hsh = (rg32(r32, p) * 0x9e3779b1u) >> (32 - DC_HB);
              const uint32_t sh = 8u * (hsh & 3u);
              const uint32_t slot_i = (atomicAdd(&S.cnt[hsh >> 2], 1u << sh) >> sh) & (uint32_t)(WAYS - 1);
              st_relaxed_u16(&S.tab[hsh * WAYS + slot_i], (uint16_t)p); // (more than WAYS positions of a step in one bucket share slots: any of them is a valid entry)
            }
            __syncthreads();
            // -- match.  Pass 1: every bucket entry that lies in the window before p and shares its first four bytes is
            // measured over the next four; of the ones equal over all eight, the nearest (up to three) wait for pass 2.
            // Everything below depends on the SET of entries in the bucket, not on their slots (which thread of a step
            // got which slot is a race).
            int mlen = 0, mdist = 0;
            uint32_t lg1 = 0x10000u, lg2 = 0x10000u, lg3 = 0x10000u; // distances of the long candidates, ascending
            const int maxlen = avail < 258 ? avail : 258;
            if (can) {
              const uint32_t maxdist = (uint32_t)min(32768, j + hist);
              const uint32_t P4 = rg32(r32, p), P4b = rg32(r32, p + 4u);
              const uint4 *bk = reinterpret_cast<const uint4 *>(S.tab + hsh * WAYS);
#pragma unroll
              for (int g = 0; g < WAYS / 8; g++) {
                const uint4 e4 = bk[g];
                const uint32_t ew[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                for (int k = 0; k < 8; k++) {
                  if (g * 8 + k < L.ways) {
                    const uint32_t q = (k & 1) ? ew[k >> 1] >> 16 : ew[k >> 1] & 0xffffu;
                    const uint32_t dist = (p - q) & 0xffffu;
                    if (dist - 1u < maxdist && rg32(r32, q) == P4) {
                      const uint32_t x = rg32(r32, q + 4u) ^ P4b;
                      if (x == 0u && maxlen > 8) {
                        if constexpr (WAYS == 16) { // (three long candidates with the deep table, two otherwise)
                          if (dist < lg1) { lg3 = lg2; lg2 = lg1; lg1 = dist; }
                          else if (dist < lg2) { if (dist != lg1) { lg3 = lg2; lg2 = dist; } }
                          else if (dist < lg3 && dist != lg2) lg3 = dist;
                        } else {
                          if (dist < lg1) { lg2 = lg1; lg1 = dist; }
                          else if (dist < lg2 && dist != lg1) lg2 = dist;
                        }
                      } else {
                        const int l = x ? 4 + ((__ffs((int)x) - 1) >> 3) : 8; // (capped to maxlen below)
                        if (l > mlen || (l == mlen && (int)dist < mdist)) { mlen = l; mdist = (int)dist; }
                      }
                    }
                  }
                }
              }
              // (measuring more than two candidates that share eight bytes buys 0.15 % of output at level 6: the deep levels take three)
            }
            // Pass 2: the long candidates, eight bytes per step; the lanes of a warp step together (a lane takes its next
            // candidate when one ends), so the loop runs as long as the warp's longest chain of them, not their sum
            {
              uint32_t q = 0, dist = 0;
              int l = 0;
              bool busy = false;
              while (__any_sync(0xffffffffu, busy || lg1 != 0x10000u)) {
                if (!busy && lg1 != 0x10000u) {
                  dist = lg1;
                  lg1 = lg2; lg2 = lg3; lg3 = 0x10000u;
                  q = (p - dist) & 0xffffu;
                  l = 8;
                  busy = true;
                }
                if (busy) {
                  bool done = l >= maxlen;
                  if (!done) {
                    const unsigned long long y = rg64(r32, q + (uint32_t)l) ^ rg64(r32, p + (uint32_t)l);
                    if (y) { l += (__ffsll((long long)y) - 1) >> 3; done = true; }
                    else l += 8;
                  }
                  if (done) {
                    if (l > maxlen) l = maxlen;
                    if (l > mlen || (l == mlen && (int)dist < mdist)) { mlen = l; mdist = (int)dist; }
                    busy = false;
                    if (mlen >= L.nice) lg1 = 0x10000u; // nice_length: good enough, stop searching (de.ml:4158)
                  }
                }
              }
            }
            if (mlen > maxlen) mlen = maxlen;
            if (mlen < DF_MIN_MATCH) mlen = 0;
            if (L.lazy) { // (a warp's last position keeps its match: its neighbour belongs to another warp)
              const int nlen = __shfl_down_sync(0xffffffffu, mlen, 1);
              if (lane < 31 && mlen && mlen < L.lazy && nlen > mlen) mlen = 0;
            }
            // -- parse.  The warp's transfer function by pointer doubling: from every entry lane, where the orbit leaves
            // the warp's 32 positions (e) and which of them it visits (M).
            const bool exists = j < hend;
            uint32_t e = exists ? (uint32_t)(lane + (mlen ? mlen : 1)) : DC_EXIT_END, M = exists ? 1u << lane : 0u;
#pragma unroll
            for (int r = 0; r < 5; r++) {
              const uint32_t e2 = __shfl_sync(0xffffffffu, e, (int)(e & 31u)), m2 = __shfl_sync(0xffffffffu, M, (int)(e & 31u));
              if (e < 32u) { e = e2; M |= m2; }
            }
            // (absolute within the step: the chain below is a load, a mask, a shift and an add per warp it enters)
            S.xfer[warp][lane] = (e >= DC_EXIT_END ? 0xffffu : 32u * (uint32_t)warp + e) | ((uint32_t)__popc(M) << 16);
            __syncthreads();
            // every warp chains the transfer functions of the warps before it (the same shared-memory words for all its
            // lanes: broadcasts), so nobody waits for a walker; the last warp also leaves the carry for the next step
            const uint32_t *xf = &S.xfer[0][0];
            uint32_t pos = (uint32_t)S.carryq[stepno & 1], tb = S.ntokq[stepno & 1];
            while (pos < 32u * (uint32_t)warp) {
              const uint32_t x = xf[pos];
              tb += x >> 16;
              pos = x & 0xffffu;
            }
            const bool entered = pos < 32u * (uint32_t)(warp + 1);
            const uint32_t marked = entered ? __shfl_sync(0xffffffffu, M, (int)(pos & 31u)) : 0u; // (pos - 32 * warp)
            if (warp == DC_NW - 1 && lane == 0) {
              uint32_t pe = pos, te = tb;
              if (entered) { const uint32_t x = xf[pos]; te += x >> 16; pe = x & 0xffffu; }
              S.carryq[(stepno + 1) & 1] = pe == 0xffffu ? 0 : (int32_t)(pe - (uint32_t)DC_T);
              S.ntokq[(stepno + 1) & 1] = te;
            }
            if ((marked >> lane) & 1u)
              tok[tb + __popc(marked & ((1u << lane) - 1u))] = mlen ? (1u << 23) | ((uint32_t)(mdist - 1) << 8) | (uint32_t)(mlen - 3) : (uint32_t)S.ring[p];
            stepno++;
            // (the next step's inserts wait behind its first barrier for this step's bucket reads: all of them happen
            // before the barrier that follows the transfer functions)
          }
        }
        __syncthreads();
        const int ntok = (int)S.ntokq[stepno & 1];
        // the tokens' symbol frequencies (De.Lz77's literals / distances counters, de.ml:2333-2351), in one pass over
        // the token list: all lanes busy, where counting inside the parse runs one lane in four
        for (int i = tid; i < ntok; i += DC_T) {
          const uint32_t t = tok[i];
          if ((t >> 23) & 1u) {
            uint32_t eb, ev, eb2, ev2;
            atomicAdd(&S.freq_l[df_len_sym((t & 0xffu) + 3u, &eb, &ev)], 1u);
            atomicAdd(&S.freq_d[df_dist_sym(((t >> 8) & 0x7fffu) + 1u, &eb2, &ev2)], 1u);
            my_extra += eb + eb2;
          } else atomicAdd(&S.freq_l[t], 1u);
        }
        my_extra = __reduce_add_sync(0xffffffffu, my_extra);
        if (lane == 0 && my_extra) atomicAdd(&S.extra_bits, my_extra);
        if (tid == 0) S.freq_l[256] = 1;
        __syncthreads();
        // ------------------------------------------------------------ 2 build (one warp: De.T.make is a serial heap)
        if (warp == 0) {
          const uint32_t bits_dyn = df_build_block<DCMP_DC_EXACT_TREE != 0>(S, !last, lane);
          uint32_t bits_fixed = 0;
          for (int s2 = lane; s2 < 286; s2 += 32) bits_fixed += S.freq_l[s2] * (s2 < 144 ? 8u : s2 < 256 ? 9u : s2 < 280 ? 7u : 8u);
          if (lane < 30) bits_fixed += S.freq_d[lane] * 5u;
          bits_fixed = __reduce_add_sync(0xffffffffu, bits_fixed) + 3u + S.extra_bits + (last ? 0u : 3u + 7u + 32u);
          const bool fixed = bits_fixed < bits_dyn;
          const uint32_t bits_stored = 8u * ((uint32_t)len + 5u * (uint32_t)((len + 65534) / 65535));
          if (lane == 0) S.blk_kind = (fixed ? bits_fixed : bits_dyn) >= bits_stored ? 0u : fixed ? 1u : 2u;
        }
        __syncthreads();
        const uint32_t kind = S.blk_kind;
        if (kind == 0u) stored = true;
        else {
          // ---------------------------------------------------------- 3 pack, all warps
          const bool fixed = kind == 1u;
          if (fixed) {
            for (int s2 = tid; s2 < 288; s2 += DC_T) {
              const uint32_t l = s2 < 144 ? 8u : s2 < 256 ? 9u : s2 < 280 ? 7u : 8u;
              const uint32_t cc = s2 < 144 ? 0x30u + (uint32_t)s2 : s2 < 256 ? 0x190u + (uint32_t)(s2 - 144) : s2 < 280 ? (uint32_t)(s2 - 256) : 0xc0u + (uint32_t)(s2 - 280);
              S.len_l[s2] = (uint8_t)l;
              S.code_l[s2] = (uint16_t)(__brev(cc) >> (32 - l));
            }
            if (tid < 32) { S.len_d[tid] = 5; S.code_d[tid] = (uint16_t)(__brev((uint32_t)tid) >> 27); }
          }
          for (int i = tid; i < DC_NW * DC_WIN; i += DC_T) (&S.win_w[0][0])[i] = 0;
          __syncthreads();
          // a warp's tokens: a contiguous range, a multiple of 32 long
          const int per = (((ntok + DC_NW - 1) / DC_NW) + 31) & ~31;
          const int t0 = min(ntok, warp * per), t1 = min(ntok, t0 + per);
          auto tok_bits = [&](uint32_t t, uint64_t &v) -> uint32_t {
            uint32_t nb;
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
            return nb;
          };
          {
            uint32_t bits = 0;
            for (int i = t0 + lane; i < t1; i += 32) { uint64_t v; bits += tok_bits(tok[i], v); }
            bits = __reduce_add_sync(0xffffffffu, bits);
            if (lane == 0) S.warp_bits[warp] = bits;
          }
          // the header, by the first lane of the first warp, at bit 0 of that warp's window
          DcBits bw{S.win_w[warp], out32, 0u, 0u, 0u, 0u};
          if (tid == 0) {
            uint32_t at = 0;
            auto put = [&](uint32_t v, uint32_t n) { bw.put(v, n, at); at += n; if ((at >> 5) >= bw.flushed + DC_WIN / 2) bw.drain_serial(at); };
            put(last ? 1u : 0u, 1);
            if (fixed) put(1u, 2);
            else {
              put(2u, 2);
              put((uint32_t)(S.hlit - 257), 5);
              put((uint32_t)(S.hdist - 1), 5);
              put((uint32_t)(S.hclen - 4), 4);
              for (int i = 0; i < S.hclen; i++) put(S.len_bl[c_bl_order[i]], 3);
              for (int i = 0; i < S.n_rle; i++) {
                const int sy = S.rle_sym[i];
                put(S.code_bl[sy], S.len_bl[sy]);
                if (sy == 16) put(S.rle_extra[i], 2);
                else if (sy == 17) put(S.rle_extra[i], 3);
                else if (sy == 18) put(S.rle_extra[i], 7);
              }
            }
            bw.drain_serial(at);
            S.hdr_bits = at;
          }
          __syncthreads();
          uint32_t start = S.hdr_bits, total = S.hdr_bits;
          for (int w = 0; w < DC_NW; w++) { if (w < warp) start += S.warp_bits[w]; total += S.warp_bits[w]; }
          const uint32_t eob_bits = S.len_l[256];
          uint32_t end_bits = total + eob_bits; // after the end-of-block code
          if (!last) end_bits = ((end_bits + 3u + 7u) & ~7u) + 32u; // + an empty stored block: 3 bits, to the byte, LEN / NLEN
          const uint32_t my_end = warp == DC_NW - 1 ? end_bits : start + S.warp_bits[warp];
          // words shared with a neighbour are OR-ed into the slot: zero them first (the header's complete words, all
          // below the first share, are out already; its last, partial word waits in warp 0's window)
          if (lane == 0) { if (warp > 0) out32[start >> 5] = 0; out32[my_end >> 5] = 0; }
          __syncthreads();
          bw.bitpos = start;
          bw.flushed = start >> 5;
          bw.first_word = warp == 0 ? 0xffffffffu : start >> 5;
          bw.last_word = my_end >> 5;
          for (int g = t0; g < t1; g += 32) {
            const int i = g + lane;
            uint64_t v = 0;
            uint32_t nb = 0;
            if (i < t1) nb = tok_bits(tok[i], v);
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
          if (warp == DC_NW - 1) {
            if (lane == 0) {
              uint32_t at = bw.bitpos;
              bw.put(S.code_l[256], eob_bits, at); at += eob_bits;
              if (!last) { at += 3; at = (at + 7u) & ~7u; bw.put(0xffff0000ull, 32, at); at += 32; }
              S.blk_bytes = (at + 7u) >> 3;
            }
            __syncwarp();
            bw.bitpos = end_bits;
          }
          bw.drain(lane, true);
          __syncthreads();
          if (tid == 0) { outs[c].bytes = S.blk_bytes; outs[c].sum_a = (uint32_t)(S.sum_a % 65521ull); outs[c].sum_b = (uint32_t)(S.sum_b % 65521ull); }
        }
      }
      if (stored) {
        // ------------------------------------------------------------ stored blocks (de.ml:2566-2581)
        if (level == 0 && want_adler) { // (no half was loaded: the sums come from the input)
          unsigned long long a = 0, b = 0;
          for (int j = tid; j < len; j += DC_T) { const uint32_t x = d[j]; a += x; b += (unsigned long long)x * (uint32_t)(len - j); }
          for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
          if (lane == 0) { atomicAdd(&S.sum_a, a); atomicAdd(&S.sum_b, b % 65521ull); }
        }
        uint32_t o = 0;
        int pos = 0;
        do {
          const int n = min(65535, len - pos);
          const bool fin = last && pos + n == len;
          if (tid == 0) {
            slot[o] = fin ? 1 : 0;
            slot[o + 1] = (uint8_t)n; slot[o + 2] = (uint8_t)(n >> 8);
            slot[o + 3] = (uint8_t)~n; slot[o + 4] = (uint8_t)(~n >> 8);
          }
          for (int k = tid; k < n; k += DC_T) slot[o + 5 + k] = d[pos + k];
          o += 5 + (uint32_t)n;
          pos += n;
        } while (pos < len);
        __syncthreads();
        if (tid == 0) { outs[c].bytes = o; outs[c].sum_a = (uint32_t)(S.sum_a % 65521ull); outs[c].sum_b = (uint32_t)(S.sum_b % 65521ull); }
      }
      carried = level != 0 && len == DF_CHUNK; // (a shorter chunk is the last of its stream)
    }
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
size_t deflate_cta_smem_bytes(int level) { return (level >= 7 ? sizeof(DcSmemT<16>) : sizeof(DcSmemT<8>)) + 128; }
static_assert(sizeof(DcSmemT<8>) + 128 + 1024 <= 227 * 1024 / 2, "two compressor CTAs per SM");
#ifndef DCMP_EMU
// DCMP_DEF_KERNEL=warp: every chunk goes to the warp-per-chunk kernel (the round-1 path; A/B measurements)
// (latched at the first call: the scratch of the warp-per-chunk kernel is sized by it)
static bool deflate_force_warp() { static const bool f = [] { const char *v = getenv("DCMP_DEF_KERNEL"); return v && v[0] == 'w'; }(); return f; }
// resident CTAs per SM of the two kernels (the persistent grids and their scratch are sized with them)
int deflate_ctas_per_sm() {
  int n = 0;
  const size_t smem = deflate_smem_bytes(1);
  cudaFuncSetAttribute(deflate_chunk_kernel<DF_HASH_BITS_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, deflate_chunk_kernel<DF_HASH_BITS_FAST>, DF_WARPS * 32, smem) != cudaSuccess || n < 1) n = 2;
  if (getenv("DCMP_DEF_CTAS")) n = atoi(getenv("DCMP_DEF_CTAS")) > 0 ? atoi(getenv("DCMP_DEF_CTAS")) : n; // tuning knob
  return n;
}
constexpr int DC_CTAS_PER_SM = 2;
// scratch: [two work counters | a token per input byte for every CTA of the CTA-per-chunk kernel | the warp-per-chunk kernel's]
size_t deflate_scratch_bytes(int sm_count) {
  const size_t warp_tok = deflate_force_warp() ? DF_CHUNK : DF_SMALL;
  return 256 + (size_t)DC_CTAS_PER_SM * sm_count * DF_CHUNK * 4 + (size_t)deflate_ctas_per_sm() * sm_count * DF_WARPS * warp_tok * 4;
}
#endif
size_t deflate_chunk_desc_bytes() { return sizeof(DfChunk); }
size_t deflate_chunk_out_bytes() { return sizeof(DfChunkOut); }
size_t deflate_slot_bytes() { return DF_SLOT; }
int deflate_chunk_size() { return DF_CHUNK; }
int deflate_small_chunk() { return DF_SMALL; }

#ifndef DCMP_EMU
int configure_deflate() { // per device (see configure_inflate_parallel)
  cudaError_t e = cudaFuncSetAttribute(deflate_chunk_kernel<DF_HASH_BITS_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)deflate_smem_bytes(1));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(deflate_chunk_kernel<DF_HASH_BITS_BEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)deflate_smem_bytes(9));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(deflate_cta_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)deflate_cta_smem_bytes(1));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(deflate_cta_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)deflate_cta_smem_bytes(9));
  return (int)e;
}
void launch_deflate_queue(const int32_t *cmds, int64_t n, int last, uint32_t *tok, uint32_t *slot, uint8_t *dst, int64_t cap, int64_t *written,
                          int32_t *status, cudaStream_t stream) {
  deflate_queue_kernel<<<1, 32, sizeof(DfSmemT<DF_HASH_BITS_BEST, uint64_t>) + 128, stream>>>(cmds, n, last, tok, slot, dst, cap, written, status);
}
void launch_deflate_tmake(const int32_t *freqs, int length, int max_bits, int32_t *lens, int32_t *codes, int32_t *max_code, cudaStream_t stream) {
  deflate_tmake_kernel<<<1, 32, sizeof(DfSmemT<DF_HASH_BITS_BEST, uint64_t>) + 128, stream>>>(freqs, length, max_bits, lens, codes, max_code);
}
// `scratch`: deflate_scratch_bytes(sm_count) bytes; `nsmall`: chunks of the batch that are small single-chunk streams
int launch_deflate(int level, int fmt, int64_t nstreams, int64_t nchunks, int64_t nsmall, const uint8_t *src_base, const void *chunks,
                   const int64_t *first_chunk, void *outs, uint8_t *slots, int64_t *chunk_dst, uint8_t *scratch,
                   int sm_count, uint8_t *dst_base, const int64_t *dst_off,
                   const int64_t *dst_cap, int64_t *written, int32_t *status, cudaStream_t stream) {
  unsigned long long *counters = reinterpret_cast<unsigned long long *>(scratch);
  cudaMemsetAsync(counters, 0, 16, stream);
  uint32_t *tok_cta = reinterpret_cast<uint32_t *>(scratch + 256);
  uint32_t *tok_warp = tok_cta + (size_t)DC_CTAS_PER_SM * sm_count * DF_CHUNK;
  const bool force_warp = deflate_force_warp();
  const int want_adler = fmt == DCMP_FMT_ZLIB ? 1 : 0;
  if (force_warp) nsmall = nchunks;
  const int tune_ways = getenv("DCMP_DC_WAYS") ? atoi(getenv("DCMP_DC_WAYS")) : 0;
  if (nchunks - nsmall > 0) {
    int64_t want = (nchunks + DC_RUN - 1) / DC_RUN;
    if (level >= 7) {
      int ctas = sm_count; // 1 per SM: 146 KiB of shared memory each
      if (ctas > want) ctas = (int)want;
      deflate_cta_kernel<16><<<(unsigned)ctas, dc_threads(16), deflate_cta_smem_bytes(level), stream>>>(
          level, want_adler, nchunks, src_base, (const DfChunk *)chunks, (DfChunkOut *)outs, slots, tok_cta, counters, tune_ways);
    } else {
      int ctas = DC_CTAS_PER_SM * sm_count;
      if (ctas > want) ctas = (int)want;
      deflate_cta_kernel<8><<<(unsigned)ctas, dc_threads(8), deflate_cta_smem_bytes(level), stream>>>(
          level, want_adler, nchunks, src_base, (const DfChunk *)chunks, (DfChunkOut *)outs, slots, tok_cta, counters, tune_ways);
    }
  }
  if (nsmall > 0) {
    const size_t smem = deflate_smem_bytes(level);
    int ctas = deflate_ctas_per_sm() * sm_count;
    int64_t want = ((force_warp ? nchunks : nsmall) + DF_WARPS - 1) / DF_WARPS;
    if (ctas > want) ctas = (int)want;
    if (deflate_hash_bits(level) == DF_HASH_BITS_BEST)
      deflate_chunk_kernel<DF_HASH_BITS_BEST><<<(unsigned)ctas, DF_WARPS * 32, smem, stream>>>(
          level, want_adler, nchunks, src_base, (const DfChunk *)chunks, (DfChunkOut *)outs, slots, tok_warp, counters + 1, force_warp ? 0 : 1);
    else
      deflate_chunk_kernel<DF_HASH_BITS_FAST><<<(unsigned)ctas, DF_WARPS * 32, smem, stream>>>(
          level, want_adler, nchunks, src_base, (const DfChunk *)chunks, (DfChunkOut *)outs, slots, tok_warp, counters + 1, force_warp ? 0 : 1);
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
