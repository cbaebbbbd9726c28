/* oracle/deflate_t.c — CPU restatement of the encoder side the reference pins with known answers:
 *   De.T.make            lib/de.ml:1828-2068  (zlib's build_tree / gen_bitlen / gen_codes: heap with depth tie-break,
 *                                              at least two codes, overflow repair, bit-reversed canonical codes)
 *   De.T.scan / symbols  lib/de.ml:2070-2192  (code-length RLE with 16 / 17 / 18)
 *   Def.dynamic_of_frequencies, bl_tree       lib/de.ml:2367-2407
 *   Def's bit writer and dynamic header       lib/de.ml:2529-2633, 2657-2897 (LSB-first, 16 bits at a time)
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Pinned by test/test.ml:507-531 (huffman_length_extra: the 16 bytes
 * ed c1 01 01 00 00 00 40 20 ff 57 1b 42 c1 ea 04) and test/test.ml:1169-1237 (tree_0; the rfc5322 histogram: literal
 * '6' = (10, 0x35f), end of block = (15, 0x6fff) — "the same tree as zlib"). */
#include <string.h>
#include "oracle.h"

enum { L_CODES = 286, D_CODES = 30, BL_CODES = 19, HEAP_SIZE = 2 * L_CODES + 1, MAX_BITS = 15 };
static const int zigzag[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15}; /* de.ml:237 */

typedef struct { int heap[HEAP_SIZE + 1]; int len, max; } heap_t;

/* de.ml:1876-1877 */
static int smaller(const int *freqs, int n, int m, const int *depth) {
  return freqs[n] < freqs[m] || (freqs[n] == freqs[m] && depth[n] <= depth[m]);
}
/* de.ml:1879-1898 */
static void pqdownheap(const int *freqs, const int *depth, heap_t *h, int k) {
  int v = h->heap[k], j = k << 1;
  while (j <= h->len) {
    if (j < h->len && smaller(freqs, h->heap[j + 1], h->heap[j], depth)) j++;
    if (smaller(freqs, v, h->heap[j], depth)) break;
    h->heap[k] = h->heap[j];
    k = j;
    j <<= 1;
  }
  h->heap[k] = v;
}
/* de.ml:1909-1924 */
static int reverse_code(int code, int len) {
  int res = 0;
  do { res |= code & 1; code >>= 1; res <<= 1; } while (--len > 0);
  return res >> 1;
}

/* T.make ~length ?max_length freqs ~bl_count (de.ml:2013-2068).  freqs: 2 * L_CODES + 1 ints, the first `length` are the
 * frequencies, the rest is the builder's workspace (as in the reference).  lengths / codes: HEAP_SIZE ints. */
int orc_t_make(int length, int max_length, int *freqs, int *tree_lengths, int *tree_codes, int *bl_count) {
  heap_t heap;
  int depth[2 * L_CODES + 1], tree_dads[HEAP_SIZE];
  memset(&heap, 0, sizeof heap);
  heap.max = HEAP_SIZE;
  memset(depth, 0, sizeof depth);
  memset(tree_dads, 0, sizeof tree_dads);
  memset(tree_lengths, 0, HEAP_SIZE * sizeof(int));
  /* Heap.populate, de.ml:1834-1856 */
  int max_code = -1;
  for (int n = 0; n < length; n++) {
    if (freqs[n] != 0) { heap.heap[++heap.len] = n; max_code = n; depth[n] = 0; }
    else tree_lengths[n] = 0;
  }
  /* Heap.pkzip, de.ml:1862-1874 */
  while (heap.len < 2) {
    int node = max_code < 2 ? ++max_code : 0;
    freqs[node] = 1;
    heap.heap[++heap.len] = node;
    depth[node] = 0;
  }
  for (int n = heap.len / 2; n >= 1; n--) pqdownheap(freqs, depth, &heap, n);
  int node = length;
  do { /* de.ml:2027-2050 */
    int n = heap.heap[1]; /* pqremove */
    heap.heap[1] = heap.heap[heap.len--];
    pqdownheap(freqs, depth, &heap, 1);
    int m = heap.heap[1];
    heap.heap[--heap.max] = n;
    heap.heap[--heap.max] = m;
    freqs[node] = freqs[n] + freqs[m];
    depth[node] = (depth[n] >= depth[m] ? depth[n] : depth[m]) + 1;
    tree_dads[n] = node;
    tree_dads[m] = node;
    heap.heap[1] = node++;
    pqdownheap(freqs, depth, &heap, 1);
  } while (heap.len >= 2);
  heap.heap[--heap.max] = heap.heap[1];
  /* generate_lengths, de.ml:1952-2009 */
  tree_lengths[heap.heap[heap.max]] = 0;
  int overflow = 0;
  for (int b = 0; b <= MAX_BITS; b++) bl_count[b] = 0;
  for (int h = heap.max + 1; h < HEAP_SIZE; h++) {
    int n = heap.heap[h];
    int bits = tree_lengths[tree_dads[n]] + 1;
    if (bits > max_length) { overflow++; bits = max_length; }
    tree_lengths[n] = bits;
    if (n <= max_code) bl_count[bits]++;
  }
  if (overflow != 0) {
    do {
      int bits = max_length - 1;
      while (bl_count[bits] == 0) bits--;
      bl_count[bits]--;
      bl_count[bits + 1] += 2;
      bl_count[max_length]--;
      overflow -= 2;
    } while (overflow > 0);
    int h = HEAP_SIZE;
    for (int bits = max_length; bits >= 1; bits--) {
      int n = bl_count[bits];
      while (n != 0) {
        int m = heap.heap[--h];
        if (m <= max_code) {
          if (tree_lengths[m] != bits) tree_lengths[m] = bits;
          n--;
        }
      }
    }
  }
  /* generate_codes, de.ml:1926-1950 */
  int next_code[MAX_BITS + 1], code = 0;
  memset(tree_codes, 0, HEAP_SIZE * sizeof(int));
  next_code[0] = 0;
  for (int bits = 1; bits <= MAX_BITS; bits++) { code = (code + bl_count[bits - 1]) << 1; next_code[bits] = code & 0xffff; }
  for (int n = 0; n <= max_code; n++) {
    int len = tree_lengths[n];
    if (len > 0) { tree_codes[n] = reverse_code(next_code[len], len); next_code[len]++; }
  }
  return max_code;
}

/* T.scan (de.ml:2070-2119): bl_freqs of the code-length alphabet for one tree */
static void t_scan(int *tree_lengths, int max_code, int *bl_freqs) {
  int prevlen = -1, nextlen = tree_lengths[0], curlen, count = 0, max_count = 7, min_count = 4;
  if (nextlen == 0) { max_count = 138; min_count = 3; }
  tree_lengths[max_code + 1] = 0xffff;
  for (int n = 0; n <= max_code; n++) {
    curlen = nextlen;
    nextlen = tree_lengths[n + 1];
    count++;
    if (count < max_count && curlen == nextlen) continue;
    else if (count < min_count) bl_freqs[curlen] += count;
    else if (curlen != 0) { if (curlen != prevlen) bl_freqs[curlen]++; bl_freqs[16]++; }
    else if (count <= 10) bl_freqs[17]++;
    else bl_freqs[18]++;
    count = 0;
    prevlen = curlen;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    else if (curlen == nextlen) { max_count = 6; min_count = 3; }
    else { max_count = 7; min_count = 4; }
  }
}
/* T.symbols (de.ml:2124-2190): the RLE symbols as (len << 15 | code) words, extra bits as their own words */
static int t_symbols(int i, const int *tree_lengths, int max_code, int *bl_symbols, const int *bl_len, const int *bl_code) {
  int prevlen = -1, nextlen = tree_lengths[0], curlen, count = 0, max_count = 7, min_count = 4;
#define SYM(s) ((bl_len[s] << MAX_BITS) | bl_code[s])
#define BITS(c, l) (((l) << MAX_BITS) | (c))
  if (nextlen == 0) { max_count = 138; min_count = 3; }
  for (int n = 0; n <= max_code; n++) {
    curlen = nextlen;
    nextlen = tree_lengths[n + 1];
    count++;
    if (count < max_count && curlen == nextlen) continue;
    else if (count < min_count) { do { bl_symbols[i++] = SYM(curlen); } while (--count != 0); }
    else if (curlen != 0) {
      if (curlen != prevlen) { bl_symbols[i++] = SYM(curlen); count--; }
      bl_symbols[i++] = SYM(16);
      bl_symbols[i++] = BITS(count - 3, 2);
    } else if (count <= 10) { bl_symbols[i++] = SYM(17); bl_symbols[i++] = BITS(count - 3, 3); }
    else { bl_symbols[i++] = SYM(18); bl_symbols[i++] = BITS(count - 11, 7); }
    count = 0;
    prevlen = curlen;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    else if (curlen == nextlen) { max_count = 6; min_count = 3; }
    else { max_count = 7; min_count = 4; }
  }
  return i;
}

/* the bit writer, de.ml:2529-2543: hold |= bits << nbits, 16 bits leave at a time (little-endian) */
typedef struct { uint8_t *o; int64_t pos, cap; uint64_t hold; int bits; int overflow; } bw_t;
static void c_bits(bw_t *e, int bits, int n) {
  e->hold |= (uint64_t)bits << e->bits;
  e->bits += n;
  while (e->bits >= 16) {
    if (e->pos + 2 > e->cap) { e->overflow = 1; return; }
    e->o[e->pos++] = (uint8_t)e->hold; e->o[e->pos++] = (uint8_t)(e->hold >> 8);
    e->hold >>= 16; e->bits -= 16;
  }
}
static void flush_bits(bw_t *e) { /* de.ml:2551-2567 */
  while (e->bits > 0) {
    if (e->pos + 1 > e->cap) { e->overflow = 1; return; }
    e->o[e->pos++] = (uint8_t)e->hold; e->hold >>= 8; e->bits -= 8;
  }
  e->bits = 0;
}

static const int base_length[29] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 255};
static const int extra_lbits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const int base_dist[30] = {0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, 12288, 16384, 24576};
static const int extra_dbits[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static int length_code(int len3) { /* _length, de.ml:240-257: len - 3 -> code */
  int c = 28;
  if (len3 == 255) return 28;
  while (base_length[c] > len3) c--;
  return c;
}
static int distance_code(int d1) { /* _distance, de.ml:259-291: dist - 1 -> code */
  int c = 29;
  while (base_dist[c] > d1) c--;
  return c;
}

/* One Dynamic block over a list of Queue commands (de.ml:2245-2266: literal = byte, end of block = 256, copy =
 * 0x2000000 | (len - 3) << 16 | (off - 1)), frequencies counted from the commands the way the tests do
 * (make_literals: end of block starts at 1; succ_literal / succ_length / succ_distance, de.ml:2333-2351).
 * Returns the bytes written, -1 when `cap` is too small.  h[0..2] get h_lit, h_dst, h_len. */
int64_t orc_def_encode_dynamic(const int32_t *cmds, int64_t n, int last, uint8_t *out, int64_t cap, int *h) {
  int lit[2 * L_CODES + 1], dst[2 * D_CODES + 1], blf[2 * BL_CODES + 1];
  int l_len[HEAP_SIZE + 1], l_code[HEAP_SIZE], d_len[HEAP_SIZE + 1], d_code[HEAP_SIZE], b_len[HEAP_SIZE + 1], b_code[HEAP_SIZE], bl_count[MAX_BITS + 1];
  memset(lit, 0, sizeof lit); memset(dst, 0, sizeof dst); memset(blf, 0, sizeof blf);
  lit[256] = 1;
  for (int64_t k = 0; k < n; k++) {
    int32_t c = cmds[k];
    if (c & 0x2000000) { lit[257 + length_code((c >> 16) & 0x1ff)]++; dst[distance_code(c & 0xffff)]++; }
    else if (c != 256) lit[c & 0xff]++;
  }
  /* Def.dynamic_of_frequencies, de.ml:2387-2407 */
  int l_max = orc_t_make(L_CODES, MAX_BITS, lit, l_len, l_code, bl_count);
  int d_max = orc_t_make(D_CODES, MAX_BITS, dst, d_len, d_code, bl_count);
  t_scan(l_len, l_max, blf);
  t_scan(d_len, d_max, blf);
  orc_t_make(BL_CODES, 7, blf, b_len, b_code, bl_count);
  int max_blindex = BL_CODES - 1;
  while (max_blindex >= 3 && b_len[zigzag[max_blindex]] == 0) max_blindex--;
  int syms[L_CODES + D_CODES + 64];
  l_len[l_max + 1] = 0xffff; d_len[d_max + 1] = 0xffff;
  int ns = t_symbols(0, l_len, l_max, syms, b_len, b_code);
  ns = t_symbols(ns, d_len, d_max, syms, b_len, b_code);
  const int h_lit = l_max + 1, h_dst = d_max + 1, h_len = max_blindex + 1;
  if (h) { h[0] = h_lit; h[1] = h_dst; h[2] = h_len; }
  /* encode_dynamic_header / zigzag / huffman, de.ml:2583-2624 */
  bw_t e = {out, 0, cap, 0, 0, 0};
  c_bits(&e, last ? 1 : 0, 1);
  c_bits(&e, 2, 2);
  c_bits(&e, h_lit - 257, 5);
  c_bits(&e, h_dst - 1, 5);
  c_bits(&e, h_len - 4, 4);
  for (int r = 0; r < h_len; r++) c_bits(&e, b_len[zigzag[r]], 3);
  for (int r = 0; r < ns; r++) c_bits(&e, syms[r] & ((1 << MAX_BITS) - 1), syms[r] >> MAX_BITS);
  /* write, de.ml:2657-2897 */
  for (int64_t k = 0; k < n; k++) {
    int32_t c = cmds[k];
    if (c & 0x2000000) {
      int len3 = (c >> 16) & 0x1ff, d1 = c & 0xffff;
      int lc = length_code(len3), dc = distance_code(d1);
      c_bits(&e, l_code[257 + lc], l_len[257 + lc]);
      c_bits(&e, len3 - base_length[lc], extra_lbits[lc]);
      c_bits(&e, d_code[dc], d_len[dc]);
      c_bits(&e, d1 - base_dist[dc], extra_dbits[dc]);
    } else if (c == 256) c_bits(&e, l_code[256], l_len[256]);
    else c_bits(&e, l_code[c & 0xff], l_len[c & 0xff]);
  }
  flush_bits(&e);
  return e.overflow ? -1 : e.pos;
}
