/* oracle/inflate_ns.c — restatement of De.Inf.Ns.inflate (one-shot inflate).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Follows, statement by
 * statement, /root/reference/lib/de.ml:
 *   constants            de.ml:225-325
 *   Inf.huffman          de.ml:523-638
 *   Inf.resolve          de.ml:640-647
 *   fixed tables         de.ml:821-833
 *   Inf.Ns               de.ml:1534-1823
 *
 * OCaml `int` is 63-bit two's complement; `hold` never exceeds 31 live bits
 * here, but `bits` may go negative on truncated input (de.ml:1643-1654 does not
 * fail) and `d.bits lsr 3` (de.ml:1805) is then a 63-bit logical shift.  Both
 * are reproduced.
 */
#include "oracle.h"
#include <string.h>
#include <setjmp.h>

static const int zigzag[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5,
                               11, 4, 12, 3, 13, 2, 14, 1, 15}; /* de.ml:237 */
static const int base_length[32] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16,
                                    20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112,
                                    128, 160, 192, 224, 255, 0, 0, 0}; /* de.ml:293 (31 entries; index masked 0x1f) */
static const int extra_lbits[32] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                    3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0}; /* de.ml:307 */
static const int extra_dbits[32] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
                                    7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 0, 0}; /* de.ml:313 */
static const int base_dist[32] = {0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96,
                                  128, 192, 256, 384, 512, 768, 1024, 1536, 2048,
                                  3072, 4096, 6144, 8192, 12288, 16384, 24576, -1, -1}; /* de.ml:321 */

#define LINK_FLAG (1 << 20) /* de.ml:520 */
#define VMASK 0x7fff        /* Lookup.mask, de.ml:360 */

enum { K_CODES = 0, K_LENS = 1, K_DISTS = 2 };

/* de.ml:523-638 */
int orc_huffman(int kind, const int *table, int off, int codes, int32_t *tbl,
                int *size_out, int *root_out, int *max_out) {
  int bl_count[16];
  int max = 15, min, left, idx, sym;
  int offs[16];
  int work[320];
  memset(bl_count, 0, sizeof bl_count);
  for (sym = 0; sym < codes; sym++) bl_count[table[off + sym]]++;
  while (max >= 1 && bl_count[max] == 0) max--;
  if (max == 0) { /* empty_table, de.ml:521,542 */
    tbl[0] = 1 << 15;
    *size_out = 1; *root_out = 1; *max_out = 1;
    return 0;
  }
  left = 1;
  for (idx = 1; idx <= 15; idx++) {
    left = (left << 1) - bl_count[idx];
    if (left < 0) return -1;
  }
  if (left > 0 && (kind == K_CODES || max != 1)) return -1;
  min = 1;
  while (min <= 15 && bl_count[min] == 0) min++;
  memset(offs, 0, sizeof offs);
  for (idx = 1; idx <= 14; idx++) offs[idx + 1] = offs[idx] + bl_count[idx];
  for (sym = 0; sym < codes; sym++) {
    int l = table[off + sym];
    if (l != 0) work[offs[l]++] = sym;
  }
  int root = kind == K_LENS ? 9 : kind == K_DISTS ? 6 : 7;
  if (root > max) root = max;
  if (root < min) root = min;
  int size = (max <= root) ? (1 << max)
             : kind == K_LENS ? 852 : kind == K_DISTS ? 592 : (1 << max);
  memset(tbl, 0, sizeof(int32_t) * (size_t)size);
  {
    int huff = 0, len = min, next = 0, curr = root, drop = 0, low = -1;
    int mask = (1 << root) - 1, fill_size = 0, finished = 0;
    sym = 0;
    while (!finished) {
      int value = work[sym];
      int entry = (len << 15) | value;
      int step = 1 << (len - drop);
      int fill, inc;
      fill_size = 1 << curr;
      fill = fill_size;
      do {
        int at;
        fill -= step;
        at = next + (huff >> drop) + fill;
        if (at >= size || at >= ORC_TBL_CAP) return -2; /* OCaml: index out of bounds */
        tbl[at] = entry;
      } while (fill != 0);
      inc = 1 << (len - 1); /* backward, de.ml:583-585 */
      while (huff & inc) inc >>= 1;
      huff = inc != 0 ? (huff & (inc - 1)) + inc : 0;
      sym++;
      bl_count[len]--;
      if (bl_count[len] == 0) {
        if (len == max) finished = 1;
        else len = table[off + work[sym]];
      }
      if (!finished && len > root && (huff & mask) != low) {
        int l;
        if (drop == 0) drop = root;
        next += fill_size;
        curr = len - drop;
        l = 1 << curr;
        while (curr + drop < max) {
          l -= bl_count[curr + drop];
          if (l <= 0) break;
          curr++;
          l <<= 1;
        }
        low = huff & mask;
        if (low >= size) return -2;
        tbl[low] = LINK_FLAG | (curr << 15) | next;
      }
    }
  }
  *size_out = size; *root_out = root; *max_out = max;
  return 0;
}

typedef struct {
  int32_t t[ORC_TBL_CAP];
  int size, m, l, root;
} lookup_t;

typedef struct {
  const uint8_t *i;
  int64_t i_pos, i_len;
  uint64_t hold;
  int64_t bits;
  uint8_t *o;
  int64_t o_pos, o_len;
  jmp_buf jb;
} ns_t;

#define FAIL(d, code) longjmp((d)->jb, (code))

static void lookup_make(lookup_t *lk, int size, int root, int max) { /* de.ml:362-364 */
  lk->size = size; lk->m = (1 << root) - 1; lk->l = max; lk->root = root;
}

/* de.ml:640-647; an out-of-range index is an uncaught OCaml exception */
static int32_t resolve(ns_t *d, const lookup_t *lk, uint64_t hold) {
  int64_t ix = (int64_t)(hold & (uint64_t)lk->m);
  int32_t value;
  if (ix >= lk->size) FAIL(d, ORC_REF_EXCEPTION);
  value = lk->t[ix];
  if ((value & LINK_FLAG) == 0) return value;
  {
    int sub = (value >> 15) & 0x1f;
    int64_t j = (value & 0x7fff) + (int64_t)((hold >> lk->root) & ((1u << sub) - 1));
    if (j >= lk->size) FAIL(d, ORC_REF_EXCEPTION);
    return lk->t[j];
  }
}

static int64_t i_rem(ns_t *d) { return d->i_len - d->i_pos; }

/* shift that tolerates OCaml's behaviour for shift counts >= 63 / negative:
 * they only occur once input is exhausted and the value shifted is 0. */
static uint64_t shl(uint64_t v, int64_t n) { return (n < 0 || n > 62) ? 0 : (v << n) & 0x7fffffffffffffffULL; }

/* de.ml:1629-1641 */
static void fill_bits_strict(ns_t *d, int n) {
  if (d->bits < n) {
    int64_t rem = i_rem(d);
    if (rem > 1) {
      uint64_t v = d->i[d->i_pos] | ((uint64_t)d->i[d->i_pos + 1] << 8);
      d->hold |= shl(v, d->bits); d->i_pos += 2; d->bits += 16;
    } else if (rem == 1) {
      d->hold |= shl(d->i[d->i_pos], d->bits); d->i_pos += 1; d->bits += 8;
    } else FAIL(d, ORC_UNEXPECTED_END_OF_INPUT);
  }
}

/* de.ml:1643-1654 */
static void fill_bits_lax(ns_t *d, int n) {
  if (d->bits < n) {
    int64_t rem = i_rem(d);
    if (rem > 1) {
      uint64_t v = d->i[d->i_pos] | ((uint64_t)d->i[d->i_pos + 1] << 8);
      d->hold |= shl(v, d->bits); d->i_pos += 2; d->bits += 16;
    } else if (rem == 1) {
      d->hold |= shl(d->i[d->i_pos], d->bits); d->i_pos += 1; d->bits += 8;
    }
  }
}

/* de.ml:1656-1661 */
static int pop_bits(ns_t *d, int n) {
  int v = (int)(d->hold & ((1u << n) - 1));
  d->hold >>= n; d->bits -= n;
  return v;
}

/* de.ml:1589-1611: both arms equal a byte-replicating forward copy */
static void blit(const uint8_t *src, int64_t so, uint8_t *dst, int64_t dof, int64_t len) {
  int64_t k;
  for (k = 0; k < len; k++) dst[dof + k] = src[so + k];
}

/* de.ml:1613-1627 */
static void flat(ns_t *d) {
  int len, nlen;
  d->i_pos -= d->bits / 8;
  d->hold = 0; d->bits = 0;
  if (i_rem(d) < 4) FAIL(d, ORC_UNEXPECTED_END_OF_INPUT);
  len = d->i[d->i_pos] | (d->i[d->i_pos + 1] << 8);
  nlen = d->i[d->i_pos + 2] | (d->i[d->i_pos + 3] << 8);
  d->i_pos += 4;
  if (nlen != 0xffff - len) FAIL(d, ORC_INVALID_COMPLEMENT_OF_LENGTH);
  if (len > i_rem(d)) FAIL(d, ORC_UNEXPECTED_END_OF_INPUT);
  if (len > d->o_len - d->o_pos) FAIL(d, ORC_UNEXPECTED_END_OF_OUTPUT);
  blit(d->i, d->i_pos, d->o, d->o_pos, len);
  d->o_pos += len; d->i_pos += len;
}

/* de.ml:1667-1712 */
static void inflate(ns_t *d, const lookup_t *lit, const lookup_t *dist) {
  for (;;) {
    int32_t code; int value, len;
    fill_bits_lax(d, lit->l);
    code = resolve(d, lit, d->hold);
    value = code & VMASK; len = code >> 15;
    d->hold >>= len; d->bits -= len;
    if (value < 256) {
      if (d->o_pos >= d->o_len) FAIL(d, ORC_UNEXPECTED_END_OF_OUTPUT);
      d->o[d->o_pos++] = (uint8_t)value;
    } else if (value == 256) {
      return;
    } else {
      int l = value - 257, extra, dsym, dd; int64_t off;
      if (l >= 32) FAIL(d, ORC_REF_EXCEPTION);
      len = extra_lbits[l];
      fill_bits_lax(d, len);
      extra = pop_bits(d, len);
      l = base_length[l & 0x1f] + 3 + extra;
      fill_bits_lax(d, dist->l);
      code = resolve(d, dist, d->hold);
      dsym = code & VMASK; len = code >> 15;
      d->hold >>= len; d->bits -= len;
      len = extra_dbits[dsym & 0x1f];
      fill_bits_lax(d, len);
      extra = pop_bits(d, len);
      if (dsym >= 32) FAIL(d, ORC_REF_EXCEPTION);
      dd = base_dist[dsym] + 1 + extra;
      if (dd == 0) FAIL(d, ORC_INVALID_DISTANCE_CODE);
      if (dd > (d->o_pos < 32768 ? d->o_pos : 32768)) FAIL(d, ORC_INVALID_DISTANCE);
      off = d->o_pos - dd;
      if (l > d->o_len - d->o_pos) FAIL(d, ORC_UNEXPECTED_END_OF_OUTPUT);
      blit(d->o, off, d->o, d->o_pos, l);
      d->o_pos += l;
    }
  }
}

static lookup_t g_fixed_lit, g_fixed_dist;
static int g_fixed_ready;

static uint8_t reverse8(uint8_t b) { /* de.ml:792-819 */
  b = (uint8_t)((b & 0xF0) >> 4 | (b & 0x0F) << 4);
  b = (uint8_t)((b & 0xCC) >> 2 | (b & 0x33) << 2);
  b = (uint8_t)((b & 0xAA) >> 1 | (b & 0x55) << 1);
  return b;
}

/* de.ml:821-833 */
__attribute__((constructor)) static void init_fixed(void) {
  int lens[288], n, size, root, max;
  if (g_fixed_ready) return;
  for (n = 0; n < 288; n++) lens[n] = n < 144 ? 8 : n < 256 ? 9 : n < 280 ? 7 : 8;
  orc_huffman(K_LENS, lens, 0, 288, g_fixed_lit.t, &size, &root, &max);
  lookup_make(&g_fixed_lit, size, root, max);
  for (n = 0; n < 32; n++) g_fixed_dist.t[n] = (5 << 15) | reverse8((uint8_t)(n << 3));
  lookup_make(&g_fixed_dist, 32, 5, 5);
  g_fixed_ready = 1;
}

/* de.ml:1718-1731 */
static void make_table(ns_t *d, const int *t, int hlit, int hdist) {
  static __thread lookup_t lit, dist;
  int size, root, max, r;
  if (t[256] == 0) FAIL(d, ORC_INVALID_DICTIONARY);
  r = orc_huffman(K_LENS, t, 0, hlit, lit.t, &size, &root, &max);
  if (r == -1) FAIL(d, ORC_INVALID_DICTIONARY);
  if (r == -2) FAIL(d, ORC_REF_EXCEPTION);
  lookup_make(&lit, size, root, max);
  r = orc_huffman(K_DISTS, t, hlit, hdist, dist.t, &size, &root, &max);
  if (r == -1) FAIL(d, ORC_INVALID_DICTIONARY);
  if (r == -2) FAIL(d, ORC_REF_EXCEPTION);
  lookup_make(&dist, size, root, max);
  inflate(d, &lit, &dist);
}

/* de.ml:1733-1769 */
static void inflate_table(ns_t *d, const int32_t *t, int tsize, int max_bits, int hlit, int hdist) {
  int res[320 + 64];
  int max_res = hlit + hdist, mask = (1 << max_bits) - 1, i = 0;
  memset(res, 0, sizeof res);
  for (;;) {
    int v, len, copy, fillv, x; int64_t ix;
    fill_bits_strict(d, max_bits);
    ix = (int64_t)(d->hold & (uint64_t)mask);
    if (ix >= tsize) FAIL(d, ORC_REF_EXCEPTION);
    v = t[ix] & VMASK; len = t[ix] >> 15;
    d->hold >>= len; d->bits -= len;
    if (v < 16) {
      res[i] = v;
      if (i + 1 < max_res) { i++; continue; }
      break;
    } else if (v == 16) {
      if (i == 0) FAIL(d, ORC_INVALID_DICTIONARY);
      fill_bits_strict(d, 2); copy = pop_bits(d, 2) + 3; fillv = res[i - 1];
    } else if (v == 17) {
      fill_bits_strict(d, 3); copy = pop_bits(d, 3) + 3; fillv = 0;
    } else if (v == 18) {
      fill_bits_strict(d, 7); copy = pop_bits(d, 7) + 11; fillv = 0;
    } else {
      FAIL(d, ORC_REF_EXCEPTION); /* assert false, de.ml:1767 */
    }
    if (i + copy > max_res) FAIL(d, ORC_INVALID_DICTIONARY);
    for (x = 0; x < copy; x++) res[i + x] = fillv;
    if (i + copy < max_res) { i += copy; continue; }
    break;
  }
  make_table(d, res, hlit, hdist);
}

/* de.ml:1771-1786 */
static void table(ns_t *d, int hlit, int hdist, int hclen) {
  int res[19], i, size, root, max, r;
  static __thread int32_t t[ORC_TBL_CAP];
  memset(res, 0, sizeof res);
  for (i = 0; i < hclen; i++) {
    fill_bits_strict(d, 3);
    res[zigzag[i]] = pop_bits(d, 3);
  }
  r = orc_huffman(K_CODES, res, 0, 19, t, &size, &root, &max);
  if (r == -1) FAIL(d, ORC_INVALID_DICTIONARY);
  if (r == -2) FAIL(d, ORC_REF_EXCEPTION);
  inflate_table(d, t, size, max, hlit, hdist);
}

/* de.ml:1788-1793 */
static void dynamic(ns_t *d) {
  int hlit, hdist, hclen;
  fill_bits_strict(d, 14);
  hlit = pop_bits(d, 5) + 257;
  hdist = pop_bits(d, 5) + 1;
  hclen = pop_bits(d, 4) + 4;
  table(d, hlit, hdist, hclen);
}

/* de.ml:1795-1805 */
static void decode(ns_t *d) {
  for (;;) {
    int last, block_type;
    fill_bits_strict(d, 3);
    last = pop_bits(d, 1) == 1;
    block_type = pop_bits(d, 2);
    switch (block_type) {
    case 0: flat(d); break;
    case 1: inflate(d, &g_fixed_lit, &g_fixed_dist); break;
    case 2: dynamic(d); break;
    default: FAIL(d, ORC_INVALID_KIND_OF_BLOCK);
    }
    if (last) {
      /* d.i_pos - (d.bits lsr 3): 63-bit logical shift of a possibly negative int */
      uint64_t b63 = (uint64_t)d->bits & 0x7fffffffffffffffULL;
      d->i_pos -= (int64_t)(b63 >> 3);
      return;
    }
  }
}

/* de.ml:1807-1822 */
int orc_inflate_ns(const uint8_t *src, int64_t src_len, uint8_t *dst,
                   int64_t dst_len, int64_t *consumed, int64_t *produced) {
  ns_t d;
  int rc;
  init_fixed();
  d.i = src; d.i_pos = 0; d.i_len = src_len;
  d.o = dst; d.o_pos = 0; d.o_len = dst_len;
  d.hold = 0; d.bits = 0;
  rc = setjmp(d.jb);
  if (rc == 0) {
    decode(&d);
    *consumed = d.i_pos; *produced = d.o_pos;
    return ORC_OK;
  }
  *consumed = d.i_pos; *produced = d.o_pos;
  return rc;
}
