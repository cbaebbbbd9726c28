// decompress_b200/csrc/inflate_tables.cuh — canonical-Huffman decode-table construction.
//
// Same accept/reject rules and the same two-level shape as the reference's
// Inf.huffman (lib/de.ml:523-638, itself zlib's inflate_table): root of 9 / 6 / 7
// bits clamped to [min,max], sub-tables sized so the remaining codes fit, 852 / 592
// entry capacity.  What differs is the entry payload (dcmp_common.cuh) and who runs
// it: one lane of the warp that owns the stream, writing straight into shared memory.
#pragma once
#include "dcmp_common.cuh"

namespace dcmp {

enum : int { TBL_OK = 0, TBL_INVALID = 1, TBL_EXCEPTION = 2 };

struct TableInfo {
  int root; // index bits of the first level (Lookup.root / Lookup.m, de.ml:358-364)
  int maxl; // longest code (Lookup.l)
};

__device__ __forceinline__ uint32_t symbol_entry(int kind, int sym) {
  if (kind == KIND_LENS) return lit_symbol_entry(sym);
  if (kind == KIND_DISTS) return dist_symbol_entry(sym);
  return mk_entry(K_LIT, (uint32_t)sym, 0, 0);
}

// lens[0..codes) are code lengths (0..15).  tbl has capacity `cap` entries.
// work: scratch for `codes` u16.  Single-thread routine.
__device__ inline int build_decode_table(int kind, const uint8_t *lens, int codes, uint32_t *tbl, int cap,
                                         uint16_t *work, TableInfo *info) {
  int bl_count[16];
  int offs[16];
#pragma unroll
  for (int i = 0; i < 16; i++) bl_count[i] = 0;
  for (int s = 0; s < codes; s++) bl_count[lens[s]]++;
  int maxl = 15;
  while (maxl >= 1 && bl_count[maxl] == 0) maxl--;
  if (maxl == 0) {
    // empty_table (de.ml:521): one entry {len 1, value 0} indexed with mask 1; index 1 is
    // out of bounds in the reference (uncaught exception).
    tbl[0] = symbol_entry(kind, 0) | 1u;
    tbl[1] = mk_entry(K_INVALID, 0, 0, 0);
    info->root = 1;
    info->maxl = 1;
    return TBL_OK;
  }
  int left = 1;
  for (int i = 1; i <= 15; i++) {
    left = (left << 1) - bl_count[i];
    if (left < 0) return TBL_INVALID;
  }
  if (left > 0 && (kind == KIND_CODES || maxl != 1)) return TBL_INVALID;
  int minl = 1;
  while (minl <= 15 && bl_count[minl] == 0) minl++;
  offs[0] = 0;
  offs[1] = 0;
  for (int i = 1; i <= 14; i++) offs[i + 1] = offs[i] + bl_count[i];
  for (int s = 0; s < codes; s++) {
    int l = lens[s];
    if (l != 0) work[offs[l]++] = (uint16_t)s;
  }
  int root = kind == KIND_LENS ? 9 : kind == KIND_DISTS ? 6 : 7;
  if (root > maxl) root = maxl;
  if (root < minl) root = minl;
  int size = (maxl <= root) ? (1 << maxl)
             : kind == KIND_LENS ? LIT_TABLE_SIZE
             : kind == KIND_DISTS ? DIST_TABLE_SIZE : (1 << maxl);
  if (size > cap) return TBL_EXCEPTION;
  // never-filled slots of an incomplete (single 1-bit code) set read as {len 0, value 0} (de.ml:582)
  uint32_t zero_entry = symbol_entry(kind, 0);
  for (int i = 0; i < size; i++) tbl[i] = zero_entry;

  int huff = 0, sym = 0, len = minl, next = 0, curr = root, drop = 0, low = -1;
  const int mask = (1 << root) - 1;
  bool finished = false;
  while (!finished) {
    uint32_t entry = symbol_entry(kind, work[sym]) | (uint32_t)len;
    int step = 1 << (len - drop);
    int fill_size = 1 << curr;
    int fill = fill_size;
    int at0 = next + (huff >> drop);
    do {
      fill -= step;
      int at = at0 + fill;
      if (at >= size) return TBL_EXCEPTION;
      tbl[at] = entry;
    } while (fill != 0);
    int inc = 1 << (len - 1);
    while (huff & inc) inc >>= 1;
    huff = inc != 0 ? (huff & (inc - 1)) + inc : 0;
    sym++;
    if (--bl_count[len] == 0) {
      if (len == maxl) finished = true;
      else len = lens[work[sym]];
    }
    if (!finished && len > root && (huff & mask) != low) {
      if (drop == 0) drop = root;
      next += fill_size;
      curr = len - drop;
      int l = 1 << curr;
      while (curr + drop < maxl) {
        l -= bl_count[curr + drop];
        if (l <= 0) break;
        curr++;
        l <<= 1;
      }
      low = huff & mask;
      if (low >= size) return TBL_EXCEPTION;
      tbl[low] = mk_entry(K_LINK, (uint32_t)next, (uint32_t)curr, 0);
    }
  }
  info->root = root;
  info->maxl = maxl;
  return TBL_OK;
}

} // namespace dcmp
