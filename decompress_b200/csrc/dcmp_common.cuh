// decompress_b200/csrc/dcmp_common.cuh — shared device-side definitions.
//
// Decode-table entry format, DEFLATE symbol constants and the bit cursor used by
// every inflate kernel.  The *semantics* follow the reference (lib/de.ml), the
// layout is ours: one u32 per entry carrying everything the hot loop needs, so a
// symbol costs one shared-memory load and no follow-up constant-table lookups
// (the reference does table -> _base_length/_extra_lbits lookups, de.ml:1683-1698).
#pragma once
#include <stdint.h>
#include "../../include/dcmp.h"

namespace dcmp {

// ---- entry format -------------------------------------------------------------
//  [3:0]   code length in bits (full length, also for second-level entries; 0 for
//          the never-filled slots of an incomplete 1-bit code, de.ml:582)
//  [7:4]   number of extra bits (length: 0..5, distance: 0..13);
//          for LINK entries: index width of the sub-table
//  [23:8]  base: literal byte | match length base (3..258) | distance base
//          (1..24577, 0 = "Invalid distance code", de.ml:321-325) | sub-table offset
//  [27:24] kind
enum : uint32_t { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_LINK = 3, K_INVALID = 4 };
__host__ __device__ constexpr uint32_t mk_entry(uint32_t kind, uint32_t base, uint32_t extra, uint32_t len) {
  return (kind << 24) | (base << 8) | (extra << 4) | len;
}
__host__ __device__ inline uint32_t e_len(uint32_t e) { return e & 15u; }
__host__ __device__ inline uint32_t e_extra(uint32_t e) { return (e >> 4) & 15u; }
__host__ __device__ inline uint32_t e_base(uint32_t e) { return (e >> 8) & 0xffffu; }
__host__ __device__ inline uint32_t e_kind(uint32_t e) { return e >> 24; }

// Table capacities: the reference allocates exactly these (de.ml:575-581) and an
// index past them is an OCaml exception there (DCMP_REF_EXCEPTION here).
constexpr int LIT_TABLE_SIZE = 852;
constexpr int DIST_TABLE_SIZE = 592;
constexpr int CODES_TABLE_SIZE = 128;
constexpr int MAX_LENS = 320; // HLIT + HDIST <= 288 + 32

enum : int { KIND_CODES = 0, KIND_LENS = 1, KIND_DISTS = 2 };

#ifdef __CUDACC__
#define DCMP_CONST __device__ __constant__
#else
#define DCMP_CONST static const
#endif

// de.ml:293-325 (base_length is stored +3, base_dist +1; dist symbols 30/31 give 0)
DCMP_CONST uint16_t c_len_base[32] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                      67, 83, 99, 115, 131, 163, 195, 227, 258, 3, 3, 3};
DCMP_CONST uint8_t c_len_extra[32] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                      3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0};
DCMP_CONST uint16_t c_dist_base[32] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                       1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0};
DCMP_CONST uint8_t c_dist_extra[32] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
                                       7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 0, 0};
DCMP_CONST uint8_t c_zigzag[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15}; // de.ml:237

// Symbol -> entry (without the code length, OR-ed in by the table builder).
__device__ __forceinline__ uint32_t lit_symbol_entry(int sym) {
  if (sym < 256) return mk_entry(K_LIT, (uint32_t)sym, 0, 0);
  if (sym == 256) return mk_entry(K_EOB, 0, 0, 0);
  int l = (sym - 257) & 31; // 286/287 decode as a length-3 match (de.ml:293-311)
  return mk_entry(K_LEN, c_len_base[l], c_len_extra[l], 0);
}
__device__ __forceinline__ uint32_t dist_symbol_entry(int sym) {
  return mk_entry(K_LIT, c_dist_base[sym & 31], c_dist_extra[sym & 31], 0);
}

} // namespace dcmp
