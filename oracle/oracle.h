/* oracle/oracle.h — CPU restatement of mirage/decompress's DEFLATE / LZO path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported
 * by, or executed from the product (decompress_b200/, libdcmp_b200.so).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use it, and there only as the checker / CPU baseline.
 *
 * Parity pin: the reference is OCaml and cannot be built in this image (no
 * ocaml/dune/opam).  The restatement is pinned against every golden vector
 * the reference's own tests hold for this path (the JSON files under tests/golden/, made by
 * tests/golden/make_golden.py from test/test.ml, test/test_ns.ml,
 * test/test_lzo.ml) and against the reference's own differential oracles:
 * system zlib (fuzz/fuzz.ml:82-88) and vendored minilzo (fuzz/fuzz_lzo.ml),
 * the latter compiled from /root/reference into oracle/_ref/.
 *
 * Every function cites the reference lines (relative to /root/reference) it
 * restates.
 */
#ifndef DCMP_ORACLE_H
#define DCMP_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: 0 = Ok, 1..7 = De.Inf.Ns.error in the order of lib/de.mli:151-157.
 * 8 = the reference would raise an uncaught OCaml exception here (array index
 * out of bounds on an empty / 1-entry lookup table, lib/de.ml:521,640-647);
 * the build treats it as malformed. */
enum {
  ORC_OK = 0,
  ORC_UNEXPECTED_END_OF_INPUT = 1,
  ORC_UNEXPECTED_END_OF_OUTPUT = 2,
  ORC_INVALID_KIND_OF_BLOCK = 3,
  ORC_INVALID_DICTIONARY = 4,
  ORC_INVALID_COMPLEMENT_OF_LENGTH = 5,
  ORC_INVALID_DISTANCE = 6,
  ORC_INVALID_DISTANCE_CODE = 7,
  ORC_REF_EXCEPTION = 8
};

/* De.Inf.Ns.inflate (lib/de.ml:1807-1822). consumed/produced are written on
 * ORC_OK; on error *produced holds the bytes written before the error. */
int orc_inflate_ns(const uint8_t *src, int64_t src_len, uint8_t *dst,
                   int64_t dst_len, int64_t *consumed, int64_t *produced);

/* Inf.huffman (lib/de.ml:523-638).  kind: 0 = CODES, 1 = LENS, 2 = DISTS.
 * tbl must hold ORC_TBL_CAP ints.  Returns 0, or -1 for Invalid_huffman, or -2
 * if the reference would index past its 852/592-entry array. */
#define ORC_TBL_CAP 4096
int orc_huffman(int kind, const int *lengths, int off, int codes, int32_t *tbl,
                int *size, int *root, int *maxlen);

/* Zl.Inf.Ns.inflate (lib/zl.ml:400-417): 9 = Invalid_header, 10 = Invalid_checksum */
enum { ORC_ZL_INVALID_HEADER = 9, ORC_ZL_INVALID_CHECKSUM = 10 };
int orc_zl_inflate_ns(const uint8_t *src, int64_t src_len, uint8_t *dst,
                      int64_t dst_len, int64_t *consumed, int64_t *produced);

/* Adler-32 / CRC-32 (checkseum; standard definitions, see oracle/README.md). */
uint32_t orc_adler32(uint32_t adler, const uint8_t *p, int64_t n);
uint32_t orc_crc32(uint32_t crc, const uint8_t *p, int64_t n);

/* Lzo.uncompress_with_buffer semantics (lib/lzo.ml:146-414), byte-replicating
 * copies.  Returns 0 ok, 1 = Malformed, 2 = Invalid_argument (output too
 * small / out-of-bounds), 3 = Invalid_dictionary. */
enum { ORC_LZO_OK = 0, ORC_LZO_MALFORMED = 1, ORC_LZO_INVALID_ARGUMENT = 2,
       ORC_LZO_INVALID_DICTIONARY = 3 };
int orc_lzo_uncompress(const uint8_t *src, int64_t src_len, uint8_t *dst,
                       int64_t dst_len, int64_t *consumed, int64_t *produced);

int orc_lzo_uncompress_with_buffer(const uint8_t *src, int64_t src_len, uint8_t *dst,
                                   int64_t dst_len, int64_t *consumed, int64_t *produced);

/* deflate_t.c: De.T.make (de.ml:2013-2068) and one Dynamic block of Def over Queue commands (de.ml:2387-2407, 2583-2897) */
int orc_t_make(int length, int max_length, int *freqs, int *tree_lengths, int *tree_codes, int *bl_count);
int64_t orc_def_encode_dynamic(const int32_t *cmds, int64_t n, int last, uint8_t *out, int64_t cap, int *h);

#ifdef __cplusplus
}
#endif
#endif
