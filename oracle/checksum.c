/* oracle/checksum.c — Adler-32 and CRC-32 (IEEE 802.3, reflected 0xEDB88320).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * The reference takes both from the un-vendored `checkseum` package
 * (decompress.opam:23, `checkseum >= 0.2.0`, no lock file => no pinned
 * version).  checkseum implements the standard RFC1950 Adler-32 and the
 * RFC1952 / ISO 3309 CRC-32; the algorithm is restated from those RFCs.  Call
 * sites in the reference: lib/de.ml:443-455,502 (window Adler), lib/zl.ml:231-245,
 * 413-415 (trailer check), lib/gz.ml:344-356,503-513 (CRC-32 + ISIZE).  Pinned
 * by the zlib/gzip goldens (test/test.ml:1455-1457, 1703-1704) and against
 * libz's adler32()/crc32() in tests/test_oracle.py.
 */
#include "oracle.h"

uint32_t orc_adler32(uint32_t adler, const uint8_t *p, int64_t n) {
  uint32_t a = adler & 0xffff, b = (adler >> 16) & 0xffff;
  while (n > 0) {
    int64_t k = n < 5552 ? n : 5552; /* largest run before 32-bit overflow */
    n -= k;
    while (k--) { a += *p++; b += a; }
    a %= 65521; b %= 65521;
  }
  return (b << 16) | a;
}

static uint32_t crc_table[256];
static int crc_ready;

static void crc_init(void) {
  uint32_t n, c; int k;
  for (n = 0; n < 256; n++) {
    c = n;
    for (k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
    crc_table[n] = c;
  }
  crc_ready = 1;
}

uint32_t orc_crc32(uint32_t crc, const uint8_t *p, int64_t n) {
  uint32_t c = crc ^ 0xffffffffu;
  if (!crc_ready) crc_init();
  while (n--) c = crc_table[(c ^ *p++) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

/* Zl.Inf.Ns.inflate, lib/zl.ml:391-417 */
int orc_zl_inflate_ns(const uint8_t *src, int64_t src_len, uint8_t *dst,
                      int64_t dst_len, int64_t *consumed, int64_t *produced) {
  int64_t i = 0, o = 0; int rc; uint32_t cmf, want, got;
  *consumed = 0; *produced = 0;
  if (src_len < 2) return ORC_UNEXPECTED_END_OF_INPUT;
  cmf = src[0] | ((uint32_t)src[1] << 8); /* unsafe_get_uint16_le */
  if ((((cmf & 0xff) << 8) + (cmf >> 8)) % 31 != 0 || (cmf & 0xf) != 8)
    return ORC_ZL_INVALID_HEADER;
  if (src_len < 6) return ORC_REF_EXCEPTION; /* bigstring_sub src 2 (len - 6): Invalid_argument */
  rc = orc_inflate_ns(src + 2, src_len - 6, dst, dst_len, &i, &o);
  *produced = o;
  if (rc != ORC_OK) return rc;
  if (src_len < i + 6) return ORC_UNEXPECTED_END_OF_INPUT;
  if (i < -2) return ORC_REF_EXCEPTION; /* negative index after the 63-bit `lsr` quirk */
  want = ((uint32_t)src[i + 2] << 24) | ((uint32_t)src[i + 3] << 16) |
         ((uint32_t)src[i + 4] << 8) | src[i + 5];
  got = orc_adler32(1, dst, o);
  if (want != got) return ORC_ZL_INVALID_CHECKSUM;
  *consumed = i + 6;
  return ORC_OK;
}
