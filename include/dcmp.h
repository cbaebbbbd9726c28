/* include/dcmp.h — C ABI of libdcmp_b200.so, the B200-native DEFLATE / LZO1X engine.
 *
 * This is the drop-in boundary for mirage/decompress's hot path.  The reference
 * has no FFI (it is pure OCaml behind lib/de.mli, lib/zl.mli, lib/gz.mli,
 * lib/lzo.mli); this header is what an OCaml shim (ocaml/de.ml, ocaml/zl.ml, ... + ocaml/dcmp_stubs.c,
 * see INTEGRATION.md) binds instead of lib/de.ml's bodies.  Each entry point
 * cites the reference interface it replaces (paths relative to the reference
 * checkout).
 *
 * Conventions: plain pointers and sizes only; every function returns an int
 * (DCMP_API_* for call-level failures, never throws); per-stream results come
 * back in `status[]` using the reference's error vocabulary.  All `*_batch`
 * calls take HOST pointers and perform the host<->device copies themselves;
 * `*_batch_device` calls take DEVICE pointers and a CUDA stream and copy
 * nothing.  There is no CPU fallback: without a CUDA device every compute entry
 * point returns DCMP_API_NO_DEVICE.
 */
#ifndef DCMP_H
#define DCMP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCMP_ABI_VERSION 1

/* ---- call-level return codes ------------------------------------------------ */
enum {
  DCMP_API_OK = 0,
  DCMP_API_NO_DEVICE = -1,     /* no CUDA device / driver */
  DCMP_API_CUDA_ERROR = -2,    /* see dcmp_last_error() */
  DCMP_API_INVALID_ARGUMENT = -3,
  DCMP_API_OUT_OF_MEMORY = -4
};

/* ---- per-stream status ------------------------------------------------------
 * 0..7: De.Inf.Ns result, in the order of lib/de.mli:150-157 (`Ok`, then the
 *       `error` constructors).  The streaming decoder's `Malformed strings
 *       (lib/de.ml:702-730) are dcmp_status_string() of the same codes.
 * 8:    inputs on which the reference raises an uncaught OCaml exception
 *       (index out of bounds on an empty/1-entry lookup table, lib/de.ml:521,
 *       640-647); reported as malformed here.
 * 9,10: Zl.Inf.Ns additions (lib/zl.ml:383); 11..13 Gz (lib/gz.ml:285-296).
 * 16..18: Lzo.error (lib/lzo.ml:4-12).
 * 32:   Def.Ns `Invalid_compression_level (lib/de.ml:3061). */
enum {
  DCMP_OK = 0,
  DCMP_UNEXPECTED_END_OF_INPUT = 1,
  DCMP_UNEXPECTED_END_OF_OUTPUT = 2,
  DCMP_INVALID_KIND_OF_BLOCK = 3,
  DCMP_INVALID_DICTIONARY = 4,
  DCMP_INVALID_COMPLEMENT_OF_LENGTH = 5,
  DCMP_INVALID_DISTANCE = 6,
  DCMP_INVALID_DISTANCE_CODE = 7,
  DCMP_REF_EXCEPTION = 8,
  DCMP_INVALID_HEADER = 9,
  DCMP_INVALID_CHECKSUM = 10,
  DCMP_INVALID_GZIP_HEADER = 11,
  DCMP_INVALID_GZIP_HEADER_CHECKSUM = 12,
  DCMP_INVALID_INPUT_SIZE = 13,
  DCMP_LZO_MALFORMED = 16,
  DCMP_LZO_INVALID_ARGUMENT = 17,
  DCMP_LZO_INVALID_DICTIONARY = 18,
  DCMP_INVALID_COMPRESSION_LEVEL = 32
};

/* ---- stream formats --------------------------------------------------------- */
enum {
  DCMP_FMT_RAW = 0,   /* RFC1951, De.Inf.Ns.inflate (lib/de.ml:1807-1822) */
  DCMP_FMT_ZLIB = 1,  /* RFC1950, Zl.Inf.Ns.inflate (lib/zl.ml:400-417) */
  DCMP_FMT_GZIP = 2   /* RFC1952, one member per unit: Gz.Inf header + CRC-32/ISIZE trailer (lib/gz.ml:463-531, 344-356),
                         parsed and checked on the device */
};

typedef struct dcmp_engine dcmp_engine;

/* ---- engine ------------------------------------------------------------------
 * One engine = one CUDA device + one stream + reusable device/pinned scratch.
 * Not thread-safe: one owner at a time, like a De.Inf.decoder (README.md:111-145). */
int dcmp_abi_version(void);
int dcmp_device_count(void);
int dcmp_engine_create(int device, dcmp_engine **out);
void dcmp_engine_destroy(dcmp_engine *e);
const char *dcmp_last_error(void);
const char *dcmp_status_string(int status); /* lib/de.ml:1557-1567, lib/zl.ml:177-183 */
/* Number of kernels this engine has launched since creation (bench.py's gpu_launches). */
int64_t dcmp_engine_launch_count(const dcmp_engine *e);
/* Counters of the two-stage dispatch since the engine was created (synchronises the device).  The fast kernels
 * (one CTA per stream / block) hand what they cannot finish — malformed or truncated input, output windows that
 * are too small, exotic code sets — to the general kernels; a benchmark whose streams are all valid must see
 * zero of them redone.  out[0] inflate streams submitted, out[1] inflate streams redone by the general kernel,
 * out[2] LZO blocks submitted, out[3] LZO blocks redone, out[4] kernels launched. */
#define DCMP_STATS_COUNT 5
int dcmp_engine_stats(dcmp_engine *e, int64_t *out, int n_out);

/* ---- batched inflate ---------------------------------------------------------
 * Replaces De.Inf.Ns.inflate (lib/de.mli:160-172, lib/de.ml:1807-1822) /
 * Zl.Inf.Ns.inflate (lib/zl.ml:400-417) applied to n independent streams.
 * Stream k reads src_base[src_off[k] .. +src_len[k]) and writes
 * dst_base[dst_off[k] .. +dst_cap[k]).  Outputs per stream: consumed[k],
 * produced[k] (the `Ok (i, o)` pair) and status[k].  For DCMP_FMT_ZLIB the
 * header and Adler-32 trailer are checked (trailer on the device), consumed
 * counts the 2 + 4 framing bytes. */
int dcmp_inflate_batch(dcmp_engine *e, int fmt, int64_t n,
                       const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                       uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                       int64_t *consumed, int64_t *produced, int32_t *status);

/* Same, all pointers are device pointers (offset tables included); enqueued on
 * `cuda_stream` (a cudaStream_t; NULL = the legacy default stream, as in CUDA);
 * returns without synchronising. */
int dcmp_inflate_batch_device(dcmp_engine *e, int fmt, int64_t n,
                              const uint8_t *d_src_base, const int64_t *d_src_off, const int64_t *d_src_len,
                              uint8_t *d_dst_base, const int64_t *d_dst_off, const int64_t *d_dst_cap,
                              int64_t *d_consumed, int64_t *d_produced, int32_t *d_status,
                              void *cuda_stream);

/* Single-call mirror of De.Inf.Ns.inflate (lib/de.ml:1807-1822): host buffers,
 * returns the status (0 = Ok) and fills consumed/produced. */
int dcmp_inf_ns_inflate(dcmp_engine *e, const uint8_t *src, int64_t src_len,
                        uint8_t *dst, int64_t dst_len, int64_t *consumed, int64_t *produced);
/* Zl.Inf.Ns.inflate (lib/zl.ml:400-417). */
int dcmp_zl_inf_ns_inflate(dcmp_engine *e, const uint8_t *src, int64_t src_len,
                           uint8_t *dst, int64_t dst_len, int64_t *consumed, int64_t *produced);

/* ---- batched LZO1X decode ---------------------------------------------------
 * Replaces Lzo.uncompress (lib/lzo.mli:15-27, lib/lzo.ml:395-403; mode 0) and
 * Lzo.uncompress_with_buffer (lib/lzo.mli:29-33, lib/lzo.ml:405-414; mode 1) applied to n independent
 * blocks.  Same slab/offset-table layout as dcmp_inflate_batch.  status[k] is DCMP_OK or one of
 * DCMP_LZO_MALFORMED / DCMP_LZO_INVALID_ARGUMENT / DCMP_LZO_INVALID_DICTIONARY; the two modes differ
 * only in how an out-of-bounds access is reported (lib/lzo.ml:399-401 vs 409-412); in mode 1 a full
 * dst is DCMP_UNEXPECTED_END_OF_OUTPUT (the reference's Buffer.t grows: enlarge dst and call again).  Back-references
 * replicate bytes (what uncompress_with_buffer, minilzo and every reference test pin). */
int dcmp_lzo_uncompress_batch(dcmp_engine *e, int mode, int64_t n,
                              const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                              uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                              int64_t *consumed, int64_t *produced, int32_t *status);
int dcmp_lzo_uncompress_batch_device(dcmp_engine *e, int mode, int64_t n,
                                     const uint8_t *d_src_base, const int64_t *d_src_off, const int64_t *d_src_len,
                                     uint8_t *d_dst_base, const int64_t *d_dst_off, const int64_t *d_dst_cap,
                                     int64_t *d_consumed, int64_t *d_produced, int32_t *d_status,
                                     void *cuda_stream);
/* Single-call mirror of Lzo.uncompress (lib/lzo.ml:395-403): returns the status, *produced = length of
 * the output prefix written. */
int dcmp_lzo_uncompress(dcmp_engine *e, const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len,
                        int64_t *produced);

/* ---- batched deflate ----------------------------------------------------------
 * Replaces De.Def.Ns.deflate (lib/de.mli:436-443, lib/de.ml:3998-4010) / Zl.Def.Ns.deflate
 * (lib/zl.ml:610-630) applied to n independent inputs, and — for one large input — the streaming
 * pair De.Lz77.compress + De.Def.encode (lib/de.ml:4294-4515, 2465-3038) as driven by
 * De.Higher.compress / Zl.Higher.compress (lib/de.ml:4517-4560, lib/zl.ml:634-648).
 * level: 0 = stored blocks only, 1..12 = compressed (out of range: every status is
 * DCMP_INVALID_COMPRESSION_LEVEL, lib/de.ml:3930).  written[k] = bytes of the deflate (or zlib)
 * stream; a dst_cap[k] that is too small gives DCMP_UNEXPECTED_END_OF_OUTPUT (lib/de.ml:3061-3076);
 * dcmp_deflate_bound(len, fmt) is always enough.  The compressed bytes are NOT those of the
 * reference (its parse is sequential, lib/de.ml:4351-4410): the contract is that De.Inf and system
 * zlib inflate them back to the input. */
int64_t dcmp_deflate_bound(int64_t len, int fmt); /* De.Def.Ns.compress_bound, lib/de.ml:3993-3996 (+6: lib/zl.ml:598) */
int dcmp_deflate_batch(dcmp_engine *e, int level, int fmt, int64_t n,
                       const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                       uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                       int64_t *written, int32_t *status);
/* Data pointers are DEVICE pointers, the four offset tables are HOST arrays (the call cuts the inputs
 * into chunks on the host); d_written / d_status are device arrays.  Synchronises `cuda_stream` once,
 * after the kernels are enqueued. */
int dcmp_deflate_batch_device(dcmp_engine *e, int level, int fmt, int64_t n,
                              const uint8_t *d_src_base, const int64_t *h_src_off, const int64_t *h_src_len,
                              uint8_t *d_dst_base, const int64_t *h_dst_off, const int64_t *h_dst_cap,
                              int64_t *d_written, int32_t *d_status, void *cuda_stream);
/* Single-call mirrors of De.Def.Ns.deflate / Zl.Def.Ns.deflate: return the status, *written = `Ok len`. */
int dcmp_def_ns_deflate(dcmp_engine *e, int level, const uint8_t *src, int64_t src_len,
                        uint8_t *dst, int64_t dst_len, int64_t *written);
int dcmp_zl_def_ns_deflate(dcmp_engine *e, int level, const uint8_t *src, int64_t src_len,
                           uint8_t *dst, int64_t dst_len, int64_t *written);

/* ---- checksums (host; checkseum's Adler32 / Crc32 as used at lib/zl.ml:231-245,
 * lib/gz.ml:344-356) ------------------------------------------------------------ */
uint32_t dcmp_adler32(uint32_t adler, const uint8_t *p, int64_t n);
uint32_t dcmp_crc32(uint32_t crc, const uint8_t *p, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* DCMP_H */
