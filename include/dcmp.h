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

#define DCMP_ABI_VERSION 2

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

/* OR-ed into DCMP_FMT_GZIP for inflate: XLEN of the FEXTRA field is little-endian, as RFC1952 and every other tool (gzip,
 * bgzip ...) write it; without it the reference's big-endian reading is kept (lib/gz.ml:455). */
#define DCMP_GZIP_RFC_XLEN 0x200

/* OR-ed into `fmt` of dcmp_deflate_batch_device (DCMP_FMT_RAW only): no block carries BFINAL and the output ends on a byte
 * boundary — one segment of a stream the caller closes itself (what the streaming encoders below do). */
#define DCMP_DEFLATE_NO_FINAL 0x100

typedef struct dcmp_engine dcmp_engine;

/* Streams: the *_batch_device entry points enqueue on the caller's CUDA stream and do not wait for the kernels.  They
 * share per-engine scratch, so consecutive calls — on the same stream or on different ones — are ordered on the device
 * one after the other (an event per engine); use several engines for concurrent batches. */

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
int dcmp_engine_device(const dcmp_engine *e);   /* the CUDA device ordinal the engine was created on */
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
 * DCMP_INVALID_COMPRESSION_LEVEL, lib/de.ml:3930).  fmt: DCMP_FMT_RAW, DCMP_FMT_ZLIB (header lib/zl.ml:511-522, Adler-32
 * trailer) or DCMP_FMT_GZIP (Gz.Def's member: header lib/gz.ml:794-813 with FLG = 0, MTIME = 0, OS = Unix; CRC-32 + ISIZE
 * trailer lib/gz.ml:725-733; both written on the device).  written[k] = bytes of the deflate (or zlib / gzip)
 * stream; a dst_cap[k] that is too small gives DCMP_UNEXPECTED_END_OF_OUTPUT (lib/de.ml:3061-3076);
 * dcmp_deflate_bound(len, fmt) is always enough.  The compressed bytes are NOT those of the
 * reference (its parse is sequential, lib/de.ml:4351-4410): the contract is that De.Inf and system
 * zlib inflate them back to the input. */
int64_t dcmp_deflate_bound(int64_t len, int fmt); /* De.Def.Ns.compress_bound, lib/de.ml:3993-3996 (+6: lib/zl.ml:598; +18 gzip) */
int dcmp_deflate_batch(dcmp_engine *e, int level, int fmt, int64_t n,
                       const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                       uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                       int64_t *written, int32_t *status);
/* Data pointers are DEVICE pointers, the four offset tables are HOST arrays (the call cuts the inputs
 * into chunks on the host and uploads the chunk table from pinned staging the engine owns); d_written / d_status are
 * device arrays.  Does not wait for `cuda_stream`: synchronise it (or order later work behind it) before reading them.
 * The host tables may be reused as soon as the call returns. */
int dcmp_deflate_batch_device(dcmp_engine *e, int level, int fmt, int64_t n,
                              const uint8_t *d_src_base, const int64_t *h_src_off, const int64_t *h_src_len,
                              uint8_t *d_dst_base, const int64_t *h_dst_off, const int64_t *h_dst_cap,
                              int64_t *d_written, int32_t *d_status, void *cuda_stream);
/* Single-call mirrors of De.Def.Ns.deflate / Zl.Def.Ns.deflate: return the status, *written = `Ok len`. */
int dcmp_def_ns_deflate(dcmp_engine *e, int level, const uint8_t *src, int64_t src_len,
                        uint8_t *dst, int64_t dst_len, int64_t *written);
int dcmp_zl_def_ns_deflate(dcmp_engine *e, int level, const uint8_t *src, int64_t src_len,
                           uint8_t *dst, int64_t dst_len, int64_t *written);

/* ---- the command-level encoder: De.Queue -> De.Def, De.T (lib/de.mli:188-263, 300-445, 677-682) -------------------------
 * dcmp_def_encode_queue: what De.Def.encode does with a De.Queue and a Dynamic block (lib/de.ml:2387-2407, 2583-2897): one
 * Dynamic-Huffman block from n commands in the Queue's own encoding (lib/de.ml:2245-2266: literal = the byte, end of
 * block = 256 (implied at the end), copy = 0x2000000 | (len - 3) << 16 | (off - 1)); BFINAL = last.  Frequencies are
 * counted from the commands (succ_literal / succ_length / succ_distance, lib/de.ml:2333-2351); trees, header and codes
 * are the reference's bit for bit (test/test.ml:507-531).  Returns the status; *written = the block's bytes.
 * dcmp_t_make: De.T.make ~length ?max_length freqs (lib/de.ml:2013-2068): per symbol the code length and the bit-reversed
 * canonical code — zlib's tree (test/test.ml:1169-1237). */
int dcmp_def_encode_queue(dcmp_engine *e, const int32_t *cmds, int64_t n, int last, uint8_t *dst, int64_t dst_len, int64_t *written);
int dcmp_t_make(dcmp_engine *e, int length, int max_length, const int32_t *freqs, int32_t *lengths, int32_t *codes, int32_t *max_code);

/* ---- streaming decoders ---------------------------------------------------------
 * The non-blocking half of the reference's interface: De.Inf.decoder / decode / src / flush / dst_rem / src_rem /
 * checksum / reset (lib/de.mli:82-148, lib/de.ml:651-753, 1427-1532) for fmt = DCMP_FMT_RAW, Zl.Inf (lib/zl.mli:30-124,
 * lib/zl.ml:141-309) for DCMP_FMT_ZLIB, Gz.Inf (lib/gz.mli:56-149, lib/gz.ml:248-634) for DCMP_FMT_GZIP.  A decoder is
 * a host state machine over a resumable device kernel (csrc/inflate_serial.cu: inflate_stream_kernel): `decode` runs it
 * on the input handed over so far, so output appears before the end of the input and memory is bounded — one
 * undecodable tail of input (a block header, a token), one launch's worth of output (<= 1 MiB) and the 32 KiB window.
 * Protocol, as in the reference (README.md:93-109):
 *   DCMP_AWAIT      the input is used up: dcmp_inf_src(d, buf, len), len = 0 for the end of the input (de.ml:739)
 *   DCMP_FLUSH      `o` holds o_len - dcmp_inf_dst_rem(d) bytes; take them and call dcmp_inf_flush(d).  For zlib / gzip
 *                   the answer repeats until flush is called (lib/zl.ml:288-291, test/test.ml:1536-1568), and the bytes
 *                   that sit in `o` when the DEFLATE data ends leave with a DCMP_FLUSH before DCMP_END (zl.ml:295-306)
 *   DCMP_END        the stream is complete (trailer checked); `o` holds the last o_len - dst_rem bytes (raw) or nothing
 *   DCMP_MALFORMED  dcmp_inf_status(d) is the DCMP_* code, dcmp_inf_error(d) the reference's message ("Invalid
 *                   checksum (expect:…, has:…)", "Invalid Zlib header", … lib/zl.ml:177-183, lib/gz.ml:285-296); the
 *                   decoder then answers DCMP_END for ever (continuation `final`, de.ml:700)
 * `buf` of dcmp_inf_src must stay valid until the next DCMP_AWAIT.  dcmp_inf_src_rem at DCMP_END = bytes of the last
 * input that follow the stream (Gz.Inf.src_rem, gz.ml:550: where a next member would start).  dcmp_inf_checksum:
 * Adler-32 (CRC-32 for gzip) of everything written to `o`; dcmp_inf_write: their count (Zl.Inf.write). */
enum { DCMP_AWAIT = 0, DCMP_FLUSH = 1, DCMP_END = 2, DCMP_MALFORMED = 3 };
typedef struct dcmp_inf dcmp_inf;
int dcmp_inf_new(dcmp_engine *e, int fmt, uint8_t *o, int64_t o_len, dcmp_inf **out);
void dcmp_inf_free(dcmp_inf *d);
int dcmp_inf_src(dcmp_inf *d, const uint8_t *buf, int64_t len);
int dcmp_inf_dst(dcmp_inf *d, uint8_t *o, int64_t o_len);   /* a new output buffer (Zl.Inf has none: convenience) */
int dcmp_inf_decode(dcmp_inf *d);                           /* DCMP_AWAIT .. DCMP_MALFORMED, or DCMP_API_* < 0 */
int dcmp_inf_status(const dcmp_inf *d);
const char *dcmp_inf_error(const dcmp_inf *d);
int64_t dcmp_inf_dst_rem(const dcmp_inf *d);
int64_t dcmp_inf_src_rem(const dcmp_inf *d);
void dcmp_inf_flush(dcmp_inf *d);
void dcmp_inf_reset(dcmp_inf *d);
uint32_t dcmp_inf_checksum(const dcmp_inf *d);
int64_t dcmp_inf_write(const dcmp_inf *d);
int64_t dcmp_inf_launches(const dcmp_inf *d);               /* kernels this decoder has launched */
int64_t dcmp_inf_buffered(const dcmp_inf *d);               /* bytes it holds: undecoded input + undelivered output */
/* Gz.Inf.os / filename / comment / extra (lib/gz.ml:612-633).  which: 0 FNAME, 1 FCOMMENT, 2 the raw FEXTRA payload
 * (sub-fields: 2-byte key, big-endian 16-bit length, gz.ml:628); returns the length, -1 for None. */
int dcmp_gz_inf_os(const dcmp_inf *d);
uint32_t dcmp_gz_inf_mtime(const dcmp_inf *d);
int64_t dcmp_gz_inf_field(const dcmp_inf *d, int which, const uint8_t **p);

/* ---- streaming encoders ---------------------------------------------------------
 * Zl.Def.encoder / encode / src / dst (lib/zl.mli:141-263, lib/zl.ml:421-592), Gz.Def (lib/gz.mli:160-265,
 * lib/gz.ml:636-900) and, for DCMP_FMT_RAW, the pair De.Lz77.compress + De.Def.encode as De.Higher.compress glues
 * it (lib/de.ml:4517-4560).  Input is gathered into segments of 4 MiB that the batch deflate kernels compress
 * (every block BFINAL = 0 and a byte-aligned end, until the last segment); framing and its checksums are host work.
 *   DCMP_AWAIT  dcmp_def_src(d, buf, len), len = 0 for the end of the input
 *   DCMP_FLUSH  `o` is full, or the trailer has been written: take o_len - dcmp_def_dst_rem(d) bytes, then dcmp_def_dst
 *   DCMP_END    everything has been flushed (`o` is empty: test/test.ml:1597-1625)
 * level 0..9 (De.Lz77.state, de.ml:4431-4437).  dcmp_def_gz_header: Gz.Def.encoder's ?ascii ?hcrc ?filename ?comment
 * ~mtime os, before the first encode (header bytes as lib/gz.ml:794-813: MTIME big-endian, FHCRC = upper half of the
 * CRC-32 of the fixed bytes + FNAME + FCOMMENT). */
typedef struct dcmp_def dcmp_def;
int dcmp_def_new(dcmp_engine *e, int fmt, int level, dcmp_def **out);
void dcmp_def_free(dcmp_def *d);
int dcmp_def_gz_header(dcmp_def *d, uint32_t mtime, int os, int ascii, int hcrc, const uint8_t *filename,
                       int64_t filename_len, const uint8_t *comment, int64_t comment_len);
int dcmp_def_src(dcmp_def *d, const uint8_t *buf, int64_t len);
int dcmp_def_dst(dcmp_def *d, uint8_t *buf, int64_t len);
int dcmp_def_encode(dcmp_def *d);                           /* DCMP_AWAIT, DCMP_FLUSH, DCMP_END, or DCMP_API_* < 0 */
int64_t dcmp_def_src_rem(const dcmp_def *d);
int64_t dcmp_def_dst_rem(const dcmp_def *d);

/* ---- multi-member gzip without decoding: a BGZF index -------------------------------------------
 * Gz.Inf answers `End after ONE member and the reference has no way to find the next boundary without decoding (lib/gz.ml:550,
 * 582-610).  BGZF members carry their own size (FEXTRA sub-field 'B','C'): this walks them on the host and fills member_off /
 * member_len / isize (up to max_members); the members then go through dcmp_inflate_batch(DCMP_FMT_GZIP | DCMP_GZIP_RFC_XLEN)
 * as independent units.  Returns the number of members, -1 if the buffer is not a sequence of BGZF members. */
int64_t dcmp_gz_bgzf_index(const uint8_t *buf, int64_t len, int64_t *member_off, int64_t *member_len, int64_t *isize, int64_t max_members);

/* ---- caller buffers ----------------------------------------------------------------
 * The *_batch calls take host pointers.  A `bigstring` (Bigarray.Array1, lib/de.mli:45) is pageable malloc memory:
 * register long-lived io buffers once (cudaHostRegister) and the copies run at the PCIe rate, asynchronously. */
int dcmp_host_register(void *p, int64_t n);
int dcmp_host_unregister(void *p);

/* ---- checksums (host; checkseum's Adler32 / Crc32 as used at lib/zl.ml:231-245,
 * lib/gz.ml:344-356) ------------------------------------------------------------ */
uint32_t dcmp_adler32(uint32_t adler, const uint8_t *p, int64_t n);
uint32_t dcmp_crc32(uint32_t crc, const uint8_t *p, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* DCMP_H */
