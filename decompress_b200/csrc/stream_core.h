// decompress_b200/csrc/stream_core.h — the non-blocking half of the reference's interface as host state machines:
//   InfCore  = De.Inf.decoder (lib/de.ml:651-753, 1427-1532), Zl.Inf (lib/zl.ml:141-309), Gz.Inf (lib/gz.ml:248-634)
//   DefCore  = De.Lz77.compress + De.Def.encode as glued by De.Higher.compress (lib/de.ml:4517-4560), Zl.Def
//              (lib/zl.ml:421-592), Gz.Def (lib/gz.ml:636-900)
// `src` / `dst` hand buffers over, `decode` / `encode` answer Await / Flush / End / Malformed, nothing blocks and
// memory is bounded: the decoder keeps at most one undecodable tail of input (a block header or a token), one
// launch's worth of output and the 32 KiB window (a device ring).  The codec itself is device code: the template
// parameter is the back end that runs it — CUDA in the product (csrc/stream.cu), the CPU SIMT emulator in the
// GPU-less tests (tests/simt_emu/emu_stream.cc).  Framing (zlib / gzip headers, trailers) and its checksums are host
// work, as in BASELINE.json's north_star ("Zl/Gz framing and checkseum stay on the host").
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/dcmp.h"
#include "inflate_stream.h"
#include "status_strings.h"

namespace dcmp {

// ---- host checksums (Checkseum.Adler32 / Crc32 at lib/zl.ml:231-245, lib/gz.ml:344-356, 503-513) -------------------
inline uint32_t host_adler32(uint32_t adler, const uint8_t *p, int64_t n) {
  uint32_t a = adler & 0xffff, b = (adler >> 16) & 0xffff;
  while (n > 0) {
    int64_t k = n < 5552 ? n : 5552;
    n -= k;
    while (k--) { a += *p++; b += a; }
    a %= 65521; b %= 65521;
  }
  return (b << 16) | a;
}
inline uint32_t host_crc32(uint32_t crc, const uint8_t *p, int64_t n) {
  static uint32_t T[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
      T[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
      for (int t = 1; t < 8; t++) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xff];
    ready = true;
  }
  uint32_t c = crc ^ 0xffffffffu;
  while (n >= 8) {
    uint32_t a, b;
    memcpy(&a, p, 4);
    memcpy(&b, p + 4, 4);
    a ^= c;
    c = T[7][a & 0xff] ^ T[6][(a >> 8) & 0xff] ^ T[5][(a >> 16) & 0xff] ^ T[4][a >> 24] ^ T[3][b & 0xff] ^
        T[2][(b >> 8) & 0xff] ^ T[1][(b >> 16) & 0xff] ^ T[0][b >> 24];
    p += 8;
    n -= 8;
  }
  while (n-- > 0) c = T[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

// ---- gzip header, as Gz.Inf parses it (lib/gz.ml:358-492) and Gz.Def writes it (lib/gz.ml:735-813) ------------------
struct GzHeader {
  int cm = 8, flg = 0, xfl = 0, os = 255;
  uint32_t mtime = 0;              // read and written BIG-endian by the reference (gz.ml:476, 797)
  bool has_extra = false, has_name = false, has_comment = false;
  std::string extra, name, comment;
  // The reference's FHCRC covers a header it re-creates from the parsed fields — the ten fixed bytes, FNAME, FCOMMENT,
  // never FEXTRA — and is the UPPER half of that CRC-32, big-endian (gz.ml:413-444, 750-776).
  uint32_t hcrc16() const {
    std::string h;
    const uint8_t fixed[10] = {0x1f, 0x8b, (uint8_t)cm, (uint8_t)flg, (uint8_t)(mtime >> 24), (uint8_t)(mtime >> 16),
                               (uint8_t)(mtime >> 8), (uint8_t)mtime, (uint8_t)xfl, (uint8_t)os};
    h.assign((const char *)fixed, 10);
    if (has_name) { h += name; h.push_back('\0'); }
    if (has_comment) { h += comment; h.push_back('\0'); }
    return host_crc32(0, (const uint8_t *)h.data(), (int64_t)h.size()) >> 16;
  }
  std::string bytes() const { // Gz.Def.encode, state Hd
    std::string h;
    const uint8_t fixed[10] = {0x1f, 0x8b, 8, (uint8_t)flg, (uint8_t)(mtime >> 24), (uint8_t)(mtime >> 16),
                               (uint8_t)(mtime >> 8), (uint8_t)mtime, (uint8_t)xfl, (uint8_t)os};
    h.assign((const char *)fixed, 10);
    if (has_name) { h += name; h.push_back('\0'); }
    if (has_comment) { h += comment; h.push_back('\0'); }
    if (flg & 2) { const uint32_t c = hcrc16(); h.push_back((char)(c >> 8)); h.push_back((char)c); }
    return h;
  }
};
enum { HDR_MORE = -1, HDR_OK = 0 }; // or a DCMP_* status
// Parse a gzip header from p[0..n): HDR_OK and *used, HDR_MORE (the header is longer than n), or the error.
inline int gz_parse_header(const uint8_t *p, int64_t n, GzHeader &h, int64_t *used) {
  if (n < 10) return HDR_MORE;
  if (p[0] != 0x1f || p[1] != 0x8b) return DCMP_INVALID_GZIP_HEADER; // (only the magic is checked, gz.ml:474)
  h.cm = p[2]; h.flg = p[3];
  h.mtime = ((uint32_t)p[4] << 24) | ((uint32_t)p[5] << 16) | ((uint32_t)p[6] << 8) | p[7];
  h.xfl = p[8]; h.os = p[9];
  int64_t i = 10;
  if (h.flg & 4) {
    if (n - i < 2) return HDR_MORE;
    const int64_t xlen = ((int64_t)p[i] << 8) | p[i + 1]; // big-endian (gz.ml:455), contrary to RFC1952
    i += 2;
    if (n - i < xlen) return HDR_MORE;
    h.has_extra = true;
    h.extra.assign((const char *)p + i, (size_t)xlen);
    i += xlen;
  }
  for (int bit = 8; bit <= 16; bit <<= 1) {
    if (!(h.flg & bit)) continue;
    int64_t j = i;
    while (j < n && p[j] != 0) j++;
    if (j >= n) return HDR_MORE;
    (bit == 8 ? h.name : h.comment).assign((const char *)p + i, (size_t)(j - i));
    (bit == 8 ? h.has_name : h.has_comment) = true;
    i = j + 1;
  }
  if (h.flg & 2) {
    if (n - i < 2) return HDR_MORE;
    const uint32_t have = ((uint32_t)p[i] << 8) | p[i + 1];
    if (h.hcrc16() != have) return DCMP_INVALID_GZIP_HEADER_CHECKSUM;
    i += 2;
  }
  *used = i;
  return HDR_OK;
}
// Zl.Def's two header bytes (lib/zl.ml:511-522)
inline uint32_t zlib_header(int level) {
  uint32_t hdr = (8u + (7u << 4)) << 8;
  const uint32_t fl = level == 0 ? 0u : level <= 5 ? 1u : level == 6 ? 2u : 3u;
  hdr |= fl << 6;
  hdr += 31u - (hdr % 31u);
  return hdr;
}

// =====================================================================================================================
// Decoder.  Backend: int run(const uint8_t *in, int64_t n, bool eof, InfStreamState &st, std::vector<uint8_t> &out)
// uploads the n input bytes, runs ONE launch of inflate_stream_kernel from `st`, brings the new state back and appends
// the bytes the launch produced to `out`; returns DCMP_API_*.  void reset() forgets the window.
template <class Backend>
struct InfCore {
  Backend be;
  int fmt;
  uint8_t *o = nullptr;            // the caller's output buffer (`o` of De.Inf.decoder)
  int64_t o_len = 0, o_pos = 0;
  const uint8_t *i = nullptr;      // the caller's current input (De.Inf.src)
  int64_t i_len = 0, i_pos = 0;
  bool eoi = false;                // `src d buf off 0` was seen (de.ml:739)
  // framing
  enum { F_HEADER, F_BODY, F_TRAILER, F_END, F_STUCK } fs = F_BODY;
  std::vector<uint8_t> hbuf;       // header / trailer bytes gathered so far
  GzHeader gz;
  bool f = false;                  // Zl / Gz: `Flush was answered and flush not called yet (zl.ml:288-291)
  // body
  std::vector<uint8_t> pend;       // input the kernel has not consumed yet
  bool fresh = false;              // pend got bytes (or the end of input) the kernel has not seen
  bool eof_seen = false;           // a launch ran with eof set
  InfStreamState st;
  std::vector<uint8_t> outq;       // produced, not yet in `o`
  size_t outq_pos = 0;
  bool body_end = false;           // the final block has been decoded
  int body_err = 0;                // DCMP_* that ends the body
  int64_t tail = 0;                // De level: bytes of the current input after the end of the stream (src_rem at `End)
  uint32_t adler = 1, crc = 0;     // of the bytes written to `o` so far
  int64_t wr = 0;                  // bytes written to `o` so far (Zl.Inf.write / Gz.Inf.write)
  int status = 0;                  // DCMP_* of `Malformed
  char msg[160] = "";

  template <class... A>
  explicit InfCore(int fmt_, A &&...a) : be(static_cast<A &&>(a)...), fmt(fmt_) { reset_state(); }
  void reset_state() {
    memset(&st, 0, sizeof st);
    st.phase = SP_HEADER;
    pend.clear(); outq.clear(); hbuf.clear();
    outq_pos = 0; fresh = false; eof_seen = false; body_end = false; body_err = 0; tail = 0;
    adler = 1; crc = 0; wr = 0; status = 0; msg[0] = 0; f = false;
    fs = fmt == DCMP_FMT_RAW ? F_BODY : F_HEADER;
    gz = GzHeader();
    o_pos = 0;
    i = nullptr; i_len = i_pos = 0; eoi = false;
  }
  void reset() { be.reset(); reset_state(); }                      // De.Inf.reset (de.ml:1513-1532)
  void set_dst(uint8_t *buf, int64_t len) { o = buf; o_len = len; o_pos = 0; }
  void src(const uint8_t *buf, int64_t len) {                      // De.Inf.src (de.ml:733-742)
    if (len == 0) { eoi = true; i = nullptr; i_len = i_pos = 0; fresh = true; }
    else { i = buf; i_len = len; i_pos = 0; }
  }
  int64_t dst_rem() const { return o_len - o_pos; }
  int64_t src_rem() const { return fs == F_END || body_end ? tail : i_len - i_pos; }
  void flush() { o_pos = 0; f = false; }
  uint32_t checksum() const { return fmt == DCMP_FMT_GZIP ? crc : adler; }

  int malformed(int code, const char *text = nullptr) {
    status = code;
    snprintf(msg, sizeof msg, "%s", text ? text : status_string(code));
    fs = F_STUCK; // the reference's continuation is `final` from here on: `End for ever (de.ml:700)
    return DCMP_MALFORMED;
  }
  // move what the current input still holds to `dst`
  void take_input(std::vector<uint8_t> &dst) {
    if (i_pos < i_len) { dst.insert(dst.end(), i + i_pos, i + i_len); i_pos = i_len; fresh = true; }
  }
  // bytes -> `o`, checksums on the way
  void deliver() {
    int64_t n = (int64_t)(outq.size() - outq_pos);
    if (n > o_len - o_pos) n = o_len - o_pos;
    if (n <= 0) return;
    memcpy(o + o_pos, outq.data() + outq_pos, (size_t)n);
    if (fmt == DCMP_FMT_GZIP) crc = host_crc32(crc, o + o_pos, n); else adler = host_adler32(adler, o + o_pos, n);
    o_pos += n; outq_pos += (size_t)n; wr += n;
    if (outq_pos == outq.size()) { outq.clear(); outq_pos = 0; }
  }
  // De.Inf.decode over the device kernel: DCMP_AWAIT / FLUSH / END / MALFORMED (status = body_err), or DCMP_API_* < 0
  int decode_body() {
    for (;;) {
      deliver();
      if (outq_pos < outq.size()) return DCMP_FLUSH;              // `o` is full and more is waiting (de.ml:1229)
      if (body_end) return DCMP_END;
      if (body_err) return DCMP_MALFORMED;
      take_input(pend);
      if (!fresh && !(eoi && !eof_seen)) return DCMP_AWAIT;        // the kernel has seen all of pend: more input, please
      fresh = false;
      if (eoi) eof_seen = true;
      const int rc = be.run(pend.data(), (int64_t)pend.size(), eoi, st, outq);
      if (rc != DCMP_API_OK) return rc;
      const int64_t used = st.result == SR_END ? (st.bitpos + 7) >> 3 : st.bitpos >> 3;
      pend.erase(pend.begin(), pend.begin() + (size_t)(used < (int64_t)pend.size() ? used : (int64_t)pend.size()));
      st.bitpos = st.result == SR_END ? 0 : st.bitpos & 7;
      if (st.result == SR_END) { body_end = true; tail = (int64_t)pend.size(); }
      else if (st.result == SR_ERROR) body_err = st.status;
      else if (st.result == SR_FULL) fresh = true;                 // (the launch stopped for its output budget only)
    }
  }
  // gather n header / trailer bytes: true when hbuf holds them
  bool gather(std::vector<uint8_t> &from, size_t n) {
    while (hbuf.size() < n && !from.empty()) { hbuf.push_back(from.front()); from.erase(from.begin()); }
    return hbuf.size() >= n;
  }

  int decode() {
    for (;;) {
      switch (fs) {
      case F_STUCK: return DCMP_END;
      case F_END: return DCMP_END;
      case F_HEADER: {
        take_input(pend);
        if (fmt == DCMP_FMT_ZLIB) { // lib/zl.ml:247-280
          if (!gather(pend, 2)) { if (eoi) return malformed(DCMP_UNEXPECTED_END_OF_INPUT); return DCMP_AWAIT; }
          const uint32_t cmf = hbuf[0], flg = hbuf[1];
          if (((cmf << 8) + flg) % 31 != 0 || (cmf & 15) != 8) return malformed(DCMP_INVALID_HEADER, "Invalid Zlib header");
        } else { // lib/gz.ml:463-492
          hbuf.insert(hbuf.end(), pend.begin(), pend.end());
          pend.clear();
          int64_t used = 0;
          GzHeader h;
          const int r = gz_parse_header(hbuf.data(), (int64_t)hbuf.size(), h, &used);
          if (r == HDR_MORE) { if (eoi) return malformed(DCMP_UNEXPECTED_END_OF_INPUT); return DCMP_AWAIT; }
          if (r == DCMP_INVALID_GZIP_HEADER) return malformed(r, "Invalid GZip header");
          if (r == DCMP_INVALID_GZIP_HEADER_CHECKSUM) return malformed(r, "Invalid GZip header checksum");
          gz = h;
          pend.assign(hbuf.begin() + used, hbuf.end());
        }
        hbuf.clear();
        fresh = true;
        fs = F_BODY;
        break;
      }
      case F_BODY: {
        if (fmt != DCMP_FMT_RAW && f) return DCMP_FLUSH;          // idempotent until flush (zl.ml:288-291, test.ml:1536-1568)
        const int r = decode_body();
        if (r < 0) return r;
        if (fmt == DCMP_FMT_RAW) {
          if (r == DCMP_MALFORMED) return malformed(body_err);
          if (r == DCMP_END) { fs = F_END; return DCMP_END; }
          return r;
        }
        if (r == DCMP_FLUSH) { f = true; return DCMP_FLUSH; }
        if (r == DCMP_AWAIT) return DCMP_AWAIT;
        if (r == DCMP_MALFORMED) return malformed(body_err);
        // `End of the DEFLATE data: what sits in `o` leaves first (zl.ml:295-306), then the trailer
        if (o_pos > 0) { f = true; return DCMP_FLUSH; }
        fs = F_TRAILER;
        hbuf.clear();
        break;
      }
      case F_TRAILER: {
        take_input(pend);
        const size_t need = fmt == DCMP_FMT_ZLIB ? 4 : 8;
        if (!gather(pend, need)) { if (eoi) return malformed(DCMP_UNEXPECTED_END_OF_INPUT); return DCMP_AWAIT; }
        tail = (int64_t)pend.size();
        if (fmt == DCMP_FMT_ZLIB) { // lib/zl.ml:231-245
          const uint32_t want = ((uint32_t)hbuf[0] << 24) | ((uint32_t)hbuf[1] << 16) | ((uint32_t)hbuf[2] << 8) | hbuf[3];
          if (want != adler) {
            char t[96];
            snprintf(t, sizeof t, "Invalid checksum (expect:%04x, has:%04x)", want, adler);
            return malformed(DCMP_INVALID_CHECKSUM, t);
          }
        } else { // lib/gz.ml:344-356: the checksum is compared first
          const uint32_t want = (uint32_t)hbuf[0] | ((uint32_t)hbuf[1] << 8) | ((uint32_t)hbuf[2] << 16) | ((uint32_t)hbuf[3] << 24);
          const uint32_t isize = (uint32_t)hbuf[4] | ((uint32_t)hbuf[5] << 8) | ((uint32_t)hbuf[6] << 16) | ((uint32_t)hbuf[7] << 24);
          char t[96];
          if (want != crc) {
            snprintf(t, sizeof t, "Invalid checksum (expect:%04x, has:%04x)", want, crc);
            return malformed(DCMP_INVALID_CHECKSUM, t);
          }
          if (isize != (uint32_t)wr) {
            snprintf(t, sizeof t, "Invalid input size (expect:%d, inflated:%d)", (int32_t)isize, (int32_t)(uint32_t)wr);
            return malformed(DCMP_INVALID_INPUT_SIZE, t);
          }
        }
        fs = F_END;
        return DCMP_END;
      }
      }
    }
  }
};

// =====================================================================================================================
// Encoder.  Backend: int deflate(int level, const uint8_t *in, int64_t n, bool final, std::vector<uint8_t> &out) deflates
// one segment on the device — every block BFINAL = 0 and a byte-aligned end unless `final` — and appends the bytes.
template <class Backend>
struct DefCore {
  Backend be;
  int fmt, level;
  GzHeader gz;                      // Gz.Def.encoder's ?ascii ?hcrc ?filename ?comment ~mtime os (gz.ml:859-900)
  const uint8_t *i = nullptr;
  int64_t i_len = 0, i_pos = 0;
  bool eoi = false;
  uint8_t *o = nullptr;
  int64_t o_len = 0, o_pos = 0;
  std::vector<uint8_t> seg;         // input gathered for the next segment
  std::vector<uint8_t> outq;
  size_t outq_pos = 0;
  bool header_done = false, body_done = false, ended = false;
  uint32_t adler = 1, crc = 0;
  int64_t rd = 0;
  int64_t seg_bytes;

  template <class... A>
  DefCore(int fmt_, int level_, int64_t seg_bytes_, A &&...a)
      : be(static_cast<A &&>(a)...), fmt(fmt_), level(level_), seg_bytes(seg_bytes_) {
    gz.xfl = level >= 9 ? 2 : 0;    // gz.ml:888-890
  }
  void src(const uint8_t *buf, int64_t len) { // Zl.Def.src / Gz.Def.src: checksums at hand-over (gz.ml:680-688)
    if (len == 0) { eoi = true; i = nullptr; i_len = i_pos = 0; return; }
    i = buf; i_len = len; i_pos = 0;
    if (fmt == DCMP_FMT_GZIP) crc = host_crc32(crc, buf, len); else adler = host_adler32(adler, buf, len);
    rd += len;
  }
  void dst(uint8_t *buf, int64_t len) { o = buf; o_len = len; o_pos = 0; }
  int64_t src_rem() const { return i_len - i_pos; }
  int64_t dst_rem() const { return o_len - o_pos; }
  void put(const void *p, size_t n) { outq.insert(outq.end(), (const uint8_t *)p, (const uint8_t *)p + n); }

  // DCMP_AWAIT / DCMP_FLUSH / DCMP_END, or DCMP_API_* < 0
  int encode() {
    for (;;) {
      // what is ready leaves first; a full `o` is a `Flush (zl.ml:488-490)
      if (outq_pos < outq.size()) {
        int64_t n = (int64_t)(outq.size() - outq_pos);
        if (n > o_len - o_pos) n = o_len - o_pos;
        if (n > 0) { memcpy(o + o_pos, outq.data() + outq_pos, (size_t)n); o_pos += n; outq_pos += (size_t)n; }
        if (outq_pos < outq.size()) return DCMP_FLUSH;
        outq.clear(); outq_pos = 0;
      }
      if (ended) {
        if (o_pos > 0) return DCMP_FLUSH;                         // the trailer leaves with a `Flush, `End finds `o` empty
        return DCMP_END;                                          // (zl.ml:494-499, test.ml:1597-1625)
      }
      if (!header_done) {
        header_done = true;
        if (fmt == DCMP_FMT_ZLIB) { const uint32_t h = zlib_header(level); const uint8_t b[2] = {(uint8_t)(h >> 8), (uint8_t)h}; put(b, 2); }
        else if (fmt == DCMP_FMT_GZIP) { const std::string h = gz.bytes(); put(h.data(), h.size()); }
        continue;
      }
      if (body_done) {
        if (fmt == DCMP_FMT_ZLIB) { const uint8_t b[4] = {(uint8_t)(adler >> 24), (uint8_t)(adler >> 16), (uint8_t)(adler >> 8), (uint8_t)adler}; put(b, 4); }
        else if (fmt == DCMP_FMT_GZIP) {
          const uint32_t sz = (uint32_t)rd;
          const uint8_t b[8] = {(uint8_t)crc, (uint8_t)(crc >> 8), (uint8_t)(crc >> 16), (uint8_t)(crc >> 24),
                                (uint8_t)sz, (uint8_t)(sz >> 8), (uint8_t)(sz >> 16), (uint8_t)(sz >> 24)};
          put(b, 8);
        }
        ended = true;
        continue;
      }
      // gather input up to one segment
      if (i_pos < i_len) {
        int64_t n = i_len - i_pos;
        if (n > seg_bytes - (int64_t)seg.size()) n = seg_bytes - (int64_t)seg.size();
        seg.insert(seg.end(), i + i_pos, i + i_pos + n);
        i_pos += n;
      }
      if ((int64_t)seg.size() < seg_bytes && !eoi) return DCMP_AWAIT;
      const bool final = eoi && i_pos == i_len;
      const int rc = be.deflate(level, seg.data(), (int64_t)seg.size(), final, outq);
      if (rc != DCMP_API_OK) return rc;
      seg.clear();
      if (final) body_done = true;
    }
  }
};

} // namespace dcmp
