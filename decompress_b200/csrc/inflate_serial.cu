// decompress_b200/csrc/inflate_serial.cu — general-purpose inflate kernel, one warp per stream.
//
// Replaces De.Inf.Ns.inflate (lib/de.ml:1534-1823) for n independent streams.  This is
// the *general* path: any block mix (stored / fixed / dynamic), any output size,
// every reference error class.  The warp executes the bit decode uniformly (all lanes
// hold the same bit cursor; table reads are shared-memory broadcasts) and spreads the
// byte work — stored-block copies, LZ77 back-references — over its 32 lanes.  The
// sub-chunk-parallel kernel for the C2 workload lives in inflate_parallel.cu; it falls
// back to this one for streams it cannot take.
//
// Input-exhaustion semantics.  The reference's observable behaviour depends only on the
// bit cursor c (bits consumed) and the input length n, not on its 16-bit refill
// schedule (proof sketch in DESIGN.md §"Bit cursor"):
//   * header-side reads (`_fill_bits`, de.ml:1629-1641) of k bits fail with
//     Unexpected_end_of_input iff 8n - c < k (k = 3, 14, 3 per HCLEN code, the CODES
//     table's max length per code-length symbol, 2/3/7 for the repeats);
//   * reads inside a compressed block (`__fill_bits`, de.ml:1643-1654) never fail; past
//     the end they see zero bits;
//   * after the final block consumed = ceil(c/8), or, when c ran past the input,
//     n - ((8n - c) lsr 3) in 63-bit OCaml arithmetic (de.ml:1805).
#include "inflate_tables.cuh"
#include "inflate_stream.h"
#include "simt.cuh"

namespace dcmp {

constexpr int SERIAL_WARPS = 4;

struct WarpScratch {
  uint32_t lit[LIT_TABLE_SIZE];
  uint32_t dist[DIST_TABLE_SIZE];
  uint32_t codes[CODES_TABLE_SIZE];
  uint16_t work[MAX_LENS];
  uint8_t lens[MAX_LENS];
};

struct BitCursor {
  const uint8_t *p;
  int64_t n;    // input length in bytes
  int64_t pos;  // next byte to load (may run past n: virtual zeros)
  uint64_t hold;
  int bits;     // loaded bits in hold
  __device__ __forceinline__ void init(const uint8_t *p_, int64_t n_) { p = p_; n = n_; pos = 0; hold = 0; bits = 0; }
  __device__ __forceinline__ void refill() { // afterwards bits >= 33
    if (bits <= 32) {
      uint32_t w;
      if (pos + 4 <= n) {
        w = (uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8) | ((uint32_t)p[pos + 2] << 16) | ((uint32_t)p[pos + 3] << 24);
      } else {
        w = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (pos + k < n) w |= (uint32_t)p[pos + k] << (8 * k);
      }
      hold |= (uint64_t)w << bits;
      bits += 32;
      pos += 4;
    }
  }
  __device__ __forceinline__ int64_t consumed_bits() const { return 8 * pos - bits; }
  __device__ __forceinline__ int64_t real_bits_left() const { return 8 * n - consumed_bits(); }
  __device__ __forceinline__ uint32_t peek(int k) const { return (uint32_t)hold & ((1u << k) - 1u); }
  __device__ __forceinline__ void drop(int k) { hold >>= k; bits -= k; }
  __device__ __forceinline__ uint32_t pop(int k) { uint32_t v = peek(k); drop(k); return v; }
  __device__ __forceinline__ void seek_byte(int64_t byte_pos) { pos = byte_pos; hold = 0; bits = 0; }
};

__device__ __forceinline__ uint32_t lookup(const uint32_t *tbl, int root, uint64_t hold) {
  uint32_t e = tbl[(uint32_t)hold & ((1u << root) - 1u)];
  if (e_kind(e) == K_LINK) e = tbl[e_base(e) + (((uint32_t)(hold >> root)) & ((1u << e_extra(e)) - 1u))];
  return e;
}

struct StreamOut {
  uint8_t *o;
  int64_t cap;
  int64_t pos;
};

// literal/length/distance loop, de.ml:1667-1712.  Returns a DCMP_* status (DCMP_OK = EOB seen).
__device__ int inflate_block(BitCursor &br, StreamOut &out, const uint32_t *lit, int lit_root,
                             const uint32_t *dist, int dist_root, int lane) {
  for (;;) {
    br.refill();
    uint32_t e = lookup(lit, lit_root, br.hold);
    uint32_t kind = e_kind(e);
    if (kind == K_INVALID) return DCMP_REF_EXCEPTION;
    br.drop(e_len(e));
    if (kind == K_LIT) {
      if (out.pos >= out.cap) return DCMP_UNEXPECTED_END_OF_OUTPUT;
      if (lane == 0) out.o[out.pos] = (uint8_t)e_base(e);
      out.pos++;
      continue;
    }
    if (kind == K_EOB) return DCMP_OK;
    int ext = e_extra(e);
    int len = (int)e_base(e) + (int)br.pop(ext);
    br.refill();
    uint32_t d = lookup(dist, dist_root, br.hold);
    if (e_kind(d) == K_INVALID) return DCMP_REF_EXCEPTION;
    br.drop(e_len(d));
    int dext = e_extra(d);
    int64_t dd = (int64_t)e_base(d) + (int64_t)br.pop(dext);
    if (e_base(d) == 0) return DCMP_INVALID_DISTANCE_CODE; // _base_dist = -1, de.ml:1698-1699
    if (dd > (out.pos < 32768 ? out.pos : 32768)) return DCMP_INVALID_DISTANCE;
    if (len > out.cap - out.pos) return DCMP_UNEXPECTED_END_OF_OUTPUT;
    __syncwarp(); // earlier stores of every lane are visible to the loads below
    {
      const uint8_t *s = out.o + out.pos - dd;
      uint8_t *t = out.o + out.pos;
      // byte k of the match is source byte (k mod dd): all sources precede out.pos
      for (int k = lane; k < len; k += 32) {
        int sk = dd >= len ? k : (int)(k % dd);
        t[k] = s[sk];
      }
    }
    out.pos += len;
  }
}

__device__ int build_dynamic_tables(WarpScratch &ws, int hlit, int hdist, int lane, TableInfo &li, TableInfo &di);
__device__ int decode_dynamic_header(BitCursor &br, WarpScratch &ws, int lane, TableInfo &li, TableInfo &di,
                                     int *hlit_out = nullptr, int *hdist_out = nullptr) {
  // de.ml:1788-1793
  br.refill();
  if (br.real_bits_left() < 14) return DCMP_UNEXPECTED_END_OF_INPUT;
  int hlit = (int)br.pop(5) + 257;
  int hdist = (int)br.pop(5) + 1;
  int hclen = (int)br.pop(4) + 4;
  // de.ml:1771-1786
  __syncwarp();
  if (lane < 19) ws.lens[lane] = 0;
  __syncwarp();
  for (int i = 0; i < hclen; i++) {
    br.refill();
    if (br.real_bits_left() < 3) return DCMP_UNEXPECTED_END_OF_INPUT;
    uint32_t v = br.pop(3);
    if (lane == 0) ws.lens[c_zigzag[i]] = (uint8_t)v;
  }
  __syncwarp();
  int rc = 0;
  TableInfo ci;
  if (lane == 0) rc = build_decode_table(KIND_CODES, ws.lens, 19, ws.codes, CODES_TABLE_SIZE, ws.work, &ci);
  rc = __shfl_sync(0xffffffffu, rc, 0);
  ci.root = __shfl_sync(0xffffffffu, ci.root, 0);
  ci.maxl = __shfl_sync(0xffffffffu, ci.maxl, 0);
  if (rc == TBL_INVALID) return DCMP_INVALID_DICTIONARY;
  if (rc == TBL_EXCEPTION) return DCMP_REF_EXCEPTION;
  __syncwarp();
  // de.ml:1733-1769
  const int max_res = hlit + hdist;
  int i = 0;
  int prev = 0;
  while (i < max_res) {
    br.refill();
    if (br.real_bits_left() < ci.maxl) return DCMP_UNEXPECTED_END_OF_INPUT;
    uint32_t e = ws.codes[br.peek(ci.maxl)];
    if (e_kind(e) == K_INVALID) return DCMP_REF_EXCEPTION;
    br.drop(e_len(e));
    int v = (int)e_base(e);
    int copy, fillv;
    if (v < 16) {
      if (lane == 0) ws.lens[i] = (uint8_t)v;
      prev = v;
      i++;
      continue;
    } else if (v == 16) {
      if (i == 0) return DCMP_INVALID_DICTIONARY;
      if (br.real_bits_left() < 2) return DCMP_UNEXPECTED_END_OF_INPUT;
      copy = (int)br.pop(2) + 3; fillv = prev;
    } else if (v == 17) {
      if (br.real_bits_left() < 3) return DCMP_UNEXPECTED_END_OF_INPUT;
      copy = (int)br.pop(3) + 3; fillv = 0;
    } else {
      if (br.real_bits_left() < 7) return DCMP_UNEXPECTED_END_OF_INPUT;
      copy = (int)br.pop(7) + 11; fillv = 0;
    }
    if (i + copy > max_res) return DCMP_INVALID_DICTIONARY;
    for (int x = lane; x < copy; x += 32) ws.lens[i + x] = (uint8_t)fillv;
    prev = fillv;
    i += copy;
  }
  if (hlit_out) { *hlit_out = hlit; *hdist_out = hdist; }
  return build_dynamic_tables(ws, hlit, hdist, lane, li, di);
}

// de.ml:1718-1731: the two decode tables of a dynamic block from its code lengths (ws.lens[0 .. hlit + hdist))
__device__ int build_dynamic_tables(WarpScratch &ws, int hlit, int hdist, int lane, TableInfo &li, TableInfo &di) {
  __syncwarp();
  if (ws.lens[256] == 0) return DCMP_INVALID_DICTIONARY;
  int rl = 0, rd = 0;
  if (lane == 0) {
    rl = build_decode_table(KIND_LENS, ws.lens, hlit, ws.lit, LIT_TABLE_SIZE, ws.work, &li);
    if (rl == TBL_OK) rd = build_decode_table(KIND_DISTS, ws.lens + hlit, hdist, ws.dist, DIST_TABLE_SIZE, ws.work, &di);
  }
  rl = __shfl_sync(0xffffffffu, rl, 0);
  rd = __shfl_sync(0xffffffffu, rd, 0);
  li.root = __shfl_sync(0xffffffffu, li.root, 0);
  li.maxl = __shfl_sync(0xffffffffu, li.maxl, 0);
  di.root = __shfl_sync(0xffffffffu, di.root, 0);
  di.maxl = __shfl_sync(0xffffffffu, di.maxl, 0);
  if (rl == TBL_INVALID || rd == TBL_INVALID) return (rl == TBL_EXCEPTION) ? DCMP_REF_EXCEPTION : DCMP_INVALID_DICTIONARY;
  if (rl == TBL_EXCEPTION || rd == TBL_EXCEPTION) return DCMP_REF_EXCEPTION;
  __syncwarp();
  return DCMP_OK;
}

__device__ void build_fixed_tables(WarpScratch &ws, int lane, TableInfo &li, TableInfo &di) {
  // de.ml:821-833
  __syncwarp();
  for (int n = lane; n < 288; n += 32) ws.lens[n] = n < 144 ? 8 : n < 256 ? 9 : n < 280 ? 7 : 8;
  if (lane < 32) ws.lens[288 + lane] = 5;
  __syncwarp();
  if (lane == 0) {
    build_decode_table(KIND_LENS, ws.lens, 288, ws.lit, LIT_TABLE_SIZE, ws.work, &li);
    build_decode_table(KIND_DISTS, ws.lens + 288, 32, ws.dist, DIST_TABLE_SIZE, ws.work, &di);
  }
  li.root = __shfl_sync(0xffffffffu, li.root, 0);
  li.maxl = __shfl_sync(0xffffffffu, li.maxl, 0);
  di.root = __shfl_sync(0xffffffffu, di.root, 0);
  di.maxl = __shfl_sync(0xffffffffu, di.maxl, 0);
  __syncwarp();
}

// De.Inf.Ns.inflate for one stream, executed by one warp.
__device__ int inflate_stream_warp(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap, WarpScratch &ws,
                                   int lane, int64_t *consumed, int64_t *produced) {
  BitCursor br;
  br.init(src, n);
  StreamOut out{dst, cap, 0};
  int status = DCMP_OK;
  TableInfo li{0, 0}, di{0, 0};
  bool fixed_ready = false; // consecutive fixed blocks reuse the tables
  for (;;) {
    // de.ml:1795-1805
    br.refill();
    if (br.real_bits_left() < 3) { status = DCMP_UNEXPECTED_END_OF_INPUT; break; }
    bool last = br.pop(1) == 1;
    uint32_t type = br.pop(2);
    if (type == 0) {
      // de.ml:1613-1627
      int64_t ip = (br.consumed_bits() + 7) >> 3;
      if (n - ip < 4) { status = DCMP_UNEXPECTED_END_OF_INPUT; break; }
      int len = src[ip] | (src[ip + 1] << 8);
      int nlen = src[ip + 2] | (src[ip + 3] << 8);
      ip += 4;
      if (nlen != 0xffff - len) { status = DCMP_INVALID_COMPLEMENT_OF_LENGTH; break; }
      if (len > n - ip) { status = DCMP_UNEXPECTED_END_OF_INPUT; break; }
      if (len > out.cap - out.pos) { status = DCMP_UNEXPECTED_END_OF_OUTPUT; break; }
      for (int k = lane; k < len; k += 32) out.o[out.pos + k] = src[ip + k];
      out.pos += len;
      br.seek_byte(ip + len);
    } else if (type == 3) {
      status = DCMP_INVALID_KIND_OF_BLOCK;
      break;
    } else {
      if (type == 1) {
        if (!fixed_ready) build_fixed_tables(ws, lane, li, di);
        fixed_ready = true;
      } else {
        fixed_ready = false;
        status = decode_dynamic_header(br, ws, lane, li, di);
        if (status != DCMP_OK) break;
      }
      status = inflate_block(br, out, ws.lit, li.root, ws.dist, di.root, lane);
      if (status != DCMP_OK) break;
    }
    if (last) break;
  }
  __syncwarp();
  int64_t c = br.consumed_bits();
  int64_t r = 8 * n - c;
  int64_t cons;
  if (status != DCMP_OK) cons = (c + 7) >> 3 > n ? n : (c + 7) >> 3;
  else if (r >= 0) cons = (c + 7) >> 3;
  else cons = n - (int64_t)(((uint64_t)r & 0x7fffffffffffffffULL) >> 3); // de.ml:1805 on a negative 63-bit int
  *consumed = cons;
  *produced = out.pos;
  return status;
}

__device__ uint32_t warp_adler32(const uint8_t *p, int64_t n, int lane) {
  // RFC1950 Adler-32 of p[0..n): lanes take contiguous slices, then combine
  // (A,B) pairs: B_total = sum_i (B_i + A_i * bytes_after_i) + n (for the initial 1).
  const uint32_t MOD = 65521u;
  int64_t per = (n + 31) / 32;
  int64_t lo = per * lane, hi = lo + per;
  if (lo > n) lo = n;
  if (hi > n) hi = n;
  uint64_t a = 0, b = 0;
  int64_t i = lo;
  while (i < hi) {
    int64_t stop = i + 2048 < hi ? i + 2048 : hi; // a,b < 2^64 trivially; reduce per chunk
    uint64_t a0 = a;
    uint64_t sa = 0, sb = 0;
    for (int64_t k = i; k < stop; k++) { sa += p[k]; sb += sa; }
    b = (b + a0 * (uint64_t)(stop - i) + sb) % MOD;
    a = (a + sa) % MOD;
    i = stop;
  }
  uint64_t after = (uint64_t)(n - hi);
  uint64_t bc = (b + (a * (after % MOD)) % MOD) % MOD;
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    bc += __shfl_xor_sync(0xffffffffu, bc, o);
  }
  uint32_t A = (uint32_t)((1 + a) % MOD);
  uint32_t B = (uint32_t)((bc + (uint64_t)(n % MOD)) % MOD);
  return (B << 16) | A;
}

__global__ void __launch_bounds__(SERIAL_WARPS * 32)
inflate_serial_kernel(int fmt, int64_t nstreams, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                      uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                      int64_t *produced, int32_t *status, int only_status, unsigned long long *redo_count) {
  __shared__ WarpScratch scratch[SERIAL_WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  int64_t k = (int64_t)blockIdx.x * SERIAL_WARPS + warp;
  if (k >= nstreams) return;
  // second stage of the dispatcher: only the streams the sub-chunk-parallel kernel gave up
  if (only_status >= 0 && status[k] != only_status) return;
  if (only_status >= 0 && redo_count && lane == 0) atomicAdd(redo_count, 1ull); // dcmp_engine_stats: streams the fast kernel gave up
  const uint8_t *src = src_base + src_off[k];
  int64_t n = src_len[k];
  uint8_t *dst = dst_base + dst_off[k];
  int64_t cap = dst_cap[k];
  int64_t cons = 0, prod = 0;
  int st;
  if (fmt == DCMP_FMT_RAW) {
    st = inflate_stream_warp(src, n, dst, cap, scratch[warp], lane, &cons, &prod);
  } else {
    // Zl.Inf.Ns.inflate, lib/zl.ml:391-417
    if (n < 2) st = DCMP_UNEXPECTED_END_OF_INPUT;
    else {
      uint32_t cmf = src[0], flg = src[1];
      if (((cmf << 8) + flg) % 31 != 0 || (cmf & 15) != 8) st = DCMP_INVALID_HEADER;
      else if (n < 6) st = DCMP_REF_EXCEPTION; // bigstring_sub src 2 (len - 6) raises
      else {
        int64_t i = 0;
        st = inflate_stream_warp(src + 2, n - 6, dst, cap, scratch[warp], lane, &i, &prod);
        if (st == DCMP_OK) {
          if (n < i + 6) st = DCMP_UNEXPECTED_END_OF_INPUT;
          else if (i < -2) st = DCMP_REF_EXCEPTION;
          else {
            uint32_t want = ((uint32_t)src[i + 2] << 24) | ((uint32_t)src[i + 3] << 16) |
                            ((uint32_t)src[i + 4] << 8) | (uint32_t)src[i + 5];
            __syncwarp();
            uint32_t got = warp_adler32(dst, prod, lane);
            if (want != got) st = DCMP_INVALID_CHECKSUM;
            else cons = i + 6;
          }
        }
      }
    }
  }
  if (lane == 0) {
    consumed[k] = cons;
    produced[k] = prod;
    status[k] = st;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Resumable inflate of ONE stream: the device side of the streaming decoders (De.Inf.decoder / Zl.Inf / Gz.Inf behind
// dcmp_inf_* in csrc/stream.cu).  One warp decodes from the saved state until the input runs out (the cursor goes
// back to the last block header / token boundary: `Await), the launch's output budget is used up (`Flush), the final
// block ends (`End) or the stream is malformed.  The window (WInf, de.ml:436-508) is a device ring that also takes
// the output; code lengths of the block in progress live in the state, so a block of any size is decoded with
// bounded memory.  At the end of the input (`eof`) a token that needs bits the input does not have is
// "Unexpected end of input" (de.ml:760-762, 938-945); before that it just waits for them.
struct RingOut {
  uint8_t *ring;
  int64_t mask;
  int64_t pos;      // total bytes produced since the stream began
  int64_t limit;    // pos may not pass this in this launch
};

enum : int { BR_EOB = -1, BR_AWAIT = -2, BR_FULL = -3 };

__device__ int inflate_block_resumable(BitCursor &br, RingOut &out, const uint32_t *lit, int lit_root, const uint32_t *dist,
                                       int dist_root, int lane, bool eof, int64_t *resume_bit) {
  const int64_t nbits = 8 * br.n;
  for (;;) {
    const int64_t t0 = br.consumed_bits();
    if (out.pos + 258 > out.limit) { *resume_bit = t0; return BR_FULL; }
    // with fewer than 48 real bits ahead (the longest token: 15 + 5 + 15 + 13) and more input to come, nothing
    // this token decodes to is final
    const bool tentative = !eof && nbits - t0 < 48;
    br.refill();
    uint32_t e = lookup(lit, lit_root, br.hold);
    uint32_t kind = e_kind(e);
    if (kind == K_INVALID) { if (tentative) { *resume_bit = t0; return BR_AWAIT; } return br.consumed_bits() >= nbits ? DCMP_UNEXPECTED_END_OF_INPUT : DCMP_REF_EXCEPTION; }
    br.drop(e_len(e));
    if (kind == K_LIT) {
      if (br.consumed_bits() > nbits) { if (!eof) { *resume_bit = t0; return BR_AWAIT; } return DCMP_UNEXPECTED_END_OF_INPUT; }
      if (lane == 0) out.ring[out.pos & out.mask] = (uint8_t)e_base(e);
      out.pos++;
      continue;
    }
    if (kind == K_EOB) {
      if (br.consumed_bits() > nbits) { if (!eof) { *resume_bit = t0; return BR_AWAIT; } return DCMP_UNEXPECTED_END_OF_INPUT; }
      return BR_EOB;
    }
    int len = (int)e_base(e) + (int)br.pop(e_extra(e));
    br.refill();
    uint32_t d = lookup(dist, dist_root, br.hold);
    if (e_kind(d) == K_INVALID) { if (tentative) { *resume_bit = t0; return BR_AWAIT; } return br.consumed_bits() >= nbits ? DCMP_UNEXPECTED_END_OF_INPUT : DCMP_REF_EXCEPTION; }
    br.drop(e_len(d));
    int64_t dd = (int64_t)e_base(d) + (int64_t)br.pop(e_extra(d));
    if (br.consumed_bits() > nbits) { if (!eof) { *resume_bit = t0; return BR_AWAIT; } return DCMP_UNEXPECTED_END_OF_INPUT; }
    if (e_base(d) == 0) return DCMP_INVALID_DISTANCE_CODE;                      // de.ml:1044
    if (dd > (out.pos < 32768 ? out.pos : 32768)) return DCMP_INVALID_DISTANCE; // de.ml:1045: d > WInf.have
    __syncwarp(); // earlier stores of every lane are visible to the loads below
    for (int k = lane; k < len; k += 32) {
      const int sk = dd >= len ? k : (int)(k % dd); // byte k repeats byte k mod dd: every source precedes out.pos
      out.ring[(out.pos + k) & out.mask] = out.ring[(out.pos - dd + sk) & out.mask];
    }
    out.pos += len;
  }
}

__global__ void __launch_bounds__(32)
inflate_stream_kernel(const uint8_t *in, int64_t n_in, int eof, uint8_t *ring, int64_t max_out, InfStreamState *S) {
  __shared__ WarpScratch ws;
  const int lane = threadIdx.x;
  BitCursor br;
  br.init(in, n_in);
  {
    const int64_t b = S->bitpos;
    br.pos = b >> 3;
    br.refill();
    br.drop((int)(b & 7)); // (a refill from pos loads 32 bits: 25..32 are left)
  }
  RingOut out{ring, STREAM_RING - 1, S->total_out, S->total_out + (max_out < STREAM_MAX_OUT ? max_out : STREAM_MAX_OUT)};
  const int64_t out0 = out.pos;
  int phase = S->phase, last = S->last, btype = S->btype, hlit = S->hlit, hdist = S->hdist;
  int64_t stored_rem = S->stored_rem;
  int64_t resume = S->bitpos;
  int result = -1, status = DCMP_OK;
  TableInfo li{0, 0}, di{0, 0};
  if (phase == SP_CODED) { // the tables of the block in progress, from its saved code lengths
    if (btype == 1) build_fixed_tables(ws, lane, li, di);
    else {
      for (int i = lane; i < hlit + hdist; i += 32) ws.lens[i] = S->lens[i];
      status = build_dynamic_tables(ws, hlit, hdist, lane, li, di);
      if (status != DCMP_OK) result = SR_ERROR; // (cannot happen: they were built from these lengths before)
    }
  }
  while (result < 0) {
    if (phase == SP_DONE) { result = SR_END; break; }
    if (phase == SP_HEADER) {
      resume = br.consumed_bits();
      br.refill();
      if (br.real_bits_left() < 3) {
        if (!eof) result = SR_AWAIT; else { result = SR_ERROR; status = DCMP_UNEXPECTED_END_OF_INPUT; }
        break;
      }
      last = (int)br.pop(1);
      const uint32_t type = br.pop(2);
      if (type == 3) { result = SR_ERROR; status = DCMP_INVALID_KIND_OF_BLOCK; break; }
      if (type == 0) {
        // de.ml:867-921: to the byte boundary, LEN, NLEN
        const int64_t ip = (br.consumed_bits() + 7) >> 3;
        if (n_in - ip < 4) {
          if (!eof) result = SR_AWAIT; else { result = SR_ERROR; status = DCMP_UNEXPECTED_END_OF_INPUT; }
          break;
        }
        const int len = in[ip] | (in[ip + 1] << 8), nlen = in[ip + 2] | (in[ip + 3] << 8);
        if (nlen != 0xffff - len) { result = SR_ERROR; status = DCMP_INVALID_COMPLEMENT_OF_LENGTH; break; }
        br.seek_byte(ip + 4);
        stored_rem = len;
        phase = SP_STORED;
        resume = 8 * (ip + 4);
      } else {
        btype = (int)type;
        if (type == 1) build_fixed_tables(ws, lane, li, di);
        else {
          status = decode_dynamic_header(br, ws, lane, li, di, &hlit, &hdist);
          if (status == DCMP_UNEXPECTED_END_OF_INPUT && !eof) { status = DCMP_OK; result = SR_AWAIT; break; } // (resume: this header)
          if (status != DCMP_OK) { result = SR_ERROR; break; }
          __syncwarp();
          for (int i = lane; i < hlit + hdist; i += 32) S->lens[i] = ws.lens[i];
        }
        phase = SP_CODED;
        resume = br.consumed_bits();
      }
    }
    if (phase == SP_STORED) {
      // de.ml:837-865: raw bytes, as many as the input and this launch's budget hold
      const int64_t ip = br.consumed_bits() >> 3;
      int64_t m = stored_rem;
      if (m > n_in - ip) m = n_in - ip;
      if (m > out.limit - out.pos) m = out.limit - out.pos;
      for (int64_t k = lane; k < m; k += 32) out.ring[(out.pos + k) & out.mask] = in[ip + k];
      out.pos += m;
      stored_rem -= m;
      br.seek_byte(ip + m);
      resume = 8 * (ip + m);
      if (stored_rem > 0) {
        if (out.pos >= out.limit) result = SR_FULL;
        else if (!eof) result = SR_AWAIT;
        else { result = SR_ERROR; status = DCMP_UNEXPECTED_END_OF_INPUT; }
        break;
      }
      phase = last ? SP_DONE : SP_HEADER;
      continue;
    }
    if (phase == SP_CODED) {
      const int r = inflate_block_resumable(br, out, ws.lit, li.root, ws.dist, di.root, lane, eof != 0, &resume);
      if (r == BR_EOB) { phase = last ? SP_DONE : SP_HEADER; resume = br.consumed_bits(); continue; }
      if (r == BR_AWAIT) { result = SR_AWAIT; break; }
      if (r == BR_FULL) { result = SR_FULL; break; }
      result = SR_ERROR;
      status = r;
      break;
    }
  }
  __syncwarp();
  if (lane == 0) {
    S->bitpos = (result == SR_END || result == SR_ERROR) ? br.consumed_bits() : resume;
    if (result == SR_END) S->bitpos = resume;
    S->total_out = out.pos;
    S->produced = out.pos - out0;
    S->phase = phase;
    S->last = last;
    S->btype = btype;
    S->hlit = hlit;
    S->hdist = hdist;
    S->stored_rem = (int32_t)stored_rem;
    S->result = result;
    S->status = status;
  }
}

#ifndef DCMP_EMU
void launch_inflate_stream(const uint8_t *d_in, int64_t n_in, int eof, uint8_t *d_ring, int64_t max_out, void *d_state,
                           cudaStream_t stream) {
  inflate_stream_kernel<<<1, 32, 0, stream>>>(d_in, n_in, eof, d_ring, max_out, (InfStreamState *)d_state);
}
#endif

#ifndef DCMP_EMU
void launch_inflate_serial(int fmt, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                           uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                           int64_t *produced, int32_t *status, int only_status, unsigned long long *redo_count,
                           cudaStream_t stream) {
  if (n <= 0) return;
  int64_t blocks = (n + SERIAL_WARPS - 1) / SERIAL_WARPS;
  inflate_serial_kernel<<<(unsigned)blocks, SERIAL_WARPS * 32, 0, stream>>>(
      fmt, n, src_base, src_off, src_len, dst_base, dst_off, dst_cap, consumed, produced, status, only_status, redo_count);
}
#endif

} // namespace dcmp
