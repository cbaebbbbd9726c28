// decompress_b200/csrc/inflate_parallel.cu — sub-chunk-parallel inflate, one CTA per stream.
//
// This is the hot path of BASELINE config C2 (batches of independent dynamic-Huffman
// streams).  It replaces the literal/length/distance state machine of De.Inf
// (lib/de.ml:1054-1261, one-shot form 1667-1712), the window (WInf, de.ml:436-508) and
// the table construction (Inf.huffman, de.ml:523-638) with a design that has no
// counterpart in the reference (which is strictly serial):
//
//   round = 128 lanes x 1056-bit sub-chunks of the compressed stream (33 words: an odd
//           word stride, so the lanes of a warp read the TMA-staged window without
//           shared-memory bank conflicts and no per-lane repacking is needed)
//   A  speculative decode   every lane decodes its sub-chunk from an arbitrary bit
//                           offset into 32-bit tokens (literal | len,dist) written to an
//                           L2-resident scratch in [token][lane] order (coalesced), and
//                           leaves a checkpoint (first token start, token index, bytes so
//                           far) for every 64-bit window it crosses;
//   B  head re-decode       lane t restarts at the true entry point handed over by lane
//                           t-1 and decodes until its first token start inside a 64-bit
//                           window equals the speculative pass's checkpoint (Huffman
//                           self-synchronisation: ~10 tokens on text, geometric tail);
//                           from there on the speculative tokens are the true ones;
//   C  chain + scan         the round is cut at the first lane that cannot be validated
//                           (end of block, token buffer full, no re-synchronisation,
//                           output budget); an exclusive scan of emitted byte counts
//                           gives every lane its output offset;
//   D  LZ77 resolve         every lane replays its tokens into the 64 KiB shared-memory
//                           ring (the sliding window); back-references into bytes another
//                           lane of the same round still owes wait on that lane's
//                           progress counter (fine-grained dataflow, no global order);
//   E  drain                finished bytes leave through TMA bulk stores (16-byte aligned
//                           body); the next round's input arrives through a TMA bulk load;
//                           Adler-32 is folded over the drained bytes.
//
// Anything that is not the happy path — stored blocks, malformed input, incomplete code
// sets, an output buffer that is too small — makes the CTA give the stream up with
// status DCMP_INTERNAL_RETRY; the host-side dispatcher (engine.cu) then runs the general
// one-warp-per-stream kernel (inflate_serial.cu), which reproduces the reference's exact
// error behaviour, on exactly those streams.  Both kernels are device code; there is no
// CPU path.
#include "inflate_tables.cuh"
#include "tma.cuh"

namespace dcmp {

constexpr int PT = 128;            // threads (= lanes = sub-chunks per round)
constexpr int SUB_WORDS = 33;      // sub-chunk size in 32-bit words (odd: conflict-free lane stride)
constexpr int SUB_BITS = 32 * SUB_WORDS;
constexpr int SPEC_CAP = 160;      // tokens a lane may produce speculatively
constexpr int HEAD_CAP = 64;       // tokens a lane may produce while re-synchronising
constexpr int TOK_SLOTS = SPEC_CAP + HEAD_CAP; // per lane, in the global scratch
constexpr int CKPT_SHIFT = 6;      // one checkpoint per 64-bit window of the sub-chunk
constexpr int NCKPT = (SUB_BITS >> CKPT_SHIFT) + 1;
constexpr int RING = 65536;        // output ring (window) in bytes
constexpr int RING_MASK = RING - 1;
constexpr int RING_SLACK = 512;
constexpr int WIN_WORDS = 3 + SUB_WORDS * PT + 4;        // staged input per round
constexpr int WIN_BYTES = (WIN_WORDS * 4 + 15) & ~15;
constexpr int HDR_BYTES = 1024 + 16;      // staged input for a block header
constexpr int LIT_ROOT = 10;
constexpr int DIST_ROOT = 9;
constexpr int BUCKET_SHIFT = 6;
constexpr int NBUCKET = (RING >> BUCKET_SHIFT) + 2;
constexpr int DCMP_INTERNAL_RETRY = 0x7f; // never leaves the library

#ifdef DCMP_EMU
struct EmuStats { long hist_nh[48]; long nosync_lanes; long rounds, valid_lanes, cut_term[5], cut_bad, cut_budget, head_tokens, spec_tokens, used_tokens, d_iters, d_wait_iters, hdr_rounds, retries; };
static EmuStats g_stats;
#define STAT(x) x
#else
#define STAT(x)
#endif

enum : int { X_NORMAL = 0, X_EOB = 1, X_CAP = 2, X_ERR = 3, X_NOSYNC = 4 };
enum : int { NEXT_CONTINUE = 0, NEXT_HEADER = 1, NEXT_FINISH = 2, NEXT_RETRY = 3 };

// token: [23] match flag; match: [22:8] distance-1, [7:0] length-3;  literal: [7:0] byte
// checkpoint: [5:0] bit offset of the first token start inside the 64-bit window,
//             [13:6] index of that token, [31:14] bytes emitted before it
__device__ __forceinline__ bool tok_is_match(uint32_t t) { return (t >> 23) & 1u; }
__device__ __forceinline__ uint32_t tok_len(uint32_t t) { return tok_is_match(t) ? (t & 0xffu) + 3u : 1u; }

struct HuffMeta {      // canonical-code description for codes longer than the root table
  uint16_t first[16];  // first code (MSB-first) of each length
  uint16_t cnt[16];
  uint16_t offs[16];   // index of the first symbol of each length in sorted[]
  int root, maxl;
};

struct ParSmem {
  uint8_t ring[RING];
  __align__(16) uint32_t win[WIN_BYTES / 4 + 4]; // staged compressed input (TMA destination)
  uint32_t lit[1 << LIT_ROOT];
  uint32_t dist[1 << DIST_ROOT];
  uint32_t codes[CODES_TABLE_SIZE];
  uint32_t code_of[MAX_LENS];    // table build scratch: bit-reversed code | len << 16 per symbol
  uint32_t ckpt[NCKPT][PT];
  uint16_t sorted_l[288];
  uint16_t sorted_d[32];
  uint8_t lens[MAX_LENS];
  HuffMeta ml, md;
  uint32_t run_cnt[16];
  // per-lane results of a round
  uint16_t exit_pos[PT];         // spec pass: bit offset (own sub-chunk coordinates) after the last token
  uint8_t exit_kind[PT];
  uint32_t lane_off[PT + 1];     // output position (stream-relative) where the lane starts writing
  volatile uint32_t progress[PT];
  uint8_t bucket[NBUCKET];
  uint32_t warp_a[PT / 32], warp_b[PT / 32], warp_c[PT / 32];
  unsigned long long adl_s[PT / 32], adl_b[PT / 32];
  // stream state (CTA-uniform)
  int64_t stream;                // index of the stream being decoded, -1 = none
  const uint8_t *src; int64_t src_len;  // raw deflate bytes
  uint8_t *dst; int64_t dst_cap;
  int64_t base_bit;              // next bit of the deflate stream to decode
  uint32_t out_pos;              // bytes produced so far
  uint32_t flushed;              // bytes already stored to global
  uint32_t ring_bias;            // ring index of output byte p is (p + ring_bias) & RING_MASK
  int last_block, next, status;
  int n_valid;                   // lanes accepted in this round
  uint32_t round_out;            // bytes produced by this round
  uint32_t adler_a, adler_b;
  uint32_t mbar_phase;
  volatile int abort_round;      // a lane met an invalid distance: everybody leaves phase D
  __align__(8) uint64_t mbar;
};
static_assert(sizeof(ParSmem) <= 113 * 1024, "two CTAs per SM need <= 113 KiB each");

// ------------------------------------------------------------------------------------------
// bit cursor over a lane's sub-chunk inside the staged window
struct LaneBits {
  uint64_t hold;
  int nbits; // valid bits in hold
  int wi;    // next word (relative to the lane's first word) to load
  int sh;    // bit offset of the sub-chunk inside its first word
  const uint32_t *w; // the lane's first word
  __device__ __forceinline__ void init(const uint32_t *lane_words, int sh_, int p0) {
    w = lane_words;
    sh = sh_;
    int b = p0 + sh;
    wi = b >> 5;
    int skip = b & 31;
    hold = (uint64_t)(w[wi] >> skip);
    nbits = 32 - skip;
    wi++;
  }
  __device__ __forceinline__ void refill() {
    if (nbits <= 32) {
      hold |= (uint64_t)w[wi] << nbits; // the window is padded: wi never leaves it
      nbits += 32;
      wi++;
    }
  }
  __device__ __forceinline__ int pos() const { return 32 * wi - nbits - sh; }
  __device__ __forceinline__ void drop(int k) { hold >>= k; nbits -= k; }
  __device__ __forceinline__ uint32_t pop(int k) {
    uint32_t v = (uint32_t)hold & ((1u << k) - 1u);
    drop(k);
    return v;
  }
};

// canonical decode for codes longer than the root table (rare): de.ml's second-level
// tables (de.ml:614-636) are replaced by a first-code comparison per length.
__device__ __noinline__ uint32_t slow_symbol(const HuffMeta &m, const uint16_t *sorted, uint32_t bits_lsb_first, int *len_out) {
  uint32_t rev = __brev(bits_lsb_first);
  for (int l = m.root + 1; l <= m.maxl; l++) {
    uint32_t c = rev >> (32 - l);
    uint32_t idx = c - m.first[l];
    if (c >= m.first[l] && idx < m.cnt[l]) {
      *len_out = l;
      return sorted[m.offs[l] + idx];
    }
  }
  *len_out = m.maxl; // unreachable for a complete code
  return 0;
}

enum : int { T_TOKEN = 0, T_EOB = 1, T_ERR = 2 };

// Decode one literal / match at the cursor.  *tok gets the token (without start tag), *nout its byte count.
__device__ __forceinline__ int decode_token(ParSmem &S, LaneBits &lb, uint32_t lmask, uint32_t dmask,
                                            uint32_t *tok, uint32_t *nout) {
  lb.refill();
  uint32_t e = S.lit[(uint32_t)lb.hold & lmask];
  if (e_kind(e) == K_LINK) {
    int l;
    uint32_t sym = slow_symbol(S.ml, S.sorted_l, (uint32_t)lb.hold, &l);
    e = lit_symbol_entry((int)sym) | (uint32_t)l;
  }
  lb.drop(e_len(e));
  uint32_t kind = e_kind(e);
  if (kind == K_LIT) {
    *tok = e_base(e);
    *nout = 1;
    return T_TOKEN;
  }
  if (kind == K_EOB) return T_EOB;
  uint32_t len = e_base(e) + lb.pop(e_extra(e));
  lb.refill();
  uint32_t d = S.dist[(uint32_t)lb.hold & dmask];
  if (e_kind(d) == K_LINK) {
    int l;
    uint32_t sym = slow_symbol(S.md, S.sorted_d, (uint32_t)lb.hold, &l);
    d = dist_symbol_entry((int)sym) | (uint32_t)l;
  }
  lb.drop(e_len(d));
  uint32_t db = e_base(d);
  uint32_t dd = db + lb.pop(e_extra(d));
  if (db == 0) return T_ERR; // distance symbols 30/31: "Invalid distance code" (de.ml:321-325,1699)
  *tok = (1u << 23) | ((dd - 1u) << 8) | (len - 3u);
  *nout = len;
  return T_TOKEN;
}

// ------------------------------------------------------------------------------------------
// Parallel construction of the root decode table + canonical metadata from code lengths.
// Accepts only complete codes (Kraft sum exactly 1); anything else is left to the general
// kernel.  Returns false (CTA-uniform) when the code set is not acceptable.
__device__ bool build_fast_table(ParSmem &S, int kind, const uint8_t *lens, int n, uint32_t *tbl, int root_max,
                                 HuffMeta &m, uint16_t *sorted, int tid) {
  __syncthreads();
  if (tid < 16) S.run_cnt[tid] = 0;
  __syncthreads();
  for (int s = tid; s < n; s += PT) {
    int l = lens[s];
    if (l) atomicAdd(&S.run_cnt[l], 1u);
  }
  __syncthreads();
  if (tid == 0) {
    int left = 1, maxl = 0;
    uint32_t code = 0, off = 0;
    m.cnt[0] = 0; m.first[0] = 0; m.offs[0] = 0;
    for (int l = 1; l <= 15; l++) {
      uint32_t c = S.run_cnt[l];
      code = (code + (l > 1 ? S.run_cnt[l - 1] : 0)) << 1;
      m.first[l] = (uint16_t)code;
      m.cnt[l] = (uint16_t)c;
      m.offs[l] = (uint16_t)off;
      off += c;
      left = (left << 1) - (int)c;
      if (left < 0) left = -100000; // over-subscribed
      if (c) maxl = l;
    }
    m.maxl = maxl;
    m.root = maxl < root_max ? maxl : root_max;
    S.warp_a[0] = (left == 0 && maxl >= 1) ? 1u : 0u; // complete codes only
  }
  __syncthreads();
  if (!S.warp_a[0]) return false;
  if (tid < 16) S.run_cnt[tid] = 0;
  __syncthreads();
  // canonical rank of every symbol among the symbols of its length (by symbol index):
  // warp 0 walks the alphabet 32 symbols at a time, __match_any_sync groups equal lengths.
  if (tid < 32) {
    for (int base = 0; base < n; base += 32) {
      int s = base + tid;
      uint32_t l = s < n ? lens[s] : 0u;
      uint32_t grp = __match_any_sync(0xffffffffu, l);
      uint32_t rank = (uint32_t)__popc(grp & ((1u << tid) - 1u));
      uint32_t before = S.run_cnt[l];
      __syncwarp();
      if (l && rank == 0) S.run_cnt[l] = before + (uint32_t)__popc(grp);
      __syncwarp();
      if (s < n) {
        if (l) {
          uint32_t r = before + rank;
          uint32_t code = (uint32_t)m.first[l] + r;
          S.code_of[s] = (__brev(code) >> (32 - l)) | (l << 16);
          sorted[m.offs[l] + r] = (uint16_t)s;
        } else {
          S.code_of[s] = 0;
        }
      }
    }
  }
  __syncthreads();
  const int root = m.root;
  for (int s = tid; s < n; s += PT) {
    uint32_t co = S.code_of[s];
    int l = (int)(co >> 16);
    if (!l) continue;
    uint32_t r = co & 0xffffu;
    if (l <= root) {
      uint32_t entry = symbol_entry(kind, s) | (uint32_t)l;
      for (uint32_t i = r; i < (1u << root); i += (1u << l)) tbl[i] = entry;
    } else {
      tbl[r & ((1u << root) - 1u)] = mk_entry(K_LINK, 0, 0, 0);
    }
  }
  __syncthreads();
  return true;
}

// ------------------------------------------------------------------------------------------
// Block header (de.ml:1788-1793, 1771-1786, 1733-1769) parsed by thread 0 from the staged
// window; returns NEXT_* / sets lens[] and hlit/hdist.
struct HdrBits {
  const uint32_t *w;
  int nwords;
  uint64_t hold;
  int nbits, wi;
  __device__ void init(const uint8_t *win, int bit0, int nbytes) {
    w = (const uint32_t *)win;
    nwords = nbytes >> 2;
    wi = bit0 >> 5;
    int skip = bit0 & 31;
    hold = (uint64_t)(w[wi] >> skip);
    nbits = 32 - skip;
    wi++;
  }
  __device__ void refill() {
    if (nbits <= 32) {
      uint32_t x = wi < nwords ? w[wi] : 0u;
      hold |= (uint64_t)x << nbits;
      nbits += 32;
      wi++;
    }
  }
  __device__ int pos() const { return 32 * wi - nbits; }
  __device__ uint32_t pop(int k) {
    uint32_t v = (uint32_t)hold & ((1u << k) - 1u);
    hold >>= k;
    nbits -= k;
    return v;
  }
};

// thread 0 only.  Returns bits consumed (>0) or -1 to give the stream to the general kernel.
__device__ int parse_dynamic_header(ParSmem &S, HdrBits &hb, int64_t bits_avail, int *hlit_out, int *hdist_out) {
  hb.refill();
  if (bits_avail - hb.pos() < 14) return -1;
  int hlit = (int)hb.pop(5) + 257, hdist = (int)hb.pop(5) + 1, hclen = (int)hb.pop(4) + 4;
  if (hlit > 286 || hdist > 30) return -1; // zlib rejects; the reference's behaviour here is exotic
  uint8_t cl[19];
  for (int i = 0; i < 19; i++) cl[i] = 0;
  for (int i = 0; i < hclen; i++) {
    hb.refill();
    cl[c_zigzag[i]] = (uint8_t)hb.pop(3);
  }
  // code-length code: must be complete (general kernel handles the rest)
  int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 19; i++) cnt[cl[i]]++;
  int left = 1, maxl = 0;
  uint32_t first[8], code = 0;
  first[0] = 0;
  for (int l = 1; l <= 7; l++) {
    code = (code + (l > 1 ? (uint32_t)cnt[l - 1] : 0u)) << 1;
    first[l] = code;
    left = (left << 1) - cnt[l];
    if (left < 0) return -1;
    if (cnt[l]) maxl = l;
  }
  if (left != 0 || maxl == 0) return -1;
  for (int s = 0; s < 19; s++) {
    int l = cl[s];
    if (!l) continue;
    uint32_t c = first[l]++;
    uint32_t r = __brev(c) >> (32 - l);
    for (uint32_t i = r; i < 128u; i += (1u << l)) S.codes[i] = mk_entry(K_LIT, (uint32_t)s, 0, (uint32_t)l);
  }
  const int total = hlit + hdist;
  int i = 0, prev = 0;
  while (i < total) {
    hb.refill();
    uint32_t e = S.codes[(uint32_t)hb.hold & 127u];
    hb.pop((int)e_len(e));
    int v = (int)e_base(e);
    if (v < 16) {
      S.lens[i++] = (uint8_t)v;
      prev = v;
      continue;
    }
    int rep, fill;
    if (v == 16) {
      if (i == 0) return -1;
      rep = 3 + (int)hb.pop(2); fill = prev;
    } else if (v == 17) {
      rep = 3 + (int)hb.pop(3); fill = 0;
    } else {
      rep = 11 + (int)hb.pop(7); fill = 0;
    }
    if (i + rep > total) return -1;
    for (int x = 0; x < rep; x++) S.lens[i + x] = (uint8_t)fill;
    prev = fill;
    i += rep;
  }
  if (bits_avail - hb.pos() < 0) return -1; // header ran past the end of the input
  if (S.lens[256] == 0) return -1;
  *hlit_out = hlit;
  *hdist_out = hdist;
  return hb.pos();
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// Stage `bytes` (16-byte multiple) from global `g` (16-byte aligned) into S.win; bytes past
// `valid` are zero-filled by the threads.  All threads call; returns after the data landed.
__device__ void stage_window(ParSmem &S, const uint8_t *g, int bytes, int valid, int tid) {
  // previous readers of S.win are done (callers sync before)
  if (valid > bytes) valid = bytes;
  if (valid < 0) valid = 0;
  for (int i = valid + tid * 16; i < bytes; i += PT * 16) *(uint4 *)((uint8_t *)S.win + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0 && valid > 0) {
    mbar_expect_tx(&S.mbar, (uint32_t)valid);
    tma_load_1d(S.win, g, (uint32_t)valid, &S.mbar);
  }
  if (valid > 0) {
    mbar_wait(&S.mbar, S.mbar_phase & 1u);
  }
  __syncthreads();
  if (tid == 0 && valid > 0) S.mbar_phase ^= 1u;
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT, 2)
inflate_parallel_kernel(int fmt, int64_t nstreams, const uint8_t *src_base, const int64_t *src_off,
                        const int64_t *src_len, uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                        int64_t *consumed, int64_t *produced, int32_t *status, unsigned long long *work_counter,
                        uint32_t *tok_scratch) {
  DCMP_DYN_SMEM(smem_raw);
  ParSmem &S = *reinterpret_cast<ParSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&S.mbar, 1);
    fence_mbar_init();
    S.mbar_phase = 0;
  }
  __syncthreads();

  for (;;) {
    // ---------------------------------------------------------------- fetch a stream
    __syncthreads();
    if (tid == 0) {
      unsigned long long k = atomicAdd(work_counter, 1ull);
      S.stream = k < (unsigned long long)nstreams ? (int64_t)k : -1;
      if (S.stream >= 0) {
        const uint8_t *p = src_base + src_off[k];
        int64_t n = src_len[k];
        int st = DCMP_OK;
        if (fmt == DCMP_FMT_ZLIB) {
          // Zl.Inf.Ns.inflate, lib/zl.ml:391-417: anything odd goes to the general kernel
          if (n < 6) st = DCMP_INTERNAL_RETRY;
          else {
            uint32_t cmf = p[0], flg = p[1];
            if (((cmf << 8) + flg) % 31 != 0 || (cmf & 15) != 8) st = DCMP_INTERNAL_RETRY;
            p += 2;
            n -= 6;
          }
        }
        S.src = p;
        S.src_len = n;
        S.dst = dst_base + dst_off[k];
        S.dst_cap = dst_cap[k];
        if (S.dst_cap > 0xfff00000LL || n > 0x0fffffffLL) st = DCMP_INTERNAL_RETRY; // 32-bit positions below
        S.base_bit = 0;
        S.out_pos = 0;
        S.flushed = 0;
        S.ring_bias = (uint32_t)((uintptr_t)S.dst & RING_MASK);
        S.last_block = 0;
        S.status = st;
        S.next = st == DCMP_OK ? NEXT_HEADER : NEXT_RETRY;
        S.adler_a = 1;
        S.adler_b = 0;
      }
    }
    __syncthreads();
    if (S.stream < 0) break;

    // ---------------------------------------------------------------- rounds
    while (S.next == NEXT_HEADER || S.next == NEXT_CONTINUE) {
      const int64_t base_bit = S.base_bit;
      const uint32_t A = (uint32_t)((uintptr_t)S.src & 15);
      const int64_t vbit = base_bit + 8 * (int64_t)A;        // bit position from the 16-aligned address below src
      const int64_t v0 = (vbit >> 3) & ~(int64_t)15;          // window start (bytes from that address)
      const int bit0 = (int)(vbit - 8 * v0);                  // 0..127
      const uint8_t *g0 = S.src - A + v0;
      const int64_t vend = ((int64_t)A + S.src_len + 15) & ~(int64_t)15; // one past the last valid granule
      if (base_bit > 8 * S.src_len) { // ran off the input: the general kernel knows what the reference does
        if (tid == 0) S.next = NEXT_RETRY;
        __syncthreads();
        break;
      }
      if (S.next == NEXT_HEADER) {
        // ------------------------------------------------------------ block header
        stage_window(S, g0, HDR_BYTES, (int)(vend - v0 > HDR_BYTES ? HDR_BYTES : vend - v0), tid);
        if (tid == 0) {
          HdrBits hb;
          hb.init((const uint8_t *)S.win, bit0, HDR_BYTES);
          int64_t avail = 8 * S.src_len - base_bit + bit0; // valid bits measured in window coordinates
          int nx = NEXT_CONTINUE;
          hb.refill();
          if (avail - hb.pos() < 3) nx = NEXT_RETRY;
          else {
            S.last_block = (int)hb.pop(1);
            uint32_t type = hb.pop(2);
            if (type == 2) {
              int hlit = 0, hdist = 0;
              int used = parse_dynamic_header(S, hb, avail, &hlit, &hdist);
              if (used < 0) nx = NEXT_RETRY;
              else {
                S.warp_b[0] = (uint32_t)hlit;
                S.warp_b[1] = (uint32_t)hdist;
                S.base_bit = base_bit + (used - bit0);
              }
            } else if (type == 1) {
              for (int n = 0; n < 288; n++) S.lens[n] = n < 144 ? 8 : n < 256 ? 9 : n < 280 ? 7 : 8; // de.ml:821-833
              for (int n = 0; n < 30; n++) S.lens[288 + n] = 5;
              // the fixed distance code has 32 codes of 5 bits; 30/31 are invalid symbols
              S.lens[288 + 30] = 5;
              S.lens[288 + 31] = 5;
              S.warp_b[0] = 288;
              S.warp_b[1] = 32;
              S.base_bit = base_bit + (hb.pos() - bit0);
            } else {
              nx = NEXT_RETRY; // stored blocks and BTYPE=3: general kernel
            }
          }
          S.next = nx;
        }
        __syncthreads();
        if (S.next == NEXT_RETRY) break;
        const int hlit = (int)S.warp_b[0], hdist = (int)S.warp_b[1];
        bool ok = build_fast_table(S, KIND_LENS, S.lens, hlit, S.lit, LIT_ROOT, S.ml, S.sorted_l, tid);
        ok = ok && build_fast_table(S, KIND_DISTS, S.lens + hlit, hdist, S.dist, DIST_ROOT, S.md, S.sorted_d, tid);
        if (!ok) {
          if (tid == 0) S.next = NEXT_RETRY;
          __syncthreads();
          break;
        }
        continue; // S.next == NEXT_CONTINUE, S.base_bit moved past the header
      }

      // -------------------------------------------------------------- data round
      {
        int64_t want = vend - v0;
        stage_window(S, g0, WIN_BYTES, (int)(want > WIN_BYTES ? WIN_BYTES : want), tid);
      }
      const int sh = bit0 & 31;
      const uint32_t *lane_words = S.win + (bit0 >> 5) + SUB_WORDS * tid;
      const uint32_t lmask = (1u << S.ml.root) - 1u, dmask = (1u << S.md.root) - 1u;
      uint32_t *tok_spec = tok_scratch + (size_t)blockIdx.x * TOK_SLOTS * PT + tid; // [slot][lane]
      uint32_t *tok_head = tok_spec + (size_t)SPEC_CAP * PT;
      // lanes whose sub-chunk starts beyond the input decode nothing
      const bool lane_has_input = base_bit + (int64_t)SUB_BITS * tid < 8 * S.src_len;
      // ---- A: speculative decode of the own sub-chunk from bit 0
      int ns = 0;
      uint32_t out_s = 0;
      int xkind = X_ERR, xpos = 0;
      int last_ck = -1;
      if (lane_has_input) {
        LaneBits lb;
        lb.init(lane_words, sh, 0);
        for (;;) {
          int p = lb.pos();
          if (p >= SUB_BITS) { xkind = X_NORMAL; xpos = p; break; }
          if (ns == SPEC_CAP) { xkind = X_CAP; xpos = p; break; }
          int ck = p >> CKPT_SHIFT;
          if (ck != last_ck) {
            S.ckpt[ck][tid] = (uint32_t)(p & 63) | ((uint32_t)ns << 6) | (out_s << 14);
            last_ck = ck;
          }
          uint32_t tok, no;
          int r = decode_token(S, lb, lmask, dmask, &tok, &no);
          if (r == T_EOB) { xkind = X_EOB; xpos = lb.pos(); break; }
          if (r == T_ERR) { xkind = X_ERR; xpos = p; break; }
          tok_spec[(size_t)ns * PT] = tok;
          ns++;
          out_s += no;
        }
      }
      S.exit_pos[tid] = (uint16_t)xpos;
      S.exit_kind[tid] = (uint8_t)xkind;
      __syncthreads();
      // ---- B: re-synchronise with the true entry point handed over by the previous lane
      int nh = 0, kmerge = 0;
      uint32_t out_valid = 0;
      bool in_ok = lane_has_input, term, merged = false;
      int tkind = xkind, tpos = xpos; // exit of this lane's TRUE sequence
      {
        int p_in = 0;
        if (tid > 0 && in_ok) {
          if (S.exit_kind[tid - 1] != X_NORMAL) in_ok = false;
          else p_in = (int)S.exit_pos[tid - 1] - SUB_BITS;
        }
        uint32_t skipped = 0, out_h = 0;
        if (in_ok) {
          if (p_in == 0 && ns > 0) {
            merged = true;
          } else {
            LaneBits lb;
            lb.init(lane_words, sh, p_in);
            int seen_ck = -1;
            for (;;) {
              int p = lb.pos();
              if (p >= SUB_BITS) { tkind = X_NOSYNC; tpos = p; break; }  // true tokens covered the whole sub-chunk
              int ck = p >> CKPT_SHIFT;
              if (ck != seen_ck) {
                seen_ck = ck;
                if (ck <= last_ck) { // the speculative pass left a checkpoint for this window
                  uint32_t c = S.ckpt[ck][tid];
                  if ((int)(c & 63u) == (p & 63)) { merged = true; kmerge = (int)((c >> 6) & 255u); skipped = c >> 14; break; }
                }
              }
              if (xkind == X_CAP && p == xpos) { merged = true; kmerge = ns; skipped = out_s; break; } // met the spec pass where its buffer filled up
              if (nh == HEAD_CAP) { in_ok = false; break; }
              uint32_t tok, no;
              int r = decode_token(S, lb, lmask, dmask, &tok, &no);
              if (r == T_EOB) { tkind = X_EOB; tpos = lb.pos(); break; }
              if (r == T_ERR) { tkind = X_ERR; tpos = p; break; }
              tok_head[(size_t)nh * PT] = tok;
              nh++;
              out_h += no;
            }
          }
        }
        STAT(if (in_ok && tid > 0) { if (merged) { g_stats.hist_nh[nh < 47 ? nh : 47]++; } else if (tkind == X_NOSYNC) g_stats.nosync_lanes++; })
        if (in_ok) {
          out_valid = out_h + (merged ? out_s - skipped : 0u);
          if (merged) { tkind = xkind; tpos = xpos; }
        }
        term = in_ok && tkind != X_NORMAL;
      }
      // ---- C: cut the round at the first lane that cannot be chained, scan output sizes
      {
        uint32_t bad = __ballot_sync(0xffffffffu, !in_ok);
        uint32_t trm = __ballot_sync(0xffffffffu, term);
        if (lane == 0) { S.warp_a[warp] = bad; S.warp_b[warp] = trm; }
      }
      __syncthreads();
      int first_bad = PT, first_term = PT;
#pragma unroll
      for (int w = PT / 32 - 1; w >= 0; w--) {
        if (S.warp_a[w]) first_bad = 32 * w + __ffs((int)S.warp_a[w]) - 1;
        if (S.warp_b[w]) first_term = 32 * w + __ffs((int)S.warp_b[w]) - 1;
      }
      int n_valid = first_term < first_bad ? first_term + 1 : first_bad;
      uint32_t mine = tid < n_valid ? out_valid : 0u;
      uint32_t incl = warp_incl_scan(mine, lane);
      __syncthreads(); // warp_a / warp_b fully read
      if (lane == 31) S.warp_c[warp] = incl;
      __syncthreads();
      uint32_t woff = 0;
#pragma unroll
      for (int w = 0; w < PT / 32; w++)
        if (w < warp) woff += S.warp_c[w];
      incl += woff;
      // output budget of the ring (it must keep min(32 KiB, produced) of history): first lane
      // whose inclusive sum exceeds it
      const uint32_t round_budget = (uint32_t)RING - umin32(S.out_pos, 32768u) - (uint32_t)RING_SLACK;
      {
        uint32_t over = __ballot_sync(0xffffffffu, tid < n_valid && incl > round_budget);
        if (lane == 0) S.warp_a[warp] = over;
      }
      __syncthreads();
      int first_over = PT;
#pragma unroll
      for (int w = PT / 32 - 1; w >= 0; w--)
        if (S.warp_a[w]) first_over = 32 * w + __ffs((int)S.warp_a[w]) - 1;
      bool cut_by_budget = first_over < n_valid;
      if (cut_by_budget) n_valid = first_over;
      const uint32_t out0 = S.out_pos;
      const uint32_t my_off = out0 + incl - mine;
      if (tid <= n_valid && tid < PT) S.lane_off[tid] = tid < n_valid ? my_off : 0u;
      __syncthreads();
      if (tid == n_valid - 1) {
        // this lane closes the round
        uint32_t round_out = incl - 0u; // inclusive sum up to the last valid lane
        S.round_out = round_out;
        S.lane_off[n_valid] = out0 + round_out;
        int nx;
        int64_t nb;
        if (!cut_by_budget && first_term < first_bad) {
          // closed by this lane's own exit
          nb = base_bit + (int64_t)SUB_BITS * tid + tpos;
          if (tkind == X_EOB) nx = S.last_block ? NEXT_FINISH : NEXT_HEADER;
          else if (tkind == X_ERR) nx = NEXT_RETRY;
          else nx = NEXT_CONTINUE; // X_CAP / X_NOSYNC: resume right after the last token
        } else {
          // the next lane could not be chained (or did not fit the budget): it opens the next round
          // at this lane's (normal) exit
          nb = base_bit + (int64_t)SUB_BITS * tid + tpos;
          nx = NEXT_CONTINUE;
        }
        if ((int64_t)out0 + round_out > S.dst_cap) nx = NEXT_RETRY; // Unexpected_end_of_output: exact point needed
        S.base_bit = nb;
        S.next = nx;
        S.n_valid = n_valid;
        STAT(g_stats.rounds++; g_stats.valid_lanes += n_valid;
             if (cut_by_budget) g_stats.cut_budget++; else if (first_term < first_bad) g_stats.cut_term[tkind]++; else g_stats.cut_bad++;)
      }
      __syncthreads();
      if (S.next == NEXT_RETRY) break;
      const uint32_t round_out = S.round_out;
      // ---- D: replay tokens into the ring
      // make sure earlier bulk stores finished reading the ring region this round overwrites
      if (tid == 0) { tma_store_wait_read(); S.abort_round = 0; }
      // bucket map: lane owning the first byte of every 64-byte bucket of this round's output
      if (tid < n_valid) {
        uint32_t a = my_off - out0, b = a + mine;
        for (uint32_t q = (a + (1u << BUCKET_SHIFT) - 1u) >> BUCKET_SHIFT; (q << BUCKET_SHIFT) < b; q++) S.bucket[q] = (uint8_t)tid;
        S.progress[tid] = my_off;
      }
      __syncthreads();
      {
        const uint32_t bias = S.ring_bias;
        bool active = tid < n_valid;
        const int total_tok = active ? nh + (merged ? ns - kmerge : 0) : 0;
        STAT(if (active) { g_stats.head_tokens += nh; g_stats.used_tokens += total_tok; } g_stats.spec_tokens += ns;)
        int ti = 0;                     // token index over head ++ spec[kmerge..]
        uint32_t pos = my_off;          // stream-relative output position
        uint32_t rem = 0, src = 0;      // pending match copy
        bool waiting = false;           // the pending match still has to see its source bytes
        uint32_t need_hi = 0;
        while (__any_sync(0xffffffffu, active)) {
          if (active) {
            STAT(g_stats.d_iters++; if (rem && waiting) g_stats.d_wait_iters++;)
            if (S.abort_round) { active = false; continue; }
            if (rem == 0) {
              if (ti == total_tok) {
                __threadfence_block();
                S.progress[tid] = pos;
                active = false;
              } else {
                uint32_t t = ti < nh ? tok_head[(size_t)ti * PT] : tok_spec[(size_t)(kmerge + (ti - nh)) * PT];
                ti++;
                if (!tok_is_match(t)) {
                  S.ring[(pos + bias) & RING_MASK] = (uint8_t)(t & 0xffu);
                  pos++;
                } else {
                  uint32_t d = ((t >> 8) & 0x7fffu) + 1u;
                  if (d > pos) { // Invalid distance (de.ml:1700): the general kernel reports it
                    S.abort_round = 1;
                    active = false;
                  } else {
                    rem = (t & 0xffu) + 3u;
                    src = pos - d;
                    // source bytes this match needs from outside itself: [src, min(src+rem, pos))
                    need_hi = umin32(src + rem, pos);
                    waiting = need_hi > out0 && src < my_off; // touches bytes other lanes write this round
                    __threadfence_block();
                    S.progress[tid] = pos; // everything before this match is in the ring
                  }
                }
              }
            }
            if (rem) {
              bool ready = true;
              if (waiting) {
                uint32_t q_hi = umin32(need_hi, my_off);
                uint32_t q_lo = umax32(src, out0);
                if (q_lo < q_hi) {
                  int v = S.bucket[(q_hi - 1u - out0) >> BUCKET_SHIFT];
                  // bucket[] names the owner of the bucket's first byte; walk up to the owner of q_hi-1
                  while (S.lane_off[v + 1] <= q_hi - 1u) v++;
                  ready = S.progress[v] >= q_hi;
                  while (ready && S.lane_off[v] > q_lo) {
                    v--;
                    ready = S.progress[v] >= S.lane_off[v + 1];
                  }
                }
                if (ready) { waiting = false; __threadfence_block(); }
              }
              if (ready) {
                uint32_t n = umin32(rem, 8u);
                for (uint32_t i = 0; i < n; i++) S.ring[(pos + i + bias) & RING_MASK] = S.ring[(src + i + bias) & RING_MASK];
                pos += n;
                src += n;
                rem -= n;
              }
            }
          }
        }
        __syncthreads();
        if (S.abort_round) {
          if (tid == 0) S.next = NEXT_RETRY;
          __syncthreads();
          break;
        }
      }
      // ---- E: drain finished bytes, fold Adler-32
      {
        const uint32_t new_pos = out0 + round_out;
        const uint32_t bias = S.ring_bias;
        if (fmt == DCMP_FMT_ZLIB && round_out) {
          // granules of 16 ring bytes covering [out0, new_pos); bytes outside are masked to zero
          uint32_t g_lo = (out0 + bias) & ~15u, g_hi = (new_pos + bias + 15u) & ~15u; // biased coordinates
          unsigned long long s_sum = 0, b_sum = 0;
          for (uint32_t g = g_lo + 16u * tid; g < g_hi; g += 16u * PT) {
            uint4 q = *(const uint4 *)(S.ring + (g & RING_MASK));
            uint32_t wv[4] = {q.x, q.y, q.z, q.w};
            uint32_t s = 0, b = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              uint32_t w = wv[k];
              // mask bytes outside [out0+bias, new_pos+bias)
              uint32_t byte0 = g + 4u * k;
              uint32_t m = 0;
#pragma unroll
              for (int j = 0; j < 4; j++) {
                uint32_t bp = byte0 + j;
                if (bp >= out0 + bias && bp < new_pos + bias) m |= 0xffu << (8 * j);
              }
              w &= m;
              s += __dp4a(w, 0x01010101u, 0u);
              // weight of byte j of word k: distance to the granule end = 16 - (4k + j)
              b += __dp4a(w, (uint32_t)(((16 - 4 * k) & 0xff) | (((15 - 4 * k) & 0xff) << 8) | (((14 - 4 * k) & 0xff) << 16) |
                                        (((13 - 4 * k) & 0xff) << 24)),
                          0u);
            }
            // bytes after this granule's end inside the round (may be negative by < 16 for the last granule)
            long long after = (long long)(new_pos + bias) - (long long)(g + 16u);
            s_sum += s;
            b_sum += (unsigned long long)((long long)b + (long long)s * after);
          }
          for (int o = 16; o > 0; o >>= 1) {
            s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o);
            b_sum += __shfl_xor_sync(0xffffffffu, b_sum, o);
          }
          if (lane == 0) { S.adl_s[warp] = s_sum; S.adl_b[warp] = b_sum; }
        }
        __syncthreads();
        if (tid == 0) {
          if (fmt == DCMP_FMT_ZLIB && round_out) {
            unsigned long long s_sum = 0, b_sum = 0;
            for (int w = 0; w < PT / 32; w++) { s_sum += S.adl_s[w]; b_sum += S.adl_b[w]; }
            const unsigned long long M = 65521ull;
            // (A,B) after appending a segment with byte sum s and suffix-weighted sum b of length n:
            // B' = B + n*A + b ; A' = A + s
            unsigned long long a = S.adler_a, b = S.adler_b;
            b = (b + (round_out % M) * a + b_sum % M) % M;
            a = (a + s_sum % M) % M;
            S.adler_a = (uint32_t)a;
            S.adler_b = (uint32_t)b;
          }
          // bulk-store the 16-byte aligned body of [flushed, new_pos); ragged ends wait for later / the finish
          fence_proxy_async();
          uint32_t lo = S.flushed, hi = new_pos;
          uintptr_t ga = (uintptr_t)S.dst;
          uint32_t lo_al = (uint32_t)((((ga + lo) + 15) & ~(uintptr_t)15) - ga);
          uint32_t hi_al = (uint32_t)(((ga + hi) & ~(uintptr_t)15) - ga);
          if (S.next == NEXT_FINISH) hi_al = hi_al; // tail bytes handled below by all threads
          if (hi_al > lo_al && lo_al >= lo) {
            // head bytes [lo, lo_al) are stored by the finish path or were stored already
            if (lo == 0 && lo_al > 0) {
              for (uint32_t i = 0; i < lo_al && i < hi; i++) S.dst[i] = S.ring[(i + bias) & RING_MASK];
            }
            uint32_t a = lo_al;
            while (a < hi_al) {
              uint32_t ri = (a + bias) & RING_MASK;
              uint32_t n = umin32(hi_al - a, (uint32_t)RING - ri); // do not cross the ring end
              tma_store_1d(S.dst + a, S.ring + ri, n);
              a += n;
            }
            tma_store_commit();
            S.flushed = hi_al;
          }
          S.out_pos = new_pos;
        }
        __syncthreads();
      }
    } // rounds

    // ---------------------------------------------------------------- finish the stream
    __syncthreads();
    {
      const int64_t k = S.stream;
      if (S.next == NEXT_FINISH) {
        // ragged tail (and a head shorter than one granule) by plain stores
        const uint32_t bias = S.ring_bias;
        for (uint32_t i = S.flushed + tid; i < S.out_pos; i += PT) S.dst[i] = S.ring[(i + bias) & RING_MASK];
        if (tid == 0) {
          tma_store_wait_all();
          int st = DCMP_OK;
          int64_t c = S.base_bit;
          int64_t cons = (c + 7) >> 3;
          if (c > 8 * S.src_len) st = DCMP_INTERNAL_RETRY; // decoded past the end: reference quirk, general kernel
          if (fmt == DCMP_FMT_ZLIB && st == DCMP_OK) {
            const uint8_t *tr = S.src + cons; // lib/zl.ml:409-415: big-endian Adler-32 right after the deflate data
            uint32_t want = ((uint32_t)tr[0] << 24) | ((uint32_t)tr[1] << 16) | ((uint32_t)tr[2] << 8) | (uint32_t)tr[3];
            uint32_t got = (S.adler_b << 16) | S.adler_a;
            if (want != got) st = DCMP_INVALID_CHECKSUM;
            else cons += 6;
            if (st != DCMP_OK) cons = 0;
          }
          consumed[k] = cons;
          produced[k] = S.out_pos;
          status[k] = st;
        }
      } else {
        if (tid == 0) {
          tma_store_wait_all();
          consumed[k] = 0;
          produced[k] = 0;
          status[k] = DCMP_INTERNAL_RETRY;
        }
      }
    }
  }
  if (tid == 0) tma_store_wait_all();
}

size_t inflate_parallel_smem_bytes() { return sizeof(ParSmem) + 128; }
// bytes of token scratch one CTA needs (global memory, stays L2-resident)
size_t inflate_parallel_scratch_bytes_per_cta() { return (size_t)TOK_SLOTS * PT * sizeof(uint32_t); }

#ifndef DCMP_EMU
int launch_inflate_parallel(int fmt, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                            uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                            int64_t *produced, int32_t *status, unsigned long long *work_counter, uint32_t *tok_scratch,
                            int ctas, cudaStream_t stream) {
  static bool configured = false;
  size_t smem = inflate_parallel_smem_bytes();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(inflate_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  if (ctas > n) ctas = (int)n;
  inflate_parallel_kernel<<<(unsigned)ctas, PT, smem, stream>>>(fmt, n, src_base, src_off, src_len, dst_base, dst_off,
                                                               dst_cap, consumed, produced, status, work_counter,
                                                               tok_scratch);
  return 0;
}
#endif

} // namespace dcmp
