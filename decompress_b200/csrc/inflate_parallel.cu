// decompress_b200/csrc/inflate_parallel.cu — sub-chunk-parallel inflate, one CTA per stream.
//
// This is the hot path of BASELINE config C2 (batches of independent dynamic-Huffman
// streams).  It replaces the literal/length/distance state machine of De.Inf
// (lib/de.ml:1054-1261, one-shot form 1667-1712), the window (WInf, de.ml:436-508) and
// the table construction (Inf.huffman, de.ml:523-638) with a design that has no
// counterpart in the reference (which is strictly serial).  Nothing in it waits on
// another lane: every phase is data-parallel over the PT lanes of the CTA.
//
//   round = PT lanes x SUB_BITS-bit sub-chunks of the compressed stream (an odd number of
//           32-bit words per lane, so the lanes of a warp read the TMA-staged window
//           without shared-memory bank conflicts)
//   A  speculative decode   every lane starts WARMUP_BITS before its sub-chunk, at an arbitrary bit
//                           offset, and throws away what it decodes there (Huffman streams
//                           self-synchronise: most lanes are on the true token boundaries when their
//                           own bits begin); from bit 0 on it keeps 32-bit tokens (literal | len,dist),
//                           written to an L2-resident scratch as 128-bit batches, and a checkpoint
//                           (position, token index, bytes so far) per 2^CKPT_SHIFT-bit window;
//   B  head re-decode       lane t restarts at the true entry point handed over by lane t-1 and
//                           decodes until a token start coincides with a checkpoint of the
//                           speculative pass (of its own sub-chunk, or of a later lane's when it
//                           crosses its own without meeting one); from there on the speculative
//                           tokens are the true ones.  With the warm-up ~5 % of the lanes need it;
//   C  chain + scan         lanes covered by a crossing head are absorbed (pointer doubling); the round
//                           is closed at the first lane that cannot be validated (end of block, token
//                           buffer full); an exclusive scan of emitted byte counts gives every lane
//                           its output offset; the output leaves in slices that fit the symbol buffer;
//   D  symbol replay        every lane replays its tokens into a 16-bit symbol buffer at its output
//                           offset.  A back-reference into the 32 KiB history (earlier rounds or
//                           slices: the stream's own output, read back through L2 with .cg loads —
//                           keeping it out of shared memory buys a third CTA per SM;
//                           DCMP_HIST_GLOBAL=0 keeps a shared byte ring instead) is copied; a
//                           back-reference into bytes produced in this slice — by another lane or
//                           by this one — is not waited for: the lane stores a marker
//                           256 + (index of the source symbol);
//   E  resolve + drain      16 symbols per thread: markers are chased to a literal (markers only
//                           point backwards, chains are short), the 16 bytes leave through one
//                           128-bit store to HBM; Adler-32 is folded over the same registers
//                           (symbol_drain.cuh).
// Block headers: the serial part of a dynamic header is parsed ahead of time — the next stream's by the
// CTA's last warp during phase B, the next block's during phase D when that warp has no lanes to replay;
// stored payloads are copied straight from the input to the output (raw_copy).
//
// Anything that is not the happy path — stored blocks, malformed input, exotic code sets, an
// output buffer that is too small — makes the CTA give the stream up with status
// DCMP_INTERNAL_RETRY; the host-side dispatcher (engine.cu) then runs the general
// one-warp-per-stream kernel (inflate_serial.cu), which reproduces the reference's exact
// error behaviour, on exactly those streams.  Both kernels are device code; there is no
// CPU path.
#include "inflate_tables.cuh"
#include "tma.cuh"

namespace dcmp {

#ifndef DCMP_PHASE_A_INLINE
#define DCMP_PHASE_A_INLINE __forceinline__
#endif
#ifndef DCMP_PT
#define DCMP_PT 256
#endif
#ifndef DCMP_SUB_WORDS
#define DCMP_SUB_WORDS 9
#endif

constexpr int PT = DCMP_PT;               // threads (= lanes = sub-chunks per round)
constexpr int NWARP = PT / 32;
constexpr int SUB_WORDS = DCMP_SUB_WORDS; // sub-chunk size in 32-bit words (odd: conflict-free lane stride)
constexpr int SUB_BITS = 32 * SUB_WORDS;
constexpr int SPEC_CAP = ((SUB_BITS / 3) + 3) & ~3; // tokens a lane may produce speculatively (< 256: checkpoint field)
constexpr int HEAD_CAP = 96;              // tokens a lane may produce while re-synchronising (p99 of the distance is ~55 on text)
constexpr int TOK_SLOTS = SPEC_CAP + HEAD_CAP; // per lane, in the global scratch
constexpr int LANE_BYTE_CAP = 4096;       // bytes a lane may produce per pass (checkpoint field: 19 bits)
#ifndef DCMP_WARMUP_BITS
#define DCMP_WARMUP_BITS 768
#endif
constexpr int WARMUP_BITS = DCMP_WARMUP_BITS; // a lane's speculative pass starts this far inside the previous sub-chunk
#ifndef DCMP_CKPT_SHIFT
#define DCMP_CKPT_SHIFT 4
#endif
#ifndef DCMP_D_PAIRS
#define DCMP_D_PAIRS 0                    // phase D: 1 = two consecutive literals of a lane are replayed in one step: a fifth fewer
                                          // warp-steps on text, but 151.0 vs 154.6 GB/s on C2 (150.4 with a three-deep token
                                          // queue): the two fetch paths diverge inside the warp
#endif
constexpr int CKPT_SHIFT = DCMP_CKPT_SHIFT; // one checkpoint per 2^CKPT_SHIFT-bit window of the sub-chunk
constexpr uint32_t CKPT_NONE = 0x80000000u; // no token of the speculative pass starts in this window (a long match can skip one)
static_assert(CKPT_SHIFT >= 3 && CKPT_SHIFT <= 5, "");
constexpr int NCKPT = (SUB_BITS >> CKPT_SHIFT) + 1;
#ifndef DCMP_RSYM_K
#define DCMP_RSYM_K 26
#endif
constexpr int RSYM = DCMP_RSYM_K * 1024;  // symbols one round may produce (+ up to 15 of alignment bias)
constexpr int SYM_SLOTS = RSYM + 32;
constexpr int WIN_WORDS = 3 + SUB_WORDS * PT + 4;        // staged input per round
constexpr int WIN_BYTES = (WIN_WORDS * 4 + 15) & ~15;
constexpr int HDR_BYTES = 624 + 16;       // staged input for a block header (the longest one: 14 + 19*3 + 316*14 bits = 562 bytes)
constexpr int LIT_ROOT = 10;
constexpr int DIST_ROOT = 9;
constexpr int LIT_SUB = 320;              // second-level entries (codes longer than the root; zlib's `enough 286 10 15` = 1332 = 1024 + 308);
constexpr int DIST_SUB = 128;             // a code set that needs more goes to the general kernel
constexpr int LIT_N = (1 << LIT_ROOT) + LIT_SUB, DIST_N = (1 << DIST_ROOT) + DIST_SUB;
constexpr uint32_t DIST_OFF = LIT_N;      // both tables live in one array: the distance table starts here
// Decode-table entry of this kernel: one shared-memory load per code, fields placed for the hot loop.
//  [4:0]   bits to drop: code length + extra bits     [5]  LEN: a match length — the next code is a distance
//  [11:8]  code length (LINK: index bits of the sub-table)
//  [30:16] base: literal byte | match length base (3..258) | distance base (1..24577) | LINK: index of the sub-table
//  [31] rare: EOB [6], LINK [7] or an unassigned slot — one sign test takes all three off the hot path
constexpr uint32_t PE_RARE = 1u << 31, PE_LEN = 1u << 5, PE_EOB = 1u << 6, PE_LINK = 1u << 7;
constexpr uint32_t PE_INVALID = PE_RARE;
__host__ __device__ constexpr uint32_t pe_make(uint32_t base, uint32_t extra, uint32_t len) {
  return (base << 16) | (len << 8) | (len + extra);
}
constexpr int DCMP_INTERNAL_RETRY = 0x7f; // never leaves the library
static_assert(SUB_WORDS % 2 == 1, "odd lane stride keeps the window reads conflict-free");
static_assert(PT % 64 == 0 && PT <= 1024, "the two halves of the CTA build one decode table each");
static_assert(SPEC_CAP < 256 && LANE_BYTE_CAP + 258 < (1 << 19), "checkpoint fields");

constexpr int MAX_CROSS = 6;              // sub-chunk boundaries one head decode may cross
// Per-round metadata that is dead before phase D writes symbols; it overlays the symbol buffer.
struct LaneDesc {          // what a lane replays in phase D, written by the lane that owns the head
  uint16_t owner;          // lane whose head slots hold the tokens
  uint8_t h0, h1;          // head tokens [h0, h1)
  uint16_t out_head;       // bytes they produce
  uint8_t spec_k;          // speculative tokens from spec_k on (if has_spec)
  uint8_t flags;           // bit0 has_spec, bit1 terminal here (head ended: kind in bits 4..6)
  uint16_t spec_skip;      // bytes the speculative pass produced before spec_k
  uint16_t term_pos;       // terminal position in the OWNER's sub-chunk coordinates
};
struct RoundMeta {
  uint32_t ckpt[NCKPT][PT];
  uint16_t cross_nh[PT][MAX_CROSS];   // head pass: tokens / bytes produced before entering the j-th next sub-chunk
  uint16_t cross_out[PT][MAX_CROSS];
  uint16_t jump[2][PT];
  uint8_t mark[PT];
  LaneDesc desc[PT];
};
#ifndef DCMP_WIN_OVERLAY
#define DCMP_WIN_OVERLAY 0                // 1: the staged input window too lives in the symbol buffer (behind the round metadata):
#endif                                    //    it is dead once phase D starts; buys the shared memory for a fourth CTA per SM
constexpr int WIN_OVL_OFF = (int)((sizeof(RoundMeta) + 15) & ~(size_t)15); // byte offset of the overlaid window in `sym`
static_assert(sizeof(RoundMeta) <= SYM_SLOTS * 2, "round metadata overlays the symbol buffer");
static_assert(!DCMP_WIN_OVERLAY || WIN_OVL_OFF + WIN_BYTES + 16 <= SYM_SLOTS * 2, "metadata + window overlay the symbol buffer");
static_assert(RSYM + 16 + 256 <= 65536, "markers are 16-bit");
static_assert(HDR_BYTES <= WIN_BYTES, "");

#ifdef DCMP_EMU
struct EmuStats {
  long hist_nh[48];
  long nosync_lanes, rounds, valid_lanes, cut_term[5], cut_bad, cut_budget, head_tokens, spec_tokens, used_tokens;
  long d_iters, d_lane_max_iters, markers, hist_bytes, own_bytes, chase_hops, chase_max, out_bytes, hdr_rounds, retries, fix_passes;
};
static EmuStats g_stats;
#define STAT(x) x
#else
#define STAT(x)
#endif
} // namespace dcmp
#include "symbol_drain.cuh"
#if !DCMP_HIST_GLOBAL
#error "inflate_parallel.cu reads the history from the output buffer (DCMP_HIST_GLOBAL=1)"
#endif
namespace dcmp {

enum : int { X_NORMAL = 0, X_EOB = 1, X_CAP = 2, X_ERR = 3, X_NOSYNC = 4 };
enum : int { NEXT_CONTINUE = 0, NEXT_HEADER = 1, NEXT_FINISH = 2, NEXT_RETRY = 3, NEXT_STORED = 4 };

// checkpoint: [4:0] low bits of the position of the first token start inside the window,
//             [12:5] index of that token, [30:13] bytes emitted before it, [31] none
__device__ __forceinline__ bool tok_is_match(uint32_t t) { return (t >> 23) & 1u; }
// one predicated 16-bit shared-memory store, without a branch
__device__ __forceinline__ void st_sym_if(bool pred, uint16_t *addr, uint32_t v) {
#ifdef DCMP_EMU
  if (pred) *addr = (uint16_t)v;
#else
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %0, 0;\n @p st.shared.u16 [%1], %2;\n}\n" ::"r"((uint32_t)pred),
               "r"((uint32_t)__cvta_generic_to_shared(addr)), "h"((uint16_t)v)
               : "memory");
#endif
}
// A token leaves for the scratch as soon as it is decoded (nothing to carry from step to step).  The scratch of a
// CTA is slot-major — token k of lane t at [k][t] — so the lanes of a warp that are at the same token count write
// (phase A) and read (phase D) one line.
struct HuffMeta {      // canonical-code description of one alphabet
  uint16_t first[16];  // first code (MSB-first) of each length
  uint16_t cnt[16];
  int root, maxl;
};

struct Prefetch {          // the stream this CTA decodes next, claimed and opened ahead of time by one warp
  int64_t stream;          // claimed index (>= nstreams: none left); -2: nothing claimed
  const uint8_t *src; int64_t src_len; uint8_t *dst; int64_t dst_cap; int st;
  int hdr;                 // 1: its first block is dynamic and the header is parsed: lens, hlit, hdist, last_block, base_bit
  int last_block;
  uint32_t hlit, hdist;
  int64_t base_bit;
  uint32_t codes[CODES_TABLE_SIZE];
  uint8_t lens[MAX_LENS];
  __align__(16) uint8_t win[HDR_BYTES];
};

struct ParSmem {
  static constexpr int THREADS = PT;
  __align__(16) uint16_t sym[SYM_SLOTS];   // phase D/E: symbols of the round; phase A/B: checkpoints [NCKPT][PT] u32
#if !DCMP_HIST_GLOBAL
  __align__(16) uint8_t ring[HIST];        // last 32 KiB of output, indexed by (global address & HIST_MASK)
#endif
#if DCMP_WIN_OVERLAY
  __align__(16) uint32_t nh_win[HDR_BYTES / 4 + 4]; // the bytes of a next-block header, saved before phase D overwrites the window
#else
  __align__(16) uint32_t win[WIN_BYTES / 4 + 4]; // staged compressed input (TMA destination)
#endif
  uint32_t tbl[LIT_N + DIST_N];  // literal/length table (root, then the sub-tables of codes longer than the root), distance table
  uint8_t lens[MAX_LENS];
  HuffMeta ml, md;
  uint32_t run_cnt[2][16];       // per table under construction (literal/length, distance)
  // per-lane results of phase A, read by phase B
  uint16_t exit_pos[PT];         // spec pass: bit offset (own sub-chunk coordinates) after the last token
  uint8_t exit_kind[PT];
  int8_t last_ck[PT];            // spec pass: last checkpoint window written
  uint8_t spec_ns[PT];           // spec pass: tokens produced
  uint16_t spec_out[PT];         // spec pass: bytes produced
  uint32_t m_owner[NWARP];       // lanes whose own head decode counts (not covered by an earlier lane's head)
  int pass_owner;                // lane whose head ran past the end of the window (-1: none)
  int pass_pos;                  // its exit (own sub-chunk coordinates)
  uint32_t warp_c[NWARP];
  unsigned long long adl_s[NWARP];
  long long adl_b[NWARP];
  uint32_t hdr_hlit, hdr_hdist, tbl_ok2[2], sub_alloc2[2];
  int first_bad, first_term;
  uint32_t slice_end;
  // stream state (CTA-uniform)
  int64_t stream;                // index of the stream being decoded, -1 = none
  const uint8_t *src; int64_t src_len;  // raw deflate bytes
  uint8_t *dst; int64_t dst_cap;
  int64_t base_bit;              // next bit of the deflate stream to decode
  uint32_t out_pos;              // bytes produced so far
  int last_block, next, status;
  uint32_t stored_rem;           // payload bytes of the current stored block still to copy
  int n_valid;                   // lanes accepted in this round
  uint32_t round_out;            // bytes produced by this round
  uint32_t adler_a, adler_b;
  int abort_round;               // a lane met an invalid distance: the general kernel reports it
  int any_cross;                 // a head of this round crossed into another lane's sub-chunk
  volatile uint32_t a_done[NWARP]; // round number up to which a warp has finished phase A (its lanes' exits and checkpoints are in place)
  int hdr_ready;                 // the first header of this stream was parsed ahead (S.pf)
  int nh_ready;                  // the header of the stream's NEXT block was parsed during phases D/E: lens, nh_*
  int nh_last;
  uint32_t nh_hlit, nh_hdist;
  int64_t nh_base_bit;
  int pf_done;                   // the next stream has been claimed during this one
  Prefetch pf;
  __align__(8) uint64_t mbar;
};
#ifndef DCMP_CTAS_PER_SM
#define DCMP_CTAS_PER_SM (DCMP_HIST_GLOBAL ? 3 : 2)
#endif
// 227 KiB of shared memory per SM, 1 KiB reserved per resident CTA
static_assert(sizeof(ParSmem) + 128 + 1024 <= 227 * 1024 / DCMP_CTAS_PER_SM, "shared memory per CTA for the intended residency");

__device__ __forceinline__ uint32_t *win_of(ParSmem &S) {
#if DCMP_WIN_OVERLAY
  return reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(S.sym) + WIN_OVL_OFF);
#else
  return S.win;
#endif
}

// ------------------------------------------------------------------------------------------
// The decoder of phases A and B.  One STEP = one table lookup: a literal/length code or — when the step before
// met a match length — a distance code.  Every step runs the same instructions whichever of the two it is, so the
// lanes of a warp stay converged (a literal takes one step, a match two; both cover 8..10 bits of input on text),
// where a decoder that finishes a whole token per iteration runs the literal and the match path one after the
// other.  The bit cursor is a bit index into the staged window: the next 32 bits are two shared-memory words and
// a funnel shift (no 64-bit hold register, no refill branch).
struct Cursor {
  int q;            // bit index into S.win of the next code
  uint32_t tb;      // 0: the next code is a literal/length code; DIST_OFF: a distance code
  uint32_t mask;    // root-table index mask of that table
  uint32_t pend;    // match length waiting for its distance (tb != 0)
  __device__ __forceinline__ void start(int q0, uint32_t lmask) { q = q0; tb = 0; mask = lmask; pend = 0; }
  __device__ __forceinline__ bool at_token() const { return tb == 0; }
};
enum : int { STEP_NONE = 0, STEP_TOKEN = 1, STEP_EOB = 2, STEP_ERR = 3 };
// a select the compiler does not turn into shifts and masks
__device__ __forceinline__ uint32_t sel_u32(uint32_t nz, uint32_t a, uint32_t b) { // nz != 0 ? a : b
#ifdef DCMP_EMU
  return nz ? a : b;
#else
  uint32_t r;
  asm("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n selp.u32 %0, %2, %3, p;\n}\n" : "=r"(r) : "r"(nz), "r"(a), "r"(b));
  return r;
#endif
}

// token: [23] match flag; match: [22:8] distance-1, [7:0] length-3; literal: [7:0] byte
template <bool KEEP>
__device__ __forceinline__ int dec_step(const uint32_t *__restrict__ tbl, const uint32_t *__restrict__ win, Cursor &c,
                                        const uint32_t lmask, const uint32_t dmask, const int lroot, const int droot,
                                        uint32_t &tok, uint32_t &nout) {
  const uint32_t *wp = win + (c.q >> 5);
  const uint32_t lo = __funnelshift_r(wp[0], wp[1], (uint32_t)c.q);
  uint32_t e = tbl[c.tb + (lo & c.mask)];
  if ((int32_t)e < 0) { // rare: end of block, a code longer than the root table, an unassigned slot
    if (!(e & PE_LINK)) {
      if (e & PE_EOB) { c.q += (int)(e & 31u); return STEP_EOB; }
      return STEP_ERR;
    }
    e = tbl[(e >> 16 & 0x7fffu) + ((lo >> (c.tb ? droot : lroot)) & ~(~0u << ((e >> 8) & 15u)))];
    if ((int32_t)e < 0) {
      if (e & PE_EOB) { c.q += (int)(e & 31u); return STEP_EOB; }
      return STEP_ERR;
    }
  }
  c.q += (int)(e & 31u);
  const uint32_t is_len = e & PE_LEN; // (never set in the distance table)
  if (KEEP) {
    // base + extra bits: the low [4:0] bits of lo, without the code ([12:8], bits 12 is clear: a wrapping shift takes
    // the field as it is)
    const uint32_t v = (e >> 16) + __funnelshift_r(lo & ~__funnelshift_l(0u, ~0u, e), 0u, e >> 8);
    tok = c.tb ? ((1u << 23) | ((v - 1u) << 8) | (c.pend - 3u)) : v;
    nout = c.tb ? c.pend : 1u;
    c.pend = v;
  }
  c.tb = sel_u32(is_len, DIST_OFF, 0u);
  c.mask = sel_u32(is_len, dmask, lmask);
  return is_len ? STEP_NONE : STEP_TOKEN;
}

// ------------------------------------------------------------------------------------------
// Parallel construction of a two-level decode table from code lengths (the shape of Inf.huffman,
// de.ml:523-638: a root table indexed by the next `root` bits, sub-tables for longer codes; here the
// root is wider and every entry carries what the hot loop needs).  Accepts complete codes (Kraft sum
// exactly 1) and, for the literal/length and distance alphabets, the two incomplete shapes the
// reference accepts (de.ml:544-550): no code at all and a single 1-bit code; their unassigned slots
// are PE_INVALID, which sends a stream that really uses them to the general kernel.  Returns false
// (CTA-uniform) otherwise, or when the sub-tables do not fit.  `scratch_all`: 2 x ((1 << LIT_ROOT) + MAX_LENS) words.
__device__ __forceinline__ uint32_t pe_symbol(int kind, int s, uint32_t l) {
  if (kind == KIND_DISTS) {
    const uint32_t b = c_dist_base[s & 31];
    return b ? pe_make(b, c_dist_extra[s & 31], l) : PE_INVALID; // symbols 30/31: "Invalid distance code" (de.ml:321-325)
  }
  if (s < 256) return pe_make((uint32_t)s, 0, l);
  if (s == 256) return PE_RARE | PE_EOB | pe_make(0, 0, l);
  const int i = (s - 257) & 31; // 286/287 decode as a length-3 match (de.ml:293-311)
  return PE_LEN | pe_make(c_len_base[i], c_len_extra[i], l);
}
// Both tables of a block are built at the same time: the lower half of the CTA builds the literal/length table, the
// upper half the distance table — same code, same barriers, half as many of them as one after the other.
__device__ bool build_fast_tables(ParSmem &S, const uint8_t *lens_l, int n_l, const uint8_t *lens_d, int n_d, uint32_t *scratch_all,
                                  int tid) {
  constexpr int NT = PT / 2;
  const int half = tid >= NT ? 1 : 0, t = tid - half * NT;
  const int kind = half ? KIND_DISTS : KIND_LENS;
  const uint8_t *lens = half ? lens_d : lens_l;
  const int n = half ? n_d : n_l;
  uint32_t *tbl = S.tbl + (half ? DIST_OFF : 0u);
  const int root_max = half ? DIST_ROOT : LIT_ROOT, sub_cap = half ? DIST_SUB : LIT_SUB;
  HuffMeta &m = half ? S.md : S.ml;
  uint32_t *scratch = scratch_all + half * ((1 << LIT_ROOT) + MAX_LENS); // (1 << root_max) + n words per table
  uint32_t *const code_of = scratch + (1 << root_max); // per symbol: bit-reversed code | len << 16
  uint32_t *run_cnt = S.run_cnt[half];
  __syncthreads();
  if (t < 16) run_cnt[t] = 0;
  __syncthreads();
  for (int s = t; s < n; s += NT) {
    int l = lens[s];
    if (l) atomicAdd(&run_cnt[l], 1u);
  }
  __syncthreads();
  if (t == 0) {
    int left = 1, maxl = 0;
    uint32_t code = 0;
    m.cnt[0] = 0; m.first[0] = 0;
    for (int l = 1; l <= 15; l++) {
      uint32_t c = run_cnt[l];
      code = (code + (l > 1 ? run_cnt[l - 1] : 0)) << 1;
      m.first[l] = (uint16_t)code;
      m.cnt[l] = (uint16_t)c;
      left = (left << 1) - (int)c;
      if (left < 0) left = -100000; // over-subscribed
      if (c) maxl = l;
    }
    bool ok = left == 0 && maxl >= 1;
    if (!ok && kind != KIND_CODES && (maxl == 0 || (maxl == 1 && left > 0))) { // none, or one 1-bit code
      ok = true;
      if (maxl == 0) maxl = 1;
    }
    m.maxl = maxl;
    m.root = maxl < root_max ? maxl : root_max;
    S.tbl_ok2[half] = ok ? 1u : 0u;
    S.sub_alloc2[half] = 0;
  }
  __syncthreads();
  if (!(S.tbl_ok2[0] && S.tbl_ok2[1])) return false; // (CTA-uniform)
  if (t < 16) run_cnt[t] = 0;
  const int root = m.root;
  const uint32_t rmask = (1u << root) - 1u;
  for (int i = t; i < (1 << root); i += NT) { tbl[i] = PE_INVALID; scratch[i] = 0; }
  __syncthreads();
  // canonical rank of every symbol among the symbols of its length (by symbol index):
  // the first warp of the half walks the alphabet 32 symbols at a time, __match_any_sync groups equal lengths.
  if (t < 32) {
    for (int base = 0; base < n; base += 32) {
      int s = base + t;
      uint32_t l = s < n ? lens[s] : 0u;
      uint32_t grp = __match_any_sync(0xffffffffu, l);
      uint32_t rank = (uint32_t)__popc(grp & ((1u << t) - 1u));
      uint32_t before = run_cnt[l];
      __syncwarp();
      if (l && rank == 0) run_cnt[l] = before + (uint32_t)__popc(grp);
      __syncwarp();
      if (s < n) {
        if (l) {
          uint32_t code = (uint32_t)m.first[l] + before + rank;
          uint32_t r = __brev(code) >> (32 - l);
          code_of[s] = r | (l << 16);
          if ((int)l > root) atomicMax(&scratch[r & rmask], l - (uint32_t)root); // index bits its sub-table needs
        } else {
          code_of[s] = 0;
        }
      }
    }
  }
  __syncthreads();
  // place the sub-tables (any order) and link them from the root
  for (int i = t; i < (1 << root); i += NT) {
    const uint32_t bits = scratch[i];
    if (bits) {
      const uint32_t off = atomicAdd(&S.sub_alloc2[half], 1u << bits);
      if (off + (1u << bits) > (uint32_t)sub_cap) S.tbl_ok2[half] = 0;
      else {
        const uint32_t at = (1u << root) + off;
        tbl[i] = PE_RARE | PE_LINK | ((at + (half ? DIST_OFF : 0u)) << 16) | (bits << 8);
        scratch[i] = at | (bits << 16);
        for (uint32_t j = 0; j < (1u << bits); j++) tbl[at + j] = PE_INVALID;
      }
    }
  }
  __syncthreads();
  if (!(S.tbl_ok2[0] && S.tbl_ok2[1])) return false;
  for (int s = t; s < n; s += NT) {
    const uint32_t co = code_of[s];
    const uint32_t l = co >> 16;
    if (!l) continue;
    const uint32_t r = co & 0xffffu;
    const uint32_t entry = pe_symbol(kind, s, l);
    if ((int)l <= root) {
      for (uint32_t i = r; i < (1u << root); i += (1u << l)) tbl[i] = entry;
    } else {
      const uint32_t sc = scratch[r & rmask];
      const uint32_t at = sc & 0xffffu, bits = sc >> 16;
      for (uint32_t j = r >> root; j < (1u << bits); j += (1u << (l - (uint32_t)root))) tbl[at + j] = entry;
    }
  }
  __syncthreads();
  return true;
}

// ------------------------------------------------------------------------------------------
// Block header (de.ml:1788-1793, 1771-1786, 1733-1769) parsed by thread 0 from the staged
// window.
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

// thread 0 only.  Returns the bit position after the header (>0) or -1 to give the stream to the
// general kernel.
__device__ int parse_dynamic_header(uint32_t *codes, uint8_t *lens, HdrBits &hb, int64_t bits_avail, int *hlit_out, int *hdist_out) {
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
    for (uint32_t i = r; i < 128u; i += (1u << l)) codes[i] = mk_entry(K_LIT, (uint32_t)s, 0, (uint32_t)l);
  }
  const int total = hlit + hdist;
  int i = 0, prev = 0;
  while (i < total) {
    hb.refill();
    uint32_t e = codes[(uint32_t)hb.hold & 127u];
    hb.pop((int)e_len(e));
    int v = (int)e_base(e);
    if (v < 16) {
      lens[i++] = (uint8_t)v;
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
    for (int x = 0; x < rep; x++) lens[i + x] = (uint8_t)fill;
    prev = fill;
    i += rep;
  }
  if (bits_avail - hb.pos() < 0) return -1; // header ran past the end of the input
  if (lens[256] == 0) return -1;
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
// `valid` are zero-filled by the threads.  All threads call; returns after the data landed.  `phase` is the
// mbarrier's parity, kept by every thread (CTA-uniform).
__device__ void stage_window(ParSmem &S, const uint8_t *g, int bytes, int valid, int tid, uint32_t &phase) {
  // previous readers of S.win are done (callers sync before)
  if (valid > bytes) valid = bytes;
  if (valid < 0) valid = 0;
  for (int i = valid + tid * 16; i < bytes; i += PT * 16) *(uint4 *)((uint8_t *)win_of(S) + i) = make_uint4(0, 0, 0, 0);
  if (valid > 0) {
    if (tid == 0) {
      mbar_expect_tx(&S.mbar, (uint32_t)valid);
      tma_load_1d(win_of(S), g, (uint32_t)valid, &S.mbar);
    }
    mbar_wait(&S.mbar, phase & 1u);
    phase ^= 1u;
  }
  __syncthreads();
}

// Resolve stream k: where its raw DEFLATE bytes and its output window are, and whether this kernel takes it.
__device__ __forceinline__ int open_stream(int fmt, int64_t k, const uint8_t *src_base, const int64_t *src_off,
                                           const int64_t *src_len, uint8_t *dst_base, const int64_t *dst_off,
                                           const int64_t *dst_cap, const uint8_t **src_out, int64_t *len_out,
                                           uint8_t **dst_out, int64_t *cap_out) {
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
  *src_out = p;
  *len_out = n;
  *dst_out = dst_base + dst_off[k];
  *cap_out = dst_cap[k];
  if (*cap_out > 0xfff00000LL || n > 0x0fffffffLL) st = DCMP_INTERNAL_RETRY; // 32-bit positions in the kernel
  return st;
}

// One warp claims the CTA's next stream and parses the header of its first block while the other warps are
// still busy with the current stream: the serial part of a dynamic header (the Huffman-coded code lengths,
// de.ml:1733-1769) then costs the CTA nothing.  All 32 lanes call.
__device__ void prefetch_next_stream(ParSmem &S, int fmt, int64_t nstreams, const uint8_t *src_base, const int64_t *src_off,
                                     const int64_t *src_len, uint8_t *dst_base, const int64_t *dst_off,
                                     const int64_t *dst_cap, unsigned long long *work_counter, int lane) {
  Prefetch &P = S.pf;
  unsigned long long k = 0;
  if (lane == 0) {
    k = atomicAdd(work_counter, 1ull);
    P.stream = (int64_t)k;
    P.hdr = 0;
    P.st = DCMP_OK;
    S.pf_done = 1;
    if (k < (unsigned long long)nstreams)
      P.st = open_stream(fmt, (int64_t)k, src_base, src_off, src_len, dst_base, dst_off, dst_cap, &P.src, &P.src_len, &P.dst, &P.dst_cap);
  }
  k = __shfl_sync(0xffffffffu, k, 0);
  __syncwarp();
  if (k >= (unsigned long long)nstreams || P.st != DCMP_OK || P.src_len <= 0) return;
  const uint32_t A = (uint32_t)((uintptr_t)P.src & 15);
  const uint8_t *g = P.src - A;
  const int64_t vend = ((int64_t)A + P.src_len + 15) & ~(int64_t)15; // one past the last granule that holds input
  for (int i = lane * 16; i < HDR_BYTES; i += 512) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i < vend) v = __ldcs(reinterpret_cast<const uint4 *>(g + i));
    *reinterpret_cast<uint4 *>(P.win + i) = v;
  }
  __syncwarp();
  if (lane == 0) {
    const int bit0 = 8 * (int)A;
    HdrBits hb;
    hb.init(P.win, bit0, HDR_BYTES);
    const int64_t avail = 8 * P.src_len + bit0; // valid bits measured in window coordinates
    hb.refill();
    if (avail - hb.pos() >= 3) {
      const int last = (int)hb.pop(1);
      if (hb.pop(2) == 2u) {
        int hlit = 0, hdist = 0;
        const int used = parse_dynamic_header(P.codes, P.lens, hb, avail, &hlit, &hdist);
        if (used >= 0) {
          P.last_block = last;
          P.hlit = (uint32_t)hlit;
          P.hdist = (uint32_t)hdist;
          P.base_bit = used - bit0;
          P.hdr = 1;
        }
      }
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Phase A of a round, one lane: speculative decode of the lane's own sub-chunk.  Results: the tokens (scratch), a
// checkpoint per CKPT window, and exit position / kind, last checkpoint window, token and byte counts in S.
__device__ DCMP_PHASE_A_INLINE void phase_a_lane(ParSmem &S, uint32_t (*ckpt)[PT], uint32_t *tok_spec, const int q0, const int tid,
                                          const bool lane_has_input) {
  const int lroot = S.ml.root, droot = S.md.root;
  const uint32_t lmask = (1u << lroot) - 1u, dmask = (1u << droot) - 1u;
  const uint32_t *const win = win_of(S);
  int ns = 0;
  uint32_t out_s = 0;
  int xkind = X_ERR, xpos = 0;
  int last_ck = -1;
  if (lane_has_input) {
#pragma unroll
    for (int c = 0; c < NCKPT; c++) ckpt[c][tid] = CKPT_NONE; // (the buffer held symbols, or an older round's checkpoints)
    Cursor cu;
    uint32_t tok = 0, no = 0;
    // Warm-up: the pass starts WARMUP_BITS before the sub-chunk and throws away what it decodes there, so that it
    // has often found the true token boundaries by the time its own bits begin (lane 0 starts on one).
    cu.start(q0 - imin(WARMUP_BITS, tid * SUB_BITS), lmask); // (never before lane 0's start, which is a token boundary)
    while (cu.q < q0 || !cu.at_token()) {
      if (dec_step<false>(S.tbl, win, cu, lmask, dmask, lroot, droot, tok, no) >= STEP_EOB) { cu.start(q0, lmask); break; }
    }
    int tstart = 0;
    for (;;) {
      if (cu.at_token()) {
        const int p = cu.q - q0;
        tstart = p;
        // one test for the three ways out: the sub-chunk is done, the token slots or the byte budget are used up
        if (((SUB_BITS - 1 - p) | (SPEC_CAP - 1 - ns) | (LANE_BYTE_CAP - 258 - (int)out_s)) < 0) {
          xkind = p >= SUB_BITS ? X_NORMAL : X_CAP;
          xpos = p;
          break;
        }
        const int ck = p >> CKPT_SHIFT;
        if (ck != last_ck) {
          ckpt[ck][tid] = (uint32_t)(p & 31) | ((uint32_t)ns << 5) | (out_s << 13);
          last_ck = ck;
        }
      }
      const int r = dec_step<true>(S.tbl, win, cu, lmask, dmask, lroot, droot, tok, no);
      if (r == STEP_NONE) continue;
      if (r == STEP_EOB) { xkind = X_EOB; xpos = cu.q - q0; break; }
      if (r == STEP_ERR) { xkind = X_ERR; xpos = tstart; break; }
      tok_spec[ns * PT] = tok;
      ns++;
      out_s += no;
    }
  }
  S.exit_pos[tid] = (uint16_t)xpos;
  S.exit_kind[tid] = (uint8_t)xkind;
  S.last_ck[tid] = (int8_t)last_ck;
  S.spec_ns[tid] = (uint8_t)ns;
  S.spec_out[tid] = (uint16_t)out_s;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT, DCMP_CTAS_PER_SM)
inflate_parallel_kernel(int fmt, int64_t nstreams, const uint8_t *src_base, const int64_t *src_off,
                        const int64_t *src_len, uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                        int64_t *consumed, int64_t *produced, int32_t *status, unsigned long long *work_counter,
                        uint32_t *tok_scratch) {
  DCMP_DYN_SMEM(smem_raw);
  ParSmem &S = *reinterpret_cast<ParSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  RoundMeta &M = *reinterpret_cast<RoundMeta *>(S.sym);
  uint32_t(*ckpt)[PT] = M.ckpt;
  const uint32_t *const win = win_of(S);

  if (tid == 0) {
    mbar_init(&S.mbar, 1);
    fence_mbar_init();
    S.pf.stream = -2;
    for (int w = 0; w < NWARP; w++) S.a_done[w] = 0;
  }
  __syncthreads();

  uint32_t round_no = 0; // data rounds this CTA has run (CTA-uniform)
  uint32_t mbar_phase = 0; // parity of the window mbarrier (CTA-uniform)
  for (;;) {
    // ---------------------------------------------------------------- fetch a stream
    __syncthreads();
    if (tid == 0) {
      const bool claimed = S.pf.stream != -2; // claimed (and opened) during the previous stream
      const unsigned long long k = claimed ? (unsigned long long)S.pf.stream : atomicAdd(work_counter, 1ull);
      S.stream = k < (unsigned long long)nstreams ? (int64_t)k : -1;
      if (S.stream >= 0) {
        int st;
        if (claimed) { S.src = S.pf.src; S.src_len = S.pf.src_len; S.dst = S.pf.dst; S.dst_cap = S.pf.dst_cap; st = S.pf.st; }
        else st = open_stream(fmt, (int64_t)k, src_base, src_off, src_len, dst_base, dst_off, dst_cap, &S.src, &S.src_len, &S.dst, &S.dst_cap);
        S.hdr_ready = claimed && st == DCMP_OK && S.pf.hdr;
        S.nh_ready = 0;
        S.base_bit = 0;
        S.out_pos = 0;
        S.last_block = 0;
        S.status = st;
        S.next = st == DCMP_OK ? NEXT_HEADER : NEXT_RETRY;
        S.adler_a = 1;
        S.adler_b = 0;
      }
      S.pf.stream = -2;
      S.pf_done = 0;
    }
    __syncthreads();
    if (S.stream < 0) break;

    // ---------------------------------------------------------------- rounds
    while (S.next == NEXT_HEADER || S.next == NEXT_CONTINUE || S.next == NEXT_STORED) {
      const int64_t base_bit = S.base_bit;
      const uint32_t A = (uint32_t)((uintptr_t)S.src & 15);
      const int64_t vbit = base_bit + 8 * (int64_t)A;        // bit position from the 16-aligned address below src
      const int64_t v0 = (vbit >> 3) & ~(int64_t)15;          // window start (bytes from that address)
      const int bit0 = (int)(vbit - 8 * v0);                  // 0..127
      const uint8_t *g0 = S.src - A + v0;
      const int64_t vend = ((int64_t)A + S.src_len + 15) & ~(int64_t)15; // one past the last valid granule
      if (base_bit > 8 * S.src_len) { // ran off the input: the general kernel knows what the reference does
        __syncthreads();
        if (tid == 0) S.next = NEXT_RETRY;
        __syncthreads();
        break;
      }
      if (S.next == NEXT_HEADER) {
        // ------------------------------------------------------------ block header
        const bool ahead = S.hdr_ready != 0; // parsed while the previous stream was being decoded
        const bool ahead2 = S.nh_ready != 0; // parsed while the previous block's last round was being replayed
        __syncthreads();
        if (ahead2) {
          if (tid == 0) {
            S.last_block = S.nh_last;
            S.hdr_hlit = S.nh_hlit;
            S.hdr_hdist = S.nh_hdist;
            S.base_bit = S.nh_base_bit;
            S.nh_ready = 0;
            S.next = NEXT_CONTINUE;
          }
        } else if (ahead) {
          for (int i = tid; i < MAX_LENS; i += PT) S.lens[i] = S.pf.lens[i];
          if (tid == 0) {
            S.last_block = S.pf.last_block;
            S.hdr_hlit = S.pf.hlit;
            S.hdr_hdist = S.pf.hdist;
            S.base_bit = S.pf.base_bit;
            S.hdr_ready = 0;
            S.next = NEXT_CONTINUE;
          }
        } else stage_window(S, g0, HDR_BYTES, (int)(vend - v0 > HDR_BYTES ? HDR_BYTES : vend - v0), tid, mbar_phase);
        if (tid == 0 && !ahead && !ahead2) {
          STAT(g_stats.hdr_rounds++;)
          HdrBits hb;
          hb.init((const uint8_t *)win_of(S), bit0, HDR_BYTES);
          int64_t avail = 8 * S.src_len - base_bit + bit0; // valid bits measured in window coordinates
          int nx = NEXT_CONTINUE;
          hb.refill();
          if (avail - hb.pos() < 3) nx = NEXT_RETRY;
          else {
            S.last_block = (int)hb.pop(1);
            uint32_t type = hb.pop(2);
            if (type == 2) {
              int hlit = 0, hdist = 0;
              int used = parse_dynamic_header(S.pf.codes, S.lens, hb, avail, &hlit, &hdist); // (pf.codes is scratch: the two parses never overlap)
              if (used < 0) nx = NEXT_RETRY;
              else {
                S.hdr_hlit = (uint32_t)hlit;
                S.hdr_hdist = (uint32_t)hdist;
                S.base_bit = base_bit + (used - bit0);
              }
            } else if (type == 1) {
              for (int n = 0; n < 288; n++) S.lens[n] = n < 144 ? 8 : n < 256 ? 9 : n < 280 ? 7 : 8; // de.ml:821-833
              for (int n = 0; n < 32; n++) S.lens[288 + n] = 5; // 32 codes of 5 bits; 30/31 are invalid symbols
              S.hdr_hlit = 288;
              S.hdr_hdist = 32;
              S.base_bit = base_bit + (hb.pos() - bit0);
            } else if (type == 0) {
              // stored block (de.ml:1613-1627): skip to the byte boundary, LEN / NLEN, then LEN raw bytes
              const int64_t sb = (base_bit + (hb.pos() - bit0) + 7) & ~(int64_t)7; // stream bit of LEN
              if (8 * S.src_len - sb < 32) nx = NEXT_RETRY;
              else {
                HdrBits h2;
                h2.init((const uint8_t *)win_of(S), (int)(sb - base_bit) + bit0, HDR_BYTES);
                h2.refill();
                const uint32_t len = h2.pop(16), nlen = h2.pop(16);
                const int64_t payload = (sb >> 3) + 4;
                if (nlen != 0xffffu - len || (int64_t)len > S.src_len - payload) nx = NEXT_RETRY; // exact error: general kernel
                else {
                  S.stored_rem = len;
                  S.base_bit = 8 * payload;
                  nx = len ? NEXT_STORED : (S.last_block ? NEXT_FINISH : NEXT_HEADER);
                }
              }
            } else {
              nx = NEXT_RETRY; // BTYPE=3: general kernel
            }
          }
          S.next = nx;
        }
        __syncthreads();
        if (S.next == NEXT_RETRY) break;
        if (S.next != NEXT_CONTINUE) continue; // stored block (or an empty one): no tables to build
        const int hlit = (int)S.hdr_hlit, hdist = (int)S.hdr_hdist;
        // (the symbol buffer is idle between rounds: it lends the builder its scratch words)
        const bool ok = build_fast_tables(S, S.lens, hlit, S.lens + hlit, hdist, reinterpret_cast<uint32_t *>(S.sym), tid);
        if (!ok) {
          __syncthreads();
          if (tid == 0) S.next = NEXT_RETRY;
          __syncthreads();
          break;
        }
        continue; // S.next == NEXT_CONTINUE, S.base_bit moved past the header
      }

      if (S.next == NEXT_STORED) {
        // ------------------------------------------------------------ stored payload: straight copy
        const uint32_t out0 = S.out_pos;
        const uint32_t n = S.stored_rem;
        if ((int64_t)out0 + n > S.dst_cap) {
          __syncthreads();
          if (tid == 0) S.next = NEXT_RETRY; // Unexpected_end_of_output: exact point needed
          __syncthreads();
          break;
        }
        raw_copy(S, fmt, S.src + (base_bit >> 3), n, S.dst + out0, out0, S.src, S.src + S.src_len, tid, lane, warp);
        if (tid == 0) {
          S.stored_rem = 0;
          S.base_bit = base_bit + 8 * (int64_t)n;
          S.next = S.last_block ? NEXT_FINISH : NEXT_HEADER;
        }
        __syncthreads();
        continue;
      }

      // -------------------------------------------------------------- data round
      {
        int64_t want = vend - v0;
        stage_window(S, g0, WIN_BYTES, (int)(want > WIN_BYTES ? WIN_BYTES : want), tid, mbar_phase);
      }
      const int q0 = bit0 + SUB_BITS * tid; // bit index of this lane's sub-chunk in the staged window
      const int lroot = S.ml.root, droot = S.md.root;
      const uint32_t lmask = (1u << lroot) - 1u, dmask = (1u << droot) - 1u;
      // this lane's token slots in the L2-resident scratch, slot-major: [HEAD_CAP head tokens | SPEC_CAP speculative tokens][PT]
      uint32_t *const tok_cta = tok_scratch + (size_t)blockIdx.x * PT * TOK_SLOTS;
      uint32_t *const tok_head = tok_cta + tid;
      uint32_t *const tok_spec = tok_cta + (size_t)HEAD_CAP * PT + tid;
      // lanes whose sub-chunk starts beyond the input decode nothing
      const bool lane_has_input = base_bit + (int64_t)SUB_BITS * tid < 8 * S.src_len;
      round_no++;
      if (tid == 0) { S.first_bad = PT; S.first_term = PT; S.abort_round = 0; S.pass_owner = -1; S.any_cross = 0; } // (read after the barrier that ends phase B)
      // The CTA's next stream is claimed and its first header parsed (a serial job of ~30k cycles) by the last warp in
      // a round that leaves it without input — the last round of a stream as a rule — while the other warps are in
      // phase A; a stream whose rounds are all full has its successor's header parsed in line.
      if (warp == NWARP - 1 && !(base_bit + (int64_t)SUB_BITS * (PT - 32) < 8 * S.src_len)) {
        const int todo = __shfl_sync(0xffffffffu, S.pf_done == 0 ? 1 : 0, 0); // (lane 0 sets the flag in there: one reader)
        if (todo) prefetch_next_stream(S, fmt, nstreams, src_base, src_off, src_len, dst_base, dst_off, dst_cap, work_counter, lane);
      }
      // ---- A: speculative decode of the own sub-chunk from bit 0 (a function of its own: the hot loop gets the
      // whole register budget, not what the rest of the kernel leaves it)
      phase_a_lane(S, ckpt, tok_spec, q0, tid, lane_has_input);
      const int ns = S.spec_ns[tid], xkind = S.exit_kind[tid], xpos = S.exit_pos[tid], last_ck = S.last_ck[tid];
      const uint32_t out_s = S.spec_out[tid];
      // No CTA-wide barrier between A and B: a head needs the exit of the lane before it and the checkpoints of the
      // lanes it runs through, so a warp waits for the warp before it only (and, in the rare case that one of its
      // heads crosses into a later warp's sub-chunks, for that warp).  Phase A never waits: no cycle.
      __syncwarp();
      if (lane == 0) { __threadfence_block(); S.a_done[warp] = round_no; }
      if (warp > 0) { while (S.a_done[warp - 1] != round_no) DCMP_SPIN(); }
      __threadfence_block();
      __syncwarp();
      // ---- B: re-synchronise with the true entry point handed over by the previous lane.
      // Lane t assumes that lane t-1's true sequence ends in that lane's speculative sequence
      // (entry = its exit) and decodes from there until the start of one of its tokens coincides
      // with a checkpoint of the speculative sequence of the sub-chunk it is in — its own, or,
      // when it crosses its whole sub-chunk without meeting it, a later lane's.  Lanes covered by
      // such a long head are "absorbed": their own head is discarded; each replays the part of the
      // covering head that starts in its sub-chunk, and the one in which that head merged adds its
      // speculative tokens from the merge point on.
      enum : int { H_BAD = 0, H_MERGE = 1, H_PASS = 2, H_TERM = 3 };
      const int ENTRY_NONE = -100000;
      int nh = 0, mk = 0, hcat = H_BAD, hkind = X_NORMAL, hpos = 0, ncross = 0;
      uint32_t out_h = 0, mskip = 0;
      {
        int entry = 0;
        if (tid > 0) entry = S.exit_kind[tid - 1] == X_NORMAL ? (int)S.exit_pos[tid - 1] - SUB_BITS : ENTRY_NONE;
        if (lane_has_input && entry != ENTRY_NONE) {
          if (tid == 0 && ns > 0) {
            hcat = H_MERGE; // lane 0's speculative pass started at the true position
          } else {
            Cursor cu;
            cu.start(q0 + entry, lmask);
            uint32_t tok = 0, no = 0;
            int u = tid, ubase = 0, lck = last_ck, seen_ck = -1, xk_u = xkind, xp_u = xpos, tstart = entry;
            for (;;) {
              if (cu.at_token()) {
                const int p = cu.q - q0;
                tstart = p;
                if (p - ubase >= SUB_BITS) { // crossed into the next sub-chunk (rare)
                  if (ncross == MAX_CROSS) { hcat = H_TERM; hkind = X_CAP; hpos = p; break; }
                  M.cross_nh[tid][ncross] = (uint16_t)nh;
                  M.cross_out[tid][ncross] = (uint16_t)out_h;
                  ncross++;
                  u++;
                  ubase += SUB_BITS;
                  seen_ck = -1;
                  if (u >= PT) { hcat = H_PASS; hpos = p; break; } // past the window: the next round starts here
                  if ((u >> 5) != warp) { while (S.a_done[u >> 5] != round_no) DCMP_SPIN(); __threadfence_block(); } // that warp's phase A
                  lck = S.last_ck[u];
                  xk_u = S.exit_kind[u];
                  xp_u = S.exit_pos[u];
                }
                const int pl = p - ubase;
                const int ck = pl >> CKPT_SHIFT;
                if (ck != seen_ck) {
                  seen_ck = ck;
                  if (ck <= lck) { // the speculative pass of lane u left a checkpoint for this window
                    const uint32_t c = ckpt[ck][u];
                    if ((c & (CKPT_NONE | 31u)) == (uint32_t)(pl & 31)) { mk = (int)((c >> 5) & 255u); mskip = c >> 13; hcat = H_MERGE; break; }
                  }
                }
                if (xk_u == X_CAP && pl == xp_u) { // met that pass where its token buffer filled up
                  mk = S.spec_ns[u]; mskip = S.spec_out[u]; hcat = H_MERGE; break;
                }
                if (nh == HEAD_CAP || out_h > (uint32_t)(LANE_BYTE_CAP - 258)) { hcat = H_TERM; hkind = X_CAP; hpos = p; break; }
              }
              const int r = dec_step<true>(S.tbl, win, cu, lmask, dmask, lroot, droot, tok, no);
              if (r == STEP_NONE) continue;
              if (r == STEP_EOB) { hcat = H_TERM; hkind = X_EOB; hpos = cu.q - q0; break; }
              if (r == STEP_ERR) { hcat = H_TERM; hkind = X_ERR; hpos = tstart; break; }
              tok_head[nh * PT] = tok;
              nh++;
              out_h += no;
            }
          }
        }
        STAT(if (hcat != H_BAD && tid > 0) { if (hcat == H_MERGE && ncross == 0) g_stats.hist_nh[nh < 47 ? nh : 47]++; else if (hcat != H_TERM) g_stats.nosync_lanes++; })
        // the lane after the ones this head covers is the next one whose own head counts
        int last = tid + ncross;           // lane in which the head ended
        if (last > PT - 1) last = PT - 1;  // (H_PASS)
        M.jump[0][tid] = (uint16_t)(last + 1);
        M.mark[tid] = tid == 0 ? 1 : 0;
        if (ncross) S.any_cross = 1;
      }
      __syncthreads();
      // ---- C: find the lanes whose head counts (the orbit of lane 0 under `jump`, by pointer doubling),
      // describe every lane's part, cut the round at the first lane that ends it, scan output sizes
      if (S.any_cross) { // (with the warm-up a head rarely leaves its sub-chunk: most rounds have every lane in the orbit)
        int cur = 0;
#pragma unroll 1
        for (int span = 1; span < PT; span <<= 1) {
          const int j = M.jump[cur][tid];
          if (j < PT) {
            // (a mark may land while another thread reads it in the same pass: early only marks members of the orbit
            // sooner — relaxed atomics, see simt.cuh)
            if (ld_relaxed_u8(&M.mark[tid])) st_relaxed_u8(&M.mark[j], 1);
            M.jump[cur ^ 1][tid] = M.jump[cur][j];
          } else {
            M.jump[cur ^ 1][tid] = (uint16_t)PT;
          }
          cur ^= 1;
          __syncthreads();
        }
      }
      const bool owner = !S.any_cross || M.mark[tid] != 0;
      {
        const uint32_t ow = __ballot_sync(0xffffffffu, owner);
        if (lane == 0) S.m_owner[warp] = ow;
      }
      if (owner) {
        // describe this lane's part and the parts of the lanes its head covers
        const bool has_head_end = hcat == H_TERM;
        for (int j = 0; j <= ncross; j++) {
          const int y = tid + j;
          if (y >= PT) break; // H_PASS: the last crossing left the window
          LaneDesc d;
          d.owner = (uint16_t)tid;
          d.h0 = (uint8_t)(j == 0 ? 0 : M.cross_nh[tid][j - 1]);
          d.h1 = (uint8_t)(j < ncross ? M.cross_nh[tid][j] : nh);
          const uint32_t o0 = j == 0 ? 0u : M.cross_out[tid][j - 1], o1 = j < ncross ? M.cross_out[tid][j] : out_h;
          d.out_head = (uint16_t)(o1 - o0);
          d.spec_k = (uint8_t)mk;
          d.spec_skip = (uint16_t)mskip;
          d.term_pos = (uint16_t)hpos;
          const bool lastlane = j == ncross;
          d.flags = (uint8_t)((lastlane && hcat == H_MERGE ? 1 : 0) | (lastlane && has_head_end ? 2 | (hkind << 4) : 0));
          M.desc[y] = d;
        }
        if (hcat == H_PASS) { S.pass_owner = tid; S.pass_pos = hpos; }
      }
      __syncthreads();
      // every lane's final contribution: head tokens [d_h0, d_h0 + d_nh) of lane d_owner, then (d_spec) its own
      // speculative tokens from d_k on
      int d_owner = tid, d_h0 = 0, d_nh = 0, d_k = 0, tkind = X_NORMAL;
      int64_t t_next = 0; // where the stream continues if this lane closes the round
      bool d_spec = false, in_ok = true, term = false, clean = false;
      uint32_t out_valid = 0;
      {
        in_ok = owner ? hcat != H_BAD : true;
        if (in_ok) {
          const LaneDesc d = M.desc[tid];
          d_owner = d.owner;
          d_h0 = d.h0;
          d_nh = (int)d.h1 - (int)d.h0;
          out_valid = d.out_head;
          if (d.flags & 1) { // the true sequence leaves this sub-chunk the way the speculative one does
            d_spec = true;
            d_k = d.spec_k;
            out_valid += out_s - d.spec_skip;
            clean = true;
            tkind = xkind;
            term = xkind != X_NORMAL;
            t_next = base_bit + (int64_t)SUB_BITS * tid + xpos;
          } else if (d.flags & 2) {
            term = true;
            tkind = (d.flags >> 4) & 7;
            t_next = base_bit + (int64_t)SUB_BITS * d_owner + d.term_pos;
          }
        }
      }
      if (!in_ok) atomicMin(&S.first_bad, tid);
      if (term) atomicMin(&S.first_term, tid);
      // (the scan runs over every lane that could be chained: the lanes in front of the cut are all among them, and
      // only their sums are used)
      uint32_t incl = warp_incl_scan(in_ok ? out_valid : 0u, lane);
      if (lane == 31) S.warp_c[warp] = incl;
      __syncthreads();
      const int first_bad = S.first_bad, first_term = S.first_term;
      int n_valid = first_term < first_bad ? first_term + 1 : first_bad;
      {
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < NWARP; w++)
          if (w < warp) woff += S.warp_c[w];
        incl += woff;
      }
      const uint32_t out0 = S.out_pos;
      if (n_valid == 0) {
        if (tid == 0) S.next = NEXT_RETRY; // truncated input: the general kernel reports what the reference does
      } else if (tid == n_valid - 1) {
        // this lane closes the round
        const uint32_t round_out = incl; // inclusive sum up to the last valid lane
        S.round_out = round_out;
        int nx = NEXT_CONTINUE;
        int64_t nb = t_next;
        if (first_term < first_bad) {
          // closed by this lane's own exit
          if (tkind == X_EOB) nx = S.last_block ? NEXT_FINISH : NEXT_HEADER;
          else if (tkind == X_ERR) nx = NEXT_RETRY;
          // X_CAP: resume right after the last token
        } else if (!clean) {
          // every lane was used and the last ones are covered by a head that ran past the window
          const int x = S.pass_owner;
          if (x < 0) nx = NEXT_RETRY; // cannot happen
          else nb = base_bit + (int64_t)SUB_BITS * x + S.pass_pos;
        }
        // otherwise the next lane could not be chained: it opens the next round at this lane's (normal) exit
        if ((int64_t)out0 + round_out > S.dst_cap) nx = NEXT_RETRY; // Unexpected_end_of_output: exact point needed
        S.base_bit = nb;
        S.next = nx;
        S.n_valid = n_valid;
        STAT(g_stats.rounds++; g_stats.valid_lanes += n_valid;
             if (first_term < first_bad) g_stats.cut_term[tkind]++; else g_stats.cut_bad++;)
      }
      __syncthreads();
      if (S.next == NEXT_RETRY) break;
      const uint32_t round_total = S.round_out;
      // The round's output leaves in slices of whole lanes that fit the symbol buffer (RSYM symbols; one
      // slice is the rule).  A back-reference from a later slice into an earlier one is history: those
      // bytes are already in HBM/L2.  From here on the round metadata (which overlays the symbol buffer)
      // is dead; what a lane replays sits in its registers.
      const uint32_t my_end = incl, my_start = incl - out_valid; // round-relative byte range of this lane
      uint32_t s0 = 0;
      bool round_failed = false;
      while (s0 < round_total) {
      uint32_t s1 = round_total;
      if (round_total - s0 > (uint32_t)RSYM) { // (CTA-uniform; the rule is one slice)
        if (tid == 0) S.slice_end = s0;
        __syncthreads();
        if (tid < n_valid && my_end <= s0 + (uint32_t)RSYM) atomicMax(&S.slice_end, my_end); // a prefix of the lanes
        __syncthreads();
        s1 = S.slice_end;
      }
      if (s1 == s0) { round_failed = true; break; } // cannot happen: one lane's output always fits
      const bool in_slice = tid < n_valid && my_start >= s0 && my_end <= s1 && out_valid > 0u;
      const uint32_t slice_out0 = out0 + s0;   // stream position of the slice's first byte
      const uint32_t round_out = s1 - s0;      // bytes of this slice
      // Symbol-buffer coordinates: index i <-> stream position slice_out0 + (i - bias) <-> global address G0 + i,
      // G0 = (dst + slice_out0) rounded down to 16 bytes, so that 16-symbol chunks are aligned 16-byte stores.
      const uint32_t bias = (uint32_t)((uintptr_t)(S.dst + slice_out0) & 15);
      uint8_t *const G0 = S.dst + slice_out0 - bias;
      const uint32_t g0_low = (uint32_t)((uintptr_t)G0 & HIST_MASK);
      const uint32_t E = bias + round_out; // one past the last symbol of the slice
      // A block ended in this round and another one follows: the last warp, when the round left it no lanes to
      // replay, parses the next header out of the staged window while the other warps are in phase D.
#if DCMP_WIN_OVERLAY
      // (the window is about to be overwritten by symbols: the header's bytes are saved first; CTA-uniform condition)
      const bool nh_try = s0 == 0 && S.next == NEXT_HEADER && n_valid <= PT - 32;
      const int nh_w0 = (int)((bit0 + (S.base_bit - base_bit)) >> 5) & ~3; // first window word saved (16-byte aligned)
      if (nh_try) {
        for (int i = tid; i < HDR_BYTES / 4; i += PT) S.nh_win[i] = nh_w0 + i < WIN_BYTES / 4 ? win[nh_w0 + i] : 0u;
        __syncthreads();
      }
#else
      const bool nh_try = s0 == 0 && S.next == NEXT_HEADER && n_valid <= PT - 32;
      const int nh_w0 = 0;
#endif
      if (nh_try && tid == PT - 32) {
        int64_t staged = 8 * (vend - v0 > WIN_BYTES ? (int64_t)WIN_BYTES : vend - v0); // window bits that hold input
        if (DCMP_WIN_OVERLAY && staged > 32 * (int64_t)nh_w0 + 8 * HDR_BYTES) staged = 32 * (int64_t)nh_w0 + 8 * HDR_BYTES;
        staged -= 32 * (int64_t)nh_w0;                                                   // (coordinates of the saved copy)
        const int64_t hbit = bit0 + (S.base_bit - base_bit) - 32 * (int64_t)nh_w0;       // where the header starts
        if (hbit + 64 < staged) {
          HdrBits hb;
#if DCMP_WIN_OVERLAY
          hb.init((const uint8_t *)S.nh_win, (int)hbit, HDR_BYTES);
#else
          hb.init((const uint8_t *)win, (int)hbit, WIN_BYTES);
#endif
          int64_t avail = 8 * S.src_len - base_bit + bit0 - 32 * (int64_t)nh_w0; // valid bits measured in the same coordinates
          if (avail > staged) avail = staged;
          hb.refill();
          const int last = (int)hb.pop(1);
          if (hb.pop(2) == 2u) {
            int hlit = 0, hdist = 0;
            const int used = parse_dynamic_header(S.pf.codes, S.lens, hb, avail, &hlit, &hdist);
            if (used >= 0) {
              S.nh_last = last;
              S.nh_hlit = (uint32_t)hlit;
              S.nh_hdist = (uint32_t)hdist;
              S.nh_base_bit = S.base_bit + (used - hbit);
              S.nh_ready = 1;
            }
          }
        }
      }
      // ---- D: replay tokens into symbols
      if (tid < 16) {
        if ((uint32_t)tid < bias) S.sym[tid] = 0;
        S.sym[E + tid] = 0;
      }
      {
        // Every lane walks its own token list: head tokens [0, nh), then the speculative tokens from
        // the merge point on.  One step of the (warp-uniform) loop emits a literal or up to eight
        // symbols of a match; the source of a match decides how:
        //   inside this round (any lane, this one included): markers 256 + source index, no load at all
        //   entirely in earlier rounds: bytes of the history ring
        //   straddling the start of the round, or overlapping itself inside the group: per symbol
        const int total_tok = in_slice ? d_nh + (d_spec ? ns - d_k : 0) : 0;
        STAT(if (in_slice) { g_stats.head_tokens += d_nh; g_stats.used_tokens += total_tok; } long iters = 0;)
#if DCMP_HIST_GLOBAL
        // history = the stream's own output in HBM/L2; symbol index s <-> address G0 + s.  L2-coherent loads
        // (.cg): the bytes were stored by this CTA in earlier rounds, a barrier ago at least.
        const uint32_t *hist32 = reinterpret_cast<const uint32_t *>(G0 - ((uintptr_t)G0 & 3)); // G0 is 16-aligned: no-op
#else
        const uint32_t *ring32 = reinterpret_cast<const uint32_t *>(S.ring);
#endif
        int ti = 0;
        int32_t i = (int32_t)(my_start - s0 + bias); // next symbol to write
        int32_t s = 0;                // next source symbol of the pending match
        uint32_t rem = 0;             // symbols the pending match still owes
        uint32_t dist = 0;
        // tokens are fetched three ahead of their use (slot-major scratch: a warp's loads of one slot are one line)
        const uint32_t *head32 = tok_cta + (size_t)d_h0 * PT + d_owner;                     // head token t: + t * PT
        const uint32_t *spec32 = tok_cta + (size_t)(HEAD_CAP + d_k - d_nh) * PT + tid;     // speculative token t (>= d_nh): + t * PT
        auto tok_at = [&](int t) { return __ldcg((t < d_nh ? head32 : spec32) + (size_t)t * PT); };
        uint32_t q0 = 0, q1 = 0, q2 = 0;
        if (0 < total_tok) q0 = tok_at(0);
        if (1 < total_tok) q1 = tok_at(1);
        if (2 < total_tok) q2 = tok_at(2);
#if DCMP_D_PAIRS
        uint32_t q3 = 0; // (a pair consumes two tokens: one more in flight keeps the lead of the loads)
        if (3 < total_tok) q3 = tok_at(3);
#endif
        bool active = total_tok > 0;
        while (__any_sync(0xffffffffu, active)) {
          STAT(if (active) iters++;)
          // -- next token, for the lanes that finished the previous one
          bool fetch = active && rem == 0;
          if (fetch && ti == total_tok) { active = false; fetch = false; }
          uint32_t t = 0, t2 = 0;
          bool pair = false; // (DCMP_D_PAIRS) two literals in a row leave in one step
          if (fetch) {
            t = q0;
#if DCMP_D_PAIRS
            pair = ti + 1 < total_tok && !tok_is_match(t) && !tok_is_match(q1);
            if (pair) {
              t2 = q1;
              q0 = q2;
              q1 = q3;
              if (ti + 4 < total_tok) q2 = tok_at(ti + 4);
              if (ti + 5 < total_tok) q3 = tok_at(ti + 5);
              ti += 2;
            } else {
              q0 = q1;
              q1 = q2;
              q2 = q3;
              if (ti + 4 < total_tok) q3 = tok_at(ti + 4);
              ti++;
            }
#else
            q0 = q1;
            q1 = q2;
            if (ti + 3 < total_tok) q2 = tok_at(ti + 3);
            ti++;
#endif
          }
          const bool new_match = fetch && tok_is_match(t);
          if (new_match) {
            dist = ((t >> 8) & 0x7fffu) + 1u;
            rem = (t & 0xffu) + 3u;
            s = i - (int32_t)dist;
            if (dist > slice_out0 + (uint32_t)(i - (int32_t)bias)) { // Invalid distance (de.ml:1700): the general kernel reports it
              S.abort_round = 1;
              active = false;
              rem = 0;
            }
          }
          // -- this step's symbols (branch-free for the three common cases)
          const bool lit = fetch && !new_match;
          const uint32_t n = lit ? (pair ? 2u : 1u) : (rem < 8u ? rem : 8u);  // 0 for idle lanes
          const bool in_round = s >= (int32_t)bias;
          const bool in_hist = s + (int32_t)n <= (int32_t)bias;
          const bool slow = rem > 0u && (in_round ? dist < n : !in_hist);
          // markers: source inside this round (another lane's output or this lane's own)
          const uint32_t m0 = (256u + (uint32_t)s) * 0x00010001u + 0x00010000u;
          // history: source entirely in earlier rounds (addresses are masked: the loads are always legal)
          const uint32_t r = g0_low + (uint32_t)s; // ring byte address (mod HIST) of the first source byte
#if !DCMP_HIST_GLOBAL
          const uint32_t wi0 = r >> 2;
#endif
          uint32_t w0 = 0, w1 = 0, w2 = 0;
          if (in_hist && rem > 0u) {
#if DCMP_HIST_GLOBAL
            const int32_t gw = s >> 2; // word index relative to G0 (negative: before this round)
            const uint32_t span = ((uint32_t)s & 3u) + n; // only words that hold a source byte are touched
            w0 = __ldcg(hist32 + gw);
            if (span > 4u) w1 = __ldcg(hist32 + gw + 1);
            if (span > 8u) w2 = __ldcg(hist32 + gw + 2);
#else
            w0 = ring32[wi0 & (HIST / 4 - 1)];
            w1 = ring32[(wi0 + 1u) & (HIST / 4 - 1)];
            w2 = ring32[(wi0 + 2u) & (HIST / 4 - 1)];
#endif
          }
          const uint32_t sh8 = 8u * (r & 3u);
          const uint32_t ha = __funnelshift_r(w0, w1, sh8), hb = __funnelshift_r(w1, w2, sh8);
          uint32_t p0 = in_round ? m0 : __byte_perm(ha, 0u, 0x4140);
          const uint32_t p1 = in_round ? m0 + 0x00020002u : __byte_perm(ha, 0u, 0x4342);
          const uint32_t p2 = in_round ? m0 + 0x00040004u : __byte_perm(hb, 0u, 0x4140);
          const uint32_t p3 = in_round ? m0 + 0x00060006u : __byte_perm(hb, 0u, 0x4342);
          if (lit) p0 = (t & 0xffu) | ((t2 & 0xffu) << 16);
          STAT(if (!slow && rem > 0) { if (in_round) g_stats.markers += n; else g_stats.hist_bytes += n; })
          if (slow) {
            // straddles the start of the round, or overlaps itself inside the group (point at the first period)
            for (uint32_t k = 0; k < n; k++) {
              const int32_t sk = s + (int32_t)(dist < n ? k % dist : k);
              uint32_t v;
              if (sk >= (int32_t)bias) { v = 256u + (uint32_t)sk; STAT(g_stats.markers++;) }
#if DCMP_HIST_GLOBAL
              else { v = __ldcg(G0 + sk); STAT(g_stats.hist_bytes++;) }
#else
              else { v = S.ring[(g0_low + (uint32_t)sk) & HIST_MASK]; STAT(g_stats.hist_bytes++;) }
#endif
              S.sym[i + (int32_t)k] = (uint16_t)v;
            }
          } else {
            // eight predicated stores (the compiler turns the plain `if`s into a cascade of eight divergent branches)
            uint16_t *const at = S.sym + i;
            st_sym_if(n > 0u, at + 0, p0);
            st_sym_if(n > 1u, at + 1, p0 >> 16);
            st_sym_if(n > 2u, at + 2, p1);
            st_sym_if(n > 3u, at + 3, p1 >> 16);
            st_sym_if(n > 4u, at + 4, p2);
            st_sym_if(n > 5u, at + 5, p2 >> 16);
            st_sym_if(n > 6u, at + 6, p3);
            st_sym_if(n > 7u, at + 7, p3 >> 16);
          }
          i += (int32_t)n;
          if (rem > 0u) { s += (int32_t)n; rem -= n; }
        }
        STAT(g_stats.d_iters += iters; if (iters > g_stats.d_lane_max_iters) g_stats.d_lane_max_iters = iters;)
      }
      __syncthreads();
      if (S.abort_round) { round_failed = true; break; }
      phase_e_drain(S, fmt, G0, g0_low, bias, E, slice_out0, round_out, tid, lane, warp);
      s0 = s1;
      } // slices
      STAT(g_stats.spec_tokens += ns;)
      if (round_failed) {
        __syncthreads();
        if (tid == 0) S.next = NEXT_RETRY;
        __syncthreads();
        break;
      }
    } // rounds

    // ---------------------------------------------------------------- finish the stream
    __syncthreads();
    if (tid == 0) {
      const int64_t k = S.stream;
      int st = DCMP_INTERNAL_RETRY;
      int64_t cons = 0, prod = 0;
      if (S.next == NEXT_FINISH) {
        st = DCMP_OK;
        int64_t c = S.base_bit;
        cons = (c + 7) >> 3;
        prod = S.out_pos;
        if (c > 8 * S.src_len) st = DCMP_INTERNAL_RETRY; // decoded past the end: reference quirk, general kernel
        if (fmt == DCMP_FMT_ZLIB && st == DCMP_OK) {
          const uint8_t *tr = S.src + cons; // lib/zl.ml:409-415: big-endian Adler-32 right after the deflate data
          uint32_t want = ((uint32_t)tr[0] << 24) | ((uint32_t)tr[1] << 16) | ((uint32_t)tr[2] << 8) | (uint32_t)tr[3];
          uint32_t got = (S.adler_b << 16) | S.adler_a;
          if (want != got) { st = DCMP_INVALID_CHECKSUM; cons = 0; }
          else cons += 6;
        }
        if (st == DCMP_INTERNAL_RETRY) { cons = 0; prod = 0; }
      }
      STAT(if (st == DCMP_INTERNAL_RETRY) g_stats.retries++;)
      consumed[k] = cons;
      produced[k] = prod;
      status[k] = st;
    }
  }
}

size_t inflate_parallel_smem_bytes() { return sizeof(ParSmem) + 128; }
int inflate_parallel_ctas_per_sm() { return DCMP_CTAS_PER_SM; }
// bytes of token scratch one CTA needs (global memory, stays L2-resident)
size_t inflate_parallel_scratch_bytes_per_cta() { return (size_t)TOK_SLOTS * PT * sizeof(uint32_t) + 64; }

#ifndef DCMP_EMU
// The opt-in to more than 48 KiB of dynamic shared memory is per device: dcmp_engine_create calls this after
// cudaSetDevice for every engine.
int configure_inflate_parallel() {
  return (int)cudaFuncSetAttribute(inflate_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inflate_parallel_smem_bytes());
}
int launch_inflate_parallel(int fmt, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                            uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                            int64_t *produced, int32_t *status, unsigned long long *work_counter, uint32_t *tok_scratch,
                            int ctas, cudaStream_t stream) {
  const size_t smem = inflate_parallel_smem_bytes();
  if (ctas > n) ctas = (int)n;
  inflate_parallel_kernel<<<(unsigned)ctas, PT, smem, stream>>>(fmt, n, src_base, src_off, src_len, dst_base, dst_off,
                                                               dst_cap, consumed, produced, status, work_counter,
                                                               tok_scratch);
  return 0;
}
#endif

} // namespace dcmp
