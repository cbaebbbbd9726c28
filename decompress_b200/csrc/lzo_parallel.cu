// decompress_b200/csrc/lzo_parallel.cu — LZO1X decode with a whole CTA per block (BASELINE config C5).
//
// Replaces Lzo.uncompress / Lzo.uncompress_with_buffer (lib/lzo.ml:395-414) on well-formed blocks.  The
// reference's interpreter (`run`/`fiber`, lzo.ml:246-393) is one dependent chain: an opcode can only be
// decoded once the previous one has said where it starts and in which state.  Here that chain is reduced
// to one shared-memory load per instruction; everything else is data-parallel over the CTA.  A round covers
// a window of the compressed block staged in shared memory by TMA:
//
//   1  candidates   every byte position of the window is decoded as if an instruction started there, in
//                   both parser states (after a literal run / after a match with trailing literals vs.
//                   not: the only state the format has, lzo.ml:327-341): next[state][p] = where the next
//                   instruction starts and in which state, or END / BAD / OUT (leaves the window);
//   2  chase        one thread follows next[] from the true start and lists the instructions of the round;
//   3  scan         the listed instructions are decoded again, in parallel: output lengths, block-wide
//                   exclusive scan, cut at the symbol budget, validity (distance, output window);
//   4  replay       literal bytes go from the staged input into a 16-bit symbol buffer, a back-reference into
//                   this round's output leaves markers 256 + (source index), one into earlier rounds reads
//                   the block's own output back through L2;
//   5  drain        markers are chased to a byte, 128-bit stores to HBM (symbol_drain.cuh, shared with
//                   inflate_parallel.cu).
// A literal run too long for a window is moved input -> output directly (raw_copy).
//
// Anything that is not the happy path — malformed or truncated input, a distance beyond the produced
// output, an output window that is too small — makes the CTA give the block up with status LZP_RETRY; the
// dispatcher (engine.cu) then runs the one-warp-per-block kernel (lzo_decode.cu) on exactly those blocks,
// which reproduces the reference's error behaviour.  Both are device code; there is no CPU path.
#include "symbol_drain.cuh"
#include "tma.cuh"

namespace dcmp {

#ifndef DCMP_LZP_PT
#define DCMP_LZP_PT 256
#endif
// Four CTAs per SM (64 registers, ~51 KB of shared memory each: an 8 KiB input window, 10 K symbols per round) beat two
// big ones (16 KiB window, 20 K symbols, 122 registers) by a quarter on C5: the kernel is latency bound — the candidate
// decode, the milestone walk and the replay's history reads all wait on memory — and residency is what hides it
// (measured, 512 blocks of 1 MiB: 2 CTAs 42.4 GB/s, 3 CTAs 40.8, 4 CTAs 54.5, 5 CTAs 48.7, 6 CTAs 43.5).
#ifndef DCMP_LZP_DEC_BITS
#define DCMP_LZP_DEC_BITS 12
#endif
#ifndef DCMP_LZP_RSYM_K
#define DCMP_LZP_RSYM_K 10
#endif
#ifndef DCMP_LZP_CTAS
#define DCMP_LZP_CTAS 4
#endif
constexpr int LZP_PT = DCMP_LZP_PT;
constexpr int LZP_RSYM = DCMP_LZP_RSYM_K * 1024;      // output symbols per round (+ up to 15 of alignment bias)
constexpr int LZP_DEC_BITS = DCMP_LZP_DEC_BITS;
constexpr int LZP_DEC = 1 << LZP_DEC_BITS; // window positions decoded as candidates (instructions may extend beyond)
constexpr int LZP_WIN = 2 * LZP_DEC;     // input bytes staged per round
constexpr int LZP_MAXI = LZP_DEC / 2 + 8;       // instructions per round (an instruction is at least two bytes long)
constexpr int LZP_MAXMS = 1024;          // milestones of the serial walk
#ifndef DCMP_LZP_JUMP_PASSES
#define DCMP_LZP_JUMP_PASSES 3
#endif
constexpr int LZP_JUMP_PASSES = DCMP_LZP_JUMP_PASSES;       // doublings of the jump table: a milestone every 8..27 instructions
constexpr int LZP_MAXZ = 80;            // zero bytes of a length extension a candidate decode follows (255 bytes each)
constexpr int LZP_RETRY = 0x7f;          // never leaves the library
constexpr uint32_t LZP_FAR = 0x7ffb, LZP_STEP = 0x7ffc, LZP_END = 0x7ffd, LZP_BAD = 0x7ffe, LZP_OUT = 0x7fff; // next[] codes above any index
constexpr uint32_t LZP_FIRST = 0x4000;   // visited[]: the block's first byte (state class 2)
static_assert(LZP_DEC_BITS <= 13 && 2 * LZP_DEC * 2 <= LZP_RSYM * 2, "the jump table overlays the symbol buffer");

struct LzpSmem {
  static constexpr int THREADS = LZP_PT;
  __align__(16) uint16_t sym[LZP_RSYM + 64];
  __align__(16) uint8_t win[LZP_WIN + 32];
  uint16_t next[2 * LZP_DEC];  // index = state class << 13 | window position; value = index of the next instruction, or LZP_*
  uint16_t visited[LZP_MAXI];  // index of every instruction of the round (| LZP_FIRST)
  uint16_t ms[LZP_MAXMS];      // every few instructions: what the serial walk visits
  int nms, npre;
  uint32_t part[LZP_PT];       // scan partials
  unsigned long long adl_s[LZP_PT / 32];
  long long adl_b[LZP_PT / 32];
  uint32_t adler_a, adler_b;   // unused (no checksum in LZO1X); the drain's interface
  uint32_t out_pos;            // bytes drained so far
  int nvis;                    // instructions listed by the chase
  int cut;                     // instructions that fit the symbol budget
  int give_up, ended;
  int64_t ip;                  // input position of the round's first instruction
  int cls;                     // its state class
  int64_t ip_after;            // where the chase stopped
  int cls_after;
  int64_t ip_end;              // input position after the terminator
  int64_t bulk_len;            // a literal run moved by raw_copy instead
  int64_t bm_len, bm_pay;      // a match longer than the symbol buffer, replicated directly: length, input position of its trailing literals
  uint32_t bm_dist;
  int bm_trail;
  uint32_t mbar_phase;
  __align__(8) uint64_t mbar;
};

struct LzpInst {
  int hdr;        // opcode + operand bytes
  int lit;        // literal run length (no match) ...
  int mlen;       // ... or match length
  uint32_t dist;
  int trail;      // literals after the match
  int ncls;       // state class of the next instruction
  uint32_t kind;  // 0, or LZP_END / LZP_BAD / LZP_OUT
};

// Decode the instruction at window position p in state class cls (0: an opcode below 16 opens a literal run,
// 1: it is a short match, lzo.ml:327-341; 2: first byte of the block, lzo.ml:372-393).  Bytes [0, wlim) of
// the window are input; at_end: the input ends at wlim (else the window does).
__device__ __forceinline__ LzpInst lzp_decode(const uint8_t *w, int p, int cls, int wlim, bool at_end) {
  LzpInst r;
  r.hdr = 0; r.lit = 0; r.mlen = 0; r.dist = 0; r.trail = 0; r.ncls = 0; r.kind = 0;
  const uint32_t miss = at_end ? LZP_BAD : LZP_OUT; // a byte that is not there: malformed, or simply not staged
  int q = p;
  if (q >= wlim) { r.kind = miss; return r; }
  const int chr = w[q++];
  bool capped = false;
  auto count = [&](int *res) -> bool { // lzo.ml:218-234
    if (q >= wlim) return false;       // (LZ_NEED_BYTE before count)
    int z = 0;
    while (q < wlim && w[q] == 0 && z < LZP_MAXZ) { q++; z++; }
    if (z == LZP_MAXZ) { capped = true; return false; }
    if (q >= wlim) return false; // (a longer one does not fit a round anyway: long literal runs are moved by raw_copy)
    *res = z * 255 + w[q++];
    return true;
  };
  if (cls == 2) {
    if (chr == 16) { r.kind = LZP_BAD; return r; }
    if (chr > 17) {
      r.lit = chr - 17;
      r.hdr = 1;
      r.ncls = 0; // state = 0
      if (q >= wlim) r.kind = miss; // guard of `go`
      return r;
    }
    cls = 0;
  }
  if (chr <= 15 && cls == 0) {
    if (chr == 0) {
      int c = 0;
      if (!count(&c)) { r.kind = capped ? LZP_OUT : miss; return r; }
      r.lit = 18 + c;
    } else r.lit = chr + 3;
    r.ncls = 1; // state = -1
  } else if (chr <= 15) {
    if (q >= wlim) { r.kind = miss; return r; }
    const int h = w[q++];
    r.dist = (uint32_t)((h << 2) + (chr >> 2) + 1);
    r.mlen = 2;
    r.trail = chr & 3;
  } else if (chr <= 63) {
    const bool m4 = chr <= 31;
    int len = m4 ? (chr & 7) : (chr & 31);
    if (len == 0) {
      int c = 0;
      if (!count(&c)) { r.kind = capped ? LZP_OUT : miss; return r; }
      len = (m4 ? 7 : 31) + c;
    }
    if (q + 2 > wlim) { r.kind = miss; return r; }
    const int s16 = w[q] | (w[q + 1] << 8);
    q += 2;
    r.dist = m4 ? (uint32_t)(16384 + (((chr & 8) >> 3) << 14) + (s16 >> 2)) : (uint32_t)((s16 >> 2) + 1);
    r.mlen = len + 2;
    r.trail = s16 & 3;
    if (m4 && r.dist == 16384) { r.hdr = q - p; r.kind = LZP_END; return r; }
  } else {
    if (q >= wlim) { r.kind = miss; return r; }
    const int h = w[q++];
    r.dist = (uint32_t)((h << 3) + ((chr >> 2) & 7) + 1);
    r.mlen = (chr >> 5) - 1 + 2;
    r.trail = chr & 3;
  }
  r.hdr = q - p;
  if (r.mlen) r.ncls = r.trail ? 1 : 0;
  if (q >= wlim) { r.kind = miss; return r; }                 // guard of `go` / of do_copy: input must go on
  if (q + r.lit + r.trail > wlim) r.kind = miss;              // payload beyond the input (malformed) or the window
  return r;
}

__global__ void __launch_bounds__(LZP_PT, DCMP_LZP_CTAS)
lzo_parallel_kernel(int mode, int64_t nblocks, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                    uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                    int64_t *produced, int32_t *status) {
  DCMP_DYN_SMEM(smem_raw);
  LzpSmem &S = *reinterpret_cast<LzpSmem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t k = blockIdx.x;
  if (k >= nblocks) return;
  const uint8_t *in = src_base + src_off[k];
  uint8_t *out = dst_base + dst_off[k];
  const int64_t ilen = src_len[k], olen = dst_cap[k];
  if (tid == 0) {
    mbar_init(&S.mbar, 1);
    fence_mbar_init();
    S.mbar_phase = 0;
    S.out_pos = 0;
    S.adler_a = 1; S.adler_b = 0;
    S.give_up = (ilen <= 0 || ilen >= (1ll << 30) || olen >= (1ll << 30)) ? 1 : 0; // 32-bit positions below
    S.ended = 0;
    S.ip = 0;
    S.cls = 2;
  }
  __syncthreads();
  while (!S.give_up && !S.ended) {
    // ---- stage the input window [w_lo, w_lo + LZP_WIN): the round's position rounded down to a 16-byte aligned address
    const int64_t ip0 = S.ip;
    const int cls0 = S.cls;
    const int64_t w_lo = ip0 - (int64_t)((uintptr_t)(in + ip0) & 15);
    const int wlim = (int)(ilen - w_lo < LZP_WIN ? ilen - w_lo : LZP_WIN); // window bytes that are input
    const bool at_end = w_lo + LZP_WIN >= ilen;
    {
      const uint8_t *g = in + w_lo;
      const int64_t vend = (int64_t)((((uintptr_t)in + (uintptr_t)ilen + 15) & ~(uintptr_t)15) - (uintptr_t)g);
      int valid = (int)(vend < LZP_WIN ? vend : LZP_WIN);
      if (valid < 0) valid = 0;
      for (int i = valid + tid * 16; i < LZP_WIN + 32; i += LZP_PT * 16) *(uint4 *)(S.win + i) = make_uint4(0, 0, 0, 0);
      if (tid == 0 && valid > 0) {
        mbar_expect_tx(&S.mbar, (uint32_t)valid);
        tma_load_1d(S.win, g, (uint32_t)valid, &S.mbar);
      }
      if (valid > 0) mbar_wait(&S.mbar, S.mbar_phase & 1u);
      __syncthreads();
      if (tid == 0 && valid > 0) S.mbar_phase ^= 1u;
    }
    // ---- 1: every position as a candidate instruction start, in both state classes
    auto link = [&](const LzpInst &r, int p) -> uint32_t { // where the instruction at p leads
      if (r.kind) return r.kind;
      const int np = p + r.hdr + r.lit + r.trail;
      return np < LZP_DEC ? (uint32_t)np | ((uint32_t)r.ncls << LZP_DEC_BITS) : LZP_FAR;
    };
    for (int p = tid; p < LZP_DEC; p += LZP_PT) {
      // the two states differ only in what an opcode below 16 means
      const uint32_t l0 = link(lzp_decode(S.win, p, 0, wlim, at_end), p);
      S.next[p] = (uint16_t)l0;
      S.next[LZP_DEC + p] = (uint16_t)(S.win[p] > 15 ? l0 : link(lzp_decode(S.win, p, 1, wlim, at_end), p));
    }
    __syncthreads();
    // jump table: J[x] = an instruction some steps ahead on x's chain (LZP_STEP: x's link is not an ordinary one).
    // Built by doubling in place: a thread may read an entry another thread is replacing in the same pass — whatever
    // it reads still points ahead on the same chain, so the accesses are relaxed atomics, not a barrier per entry —
    // in the symbol buffer, which is idle until the replay.
    uint16_t *J = S.sym;
    for (int x = tid; x < 2 * LZP_DEC; x += LZP_PT) {
      const uint32_t e1 = S.next[x];
      J[x] = (uint16_t)(e1 < 2u * LZP_DEC ? e1 : LZP_STEP);
    }
    __syncthreads();
    for (int pass = 0; pass < LZP_JUMP_PASSES; pass++) {
      for (int x = tid; x < 2 * LZP_DEC; x += LZP_PT) {
        const uint32_t j = ld_relaxed_u16(J + x);
        if (j < 2u * LZP_DEC) {
          const uint32_t jj = ld_relaxed_u16(J + j);
          if (jj < 2u * LZP_DEC) st_relaxed_u16(J + x, (uint16_t)jj);
        }
      }
      __syncthreads();
    }
    // ---- 2: the chain of real instructions.  One thread walks the milestones, all threads fill in between.
    const uint32_t out0 = S.out_pos;
    const int p0 = (int)(ip0 - w_lo);
    int64_t bulk = 0, ip_after = -1;
    int cls_after = 0, nlist = 0; // (thread 0) instructions listed so far
    uint32_t xcur = 0;
    auto single = [&](uint32_t e, uint32_t vis) { // (thread 0) one instruction whose link is e; returns false when the chain stops
      if (e == LZP_OUT || e == LZP_BAD) {
        if (nlist == 0) {
          // the round's first instruction does not fit the window: a long literal run is moved directly
          // (its header is re-read from global memory byte-wise, the zero run of its length may be long)
          S.give_up = 1;
          if (e == LZP_OUT && cls0 != 1) { // (the block's first byte, if it is 0, opens an ordinary literal run)
            int64_t q = w_lo + p0;
            const int chr = in[q++];
            if (chr == 0) {
              int64_t z = 0;
              while (q < ilen && in[q] == 0) { q++; z++; }
              if (q < ilen) {
                const int64_t lit = 18 + z * 255 + in[q++];
                if (q < ilen && lit <= ilen - q && (int64_t)out0 + lit <= olen) {
                  bulk = lit;
                  S.ip = q;            // payload start; the round after continues behind it
                  S.give_up = 0;
                }
              }
            }
          }
          if (e == LZP_OUT && S.give_up) {
            // ... or a match whose length extension is longer than a candidate decode follows (run-length data)
            int64_t q = w_lo + p0;
            const int chr = in[q++];
            const bool m4 = chr >= 16 && chr <= 31 && (chr & 7) == 0, m3 = chr >= 32 && chr <= 63 && (chr & 31) == 0;
            if (m3 || m4) {
              int64_t z = 0;
              while (q < ilen && in[q] == 0) { q++; z++; }
              if (q + 3 <= ilen) {
                const int64_t len = (m4 ? 7 : 31) + z * 255 + in[q++];
                const int s16 = in[q] | (in[q + 1] << 8);
                q += 2;
                const int64_t dist = m4 ? 16384 + (((chr & 8) >> 3) << 14) + (s16 >> 2) : (s16 >> 2) + 1;
                const int trail = s16 & 3;
                if (!(m4 && dist == 16384) && q < ilen && trail <= ilen - q && dist <= (int64_t)out0 &&
                    (int64_t)out0 + len + 2 + trail <= olen) {
                  S.bm_len = len + 2; S.bm_dist = (uint32_t)dist; S.bm_trail = trail; S.bm_pay = q;
                  S.give_up = 0;
                }
              }
            }
          }
        }
        return false;
      }
      if (e == LZP_END || e == LZP_FAR) {
        const LzpInst r = lzp_decode(S.win, (int)(vis & (LZP_DEC - 1)), (vis & LZP_FIRST) ? 2 : (int)((vis >> LZP_DEC_BITS) & 1), wlim, at_end);
        if (e == LZP_END) {
          S.ended = 1;
          S.ip_end = w_lo + (int64_t)(vis & (LZP_DEC - 1)) + r.hdr; // the terminator produces nothing: not listed
        } else {
          S.visited[nlist++] = (uint16_t)vis;
          ip_after = w_lo + (int64_t)(vis & (LZP_DEC - 1)) + r.hdr + r.lit + r.trail;
          cls_after = r.ncls;
        }
        return false;
      }
      S.visited[nlist++] = (uint16_t)vis;
      xcur = e;
      return true;
    };
    if (tid == 0) {
      S.bm_len = 0;
      xcur = (uint32_t)p0 | ((uint32_t)(cls0 & 1) << LZP_DEC_BITS);
      bool stop = false;
      if (cls0 == 2) stop = !single(link(lzp_decode(S.win, p0, 2, wlim, at_end), p0), (uint32_t)p0 | LZP_FIRST);
      int m = 0;
      if (!stop) {
        for (;;) {
          S.ms[m++] = (uint16_t)xcur;
          const uint32_t j = J[xcur];
          if (j >= 2u * LZP_DEC || m == LZP_MAXMS) break;
          xcur = j;
        }
      }
      S.nms = m;
      S.npre = nlist;
    }
    __syncthreads();
    {
      // between two milestones: count, scan, list
      const int m = S.nms, npre = S.npre;
      const int nseg = m > 0 ? m - 1 : 0;
      const int per = (nseg + LZP_PT - 1) / LZP_PT;
      const int i0 = tid * per;
      uint32_t cnt = 0;
      for (int i = i0; i < i0 + per && i < nseg; i++) {
        const uint32_t stop_at = S.ms[i + 1];
        for (uint32_t y = S.ms[i]; y != stop_at && y < 2u * LZP_DEC && cnt <= (uint32_t)LZP_MAXI; y = S.next[y]) cnt++;
      }
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (lane == 31) S.part[warp] = incl;
      __syncthreads();
      uint32_t at = (uint32_t)npre + incl - cnt;
      uint32_t total = 0;
      for (int w2 = 0; w2 < LZP_PT / 32; w2++) { if (w2 < warp) at += S.part[w2]; total += S.part[w2]; }
      for (int i = i0; i < i0 + per && i < nseg; i++) {
        const uint32_t stop_at = S.ms[i + 1];
        for (uint32_t y = S.ms[i]; y != stop_at && y < 2u * LZP_DEC && at < (uint32_t)LZP_MAXI; y = S.next[y]) S.visited[at++] = (uint16_t)y;
      }
      if (tid == 0) {
        nlist = npre + (int)total;
        if (nlist >= LZP_MAXI) { nlist = 0; S.give_up = 1; } // cannot happen: an instruction is at least two bytes long
        if (m > 0) {
          const uint32_t xl = S.ms[m - 1]; // the last milestone: its link ends the chain, unless the milestone list was full
          xcur = xl;
          if (J[xl] >= 2u * LZP_DEC) {
            // (its link may still be an ordinary one when a doubling raced: then walk on, instruction by instruction)
            while (nlist < LZP_MAXI - 1 && single(S.next[xcur], xcur)) {}
          }
        }
        S.nvis = nlist;
        S.bulk_len = bulk;
        S.ip_after = ip_after >= 0 ? ip_after : w_lo + (int64_t)(xcur & (LZP_DEC - 1));
        S.cls_after = ip_after >= 0 ? cls_after : (int)(xcur >> LZP_DEC_BITS);
        if (nlist == 0 && !bulk && !S.bm_len && !S.ended) S.give_up = 1; // no progress: cannot happen on the happy path
      }
    }
    __syncthreads();
    if (S.give_up) break;
    if (S.nvis == 0 && S.ended) break; // the terminator was the round's first instruction
    if (S.bulk_len) {
      // ---- a literal run longer than a window: input -> output, in pieces the 32-bit drain can take
      int64_t left = S.bulk_len;
      const uint8_t *sp = in + S.ip;
      while (left > 0) {
        const uint32_t piece = (uint32_t)(left < (1 << 20) ? left : (1 << 20));
        const uint32_t o = S.out_pos;
        raw_copy(S, DCMP_FMT_RAW, sp, piece, out + o, o, in, in + ilen, tid, lane, warp);
        sp += piece;
        left -= piece;
      }
      if (tid == 0) { S.ip += S.bulk_len; S.cls = 1; }
      __syncthreads();
      continue;
    }
    auto bulk_match = [&]() {
      // A match longer than the symbol buffer is replicated output -> output directly: byte k repeats byte k mod dist
      // of the period in front of it, which is complete, so the copy has no order (lzo.ml:206-216 semantics); then
      // its trailing literals.  Arguments in S.bm_*; all threads call.
      const uint32_t o_abs = S.out_pos;
      uint8_t *const o = out + o_abs;
      const uint32_t dist = S.bm_dist, mlen = (uint32_t)S.bm_len;
      const int trail = S.bm_trail;
      const int64_t pay = S.bm_pay;
      for (uint32_t kk = (uint32_t)tid; kk < mlen; kk += LZP_PT) o[kk] = __ldcg(o - dist + kk % dist);
      if (tid < trail) o[mlen + tid] = in[pay + tid];
      __syncthreads();
      if (tid == 0) {
        S.out_pos = o_abs + mlen + (uint32_t)trail;
        S.ip = pay + trail;
        S.cls = trail ? 1 : 0;
        S.ended = 0; // (a terminator seen further down the chain is met again)
        S.bm_len = 0;
      }
      __syncthreads();
    };
    if (S.bm_len) { bulk_match(); continue; }
    // ---- 3: output lengths, scan, cut at the budget, validity (a thread owns `per` consecutive instructions and
    // decodes them again in every pass: cheaper than keeping them)
    const int n = S.nvis;
    const int per = (n + LZP_PT - 1) / LZP_PT; // consecutive instructions per thread
    const int j0 = tid * per;
    auto inst = [&](int j) {
      const uint32_t vv = S.visited[j];
      return lzp_decode(S.win, (int)(vv & (LZP_DEC - 1)), (vv & LZP_FIRST) ? 2 : (int)((vv >> LZP_DEC_BITS) & 1), wlim, at_end);
    };
    auto out_len = [](const LzpInst &r) {
      const uint32_t ol = (uint32_t)(r.lit + r.mlen + r.trail);
      return ol > (uint32_t)LZP_RSYM ? (uint32_t)LZP_RSYM + 1u : ol; // never fits: the round is cut before it
    };
    uint32_t len_sum = 0;
#pragma unroll 1
    for (int j = j0; j < j0 + per && j < n; j++) len_sum += out_len(inst(j));
    uint32_t incl = len_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) S.part[warp] = incl;
    if (tid == 0) S.cut = n;
    __syncthreads();
    uint32_t base = incl - len_sum;
    for (int w2 = 0; w2 < warp; w2++) base += S.part[w2];
    // the first instruction that overflows the symbol budget ends the round (the next one starts with it)
    {
      uint32_t o = base;
#pragma unroll 1
      for (int j = j0; j < j0 + per && j < n && o <= (uint32_t)LZP_RSYM; j++) {
        const uint32_t ol = out_len(inst(j));
        if (o + ol > (uint32_t)LZP_RSYM) atomicMin(&S.cut, j);
        o += ol;
      }
    }
    __syncthreads();
    const int ncut = S.cut;
    if (ncut == 0) {
      // the round's first instruction alone is larger than the symbol buffer: a very long match (run-length data)
      const LzpInst r = inst(0);
      const bool ok = !r.kind && r.mlen && (int64_t)r.dist <= (int64_t)out0 && (int64_t)out0 + r.mlen + r.trail <= olen;
      __syncthreads();
      if (tid == 0) {
        if (ok) {
          S.bm_len = r.mlen; S.bm_dist = r.dist; S.bm_trail = r.trail;
          S.bm_pay = w_lo + (int64_t)(S.visited[0] & (LZP_DEC - 1)) + r.hdr;
        } else S.give_up = 1;
      }
      __syncthreads();
      if (S.give_up) break;
      bulk_match();
      continue;
    }
    // validity: distance within the produced output (Invalid_dictionary / Out_of_bound otherwise), output window
    {
      uint32_t o = base;
#pragma unroll 1
      for (int j = j0; j < j0 + per && j < ncut; j++) {
        const LzpInst r = inst(j);
        const int64_t o_abs = (int64_t)out0 + o;
        const uint32_t ol = (uint32_t)(r.lit + r.mlen + r.trail);
        if (r.kind || (r.mlen && (int64_t)r.dist > o_abs) || o_abs + ol > olen) S.give_up = 1;
        o += ol;
        if (j == ncut - 1) S.part[LZP_PT - 1] = o; // bytes of the round (slots beyond the warp count are free)
      }
    }
    __syncthreads();
    if (S.give_up) break;
    const uint32_t round_out = S.part[LZP_PT - 1];
    const uint32_t bias = (uint32_t)((uintptr_t)(out + out0) & 15);
    uint8_t *const G0 = out + out0 - bias;
    const uint32_t g0_low = (uint32_t)((uintptr_t)G0 & HIST_MASK);
    const uint32_t E = bias + round_out;
    // ---- 4: replay into symbols
    if (tid < 16) {
      if ((uint32_t)tid < bias) S.sym[tid] = 0;
      S.sym[E + tid] = 0;
    }
    {
      int32_t i = (int32_t)(bias + base);
#pragma unroll 1
      for (int j = j0; j < j0 + per && j < ncut; j++) {
        const LzpInst r = inst(j);
        const uint8_t *pay = S.win + (S.visited[j] & (LZP_DEC - 1)) + r.hdr; // literal bytes follow the header
        if (r.mlen) {
          const int32_t s = i - (int32_t)r.dist; // symbol index of the first source byte (below bias: earlier rounds)
          const uint32_t len = (uint32_t)r.mlen, dist = r.dist;
          const bool overlap = dist < len; // an overlapping copy points at its first period
          for (uint32_t j0 = 0; j0 < len; j0 += 8) { // eight at a time: the history reads (L2) leave together
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const uint32_t jj = j0 + (uint32_t)u;
              const int32_t sj = s + (int32_t)(overlap ? jj % dist : jj);
              v[u] = 256u + (uint32_t)sj;
              if (jj < len && sj < (int32_t)bias) v[u] = __ldcg(G0 + sj);
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
              if (j0 + (uint32_t)u < len) S.sym[i + (int32_t)j0 + u] = (uint16_t)v[u];
          }
          i += (int32_t)len;
          for (int jj = 0; jj < r.trail; jj++) S.sym[i + jj] = pay[jj];
          i += r.trail;
        } else {
          for (int jj = 0; jj < r.lit; jj++) S.sym[i + jj] = pay[jj];
          i += r.lit;
        }
      }
    }
    if (tid == 0) {
      if (ncut < n) { // cut by the budget: the next round starts at the first instruction left out
        const uint32_t vv = S.visited[ncut];
        S.ip = w_lo + (int64_t)(vv & (LZP_DEC - 1));
        S.cls = (int)((vv >> LZP_DEC_BITS) & 1);
        S.ended = 0;
      } else {
        S.ip = S.ip_after;
        S.cls = S.cls_after;
      }
    }
    __syncthreads();
    // ---- 5: resolve markers, drain
    phase_e_drain(S, DCMP_FMT_RAW, G0, g0_low, bias, E, out0, round_out, tid, lane, warp);
  }
  __syncthreads();
  if (tid == 0) {
    if (S.ended && !S.give_up) {
      consumed[k] = S.ip_end;
      produced[k] = S.out_pos;
      status[k] = DCMP_OK;
    } else {
      consumed[k] = 0;
      produced[k] = 0;
      status[k] = LZP_RETRY;
    }
  }
  (void)mode;
}

size_t lzo_parallel_smem_bytes() { return sizeof(LzpSmem) + 128; }

#ifndef DCMP_EMU
int configure_lzo_parallel() { // per device (see configure_inflate_parallel)
  return (int)cudaFuncSetAttribute(lzo_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lzo_parallel_smem_bytes());
}
int launch_lzo_parallel(int mode, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                        uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                        int64_t *produced, int32_t *status, cudaStream_t stream) {
  if (n <= 0) return 0;
  const size_t smem = lzo_parallel_smem_bytes();
  lzo_parallel_kernel<<<(unsigned)n, LZP_PT, smem, stream>>>(mode, n, src_base, src_off, src_len, dst_base, dst_off, dst_cap,
                                                            consumed, produced, status);
  return 0;
}
#endif

} // namespace dcmp
