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
// Two round shapes of the same kernel (lzo_parallel_body.inc):
//   lzp_small  4096 candidates and 10 K symbols per round, four CTAs per SM — a batch that fills the GPU (C5 at 512+ blocks:
//              +45 % over the large shape);
//   lzp_large  8192 candidates and 20 K symbols per round, two CTAs per SM — a batch of at most two blocks per SM, where
//              the time is the latency of one block's rounds and half as many of them is what counts (C5 at 256 blocks).
namespace lzp_small {
#define LZP_CFG_DEC_BITS DCMP_LZP_DEC_BITS
#define LZP_CFG_RSYM_K DCMP_LZP_RSYM_K
#define LZP_CFG_CTAS DCMP_LZP_CTAS
#include "lzo_parallel_body.inc"
#undef LZP_CFG_DEC_BITS
#undef LZP_CFG_RSYM_K
#undef LZP_CFG_CTAS
} // namespace lzp_small
namespace lzp_large {
#define LZP_CFG_DEC_BITS 13
#define LZP_CFG_RSYM_K 20
#define LZP_CFG_CTAS 2
#include "lzo_parallel_body.inc"
#undef LZP_CFG_DEC_BITS
#undef LZP_CFG_RSYM_K
#undef LZP_CFG_CTAS
} // namespace lzp_large
constexpr int LZP_PT = lzp_small::LZP_PT;
constexpr int LZP_RETRY = lzp_small::LZP_RETRY;
static_assert(lzp_large::LZP_PT == LZP_PT && lzp_large::LZP_RETRY == LZP_RETRY, "");

// which shape a batch of n blocks gets on a GPU of `sms` SMs (0: small rounds, 1: large rounds)
inline int lzo_parallel_shape(int64_t n, int sms) {
  const char *v = getenv("DCMP_LZP_SHAPE"); // (measurements: s / l)
  if (v && (v[0] == 's' || v[0] == 'l')) return v[0] == 'l';
  return n <= 2 * (int64_t)sms;
}
size_t lzo_parallel_smem_bytes(int shape) { return (shape ? sizeof(lzp_large::LzpSmem) : sizeof(lzp_small::LzpSmem)) + 128; }

#ifndef DCMP_EMU
int configure_lzo_parallel() { // per device (see configure_inflate_parallel)
  cudaError_t e = cudaFuncSetAttribute(lzp_small::lzo_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lzo_parallel_smem_bytes(0));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lzp_large::lzo_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lzo_parallel_smem_bytes(1));
  return (int)e;
}
int launch_lzo_parallel(int mode, int64_t n, int sms, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                        uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                        int64_t *produced, int32_t *status, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (lzo_parallel_shape(n, sms))
    lzp_large::lzo_parallel_kernel<<<(unsigned)n, LZP_PT, lzo_parallel_smem_bytes(1), stream>>>(mode, n, src_base, src_off, src_len, dst_base, dst_off,
                                                                                               dst_cap, consumed, produced, status);
  else
    lzp_small::lzo_parallel_kernel<<<(unsigned)n, LZP_PT, lzo_parallel_smem_bytes(0), stream>>>(mode, n, src_base, src_off, src_len, dst_base, dst_off,
                                                                                               dst_cap, consumed, produced, status);
  return 0;
}
#endif

} // namespace dcmp
