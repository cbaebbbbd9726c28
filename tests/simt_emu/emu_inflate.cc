// tests/simt_emu/emu_inflate.cc — runs the REAL kernel sources (decompress_b200/csrc/*.cu,
// compiled as C++ with -DDCMP_EMU) on the CPU SIMT emulator so their logic can be debugged
// without a GPU.  TEST INFRASTRUCTURE ONLY: built into tests/_build/, never part of the product.
#include "simt_emu.h"
#include <algorithm>
#include "../../decompress_b200/csrc/inflate_serial.cu"
#include "../../decompress_b200/csrc/inflate_parallel.cu"
#include "../../decompress_b200/csrc/lzo_decode.cu"
#include "../../decompress_b200/csrc/lzo_parallel.cu"
#include "../../decompress_b200/csrc/deflate.cu"

static int64_t dcmp_deflate_bound_emu(int64_t len) { // De.Def.Ns.compress_bound (de.ml:3993-3996)
  int64_t blocks = std::max<int64_t>(1, (len + 9999) / 10000);
  return 5 * blocks + len + 1 + 8;
}

extern "C" {

// mode 0: parallel kernel, then the general kernel on the streams it gave up (like engine.cu)
// mode 1: parallel kernel only (status 0x7f marks given-up streams)
// mode 2: general kernel only
int emu_inflate_batch(int mode, int fmt, int64_t n, const uint8_t *src, const int64_t *soff, const int64_t *slen,
                      uint8_t *dst, const int64_t *doff, const int64_t *dcap, int64_t *cons, int64_t *prod,
                      int32_t *status, int ctas) {
  using namespace dcmp;
  if (mode != 2) {
    unsigned long long counter = 0;
    if (ctas < 1) ctas = 1;
    std::vector<uint32_t> scratch(inflate_parallel_scratch_bytes_per_cta() / 4 * (size_t)ctas);
    uint32_t *tok_scratch = scratch.data();
    simt::launch(simt::dim3((unsigned)ctas), simt::dim3(PT), inflate_parallel_smem_bytes(), [&]() {
      inflate_parallel_kernel(fmt, n, src, soff, slen, dst, doff, dcap, cons, prod, status, &counter, tok_scratch);
    });
  }
  if (mode == 1) return 0;
  int retried = 0;
  for (int64_t k = 0; k < n; k++)
    if (mode == 2 || status[k] == DCMP_INTERNAL_RETRY) retried++;
  if (retried) {
    simt::launch(simt::dim3((unsigned)((n + SERIAL_WARPS - 1) / SERIAL_WARPS)), simt::dim3(SERIAL_WARPS * 32), 0, [&]() {
      inflate_serial_kernel(fmt, n, src, soff, slen, dst, doff, dcap, cons, prod, status, mode == 2 ? -1 : DCMP_INTERNAL_RETRY, nullptr);
    });
  }
  return retried;
}

// lzo_mode: 0 = Lzo.uncompress, 1 = uncompress_with_buffer; which: 0 = CTA-per-block kernel then the general one on
// the blocks it gave up (like engine.cu), 2 = general kernel only.  Returns the number of blocks given up.
int emu_lzo_batch2(int lzo_mode, int which, int64_t n, const uint8_t *src, const int64_t *soff, const int64_t *slen, uint8_t *dst,
                   const int64_t *doff, const int64_t *dcap, int64_t *cons, int64_t *prod, int32_t *status) {
  using namespace dcmp;
  int retried = 0;
  if (which != 2) {
    // which: 0 = the CTA-per-block kernel in its small-round shape (+ the general kernel for what it gives up), 3 = in its
    // large-round shape, 2 = the general kernel alone
    const int shape = which == 3 ? 1 : 0;
    simt::launch(simt::dim3((unsigned)n), simt::dim3(LZP_PT), lzo_parallel_smem_bytes(shape), [&]() {
      if (shape) lzp_large::lzo_parallel_kernel(lzo_mode, n, src, soff, slen, dst, doff, dcap, cons, prod, status);
      else lzp_small::lzo_parallel_kernel(lzo_mode, n, src, soff, slen, dst, doff, dcap, cons, prod, status);
    });
    for (int64_t k = 0; k < n; k++) retried += status[k] == LZP_RETRY;
    if (!retried) return 0;
  }
  simt::launch(simt::dim3((unsigned)((n + LZO_WARPS - 1) / LZO_WARPS)), simt::dim3(LZO_WARPS * 32), (size_t)LZO_WARPS * (LZO_RING + LZO_CACHE), [&]() {
    lzo_decode_kernel(lzo_mode, n, src, soff, slen, dst, doff, dcap, cons, prod, status, which == 2 ? -1 : LZP_RETRY, nullptr);
  });
  return retried;
}
int emu_lzo_batch(int mode, int64_t n, const uint8_t *src, const int64_t *soff, const int64_t *slen, uint8_t *dst,
                  const int64_t *doff, const int64_t *dcap, int64_t *cons, int64_t *prod, int32_t *status) {
  return emu_lzo_batch2(mode, 0, n, src, soff, slen, dst, doff, dcap, cons, prod, status);
}

int emu_deflate_batch(int level, int fmt, int64_t n, const uint8_t *src, const int64_t *soff, const int64_t *slen, uint8_t *dst,
                      const int64_t *doff, const int64_t *dcap, int64_t *written, int32_t *status) {
  using namespace dcmp;
  const bool no_final = (fmt & DCMP_DEFLATE_NO_FINAL) != 0; // (engine.cu: one segment of a stream the caller closes)
  fmt &= ~DCMP_DEFLATE_NO_FINAL;
  std::vector<DfChunk> chunks;
  std::vector<int64_t> first(n + 1);
  for (int64_t k = 0; k < n; k++) {
    first[k] = (int64_t)chunks.size();
    int64_t pos = 0;
    do {
      int64_t m = std::min<int64_t>(DF_CHUNK, slen[k] - pos);
      chunks.push_back(DfChunk{soff[k] + pos, (int32_t)m, (int32_t)k, (!no_final && pos + m == slen[k]) ? 1 : 0, (int32_t)std::min<int64_t>(pos, 32768)});
      pos += m;
    } while (pos < slen[k]);
  }
  first[n] = (int64_t)chunks.size();
  int64_t nc = (int64_t)chunks.size();
  std::vector<DfChunkOut> outs(nc);
  std::vector<uint8_t> slots((size_t)nc * DF_SLOT);
  std::vector<int64_t> cdst(nc);
  // the two compressors share the batch as in launch_deflate: small single-chunk streams go to the warp-per-chunk kernel,
  // everything else to the CTA-per-chunk kernel (DCMP_EMU_DEF_KERNEL=warp: all to the former)
  const char *fk = getenv("DCMP_EMU_DEF_KERNEL");
  const bool force_warp = fk && fk[0] == 'w';
  int64_t nsmall = 0;
  for (const auto &c : chunks) nsmall += (c.len <= DF_SMALL && c.hist == 0) ? 1 : 0;
  if (force_warp) nsmall = nc;
  int ctas = 2;
  unsigned long long counters[2] = {0, 0};
  if (nc - nsmall > 0) {
    std::vector<uint32_t> tokc((size_t)ctas * DF_CHUNK);
    simt::launch(simt::dim3(ctas), simt::dim3(dc_threads(level >= 7 ? 16 : 8)), deflate_cta_smem_bytes(level), [&]() {
      if (level >= 7) deflate_cta_kernel<16>(level, fmt == DCMP_FMT_ZLIB, nc, src, chunks.data(), outs.data(), slots.data(), tokc.data(), &counters[0], 0);
      else deflate_cta_kernel<8>(level, fmt == DCMP_FMT_ZLIB, nc, src, chunks.data(), outs.data(), slots.data(), tokc.data(), &counters[0], 0);
    });
  }
  if (nsmall > 0) {
    std::vector<uint32_t> tok((size_t)ctas * DF_WARPS * (force_warp ? DF_CHUNK : DF_SMALL));
    simt::launch(simt::dim3(ctas), simt::dim3(DF_WARPS * 32), deflate_smem_bytes(level), [&]() {
      if (level >= 7) deflate_chunk_kernel<DF_HASH_BITS_BEST>(level, fmt == DCMP_FMT_ZLIB, nc, src, chunks.data(), outs.data(), slots.data(), tok.data(), &counters[1], force_warp ? 0 : 1);
      else deflate_chunk_kernel<DF_HASH_BITS_FAST>(level, fmt == DCMP_FMT_ZLIB, nc, src, chunks.data(), outs.data(), slots.data(), tok.data(), &counters[1], force_warp ? 0 : 1);
    });
  }
  simt::launch(simt::dim3((unsigned)((n + 3) / 4)), simt::dim3(128), 0, [&]() {
    deflate_plan_kernel(fmt, level, n, first.data(), outs.data(), chunks.data(), dst, doff, dcap, cdst.data(), written, status);
  });
  simt::launch(simt::dim3((unsigned)nc), simt::dim3(256), 0, [&]() {
    deflate_gather_kernel(chunks.data(), outs.data(), slots.data(), cdst.data(), dst, doff);
  });
  return 0;
}

// De.Def.encode over a Queue / De.T.make on the emulator (deflate_queue_kernel, deflate_tmake_kernel)
int emu_def_encode_queue(const int32_t *cmds, int64_t n, int last, uint8_t *dst, int64_t cap, int64_t *written) {
  using namespace dcmp;
  std::vector<uint32_t> tok((size_t)n + 8), slot((size_t)n * 2 + 1024);
  int32_t status = -1;
  simt::launch(simt::dim3(1), simt::dim3(32), sizeof(DfSmemT<DF_HASH_BITS_BEST, uint64_t>) + 128, [&]() {
    deflate_queue_kernel(cmds, n, last, tok.data(), slot.data(), dst, cap, written, &status);
  });
  return status;
}
int emu_t_make(int length, int max_bits, const int32_t *freqs, int32_t *lens, int32_t *codes) {
  using namespace dcmp;
  int32_t mc = -1;
  simt::launch(simt::dim3(1), simt::dim3(32), sizeof(DfSmemT<DF_HASH_BITS_BEST, uint64_t>) + 128, [&]() {
    deflate_tmake_kernel(freqs, length, max_bits, lens, codes, &mc);
  });
  return mc;
}

int emu_t_make_fast(int length, int max_bits, const int32_t *freqs, int32_t *lens, int32_t *codes) {
  using namespace dcmp;
  int32_t mc = -1;
  simt::launch(simt::dim3(1), simt::dim3(32), 16384, [&]() { deflate_tmake_fast_kernel(freqs, length, max_bits, lens, codes, &mc); });
  return mc;
}

long emu_switches() { return simt::g_switches; }
} // extern "C"
#include "emu_stream.h"
int EmuDefBackend::deflate(int level, const uint8_t *in, int64_t n, bool final, std::vector<uint8_t> &out) {
  int64_t zero = 0, len = n, cap = dcmp_deflate_bound_emu(n) + 64 * ((n >> 16) + 2), written = 0;
  int32_t status = -1;
  std::vector<uint8_t> dst((size_t)cap + 16);
  std::vector<uint8_t> src(in, in + n);
  src.resize((size_t)n + 16);
  emu_deflate_batch(level, DCMP_FMT_RAW | (final ? 0 : DCMP_DEFLATE_NO_FINAL), 1, src.data(), &zero, &len, dst.data(), &zero, &cap, &written, &status);
  if (status != DCMP_OK) return DCMP_API_CUDA_ERROR;
  out.insert(out.end(), dst.begin(), dst.begin() + written);
  return DCMP_API_OK;
}
extern "C" {
void emu_stats(long *out) { memcpy(out, &dcmp::g_stats, sizeof(dcmp::g_stats)); memset(&dcmp::g_stats, 0, sizeof(dcmp::g_stats)); }
}
