// tests/simt_emu/emu_stream.h — the streaming state machines of the product (decompress_b200/csrc/stream_core.h) over
// the CPU SIMT emulator instead of CUDA: the same InfCore / DefCore templates, back ends that run the REAL kernels
// (inflate_stream_kernel, the deflate kernels) on fibers.  TEST INFRASTRUCTURE ONLY (included by emu_inflate.cc): it
// lets the GPU-less suite hold the Await / Flush / End / Malformed protocol to the reference's sequences.
#include "../../decompress_b200/csrc/stream_core.h"

struct EmuInfBackend {
  std::vector<uint8_t> ring;
  long launches = 0;
  EmuInfBackend() : ring((size_t)dcmp::STREAM_RING) {}
  void reset() {}
  int run(const uint8_t *in, int64_t n, bool eof, dcmp::InfStreamState &st, std::vector<uint8_t> &out) {
    using namespace dcmp;
    std::vector<uint8_t> copy(in, in + n); // (the kernel reads a device copy in the product: nothing aliases the caller's buffer)
    copy.resize((size_t)n + 16);
    launches++;
    simt::launch(simt::dim3(1), simt::dim3(32), 0, [&]() {
      inflate_stream_kernel(copy.data(), n, eof ? 1 : 0, ring.data(), STREAM_MAX_OUT, &st);
    });
    for (int64_t k = st.total_out - st.produced; k < st.total_out; k++) out.push_back(ring[(size_t)(k & (STREAM_RING - 1))]);
    return DCMP_API_OK;
  }
};
struct EmuDefBackend {
  int deflate(int level, const uint8_t *in, int64_t n, bool final, std::vector<uint8_t> &out);
};

struct emu_inf { dcmp::InfCore<EmuInfBackend> core; explicit emu_inf(int fmt) : core(fmt) {} };
struct emu_def { dcmp::DefCore<EmuDefBackend> core; emu_def(int fmt, int level, int64_t seg) : core(fmt, level, seg) {} };

extern "C" {
emu_inf *emu_inf_new(int fmt, uint8_t *o, int64_t o_len) { emu_inf *d = new emu_inf(fmt); d->core.set_dst(o, o_len); return d; }
void emu_inf_free(emu_inf *d) { delete d; }
void emu_inf_src(emu_inf *d, const uint8_t *buf, int64_t len) { d->core.src(buf, len); }
int emu_inf_decode(emu_inf *d) { return d->core.decode(); }
int emu_inf_status(emu_inf *d) { return d->core.status; }
const char *emu_inf_error(emu_inf *d) { return d->core.msg; }
int64_t emu_inf_dst_rem(emu_inf *d) { return d->core.dst_rem(); }
int64_t emu_inf_src_rem(emu_inf *d) { return d->core.src_rem(); }
void emu_inf_flush(emu_inf *d) { d->core.flush(); }
void emu_inf_reset(emu_inf *d) { uint8_t *o = d->core.o; int64_t n = d->core.o_len; d->core.reset(); d->core.set_dst(o, n); }
uint32_t emu_inf_checksum(emu_inf *d) { return d->core.checksum(); }
int64_t emu_inf_write(emu_inf *d) { return d->core.wr; }
int64_t emu_inf_launches(emu_inf *d) { return d->core.be.launches; }
int64_t emu_inf_buffered(emu_inf *d) { return (int64_t)(d->core.pend.size() + d->core.outq.size()); }
int emu_gz_inf_os(emu_inf *d) { return d->core.gz.os; }
int64_t emu_gz_inf_field(emu_inf *d, int which, const uint8_t **p) {
  const dcmp::GzHeader &g = d->core.gz;
  const std::string *s = which == 0 ? (g.has_name ? &g.name : nullptr) : which == 1 ? (g.has_comment ? &g.comment : nullptr)
                                    : which == 2 ? (g.has_extra ? &g.extra : nullptr) : nullptr;
  if (!s) return -1;
  *p = (const uint8_t *)s->data();
  return (int64_t)s->size();
}
emu_def *emu_def_new(int fmt, int level, int64_t seg) { return new emu_def(fmt, level, seg); }
void emu_def_free(emu_def *d) { delete d; }
void emu_def_gz_header(emu_def *d, uint32_t mtime, int os, int ascii, int hcrc, const uint8_t *fn, int64_t fn_len, const uint8_t *cm,
                       int64_t cm_len) {
  dcmp::GzHeader &g = d->core.gz;
  g.mtime = mtime; g.os = os & 0xff;
  g.has_name = fn != nullptr; g.has_comment = cm != nullptr;
  if (fn) g.name.assign((const char *)fn, (size_t)fn_len);
  if (cm) g.comment.assign((const char *)cm, (size_t)cm_len);
  g.flg = (ascii ? 1 : 0) | (hcrc ? 2 : 0) | (fn ? 8 : 0) | (cm ? 16 : 0);
}
void emu_def_src(emu_def *d, const uint8_t *buf, int64_t len) { d->core.src(buf, len); }
void emu_def_dst(emu_def *d, uint8_t *buf, int64_t len) { d->core.dst(buf, len); }
int emu_def_encode(emu_def *d) { return d->core.encode(); }
int64_t emu_def_src_rem(emu_def *d) { return d->core.src_rem(); }
int64_t emu_def_dst_rem(emu_def *d) { return d->core.dst_rem(); }
}
