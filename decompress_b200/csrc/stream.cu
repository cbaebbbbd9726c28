// decompress_b200/csrc/stream.cu — the streaming half of the C ABI (include/dcmp.h: dcmp_inf_*, dcmp_def_*): the
// reference's non-blocking decoders and encoders (De.Inf.decoder, Zl.Inf, Gz.Inf, Zl.Def, Gz.Def, De.Higher) as
// host state machines (stream_core.h) whose codec work is device code — inflate_stream_kernel for the decoders, the
// batch deflate kernels for the encoders.  No CPU path: without a device the constructors fail with DCMP_API_NO_DEVICE.
#include <cuda_runtime.h>
#include <new>
#include "stream_core.h"

namespace dcmp {
void launch_inflate_stream(const uint8_t *d_in, int64_t n_in, int eof, uint8_t *d_ring, int64_t max_out, void *d_state,
                           cudaStream_t stream);
}
extern "C" int dcmp_engine_device(const dcmp_engine *e);

namespace {
using namespace dcmp;

#define SCU(call)                                                        \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) return e_ == cudaErrorMemoryAllocation ? DCMP_API_OUT_OF_MEMORY : DCMP_API_CUDA_ERROR; \
  } while (0)

// One resumable stream on the device: input staging, the ring (window + output of a launch), the state.
struct CudaInfBackend {
  dcmp_engine *eng;
  int device;
  cudaStream_t stream = nullptr;
  uint8_t *d_in = nullptr, *d_ring = nullptr;
  InfStreamState *d_state = nullptr;
  size_t in_cap = 0;
  int64_t launches = 0;
  explicit CudaInfBackend(dcmp_engine *e) : eng(e), device(dcmp_engine_device(e)) {}
  ~CudaInfBackend() {
    cudaSetDevice(device);
    if (d_in) cudaFree(d_in);
    if (d_ring) cudaFree(d_ring);
    if (d_state) cudaFree(d_state);
    if (stream) cudaStreamDestroy(stream);
  }
  int init() {
    SCU(cudaSetDevice(device));
    SCU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    SCU(cudaMalloc(&d_ring, (size_t)STREAM_RING));
    SCU(cudaMalloc(&d_state, sizeof(InfStreamState)));
    return DCMP_API_OK;
  }
  void reset() {}
  int run(const uint8_t *in, int64_t n, bool eof, InfStreamState &st, std::vector<uint8_t> &out) {
    SCU(cudaSetDevice(device));
    if ((size_t)n + 16 > in_cap) {
      if (d_in) { SCU(cudaFree(d_in)); d_in = nullptr; }
      in_cap = (size_t)n + (size_t)(n >> 2) + 65536;
      SCU(cudaMalloc(&d_in, in_cap));
    }
    if (n > 0) SCU(cudaMemcpyAsync(d_in, in, (size_t)n, cudaMemcpyHostToDevice, stream));
    SCU(cudaMemcpyAsync(d_state, &st, sizeof st, cudaMemcpyHostToDevice, stream));
    launch_inflate_stream(d_in, n, eof ? 1 : 0, d_ring, STREAM_MAX_OUT, d_state, stream);
    launches++;
    SCU(cudaGetLastError());
    SCU(cudaMemcpyAsync(&st, d_state, sizeof st, cudaMemcpyDeviceToHost, stream));
    SCU(cudaStreamSynchronize(stream));
    if (st.produced > 0) { // the ring holds them at [total_out - produced, total_out) mod STREAM_RING
      const size_t at = out.size();
      out.resize(at + (size_t)st.produced);
      const int64_t lo = (st.total_out - st.produced) & (STREAM_RING - 1);
      const int64_t first = st.produced < STREAM_RING - lo ? st.produced : STREAM_RING - lo;
      SCU(cudaMemcpyAsync(out.data() + at, d_ring + lo, (size_t)first, cudaMemcpyDeviceToHost, stream));
      if (first < st.produced)
        SCU(cudaMemcpyAsync(out.data() + at + first, d_ring, (size_t)(st.produced - first), cudaMemcpyDeviceToHost, stream));
      SCU(cudaStreamSynchronize(stream));
    }
    return DCMP_API_OK;
  }
};

// One segment through the batch deflate kernels (dcmp_deflate_batch_device), device buffers of its own.
struct CudaDefBackend {
  dcmp_engine *eng;
  int device;
  cudaStream_t stream = nullptr;
  uint8_t *d_buf = nullptr;
  size_t cap = 0;
  explicit CudaDefBackend(dcmp_engine *e) : eng(e), device(dcmp_engine_device(e)) {}
  ~CudaDefBackend() {
    cudaSetDevice(device);
    if (d_buf) cudaFree(d_buf);
    if (stream) cudaStreamDestroy(stream);
  }
  int init() {
    SCU(cudaSetDevice(device));
    SCU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return DCMP_API_OK;
  }
  int deflate(int level, const uint8_t *in, int64_t n, bool final, std::vector<uint8_t> &out) {
    SCU(cudaSetDevice(device));
    const int64_t bound = dcmp_deflate_bound(n, DCMP_FMT_RAW) + 64 * ((n >> 16) + 2); // + the byte-aligning markers between chunks
    const size_t o_dst = ((size_t)n + 255) & ~(size_t)255, o_res = (o_dst + (size_t)bound + 255) & ~(size_t)255;
    if (o_res + 64 > cap) {
      if (d_buf) { SCU(cudaFree(d_buf)); d_buf = nullptr; }
      cap = o_res + 64 + (o_res >> 2);
      SCU(cudaMalloc(&d_buf, cap));
    }
    if (n > 0) SCU(cudaMemcpyAsync(d_buf, in, (size_t)n, cudaMemcpyHostToDevice, stream));
    int64_t zero = 0, len = n, dcap = bound;
    int64_t *d_written = (int64_t *)(d_buf + o_res);
    int32_t *d_status = (int32_t *)(d_written + 1);
    const int rc = dcmp_deflate_batch_device(eng, level, DCMP_FMT_RAW | (final ? 0 : DCMP_DEFLATE_NO_FINAL), 1, d_buf, &zero, &len,
                                             d_buf + o_dst, &zero, &dcap, d_written, d_status, (void *)stream);
    if (rc != DCMP_API_OK) return rc;
    int64_t res[2] = {0, 0};
    SCU(cudaMemcpyAsync(res, d_written, 12, cudaMemcpyDeviceToHost, stream));
    SCU(cudaStreamSynchronize(stream));
    const int32_t st = (int32_t)(res[1] & 0xffffffff);
    if (st != DCMP_OK || res[0] < 0 || res[0] > bound) return DCMP_API_CUDA_ERROR;
    const size_t at = out.size();
    out.resize(at + (size_t)res[0]);
    if (res[0] > 0) SCU(cudaMemcpyAsync(out.data() + at, d_buf + o_dst, (size_t)res[0], cudaMemcpyDeviceToHost, stream));
    SCU(cudaStreamSynchronize(stream));
    return DCMP_API_OK;
  }
};
} // namespace

struct dcmp_inf { dcmp::InfCore<CudaInfBackend> core; dcmp_inf(int fmt, dcmp_engine *e) : core(fmt, e) {} };
struct dcmp_def { dcmp::DefCore<CudaDefBackend> core; dcmp_def(int fmt, int level, dcmp_engine *e) : core(fmt, level, (int64_t)4 << 20, e) {} };

extern "C" {

int dcmp_inf_new(dcmp_engine *e, int fmt, uint8_t *o, int64_t o_len, dcmp_inf **out) {
  if (!e || !out || o_len < 0 || (o_len > 0 && !o)) return DCMP_API_INVALID_ARGUMENT;
  if (fmt != DCMP_FMT_RAW && fmt != DCMP_FMT_ZLIB && fmt != DCMP_FMT_GZIP) return DCMP_API_INVALID_ARGUMENT;
  dcmp_inf *d = new (std::nothrow) dcmp_inf(fmt, e);
  if (!d) return DCMP_API_OUT_OF_MEMORY;
  const int rc = d->core.be.init();
  if (rc != DCMP_API_OK) { delete d; return rc; }
  d->core.set_dst(o, o_len);
  *out = d;
  return DCMP_API_OK;
}
void dcmp_inf_free(dcmp_inf *d) { delete d; }
int dcmp_inf_src(dcmp_inf *d, const uint8_t *buf, int64_t len) {
  if (!d || len < 0 || (len > 0 && !buf)) return DCMP_API_INVALID_ARGUMENT; // `Invalid_argument "Invalid bounds", de.ml:736-737
  d->core.src(buf, len);
  return DCMP_API_OK;
}
int dcmp_inf_decode(dcmp_inf *d) { return d ? d->core.decode() : DCMP_API_INVALID_ARGUMENT; }
int dcmp_inf_status(const dcmp_inf *d) { return d ? d->core.status : DCMP_API_INVALID_ARGUMENT; }
const char *dcmp_inf_error(const dcmp_inf *d) { return d ? d->core.msg : ""; }
int64_t dcmp_inf_dst_rem(const dcmp_inf *d) { return d ? d->core.dst_rem() : 0; }
int64_t dcmp_inf_src_rem(const dcmp_inf *d) { return d ? d->core.src_rem() : 0; }
void dcmp_inf_flush(dcmp_inf *d) { if (d) d->core.flush(); }
void dcmp_inf_reset(dcmp_inf *d) { if (d) { uint8_t *o = d->core.o; int64_t n = d->core.o_len; d->core.reset(); d->core.set_dst(o, n); } }
int dcmp_inf_dst(dcmp_inf *d, uint8_t *o, int64_t o_len) {
  if (!d || o_len < 0 || (o_len > 0 && !o)) return DCMP_API_INVALID_ARGUMENT;
  d->core.set_dst(o, o_len);
  return DCMP_API_OK;
}
uint32_t dcmp_inf_checksum(const dcmp_inf *d) { return d ? d->core.checksum() : 0; }
int64_t dcmp_inf_write(const dcmp_inf *d) { return d ? d->core.wr : 0; }
int64_t dcmp_inf_launches(const dcmp_inf *d) { return d ? d->core.be.launches : 0; }
int64_t dcmp_inf_buffered(const dcmp_inf *d) { return d ? (int64_t)(d->core.pend.size() + d->core.outq.size()) : 0; }
int dcmp_gz_inf_os(const dcmp_inf *d) { return d ? d->core.gz.os : -1; }
uint32_t dcmp_gz_inf_mtime(const dcmp_inf *d) { return d ? d->core.gz.mtime : 0; }
int64_t dcmp_gz_inf_field(const dcmp_inf *d, int which, const uint8_t **p) {
  if (!d || !p) return -1;
  const dcmp::GzHeader &g = d->core.gz;
  const std::string *s = which == 0 ? (g.has_name ? &g.name : nullptr) : which == 1 ? (g.has_comment ? &g.comment : nullptr)
                                    : which == 2 ? (g.has_extra ? &g.extra : nullptr) : nullptr;
  if (!s) return -1;
  *p = (const uint8_t *)s->data();
  return (int64_t)s->size();
}

int dcmp_def_new(dcmp_engine *e, int fmt, int level, dcmp_def **out) {
  if (!e || !out) return DCMP_API_INVALID_ARGUMENT;
  if (fmt != DCMP_FMT_RAW && fmt != DCMP_FMT_ZLIB && fmt != DCMP_FMT_GZIP) return DCMP_API_INVALID_ARGUMENT;
  if (level < 0 || level > 9) return DCMP_API_INVALID_ARGUMENT; // Lz77.state: "Invalid level", de.ml:4431-4437
  dcmp_def *d = new (std::nothrow) dcmp_def(fmt, level, e);
  if (!d) return DCMP_API_OUT_OF_MEMORY;
  const int rc = d->core.be.init();
  if (rc != DCMP_API_OK) { delete d; return rc; }
  *out = d;
  return DCMP_API_OK;
}
void dcmp_def_free(dcmp_def *d) { delete d; }
int dcmp_def_gz_header(dcmp_def *d, uint32_t mtime, int os, int ascii, int hcrc, const uint8_t *filename, int64_t filename_len,
                       const uint8_t *comment, int64_t comment_len) {
  if (!d || d->core.header_done) return DCMP_API_INVALID_ARGUMENT;
  dcmp::GzHeader &g = d->core.gz;
  g.mtime = mtime;
  g.os = os & 0xff;
  g.has_name = filename != nullptr; g.has_comment = comment != nullptr;
  if (filename) g.name.assign((const char *)filename, (size_t)filename_len);
  if (comment) g.comment.assign((const char *)comment, (size_t)comment_len);
  g.flg = (ascii ? 1 : 0) | (hcrc ? 2 : 0) | (filename ? 8 : 0) | (comment ? 16 : 0); // gz.ml:848-857
  return DCMP_API_OK;
}
int dcmp_def_src(dcmp_def *d, const uint8_t *buf, int64_t len) {
  if (!d || len < 0 || (len > 0 && !buf)) return DCMP_API_INVALID_ARGUMENT;
  d->core.src(buf, len);
  return DCMP_API_OK;
}
int dcmp_def_dst(dcmp_def *d, uint8_t *buf, int64_t len) {
  if (!d || len < 0 || (len > 0 && !buf)) return DCMP_API_INVALID_ARGUMENT;
  d->core.dst(buf, len);
  return DCMP_API_OK;
}
int dcmp_def_encode(dcmp_def *d) { return d ? d->core.encode() : DCMP_API_INVALID_ARGUMENT; }
int64_t dcmp_def_src_rem(const dcmp_def *d) { return d ? d->core.src_rem() : 0; }
int64_t dcmp_def_dst_rem(const dcmp_def *d) { return d ? d->core.dst_rem() : 0; }

} // extern "C"
