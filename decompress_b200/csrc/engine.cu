// decompress_b200/csrc/engine.cu — the C ABI (include/dcmp.h) over the CUDA kernels.
//
// Host-side plumbing only: device/stream ownership, scratch buffers, host<->device
// staging for the *_batch entry points, kernel dispatch.  No codec logic lives here
// and nothing here can run a stream on the CPU: without a device every compute call
// fails with DCMP_API_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../include/dcmp.h"
#include "status_strings.h"

namespace dcmp {
void launch_inflate_serial(int fmt, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                           uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                           int64_t *produced, int32_t *status, int only_status, unsigned long long *redo_count,
                           cudaStream_t stream);
int launch_inflate_parallel(int fmt, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                            uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                            int64_t *produced, int32_t *status, unsigned long long *work_counter, uint32_t *tok_scratch,
                            int ctas, cudaStream_t stream);
size_t inflate_parallel_scratch_bytes_per_cta();
int configure_inflate_parallel();
int configure_lzo_parallel();
int configure_lzo_decode();
int configure_deflate();
int inflate_parallel_ctas_per_sm();
void launch_gzip_header(int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len, int64_t *body_off,
                        int64_t *body_len, int32_t *pre_status, int rfc_xlen, cudaStream_t stream);
void launch_gzip_trailer(int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *body_off, const int64_t *body_len,
                         const int32_t *pre_status, const uint8_t *dst_base, const int64_t *dst_off, int64_t *consumed,
                         int64_t *produced, int32_t *status, cudaStream_t stream);
void launch_lzo_decode(int mode, int64_t n, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                       uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                       int64_t *produced, int32_t *status, int only_status, unsigned long long *redo_count,
                       cudaStream_t stream);
int launch_lzo_parallel(int mode, int64_t n, int sms, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                        uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *consumed,
                        int64_t *produced, int32_t *status, cudaStream_t stream);
int launch_deflate(int level, int fmt, int64_t nstreams, int64_t nchunks, int64_t nsmall, const uint8_t *src_base, const void *chunks,
                   const int64_t *first_chunk, void *outs, uint8_t *slots, int64_t *chunk_dst, uint8_t *scratch,
                   int sm_count, uint8_t *dst_base, const int64_t *dst_off,
                   const int64_t *dst_cap, int64_t *written, int32_t *status, cudaStream_t stream);
void launch_gzip_wrap(int64_t n, int level, int os, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                      uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *written, int32_t *status,
                      cudaStream_t stream);
void launch_deflate_queue(const int32_t *cmds, int64_t n, int last, uint32_t *tok, uint32_t *slot, uint8_t *dst, int64_t cap, int64_t *written,
                          int32_t *status, cudaStream_t stream);
void launch_deflate_tmake(const int32_t *freqs, int length, int max_bits, int32_t *lens, int32_t *codes, int32_t *max_code, cudaStream_t stream);
size_t deflate_scratch_bytes(int sm_count);
int deflate_small_chunk();
size_t deflate_slot_bytes();
int deflate_chunk_size();
struct DfChunkHost { int64_t src_off; int32_t len; int32_t stream; int32_t flags; int32_t hist; }; // = deflate.cu's DfChunk
constexpr int INTERNAL_RETRY = 0x7f; // inflate_parallel.cu: "this stream is for the general kernel"
}

static thread_local char g_err[512] = "";
static void set_err(const char *what, cudaError_t e) {
  snprintf(g_err, sizeof g_err, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) {                             \
      set_err(#call, e_);                                \
      return e_ == cudaErrorMemoryAllocation ? DCMP_API_OUT_OF_MEMORY : DCMP_API_CUDA_ERROR; \
    }                                                    \
  } while (0)

struct dcmp_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint8_t *d_buf = nullptr; // one growable device arena: [src | dst | tables]
  size_t d_cap = 0;
  int64_t launches = 0;
  int64_t n_inflate = 0, n_lzo = 0;  // units submitted to the two-stage dispatch (dcmp_engine_stats)
  int sm_count = 0;
  int par_ctas = 0;                  // persistent CTAs of the sub-chunk-parallel kernel (2 per SM)
  uint8_t *d_scratch = nullptr;      // [work counter, +64: inflate streams redone, +72: LZO blocks redone (256 B) | token scratch of every CTA]
  int mode = 0;                      // 0 = parallel + general, 1 = general kernel only (DCMP_INFLATE_MODE=serial)
  uint8_t *d_gz = nullptr;           // gzip: rebased body offsets / lengths / header status per member
  size_t gz_cap = 0;
  std::vector<void *> retired;       // scratch that was outgrown while launches on other streams might still read it: freed with the engine
  // The *_batch_device calls share per-engine scratch (work counters, token scratch, the gzip tables, the deflate arena):
  // every call waits for the previous one — whatever stream that was enqueued on — before it touches them, and records
  // where it ends.  Two calls on two streams are therefore serialised on the device, never corrupted.
  cudaEvent_t ev_last = nullptr;
  int chain_begin(cudaStream_t st) {
    if (!ev_last) CU(cudaEventCreateWithFlags(&ev_last, cudaEventDisableTiming));
    else CU(cudaStreamWaitEvent(st, ev_last, 0));
    return DCMP_API_OK;
  }
  int chain_end(cudaStream_t st) {
    CU(cudaEventRecord(ev_last, st));
    return DCMP_API_OK;
  }
  // deflate: persistent CTAs (3 per SM), their token scratch, and a growable arena for the per-chunk state
  uint8_t *d_def_scratch = nullptr;  // [work counters (256 B) | token scratch of every CTA of the two kernels]
  uint8_t *d_def_arena = nullptr;
  size_t def_arena_cap = 0;
  // the chunk table and the per-stream tables of a deflate call are built in pinned host memory the engine owns and
  // uploaded asynchronously: the call does not wait for the stream; the next call waits for the uploads of this one
  // (ev_tab) before it overwrites them
  uint8_t *h_def_tab = nullptr;
  size_t def_tab_cap = 0;
  cudaEvent_t ev_tab = nullptr;
  bool tab_in_flight = false;
  int reserve_tab(size_t bytes) {
    if (!ev_tab) CU(cudaEventCreateWithFlags(&ev_tab, cudaEventDisableTiming));
    if (tab_in_flight) { CU(cudaEventSynchronize(ev_tab)); tab_in_flight = false; }
    if (bytes <= def_tab_cap) return DCMP_API_OK;
    if (h_def_tab) { CU(cudaFreeHost(h_def_tab)); h_def_tab = nullptr; def_tab_cap = 0; }
    CU(cudaMallocHost((void **)&h_def_tab, bytes + (bytes >> 2) + 4096));
    def_tab_cap = bytes + (bytes >> 2) + 4096;
    return DCMP_API_OK;
  }
  // host-facing calls: three-stage pipeline (H2D | kernels | D2H) over slices of the batch
  static const int MAX_SLICES = 16;
  cudaStream_t s_kernel = nullptr, s_out = nullptr, s_meta = nullptr;
  cudaEvent_t ev_in[MAX_SLICES] = {}, ev_k[MAX_SLICES] = {}, ev_meta[MAX_SLICES] = {};
  bool pipeline_ready = false;
  int ensure_pipeline() {
    if (pipeline_ready) return DCMP_API_OK;
    CU(cudaStreamCreateWithFlags(&s_kernel, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s_meta, cudaStreamNonBlocking));
    for (int i = 0; i < MAX_SLICES; i++) {
      CU(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_meta[i], cudaEventDisableTiming));
    }
    pipeline_ready = true;
    return DCMP_API_OK;
  }
  int reserve(size_t bytes) {
    if (bytes <= d_cap) return DCMP_API_OK;
    if (d_buf) { CU(cudaFree(d_buf)); d_buf = nullptr; d_cap = 0; }
    size_t want = bytes + (bytes >> 3) + (1 << 20);
    CU(cudaMalloc(&d_buf, want));
    d_cap = want;
    return DCMP_API_OK;
  }
};

// The outputs of a batch packed back to back (one CTA per stream): what the host-facing deflate call copies to the host in
// ONE transfer when the batch is many small streams, instead of one cudaMemcpyAsync per stream.
__global__ void __launch_bounds__(128)
dcmp_pack_outputs_kernel(const uint8_t *base, const int64_t *off, const int64_t *len, const int64_t *packed_off, uint8_t *packed) {
  const int64_t k = blockIdx.x;
  const int64_t n = len[k];
  const uint8_t *s = base + off[k];
  uint8_t *d = packed + packed_off[k];
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

extern "C" {

int dcmp_abi_version(void) { return DCMP_ABI_VERSION; }

int dcmp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char *dcmp_last_error(void) { return g_err; }

const char *dcmp_status_string(int s) { return dcmp::status_string(s); }

int dcmp_engine_create(int device, dcmp_engine **out) {
  if (!out) return DCMP_API_INVALID_ARGUMENT;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    snprintf(g_err, sizeof g_err, "no CUDA device (%s); libdcmp_b200 has no CPU fallback", cudaGetErrorName(e));
    return DCMP_API_NO_DEVICE;
  }
  if (device < 0 || device >= n) return DCMP_API_INVALID_ARGUMENT;
  CU(cudaSetDevice(device));
  dcmp_engine *eng = new dcmp_engine();
  eng->device = device;
  cudaError_t se = cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking);
  if (se != cudaSuccess) { set_err("cudaStreamCreate", se); delete eng; return DCMP_API_CUDA_ERROR; }
  cudaDeviceGetAttribute(&eng->sm_count, cudaDevAttrMultiProcessorCount, device);
  eng->par_ctas = dcmp::inflate_parallel_ctas_per_sm() * (eng->sm_count > 0 ? eng->sm_count : 148);
  se = cudaMalloc(&eng->d_scratch, 256 + (size_t)eng->par_ctas * dcmp::inflate_parallel_scratch_bytes_per_cta());
  if (se != cudaSuccess) { set_err("cudaMalloc(scratch)", se); cudaStreamDestroy(eng->stream); delete eng; return DCMP_API_OUT_OF_MEMORY; }
  cudaMemset(eng->d_scratch, 0, 256);
  // kernels that need more than 48 KiB of dynamic shared memory opt in per device
  int ce = dcmp::configure_inflate_parallel();
  if (ce == 0) ce = dcmp::configure_lzo_parallel();
  if (ce == 0) ce = dcmp::configure_lzo_decode();
  if (ce == 0) ce = dcmp::configure_deflate();
  if (ce != 0) {
    set_err("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", (cudaError_t)ce);
    cudaFree(eng->d_scratch); cudaStreamDestroy(eng->stream); delete eng;
    return DCMP_API_CUDA_ERROR;
  }
  const char *m = getenv("DCMP_INFLATE_MODE");
  eng->mode = (m && strcmp(m, "serial") == 0) ? 1 : 0;
  *out = eng;
  return DCMP_API_OK;
}

void dcmp_engine_destroy(dcmp_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->d_buf) cudaFree(e->d_buf);
  if (e->d_scratch) cudaFree(e->d_scratch);
  if (e->d_gz) cudaFree(e->d_gz);
  for (void *p : e->retired) cudaFree(p);
  if (e->ev_last) cudaEventDestroy(e->ev_last);
  if (e->d_def_scratch) cudaFree(e->d_def_scratch);
  if (e->d_def_arena) cudaFree(e->d_def_arena);
  if (e->h_def_tab) cudaFreeHost(e->h_def_tab);
  if (e->ev_tab) cudaEventDestroy(e->ev_tab);
  if (e->pipeline_ready) {
    for (int i = 0; i < dcmp_engine::MAX_SLICES; i++) { cudaEventDestroy(e->ev_in[i]); cudaEventDestroy(e->ev_k[i]); cudaEventDestroy(e->ev_meta[i]); }
    cudaStreamDestroy(e->s_kernel);
    cudaStreamDestroy(e->s_out);
    cudaStreamDestroy(e->s_meta);
  }
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int64_t dcmp_engine_launch_count(const dcmp_engine *e) { return e ? e->launches : 0; }
int dcmp_engine_device(const dcmp_engine *e) { return e ? e->device : -1; }

int dcmp_engine_stats(dcmp_engine *e, int64_t *out, int n_out) {
  if (!e || !out || n_out < 0) return DCMP_API_INVALID_ARGUMENT;
  CU(cudaSetDevice(e->device));
  CU(cudaDeviceSynchronize()); // (the counters are written by kernels on whatever stream the caller used)
  unsigned long long redo[2] = {0, 0};
  CU(cudaMemcpy(redo, e->d_scratch + 64, sizeof redo, cudaMemcpyDeviceToHost));
  const int64_t v[DCMP_STATS_COUNT] = {e->n_inflate, (int64_t)redo[0], e->n_lzo, (int64_t)redo[1], e->launches};
  for (int i = 0; i < n_out && i < DCMP_STATS_COUNT; i++) out[i] = v[i];
  return DCMP_API_OK;
}

int dcmp_inflate_batch_device(dcmp_engine *e, int fmt, int64_t n, const uint8_t *d_src_base, const int64_t *d_src_off,
                              const int64_t *d_src_len, uint8_t *d_dst_base, const int64_t *d_dst_off,
                              const int64_t *d_dst_cap, int64_t *d_consumed, int64_t *d_produced, int32_t *d_status,
                              void *cuda_stream) {
  if (!e) return DCMP_API_INVALID_ARGUMENT;
  const int rfc_xlen = (fmt & DCMP_GZIP_RFC_XLEN) ? 1 : 0;
  fmt &= ~DCMP_GZIP_RFC_XLEN;
  if (fmt != DCMP_FMT_RAW && fmt != DCMP_FMT_ZLIB && fmt != DCMP_FMT_GZIP) return DCMP_API_INVALID_ARGUMENT;
  if (n < 0) return DCMP_API_INVALID_ARGUMENT;
  if (n == 0) return DCMP_API_OK;
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)cuda_stream; // NULL = the legacy default stream, as everywhere in CUDA
  { const int rc0 = e->chain_begin(st); if (rc0 != DCMP_API_OK) return rc0; }
  // gzip: one member per unit.  The header is parsed on the device, the DEFLATE body goes through the raw
  // kernels with rebased offset tables, then CRC-32 / ISIZE are checked on the device (gzip_frame.cu).
  const int64_t *g_src_off = d_src_off;
  int64_t *gz_off = nullptr, *gz_len = nullptr;
  int32_t *gz_pre = nullptr;
  if (fmt == DCMP_FMT_GZIP) {
    const size_t need = (size_t)n * 20 + 256;
    if (need > e->gz_cap) {
      // (earlier launches may still read the old tables: they are kept until the engine goes, no device-wide synchronise)
      if (e->d_gz) { e->retired.push_back(e->d_gz); e->d_gz = nullptr; e->gz_cap = 0; }
      CU(cudaMalloc(&e->d_gz, need * 2));
      e->gz_cap = need * 2;
    }
    gz_off = (int64_t *)e->d_gz;
    gz_len = gz_off + n;
    gz_pre = (int32_t *)(gz_len + n);
    dcmp::launch_gzip_header(n, d_src_base, d_src_off, d_src_len, gz_off, gz_len, gz_pre, rfc_xlen, st);
    e->launches += 1;
    d_src_off = gz_off;
    d_src_len = gz_len;
    fmt = DCMP_FMT_RAW;
  }
  if (e->mode == 0) {
    // stage 1: sub-chunk-parallel kernel, persistent CTAs pulling streams from a device counter;
    // stage 2: the general kernel on exactly the streams stage 1 marked INTERNAL_RETRY.
    unsigned long long *counter = (unsigned long long *)e->d_scratch;
    CU(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    int rc = dcmp::launch_inflate_parallel(fmt, n, d_src_base, d_src_off, d_src_len, d_dst_base, d_dst_off, d_dst_cap,
                                           d_consumed, d_produced, d_status, counter, (uint32_t *)(e->d_scratch + 256),
                                           e->par_ctas, st);
    if (rc != 0) { set_err("launch_inflate_parallel", (cudaError_t)rc); return DCMP_API_CUDA_ERROR; }
    CU(cudaGetLastError());
    dcmp::launch_inflate_serial(fmt, n, d_src_base, d_src_off, d_src_len, d_dst_base, d_dst_off, d_dst_cap, d_consumed,
                                d_produced, d_status, dcmp::INTERNAL_RETRY, (unsigned long long *)(e->d_scratch + 64), st);
    e->launches += 2;
    e->n_inflate += n;
  } else {
    dcmp::launch_inflate_serial(fmt, n, d_src_base, d_src_off, d_src_len, d_dst_base, d_dst_off, d_dst_cap, d_consumed,
                                d_produced, d_status, -1, nullptr, st);
    e->launches += 1;
  }
  if (gz_pre) {
    dcmp::launch_gzip_trailer(n, d_src_base, g_src_off, gz_off, gz_len, gz_pre, d_dst_base, d_dst_off, d_consumed, d_produced,
                              d_status, st);
    e->launches += 1;
  }
  CU(cudaGetLastError());
  return e->chain_end(st);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Host-buffer driver shared by every *_batch entry point: uploads, runs `launch` (one of the *_batch_device
// functions, bound to its codec arguments) and downloads.
typedef int (*device_batch_fn)(dcmp_engine *e, int arg, int64_t n, const uint8_t *d_src_base, const int64_t *d_src_off,
                               const int64_t *d_src_len, uint8_t *d_dst_base, const int64_t *d_dst_off,
                               const int64_t *d_dst_cap, int64_t *d_consumed, int64_t *d_produced, int32_t *d_status,
                               void *cuda_stream);

static int host_batch(dcmp_engine *e, device_batch_fn launch, int fmt, int64_t n, const uint8_t *src_base,
                      const int64_t *src_off, const int64_t *src_len, uint8_t *dst_base, const int64_t *dst_off,
                      const int64_t *dst_cap, int64_t *consumed, int64_t *produced, int32_t *status) {
  if (!e) return DCMP_API_INVALID_ARGUMENT;
  if (n < 0) return DCMP_API_INVALID_ARGUMENT;
  if (n == 0) return DCMP_API_OK;
  if (!src_off || !src_len || !dst_off || !dst_cap || !consumed || !produced || !status)
    return DCMP_API_INVALID_ARGUMENT;
  CU(cudaSetDevice(e->device));
  // Spans of the host slabs actually referenced.
  int64_t s_lo = INT64_MAX, s_hi = 0, d_lo = INT64_MAX, d_hi = 0, work = 0;
  for (int64_t k = 0; k < n; k++) {
    if (src_len[k] < 0 || dst_cap[k] < 0 || src_off[k] < 0 || dst_off[k] < 0) return DCMP_API_INVALID_ARGUMENT;
    s_lo = std::min(s_lo, src_off[k]); s_hi = std::max(s_hi, src_off[k] + src_len[k]);
    d_lo = std::min(d_lo, dst_off[k]); d_hi = std::max(d_hi, dst_off[k] + dst_cap[k]);
    work += src_len[k] + dst_cap[k];
  }
  size_t s_bytes = (size_t)(s_hi - s_lo), d_bytes = (size_t)(d_hi - d_lo);
  if ((s_bytes && !src_base) || (d_bytes && !dst_base)) return DCMP_API_INVALID_ARGUMENT;
  size_t o_src = 0;
  size_t o_dst = align_up(o_src + s_bytes + 16, 256);
  size_t o_tab = align_up(o_dst + d_bytes + 16, 256);
  size_t tab_bytes = (size_t)n * (8 * 6 + 4);
  int rc = e->reserve(o_tab + tab_bytes + 256);
  if (rc != DCMP_API_OK) return rc;
  rc = e->ensure_pipeline();
  if (rc != DCMP_API_OK) return rc;
  // Offset tables rebased to the spans, packed into one upload.
  std::vector<int64_t> tab((size_t)n * 4);
  for (int64_t k = 0; k < n; k++) {
    tab[k] = src_off[k] - s_lo;
    tab[n + k] = src_len[k];
    tab[2 * n + k] = dst_off[k] - d_lo;
    tab[3 * n + k] = dst_cap[k];
  }
  uint8_t *d_src = e->d_buf + o_src, *d_dst = e->d_buf + o_dst;
  int64_t *d_tab = (int64_t *)(e->d_buf + o_tab);
  int64_t *d_cons = d_tab + 4 * n, *d_prod = d_tab + 5 * n;
  int32_t *d_stat = (int32_t *)(d_tab + 6 * n);
  // The batch moves through three stages on three CUDA streams — host->device copy of the compressed
  // bytes, the kernels, device->host copy of the inflated bytes — in slices of consecutive streams, so
  // that PCIe in both directions and the SMs work at the same time.  A slice's input is the byte span
  // its streams reference; if the caller's layout is not monotonic the batch runs as one slice.
  // The slices DOUBLE in size: the first one is small, so that its output starts down the link a fraction of a
  // millisecond into the call, and every later one is as big as everything before it, so that its kernel is done long before
  // the copy engine is through with its predecessors (C2, end to end against a 52 GB/s ceiling of the link: even slices of
  // 48 MB 42.8 GB/s — the first device-to-host byte left 0.7 ms into a 6.2 ms call — doubling slices from 2 / 4 / 8 / 16 / 32 MB
  // 42.7 / 43.8 / 44.5 / 45.2 / 43.6; gpurun_out/r2_e2e_slices.txt).
  static const int64_t first_mb = getenv("DCMP_SLICE_MB") ? std::max(1, atoi(getenv("DCMP_SLICE_MB"))) : 16; // tuning knobs
  static const bool even = getenv("DCMP_SLICE_EVEN") != nullptr;   // round 1's even slices of DCMP_SLICE_MB (A/B)
  int nslice = 1;
  if (even) nslice = (int)std::min<int64_t>(dcmp_engine::MAX_SLICES, std::max<int64_t>(1, work / (first_mb << 20)));
  else while (nslice < 8 && work / ((1ll << (nslice + 1)) - 1) >= (first_mb << 20)) nslice++;
  if (n < 4 * nslice) nslice = 1;
  bool monotonic = true;
  for (int64_t k = 1; k < n && monotonic; k++)
    monotonic = src_off[k] >= src_off[k - 1] + src_len[k - 1] && dst_off[k] >= dst_off[k - 1] + dst_cap[k - 1];
  if (!monotonic) nslice = 1;
  std::vector<int64_t> cut(nslice + 1, 0);
  {
    int64_t acc = 0;
    int c = 1;
    const long double whole = even ? (long double)nslice : (long double)((1ll << nslice) - 1);
    for (int64_t k = 0; k < n && c < nslice; k++) {
      acc += src_len[k] + dst_cap[k];
      const long double part = even ? (long double)c : (long double)((1ll << c) - 1);
      if ((long double)acc >= (long double)work * part / whole) cut[c++] = k + 1;
    }
    while (c < nslice) cut[c++] = n;
    cut[nslice] = n;
  }
  cudaStream_t s_in = e->stream, s_k = e->s_kernel, s_out = e->s_out;
  CU(cudaMemcpyAsync(d_tab, tab.data(), (size_t)n * 32, cudaMemcpyHostToDevice, s_in));
  for (int c = 0; c < nslice; c++) {
    int64_t k0 = cut[c], k1 = cut[c + 1];
    if (k1 > k0) {
      int64_t lo = nslice == 1 ? s_lo : src_off[k0], hi = nslice == 1 ? s_hi : src_off[k1 - 1] + src_len[k1 - 1];
      if (hi > lo) CU(cudaMemcpyAsync(d_src + (lo - s_lo), src_base + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, s_in));
    }
    CU(cudaEventRecord(e->ev_in[c], s_in));
    CU(cudaStreamWaitEvent(s_k, e->ev_in[c], 0));
    if (k1 > k0) {
      rc = launch(e, fmt, k1 - k0, d_src, d_tab + k0, d_tab + n + k0, d_dst, d_tab + 2 * n + k0,
                  d_tab + 3 * n + k0, d_cons + k0, d_prod + k0, d_stat + k0, (void *)s_k);
      if (rc != DCMP_API_OK) return rc;
    }
    CU(cudaEventRecord(e->ev_k[c], s_k));
  }
  // Copy back exactly what each stream produced and nothing else (the reference never
  // touches dst beyond o_pos).  Adjacent, completely filled windows are merged into one
  // transfer, so the common layout (C2: dense slab of full windows) is one copy per slice.
  for (int c = 0; c < nslice; c++) {
    int64_t k0 = cut[c], k1 = cut[c + 1], m = k1 - k0;
    if (m <= 0) continue;
    // the small result rows travel on their own stream, so that they do not queue behind the previous
    // slice's payload copy: the host learns what to fetch for slice c while slice c-1 is still leaving
    CU(cudaStreamWaitEvent(e->s_meta, e->ev_k[c], 0));
    CU(cudaMemcpyAsync(consumed + k0, d_cons + k0, (size_t)m * 8, cudaMemcpyDeviceToHost, e->s_meta));
    CU(cudaMemcpyAsync(produced + k0, d_prod + k0, (size_t)m * 8, cudaMemcpyDeviceToHost, e->s_meta));
    CU(cudaMemcpyAsync(status + k0, d_stat + k0, (size_t)m * 4, cudaMemcpyDeviceToHost, e->s_meta));
    CU(cudaEventRecord(e->ev_meta[c], e->s_meta));
    CU(cudaEventSynchronize(e->ev_meta[c]));
    CU(cudaStreamWaitEvent(s_out, e->ev_k[c], 0));
    int64_t run_lo = -1, run_hi = -1;
    for (int64_t k = k0; k <= k1; k++) {
      int64_t lo = 0, hi = 0;
      if (k < k1) {
        int64_t p = std::min(std::max<int64_t>(produced[k], 0), dst_cap[k]);
        lo = dst_off[k];
        hi = lo + p;
        if (hi == lo) continue;
      }
      if (k < k1 && run_hi == lo) { run_hi = hi; continue; }
      if (run_hi > run_lo)
        CU(cudaMemcpyAsync(dst_base + run_lo, d_dst + (run_lo - d_lo), (size_t)(run_hi - run_lo), cudaMemcpyDeviceToHost, s_out));
      run_lo = lo; run_hi = hi;
    }
  }
  CU(cudaStreamSynchronize(s_out));
  CU(cudaStreamSynchronize(s_k));
  CU(cudaStreamSynchronize(s_in));
  return DCMP_API_OK;
}

int dcmp_inflate_batch(dcmp_engine *e, int fmt, int64_t n, const uint8_t *src_base, const int64_t *src_off,
                       const int64_t *src_len, uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                       int64_t *consumed, int64_t *produced, int32_t *status) {
  if ((fmt & ~DCMP_GZIP_RFC_XLEN) != DCMP_FMT_RAW && (fmt & ~DCMP_GZIP_RFC_XLEN) != DCMP_FMT_ZLIB && (fmt & ~DCMP_GZIP_RFC_XLEN) != DCMP_FMT_GZIP)
    return DCMP_API_INVALID_ARGUMENT;
  return host_batch(e, dcmp_inflate_batch_device, fmt, n, src_base, src_off, src_len, dst_base, dst_off, dst_cap,
                    consumed, produced, status);
}

int dcmp_lzo_uncompress_batch_device(dcmp_engine *e, int mode, int64_t n, const uint8_t *d_src_base,
                                     const int64_t *d_src_off, const int64_t *d_src_len, uint8_t *d_dst_base,
                                     const int64_t *d_dst_off, const int64_t *d_dst_cap, int64_t *d_consumed,
                                     int64_t *d_produced, int32_t *d_status, void *cuda_stream) {
  if (!e || n < 0 || (mode != 0 && mode != 1)) return DCMP_API_INVALID_ARGUMENT;
  if (n == 0) return DCMP_API_OK;
  CU(cudaSetDevice(e->device));
  { const int rc0 = e->chain_begin((cudaStream_t)cuda_stream); if (rc0 != DCMP_API_OK) return rc0; }
  if (e->mode == 0) {
    // stage 1: one CTA per block (serial opcode parse, CTA-wide copies); stage 2: the one-warp-per-block
    // kernel on exactly the blocks stage 1 gave up (malformed input, short output window, ...)
    int rc = dcmp::launch_lzo_parallel(mode, n, e->sm_count > 0 ? e->sm_count : 148, d_src_base, d_src_off, d_src_len, d_dst_base, d_dst_off, d_dst_cap, d_consumed,
                                       d_produced, d_status, (cudaStream_t)cuda_stream);
    if (rc != 0) { set_err("launch_lzo_parallel", (cudaError_t)rc); return DCMP_API_CUDA_ERROR; }
    static const bool no_fallback = getenv("DCMP_LZO_NO_FALLBACK") != nullptr; // debugging aid: given-up blocks keep status 0x7f
    if (!no_fallback)
      dcmp::launch_lzo_decode(mode, n, d_src_base, d_src_off, d_src_len, d_dst_base, d_dst_off, d_dst_cap, d_consumed,
                              d_produced, d_status, dcmp::INTERNAL_RETRY, (unsigned long long *)(e->d_scratch + 72),
                              (cudaStream_t)cuda_stream);
    e->launches += 2;
    e->n_lzo += n;
  } else {
    dcmp::launch_lzo_decode(mode, n, d_src_base, d_src_off, d_src_len, d_dst_base, d_dst_off, d_dst_cap, d_consumed,
                            d_produced, d_status, -1, nullptr, (cudaStream_t)cuda_stream);
    e->launches += 1;
  }
  CU(cudaGetLastError());
  return e->chain_end((cudaStream_t)cuda_stream);
}

int dcmp_lzo_uncompress_batch(dcmp_engine *e, int mode, int64_t n, const uint8_t *src_base, const int64_t *src_off,
                              const int64_t *src_len, uint8_t *dst_base, const int64_t *dst_off,
                              const int64_t *dst_cap, int64_t *consumed, int64_t *produced, int32_t *status) {
  if (mode != 0 && mode != 1) return DCMP_API_INVALID_ARGUMENT;
  return host_batch(e, dcmp_lzo_uncompress_batch_device, mode, n, src_base, src_off, src_len, dst_base, dst_off,
                    dst_cap, consumed, produced, status);
}

int dcmp_lzo_uncompress(dcmp_engine *e, const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len,
                        int64_t *produced) {
  int64_t zero = 0, consumed = 0;
  int32_t status = 0;
  int rc = dcmp_lzo_uncompress_batch(e, 0, 1, src, &zero, &src_len, dst, &zero, &dst_len, &consumed, produced, &status);
  return rc != DCMP_API_OK ? rc : status;
}

int dcmp_inf_ns_inflate(dcmp_engine *e, const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len,
                        int64_t *consumed, int64_t *produced) {
  int64_t zero = 0;
  int32_t status = 0;
  int rc = dcmp_inflate_batch(e, DCMP_FMT_RAW, 1, src, &zero, &src_len, dst, &zero, &dst_len, consumed, produced, &status);
  return rc != DCMP_API_OK ? rc : status;
}

int dcmp_zl_inf_ns_inflate(dcmp_engine *e, const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len,
                           int64_t *consumed, int64_t *produced) {
  int64_t zero = 0;
  int32_t status = 0;
  int rc = dcmp_inflate_batch(e, DCMP_FMT_ZLIB, 1, src, &zero, &src_len, dst, &zero, &dst_len, consumed, produced, &status);
  return rc != DCMP_API_OK ? rc : status;
}

// ---- deflate ------------------------------------------------------------------------------------
int64_t dcmp_deflate_bound(int64_t len, int fmt) {
  // De.Def.Ns.compress_bound (lib/de.ml:3993-3996), + 6 for zlib (lib/zl.ml:598)
  if (len < 0) return -1;
  int64_t blocks = std::max<int64_t>(1, (len + 9999) / 10000);
  return 5 * blocks + len + 1 + 8 + (fmt == DCMP_FMT_ZLIB ? 6 : fmt == DCMP_FMT_GZIP ? 18 : 0);
}

int dcmp_deflate_batch_device(dcmp_engine *e, int level, int fmt, int64_t n, const uint8_t *d_src_base,
                              const int64_t *h_src_off, const int64_t *h_src_len, uint8_t *d_dst_base,
                              const int64_t *h_dst_off, const int64_t *h_dst_cap, int64_t *d_written,
                              int32_t *d_status, void *cuda_stream) {
  if (!e || n < 0) return DCMP_API_INVALID_ARGUMENT;
  // DCMP_DEFLATE_NO_FINAL: every block leaves with BFINAL = 0 and the output ends on a byte boundary (the streaming
  // encoders, csrc/stream.cu, emit a stream segment by segment and close it themselves)
  const bool no_final = (fmt & DCMP_DEFLATE_NO_FINAL) != 0;
  fmt &= ~DCMP_DEFLATE_NO_FINAL;
  if (fmt != DCMP_FMT_RAW && fmt != DCMP_FMT_ZLIB && fmt != DCMP_FMT_GZIP) return DCMP_API_INVALID_ARGUMENT;
  if (no_final && fmt != DCMP_FMT_RAW) return DCMP_API_INVALID_ARGUMENT;
  if (n == 0) return DCMP_API_OK;
  if (!h_src_off || !h_src_len || !h_dst_off || !h_dst_cap || !d_written || !d_status) return DCMP_API_INVALID_ARGUMENT;
  // gzip: the body is written 10 bytes into the window, the header and the CRC-32 / ISIZE trailer are added around
  // it on the device (gzip_frame.cu)
  const bool gz = fmt == DCMP_FMT_GZIP;
  std::vector<int64_t> gz_off, gz_cap;
  const int64_t *body_off = h_dst_off, *body_cap = h_dst_cap;
  if (gz) {
    gz_off.resize((size_t)n); gz_cap.resize((size_t)n);
    for (int64_t k = 0; k < n; k++) { gz_off[k] = h_dst_off[k] + 10; gz_cap[k] = std::max<int64_t>(0, h_dst_cap[k] - 18); }
    body_off = gz_off.data(); body_cap = gz_cap.data();
    fmt = DCMP_FMT_RAW;
  }
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (level < 0 || level > 12) {
    // De.Def.Ns.encoder (lib/de.ml:3930): Invalid_compression_level for every stream
    std::vector<int32_t> bad((size_t)n, DCMP_INVALID_COMPRESSION_LEVEL);
    CU(cudaMemsetAsync(d_written, 0, (size_t)n * 8, st));
    CU(cudaMemcpyAsync(d_status, bad.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return DCMP_API_OK;
  }
  // chunk table: every stream is cut into independent chunks (an empty stream still has one)
  const int64_t CH = dcmp::deflate_chunk_size();
  std::vector<dcmp::DfChunkHost> chunks;
  std::vector<int64_t> first((size_t)n + 1);
  for (int64_t k = 0; k < n; k++) {
    if (h_src_len[k] < 0 || h_dst_cap[k] < 0) return DCMP_API_INVALID_ARGUMENT;
    first[k] = (int64_t)chunks.size();
    int64_t pos = 0;
    do {
      int64_t m = std::min<int64_t>(CH, h_src_len[k] - pos);
      dcmp::DfChunkHost c{h_src_off[k] + pos, (int32_t)m, (int32_t)k, (!no_final && pos + m == h_src_len[k]) ? 1 : 0, (int32_t)std::min<int64_t>(pos, 32768)};
      chunks.push_back(c);
      pos += m;
    } while (pos < h_src_len[k]);
  }
  first[n] = (int64_t)chunks.size();
  const int64_t nc = (int64_t)chunks.size();
  int64_t nsmall = 0; // small single-chunk streams: the warp-per-chunk kernel's share of the batch
  for (const auto &c : chunks) nsmall += (c.len <= dcmp::deflate_small_chunk() && c.hist == 0) ? 1 : 0;
  const int sms = e->sm_count > 0 ? e->sm_count : 148;
  if (!e->d_def_scratch) CU(cudaMalloc(&e->d_def_scratch, dcmp::deflate_scratch_bytes(sms)));
  { const int rc0 = e->chain_begin(st); if (rc0 != DCMP_API_OK) return rc0; }
  const size_t slot = dcmp::deflate_slot_bytes();
  size_t o_chunks = 0;
  size_t o_outs = align_up(o_chunks + (size_t)nc * sizeof(dcmp::DfChunkHost), 256);
  size_t o_cdst = align_up(o_outs + (size_t)nc * 16, 256);
  size_t o_first = align_up(o_cdst + (size_t)nc * 8, 256);
  size_t o_doff = align_up(o_first + (size_t)(n + 1) * 8, 256);
  size_t o_dcap = align_up(o_doff + (size_t)n * 8, 256);
  size_t o_soff = align_up(o_dcap + (size_t)n * 8, 256);   // gzip only: the input table and the caller's windows
  size_t o_slen = align_up(o_soff + (gz ? (size_t)n * 8 : 0), 256);
  size_t o_goff = align_up(o_slen + (gz ? (size_t)n * 8 : 0), 256);
  size_t o_gcap = align_up(o_goff + (gz ? (size_t)n * 8 : 0), 256);
  size_t o_slots = align_up(o_gcap + (gz ? (size_t)n * 8 : 0), 256);
  size_t need = o_slots + (size_t)nc * slot + 256;
  if (need > e->def_arena_cap) {
    if (e->d_def_arena) { e->retired.push_back(e->d_def_arena); e->d_def_arena = nullptr; e->def_arena_cap = 0; }
    CU(cudaMalloc(&e->d_def_arena, need + (need >> 3)));
    e->def_arena_cap = need + (need >> 3);
  }
  uint8_t *A = e->d_def_arena;
  // one pinned image of [chunks .. slots) and one upload: the tables sit in the arena exactly as in the image
  {
    const int rt = e->reserve_tab(o_slots);
    if (rt != DCMP_API_OK) return rt;
    uint8_t *H = e->h_def_tab;
    memcpy(H + o_chunks, chunks.data(), (size_t)nc * sizeof(dcmp::DfChunkHost));
    memcpy(H + o_first, first.data(), (size_t)(n + 1) * 8);
    memcpy(H + o_doff, body_off, (size_t)n * 8);
    memcpy(H + o_dcap, body_cap, (size_t)n * 8);
    if (gz) {
      memcpy(H + o_soff, h_src_off, (size_t)n * 8);
      memcpy(H + o_slen, h_src_len, (size_t)n * 8);
      memcpy(H + o_goff, h_dst_off, (size_t)n * 8);
      memcpy(H + o_gcap, h_dst_cap, (size_t)n * 8);
    }
    // (the outs / chunk_dst regions in between are written by the kernels: they are skipped)
    CU(cudaMemcpyAsync(A + o_chunks, H + o_chunks, o_outs - o_chunks, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(A + o_first, H + o_first, o_slots - o_first, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(e->ev_tab, st));
    e->tab_in_flight = true;
  }
  int rc = dcmp::launch_deflate(level, fmt, n, nc, nsmall, d_src_base, A + o_chunks, (const int64_t *)(A + o_first), A + o_outs,
                                A + o_slots, (int64_t *)(A + o_cdst), e->d_def_scratch,
                                sms, d_dst_base, (const int64_t *)(A + o_doff), (const int64_t *)(A + o_dcap),
                                d_written, d_status, st);
  if (rc != 0) { set_err("launch_deflate", (cudaError_t)rc); return DCMP_API_CUDA_ERROR; }
  e->launches += 3;
  if (gz) {
    dcmp::launch_gzip_wrap(n, level, 3 /* Unix */, d_src_base, (const int64_t *)(A + o_soff), (const int64_t *)(A + o_slen), d_dst_base,
                           (const int64_t *)(A + o_goff), (const int64_t *)(A + o_gcap), d_written, d_status, st);
    e->launches += 1;
  }
  CU(cudaGetLastError());
  { const int rc1 = e->chain_end(st); if (rc1 != DCMP_API_OK) return rc1; }
  return DCMP_API_OK; // (nothing here waits for the stream: the tables were uploaded from the engine's pinned staging)
}

int dcmp_deflate_batch(dcmp_engine *e, int level, int fmt, int64_t n, const uint8_t *src_base, const int64_t *src_off,
                       const int64_t *src_len, uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                       int64_t *written, int32_t *status) {
  if (!e || n < 0) return DCMP_API_INVALID_ARGUMENT;
  if (n == 0) return DCMP_API_OK;
  if (!src_off || !src_len || !dst_off || !dst_cap || !written || !status) return DCMP_API_INVALID_ARGUMENT;
  CU(cudaSetDevice(e->device));
  int64_t s_lo = INT64_MAX, s_hi = 0, d_lo = INT64_MAX, d_hi = 0;
  for (int64_t k = 0; k < n; k++) {
    if (src_len[k] < 0 || dst_cap[k] < 0 || src_off[k] < 0 || dst_off[k] < 0) return DCMP_API_INVALID_ARGUMENT;
    s_lo = std::min(s_lo, src_off[k]); s_hi = std::max(s_hi, src_off[k] + src_len[k]);
    d_lo = std::min(d_lo, dst_off[k]); d_hi = std::max(d_hi, dst_off[k] + dst_cap[k]);
  }
  size_t s_bytes = (size_t)(s_hi - s_lo), d_bytes = (size_t)(d_hi - d_lo);
  if ((s_bytes && !src_base) || (d_bytes && !dst_base)) return DCMP_API_INVALID_ARGUMENT;
  // many small streams: their outputs are packed on the device and come back in one transfer (below)
  const bool pack = n >= 64 && d_bytes / (size_t)n < (256u << 10) && d_bytes <= ((size_t)1 << 30); // (the packed copy is staged in host memory once more)
  size_t o_dst = align_up(s_bytes + 16, 256);
  size_t o_res = align_up(o_dst + d_bytes + 16, 256);
  size_t o_tab = align_up(o_res + (size_t)n * 12 + 16, 256);            // pack: offsets in the window, packed offsets
  size_t o_pack = align_up(o_tab + (pack ? (size_t)n * 16 : 0), 256);
  int rc = e->reserve(o_pack + (pack ? d_bytes : 0) + 256);
  if (rc != DCMP_API_OK) return rc;
  uint8_t *d_src = e->d_buf, *d_dst = e->d_buf + o_dst;
  int64_t *d_written = (int64_t *)(e->d_buf + o_res);
  int32_t *d_status = (int32_t *)(d_written + n);
  std::vector<int64_t> so((size_t)n), dof((size_t)n);
  for (int64_t k = 0; k < n; k++) { so[k] = src_off[k] - s_lo; dof[k] = dst_off[k] - d_lo; }
  cudaStream_t st = e->stream;
  if (s_bytes) CU(cudaMemcpyAsync(d_src, src_base + s_lo, s_bytes, cudaMemcpyHostToDevice, st));
  rc = dcmp_deflate_batch_device(e, level, fmt, n, d_src, so.data(), src_len, d_dst, dof.data(), dst_cap, d_written,
                                 d_status, (void *)st);
  if (rc != DCMP_API_OK) return rc;
  CU(cudaMemcpyAsync(written, d_written, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (pack) {
    std::vector<int64_t> tab((size_t)n * 2);
    int64_t total = 0;
    for (int64_t k = 0; k < n; k++) {
      tab[k] = dof[k];
      tab[n + k] = total;
      if (status[k] == DCMP_OK && written[k] > 0) total += written[k];
    }
    if (total > 0) {
      int64_t *d_tab = (int64_t *)(e->d_buf + o_tab);
      uint8_t *d_pack = e->d_buf + o_pack;
      CU(cudaMemcpyAsync(d_tab, tab.data(), (size_t)n * 16, cudaMemcpyHostToDevice, st));
      dcmp_pack_outputs_kernel<<<(unsigned)n, 128, 0, st>>>(d_dst, d_tab, d_written, d_tab + n, d_pack);
      CU(cudaGetLastError());
      e->launches += 1;
      std::vector<uint8_t> host((size_t)total);
      CU(cudaMemcpyAsync(host.data(), d_pack, (size_t)total, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      for (int64_t k = 0; k < n; k++)
        if (status[k] == DCMP_OK && written[k] > 0) memcpy(dst_base + dst_off[k], host.data() + tab[n + k], (size_t)written[k]);
    }
    return DCMP_API_OK;
  }
  for (int64_t k = 0; k < n; k++)
    if (status[k] == DCMP_OK && written[k] > 0)
      CU(cudaMemcpyAsync(dst_base + dst_off[k], d_dst + dof[k], (size_t)written[k], cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return DCMP_API_OK;
}

// De.Def.encode over a De.Queue (one Dynamic block) and De.T.make, on the device: the command-level half of the encoder,
// held to the reference's encoder KATs (test/test.ml:507-531, 1169-1237)
int dcmp_def_encode_queue(dcmp_engine *e, const int32_t *cmds, int64_t n, int last, uint8_t *dst, int64_t dst_len, int64_t *written) {
  if (!e || n < 0 || dst_len < 0 || !written || (n > 0 && !cmds) || (dst_len > 0 && !dst)) return DCMP_API_INVALID_ARGUMENT;
  CU(cudaSetDevice(e->device));
  const size_t o_tok = align_up((size_t)n * 4 + 16, 256), o_slot = align_up(o_tok + (size_t)n * 4 + 16, 256);
  const size_t slot_bytes = align_up((size_t)n * 8 + 2048, 256), o_dst = o_slot + slot_bytes, o_res = align_up(o_dst + (size_t)dst_len + 16, 256);
  int rc = e->reserve(o_res + 64);
  if (rc != DCMP_API_OK) return rc;
  uint8_t *B = e->d_buf;
  cudaStream_t st = e->stream;
  if (n > 0) CU(cudaMemcpyAsync(B, cmds, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  dcmp::launch_deflate_queue((const int32_t *)B, n, last, (uint32_t *)(B + o_tok), (uint32_t *)(B + o_slot), B + o_dst, dst_len,
                             (int64_t *)(B + o_res), (int32_t *)(B + o_res + 8), st);
  e->launches += 1;
  CU(cudaGetLastError());
  int64_t res[2] = {0, 0};
  CU(cudaMemcpyAsync(res, B + o_res, 12, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const int status = (int)(int32_t)(res[1] & 0xffffffff);
  *written = res[0];
  if (status == DCMP_OK && res[0] > 0) {
    CU(cudaMemcpyAsync(dst, B + o_dst, (size_t)res[0], cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return status;
}

int dcmp_t_make(dcmp_engine *e, int length, int max_length, const int32_t *freqs, int32_t *lengths, int32_t *codes, int32_t *max_code) {
  if (!e || length < 1 || length > 286 || max_length < 1 || max_length > 15 || !freqs || !lengths || !codes || !max_code)
    return DCMP_API_INVALID_ARGUMENT;
  CU(cudaSetDevice(e->device));
  int rc = e->reserve(4096 * 4);
  if (rc != DCMP_API_OK) return rc;
  int32_t *B = (int32_t *)e->d_buf;
  cudaStream_t st = e->stream;
  CU(cudaMemcpyAsync(B, freqs, (size_t)length * 4, cudaMemcpyHostToDevice, st));
  dcmp::launch_deflate_tmake(B, length, max_length, B + 512, B + 1024, B + 1536, st);
  e->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(lengths, B + 512, (size_t)length * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(codes, B + 1024, (size_t)length * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(max_code, B + 1536, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return DCMP_API_OK;
}

static int def_ns(dcmp_engine *e, int level, int fmt, const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len,
                  int64_t *written) {
  int64_t zero = 0;
  int32_t status = 0;
  int rc = dcmp_deflate_batch(e, level, fmt, 1, src, &zero, &src_len, dst, &zero, &dst_len, written, &status);
  return rc != DCMP_API_OK ? rc : status;
}
int dcmp_def_ns_deflate(dcmp_engine *e, int level, const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len,
                        int64_t *written) {
  return def_ns(e, level, DCMP_FMT_RAW, src, src_len, dst, dst_len, written);
}
int dcmp_zl_def_ns_deflate(dcmp_engine *e, int level, const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len,
                           int64_t *written) {
  return def_ns(e, level, DCMP_FMT_ZLIB, src, src_len, dst, dst_len, written);
}

// A member index without decoding: BGZF — the blocked gzip of htslib / bgzip — stores every member's size in
// an FEXTRA sub-field 'B','C' (BSIZE = member bytes - 1, little-endian as RFC1952 has it), so a multi-member file can be cut
// into units on the host by hopping from header to header; ISIZE (the member's last four bytes) gives the output windows.
// Host framing only.  Returns the number of members (an empty EOF member included), -1 if a member is not BGZF-framed.
int64_t dcmp_gz_bgzf_index(const uint8_t *buf, int64_t len, int64_t *member_off, int64_t *member_len, int64_t *isize, int64_t max_members) {
  if (!buf || len < 0) return -1;
  int64_t pos = 0, n = 0;
  while (pos < len) {
    if (len - pos < 18 || buf[pos] != 0x1f || buf[pos + 1] != 0x8b || buf[pos + 2] != 8 || !(buf[pos + 3] & 4)) return -1;
    const int64_t xlen = (int64_t)buf[pos + 10] | ((int64_t)buf[pos + 11] << 8);
    if (len - pos < 12 + xlen) return -1;
    int64_t bsize = -1;
    for (int64_t i = pos + 12; i + 4 <= pos + 12 + xlen;) {
      const int64_t slen = (int64_t)buf[i + 2] | ((int64_t)buf[i + 3] << 8);
      if (buf[i] == 'B' && buf[i + 1] == 'C' && slen == 2 && i + 6 <= pos + 12 + xlen) bsize = ((int64_t)buf[i + 4] | ((int64_t)buf[i + 5] << 8)) + 1;
      i += 4 + slen;
    }
    if (bsize < 12 + xlen + 8 || pos + bsize > len) return -1;
    if (n < max_members) {
      if (member_off) member_off[n] = pos;
      if (member_len) member_len[n] = bsize;
      if (isize) isize[n] = (int64_t)buf[pos + bsize - 4] | ((int64_t)buf[pos + bsize - 3] << 8) | ((int64_t)buf[pos + bsize - 2] << 16) | ((int64_t)buf[pos + bsize - 1] << 24);
    }
    n++;
    pos += bsize;
  }
  return n;
}

// Page-lock a caller's buffer (an OCaml Bigarray is pageable malloc memory): host<->device copies of a registered
// buffer run at the PCIe rate and asynchronously, those of a pageable one go through the driver's staging buffer.
int dcmp_host_register(void *p, int64_t n) {
  if (!p || n <= 0) return DCMP_API_INVALID_ARGUMENT;
  cudaError_t e = cudaHostRegister(p, (size_t)n, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return DCMP_API_OK; }
  CU(e);
  return DCMP_API_OK;
}
int dcmp_host_unregister(void *p) {
  if (!p) return DCMP_API_INVALID_ARGUMENT;
  cudaError_t e = cudaHostUnregister(p);
  if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return DCMP_API_OK; }
  CU(e);
  return DCMP_API_OK;
}

uint32_t dcmp_adler32(uint32_t adler, const uint8_t *p, int64_t n) {
  uint32_t a = adler & 0xffff, b = (adler >> 16) & 0xffff;
  while (n > 0) {
    int64_t k = n < 5552 ? n : 5552;
    n -= k;
    while (k--) { a += *p++; b += a; }
    a %= 65521; b %= 65521;
  }
  return (b << 16) | a;
}

static uint32_t crc32_fast(uint32_t crc, const uint8_t *p, int64_t n) {
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

uint32_t dcmp_crc32(uint32_t crc, const uint8_t *p, int64_t n) {
  return crc32_fast(crc, p, n);
}
} // extern "C"
