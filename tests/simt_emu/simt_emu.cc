// tests/simt_emu/simt_emu.cc — fiber scheduler of the SIMT emulator (TEST INFRASTRUCTURE ONLY).
#include "simt_emu.h"

namespace simt {

Block *g_blk = nullptr;
dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
unsigned char *g_dyn_smem = nullptr;
long g_switches = 0;

static const size_t STACK = 256 * 1024;

static void fiber_entry() {
  Block *b = g_blk;
  b->body();
  b->fibers[b->cur].state = DONE;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

static void set_thread(Block &b, int t) {
  b.cur = t;
  g_threadIdx = dim3((unsigned)t % g_blockDim.x, ((unsigned)t / g_blockDim.x) % g_blockDim.y,
                     (unsigned)t / (g_blockDim.x * g_blockDim.y));
}

void run_block(Block &b) {
  const int n = b.nthreads;
  for (;;) {
    bool progressed = false;
    int done = 0;
    // 1. run every runnable fiber until it blocks or exits
    for (int t = 0; t < n; t++) {
      Fiber &f = b.fibers[t];
      if (f.state == DONE) { done++; continue; }
      if (f.state != RUN) continue;
      set_thread(b, t);
      swapcontext(&b.sched, &f.ctx);
      progressed = true;
      if (f.state == DONE) done++;
    }
    if (done == n) return;
    // 2. release warp barriers whose (live) participants have all arrived
    const int nwarps = (n + 31) / 32;
    for (int w = 0; w < nwarps; w++) {
      unsigned waiting = 0, live = 0, want = 0;
      for (int l = 0; l < 32; l++) {
        int t = w * 32 + l;
        if (t >= n) break;
        Fiber &f = b.fibers[t];
        if (f.state == DONE) continue;
        live |= 1u << l;
        if (f.state == WAIT_WARP) { waiting |= 1u << l; want |= f.warp_mask; }
      }
      if (!waiting) continue;
      unsigned need = want & live;
      if ((waiting & need) == need) {
        for (int l = 0; l < 32; l++)
          if ((waiting >> l) & 1) b.fibers[w * 32 + l].state = RUN;
        progressed = true;
      }
    }
    // 3. release the CTA barrier when every live thread waits on it
    int at_cta = 0, live_n = 0;
    for (int t = 0; t < n; t++) {
      if (b.fibers[t].state == DONE) continue;
      live_n++;
      if (b.fibers[t].state == WAIT_CTA) at_cta++;
    }
    if (live_n && at_cta == live_n) {
      for (int t = 0; t < n; t++)
        if (b.fibers[t].state == WAIT_CTA) b.fibers[t].state = RUN;
      progressed = true;
    }
    if (!progressed) {
      fprintf(stderr, "simt: deadlock in block (%u,%u): ", g_blockIdx.x, g_blockIdx.y);
      for (int t = 0; t < n; t++) fprintf(stderr, "%d", b.fibers[t].state);
      fprintf(stderr, "\n");
      abort();
    }
  }
}

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body) {
  Block b;
  b.nthreads = (int)(block.x * block.y * block.z);
  b.fibers.resize(b.nthreads);
  b.slot.assign(b.nthreads, 0);
  b.body = body;
  g_blockDim = block;
  g_gridDim = grid;
  std::vector<unsigned char> smem(dyn_smem + 64);
  g_dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
  for (auto &f : b.fibers) f.stack = (char *)malloc(STACK);
  Block *saved = g_blk;
  g_blk = &b;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        g_blockIdx = dim3(bx, by, bz);
        for (int t = 0; t < b.nthreads; t++) {
          Fiber &f = b.fibers[t];
          f.state = RUN;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = STACK;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, fiber_entry, 0);
        }
        run_block(b);
      }
  for (auto &f : b.fibers) free(f.stack);
  g_blk = saved;
}

} // namespace simt
