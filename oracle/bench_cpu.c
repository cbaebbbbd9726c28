/* oracle/bench_cpu.c — multi-threaded driver that times the CPU restatement on a batch of
 * streams (one stream per task, all host cores).  Used ONLY by bench.py's cpu_baseline leg
 * and `bench.py --impl reference`.  TEST/BENCH INFRASTRUCTURE, never linked into the product.
 * Method mirrors bench/b.ml:11-24 (wall clock around the whole batch, uncompressed bytes / time). */
#define _GNU_SOURCE
#include "oracle.h"
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <time.h>

typedef struct {
  int fmt; int64_t n;
  const uint8_t *src; const int64_t *soff, *slen;
  uint8_t *dst; const int64_t *doff, *dcap;
  int64_t *cons, *prod; int32_t *status;
  atomic_long next;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  for (;;) {
    long k = atomic_fetch_add(&j->next, 1);
    if (k >= j->n) break;
    if (j->fmt == 2)
      j->status[k] = orc_lzo_uncompress(j->src + j->soff[k], j->slen[k], j->dst + j->doff[k], j->dcap[k],
                                        &j->cons[k], &j->prod[k]);
    else if (j->fmt == 1)
      j->status[k] = orc_zl_inflate_ns(j->src + j->soff[k], j->slen[k], j->dst + j->doff[k], j->dcap[k],
                                       &j->cons[k], &j->prod[k]);
    else
      j->status[k] = orc_inflate_ns(j->src + j->soff[k], j->slen[k], j->dst + j->doff[k], j->dcap[k],
                                    &j->cons[k], &j->prod[k]);
  }
  return 0;
}

/* Returns wall seconds for one pass over the batch with `threads` workers. */
double orc_bench_batch(int fmt, int64_t n, const uint8_t *src, const int64_t *soff, const int64_t *slen,
                       uint8_t *dst, const int64_t *doff, const int64_t *dcap, int64_t *cons, int64_t *prod,
                       int32_t *status, int threads) {
  job_t j = {fmt, n, src, soff, slen, dst, doff, dcap, cons, prod, status, 0};
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  struct timespec a, b;
  int t;
  clock_gettime(CLOCK_MONOTONIC, &a);
  for (t = 0; t < threads; t++) pthread_create(&th[t], 0, worker, &j);
  for (t = 0; t < threads; t++) pthread_join(th[t], 0);
  clock_gettime(CLOCK_MONOTONIC, &b);
  free(th);
  return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
