/* 64 host threads issuing single-query ehb_index_search calls against one index: the honest stand-in for
 * the cgo provider, whose Nearest() is called goroutine-per-request with one vector
 * (serving/serving.go:744-771, provider/online.go:61-64) — there is no Go toolchain in this image.
 * Checks every answer against one batched search, then measures queries/s
 *   (a) one thread, calls back to back            = the reference's fully serialised service (server.cc:175)
 *   (b) 64 threads, combining queue off           = re-entrant searches on the slot pool only
 *   (c) 64 threads, combining queue on (default)  = concurrent callers share batched launches
 * and prints "speedup <c/a>".  Plain C against include/ehb200.h. */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ehb200.h"

#define N 200000
#define D 128
#define K 10
#define EF 64
#define THREADS 64
#define PER_THREAD 400

static ehb_index* ix;
static float* queries;          /* [THREADS*PER_THREAD][D] */
static uint64_t* want;          /* [THREADS*PER_THREAD][K] */
static uint64_t* got;
static int check_results;
static int failures;

static uint64_t rng_state = 88172645463325252ull;
static float frand(void) { /* xorshift -> roughly N(0,1) by summing */
  float s = 0.f;
  for (int i = 0; i < 4; ++i) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    s += (float)(rng_state >> 40) / (float)(1 << 24);
  }
  return (s - 2.0f) * 1.7320508f;
}
static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
#define CHECK(x)                                                  \
  do {                                                            \
    int rc_ = (x);                                                \
    if (rc_ != EHB_OK) {                                          \
      printf("FAILED %s: %d %s\n", #x, rc_, ehb_last_error());    \
      exit(1);                                                    \
    }                                                             \
  } while (0)

static void* worker(void* arg) {
  long t = (long)arg;
  uint64_t lab[K];
  float dist[K];
  uint32_t cnt;
  for (int j = 0; j < PER_THREAD; ++j) {
    size_t i = (size_t)t * PER_THREAD + j;
    int rc = ehb_index_search(ix, 1, queries + i * D, K, EF, lab, dist, &cnt);
    if (rc != EHB_OK || cnt != K) {
      __sync_fetch_and_add(&failures, 1);
      continue;
    }
    if (check_results) memcpy(got + i * K, lab, sizeof(lab));
  }
  return NULL;
}

static double run_threads(int nthreads) {
  pthread_t th[THREADS];
  double t0 = now();
  for (long t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, (void*)t);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  return (double)nthreads * PER_THREAD / (now() - t0);
}

int main(void) {
  ehb_params p;
  ehb_params_default(&p, D);
  p.capacity = N;
  CHECK(ehb_index_create(&p, &ix));
  float* base = (float*)malloc((size_t)N * D * sizeof(float));
  for (size_t i = 0; i < (size_t)N * D; ++i) base[i] = frand();
  CHECK(ehb_index_add(ix, N, base, NULL));
  CHECK(ehb_index_build(ix));
  free(base);
  const size_t nq = (size_t)THREADS * PER_THREAD;
  queries = (float*)malloc(nq * D * sizeof(float));
  for (size_t i = 0; i < nq * D; ++i) queries[i] = frand();
  want = (uint64_t*)malloc(nq * K * sizeof(uint64_t));
  got = (uint64_t*)calloc(nq * K, sizeof(uint64_t));

  /* correctness: one warp per query everywhere, so batch size cannot change an answer */
  CHECK(ehb_index_set_search_width(ix, 1));
  CHECK(ehb_index_search(ix, nq, queries, K, EF, want, NULL, NULL));
  check_results = 1;
  run_threads(THREADS);
  if (failures || memcmp(want, got, nq * K * sizeof(uint64_t)) != 0) {
    printf("MISMATCH failures=%d\n", failures);
    return 1;
  }
  check_results = 0;
  CHECK(ehb_index_set_search_width(ix, 0));

  CHECK(ehb_index_set_option(ix, "combine", 0));
  run_threads(1); /* warm */
  double serial = run_threads(1);
  double slots = run_threads(THREADS);
  CHECK(ehb_index_set_option(ix, "combine", 1));
  run_threads(THREADS); /* warm */
  double combined = run_threads(THREADS);
  ehb_stats st;
  CHECK(ehb_index_stats(ix, &st));
  printf("serial_1thread_qps %.0f\nslots_%dthreads_qps %.0f\ncombined_%dthreads_qps %.0f\n", serial, THREADS, slots,
         THREADS, combined);
  printf("combined_batches %llu combined_queries %llu (mean batch %.1f)\n", (unsigned long long)st.combined_batches,
         (unsigned long long)st.combined_queries, (double)st.combined_queries / (double)st.combined_batches);
  printf("speedup %.2f\n", combined / serial);
  printf(failures ? "FAILED\n" : "OK\n");
  ehb_index_destroy(ix);
  return failures ? 1 : 0;
}
