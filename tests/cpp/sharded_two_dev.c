/* Drives a range-sharded index with n_dev = 2 through include/ehb200.h ("device_ids[],
 * n_dev" on create): add routed by label range, concurrent shard builds, search on both shards with their
 * kernels storing into device 0's gather buffer over NVLink, merge on device 0.  Checks the exact path
 * against a host brute force and the graph path's recall.  On a one-GPU box both shards share device 0. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ehb200.h"

#define N 40000
#define D 64
#define NQ 100
#define K 10

#define CHECK(x)                                               \
  do {                                                         \
    int rc_ = (x);                                             \
    if (rc_ != EHB_OK) {                                       \
      printf("FAILED %s: %d %s\n", #x, rc_, ehb_last_error()); \
      return 1;                                                \
    }                                                          \
  } while (0)

static uint64_t st = 0x9E3779B97F4A7C15ull;
static float frand(void) {
  st ^= st << 13;
  st ^= st >> 7;
  st ^= st << 17;
  return (float)(st >> 40) / (float)(1 << 24) - 0.5f;
}

int main(void) {
  int32_t ndev = 0;
  CHECK(ehb_device_count(&ndev));
  int32_t devs[2] = {0, ndev > 1 ? 1 : 0};
  float* base = (float*)malloc((size_t)N * D * 4);
  float* q = (float*)malloc((size_t)NQ * D * 4);
  for (size_t i = 0; i < (size_t)N * D; ++i) base[i] = frand();
  for (size_t i = 0; i < (size_t)NQ * D; ++i) q[i] = frand();
  ehb_params p;
  ehb_params_default(&p, D);
  p.capacity = 1024;
  ehb_sharded* sh = NULL;
  CHECK(ehb_sharded_create(&p, devs, 2, N / 2, &sh));
  CHECK(ehb_sharded_add(sh, N, base, NULL)); /* labels 0..N-1: [0, N/2) -> shard 0, [N/2, N) -> shard 1 */
  CHECK(ehb_sharded_build(sh));
  uint64_t size = 0;
  CHECK(ehb_sharded_size(sh, &size));
  uint32_t shards = 0;
  CHECK(ehb_sharded_n_shards(sh, &shards));
  uint64_t lab[NQ * K], gl[NQ * K];
  float dist[NQ * K];
  uint32_t cnt[NQ];
  CHECK(ehb_sharded_search_bruteforce(sh, NQ, q, K, EHB_FP32, lab, dist, cnt));
  /* host reference: squared L2, one fma chain, order (distance, label) */
  int bad = 0;
  for (int i = 0; i < NQ; ++i) {
    float bd[K];
    uint64_t bl[K];
    int m = 0;
    for (uint64_t n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < D; ++k) {
        float t = q[i * D + k] - base[n * D + k];
        acc = fmaf(t, t, acc);
      }
      int pos = m;
      while (pos > 0 && bd[pos - 1] > acc) --pos;
      if (pos < K) {
        int last = m < K ? m : K - 1;
        for (int j = last; j > pos; --j) bd[j] = bd[j - 1], bl[j] = bl[j - 1];
        bd[pos] = acc, bl[pos] = n;
        if (m < K) ++m;
      }
    }
    for (int j = 0; j < K; ++j)
      if (lab[i * K + j] != bl[j] || dist[i * K + j] != bd[j]) ++bad;
    if (cnt[i] != K) ++bad;
  }
  CHECK(ehb_sharded_search(sh, NQ, q, K, 100, gl, NULL, cnt));
  int hit = 0;
  for (int i = 0; i < NQ; ++i)
    for (int j = 0; j < K; ++j)
      for (int l = 0; l < K; ++l)
        if (gl[i * K + j] == lab[i * K + l]) ++hit;
  double recall = (double)hit / (NQ * K);
  float v[D];
  CHECK(ehb_sharded_get(sh, N - 3, v));
  if (memcmp(v, base + (size_t)(N - 3) * D, sizeof(v)) != 0) ++bad;
  printf("devices {%d,%d} shards %u size %llu exact_mismatches %d graph_recall %.3f\n", devs[0], devs[1], shards,
         (unsigned long long)size, bad, recall);
  CHECK(ehb_sharded_destroy(sh));
  if (bad || size != N || shards != 2 || recall < 0.9) {
    printf("FAILED\n");
    return 1;
  }
  printf("OK\n");
  return 0;
}
