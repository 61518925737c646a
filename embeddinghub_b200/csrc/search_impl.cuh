// K2 kernel template + launcher (included by the per-shape translation units).
#pragma once
#include "kernels.h"

namespace ehb {

// One warp walks one query (device body shared by the two kernels below).
template <int LPV, int NQ, int KPL, bool HASDEL, int UDIV>
__device__ __forceinline__ void search_body(const GraphView& g, const WalkCfg& cfg, const float* __restrict__ queries,
                                            uint32_t nq, uint32_t k, uint32_t ef, const ResultSink& sink,
                                            uint32_t* __restrict__ out_counts, uint32_t* __restrict__ stats,
                                            uint32_t warp_smem) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t w = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + w;
  if (q >= nq) return;
  WarpCtx c;
  ctx_init(c, smem + (size_t)w * warp_smem, cfg, g.dpad);
  float4 qr[NQ];
  load_query_regs<LPV, NQ>(qr, queries + (size_t)q * g.dim, g.dim, c.lane);
  WalkCounters wc = {0, 0, 0, 0};
  UList<KPL> ul;
  ul_clear<KPL>(ul, ef, c.lane);
  if (g.n != 0) {
    uint32_t cur = g.entry;
    if (c.lane == 0) c.cand_id[0] = cur;
    __syncwarp();
    eval_candidates<LPV, NQ, UDIV>(c, g.vecs, qr, 1, g.metric);
    float curdist = c.cand_dist[0];
    __syncwarp();
    wc.evals = 1;
    greedy_descent<LPV, NQ, UDIV>(c, g, qr, cur, curdist, g.max_level, 0, wc);
    beam_search<LPV, NQ, KPL, true, HASDEL, UDIV>(c, g, qr, ul, cur, curdist, 0, ef, kInvalid, wc);
  }
  // nearest-first output: extract the k closest in ascending order into registers (element i -> lane i & 31,
  // slot i >> 5), then store them to every destination of the sink with coalesced stores
  uint64_t rk[KPL];
#pragma unroll
  for (int s = 0; s < KPL; ++s) rk[s] = kMaxKey;
  uint32_t found = 0;
  for (uint32_t i = 0; i < k; ++i) {
    uint64_t key = ul_extract_min<KPL>(ul, c.lane);
    if (key == kMaxKey) break;
#pragma unroll
    for (int s = 0; s < KPL; ++s)
      if ((i >> 5) == (uint32_t)s && (i & 31u) == c.lane) rk[s] = key;
    found++;
  }
#pragma unroll
  for (int s = 0; s < KPL; ++s) {
    const uint32_t idx = (uint32_t)s * 32u + c.lane;
    if ((uint32_t)s * 32u < k && idx < k) {
      const bool ok = rk[s] != kMaxKey;
      const uint64_t lab = ok ? g.labels[key_id(rk[s])] : 0xFFFFFFFFFFFFFFFFull;
      const float dist = ok ? key_dist(rk[s]) : INFINITY;
      const size_t at = (size_t)q * k + idx;
      for (uint32_t t = 0; t < sink.n; ++t) {
        sink.labels[t][at] = lab;
        if (sink.dists[t]) sink.dists[t][at] = dist;
      }
    }
  }
  if (sink.qs) {  // sharded: the warp that completes a slice raises its flag on every peer
    __threadfence_system();
    __syncwarp();
    if (c.lane == 0) {
      const uint32_t slice = q / sink.qs;
      const uint32_t size = min(sink.qs, nq - slice * sink.qs);
      if (atomicAdd(&sink.slice_count[slice], 1u) + 1u == size) {
        sink.slice_count[slice] = 0;  // ready for the next step (which starts after this kernel)
        __threadfence_system();       // the other warps fenced before their atomicAdd: fence-fence ordering
        for (uint32_t t = 1; t < sink.n; ++t) st_release_sys(sink.flags[t] + slice, sink.epoch);
      }
    }
  }
  if (c.lane == 0) {
    if (out_counts) out_counts[q] = found;
    if (stats) ((uint4*)stats)[q] = make_uint4(wc.hops_upper, wc.hops_base, wc.evals, wc.overflow);
  }
}

// Register budget of the default form: ptxas chooses (128 registers for d = 128 / ef = 256, 16 vectors in
// flight per warp).  An explicit minBlocksPerSM changes its heuristics — even "1" gives 143 registers, and
// capping this body at 128 serialised the load batches: 9.1 -> 22.7 ms on the C5 shape — so none is given.
template <int LPV, int NQ, int KPL, bool HASDEL>
__global__ void __launch_bounds__(128) hnsw_search_kernel(GraphView g, WalkCfg cfg, const float* __restrict__ queries,
                                                          uint32_t nq, uint32_t k, uint32_t ef,
                                                          const __grid_constant__ ResultSink sink,
                                                          uint32_t* __restrict__ out_counts,
                                                          uint32_t* __restrict__ stats, uint32_t warp_smem) {
  search_body<LPV, NQ, KPL, HASDEL, 1>(g, cfg, queries, nq, k, ef, sink, out_counts, stats, warp_smem);
}

// "Dense" form for big batches of short rows (LPV = 8, d <= 128): 8 vectors in flight per warp instead of 16
// and a 96-register budget -> 20 resident warps per SM instead of 16 (the visited table shrinks to match,
// api.cu walk_cfg).  Measured on the C5 shape (N = 1M, Q = 10k, ef = 256): 8.66 ms vs 9.08 ms; with the
// same 8-vector batches but 16 warps: 9.48 ms (profiles/r02_ab_walk_occupancy.txt).
template <int LPV, int NQ, int KPL>
__global__ void __launch_bounds__(128, 5) hnsw_search_dense_kernel(GraphView g, WalkCfg cfg,
                                                                   const float* __restrict__ queries, uint32_t nq,
                                                                   uint32_t k, uint32_t ef,
                                                                   const __grid_constant__ ResultSink sink,
                                                                   uint32_t* __restrict__ out_counts,
                                                                   uint32_t* __restrict__ stats, uint32_t warp_smem) {
  search_body<LPV, NQ, KPL, false, 2>(g, cfg, queries, nq, k, ef, sink, out_counts, stats, warp_smem);
}

template <int LPV, int NQ, int KPL, bool HASDEL>
cudaError_t launch_search_t(const GraphView& g, const WalkCfg& cfg, const float* queries, uint32_t nq, uint32_t k,
                            uint32_t ef, const ResultSink& sink, uint32_t* out_counts, uint32_t* stats, uint32_t wpb,
                            cudaStream_t s) {
  uint32_t wsm = warp_smem_bytes(cfg, g.dpad);
  size_t smem = (size_t)wsm * wpb;
  dim3 grid((nq + wpb - 1) / wpb), block(32 * wpb);
  void (*kern)(GraphView, WalkCfg, const float*, uint32_t, uint32_t, uint32_t, const ResultSink, uint32_t*, uint32_t*,
               uint32_t) = hnsw_search_kernel<LPV, NQ, KPL, HASDEL>;
  if constexpr (LPV == 8 && NQ <= 4 && !HASDEL) {
    if (cfg.dense) kern = hnsw_search_dense_kernel<LPV, NQ, KPL>;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, block, smem, s>>>(g, cfg, queries, nq, k, ef, sink, out_counts, stats, wsm);
  return cudaGetLastError();
}

template <int LPV, int NQ>
cudaError_t launch_search_kpl(const GraphView& g, const WalkCfg& cfg, const float* queries, uint32_t nq, uint32_t k,
                              uint32_t ef, const ResultSink& sink, uint32_t* out_counts, uint32_t* stats, uint32_t wpb,
                              cudaStream_t s) {
  // (Keeping a whole 2M-neighbour hop in flight per batch (~168 registers) was measured on C2:
  //  0.446 ms vs 0.423 ms — no gain, so batches stay at 16 vectors.)
#define EHB_KPL(K)                                                                                          \
  return g.deleted ? launch_search_t<LPV, NQ, K, true>(g, cfg, queries, nq, k, ef, sink, out_counts, stats, wpb, s) \
                   : launch_search_t<LPV, NQ, K, false>(g, cfg, queries, nq, k, ef, sink, out_counts, stats, wpb, s)
  if (ef <= 64) EHB_KPL(2);
  if (ef <= 128) EHB_KPL(4);
  if (ef <= 256) EHB_KPL(8);
  EHB_KPL(16);
#undef EHB_KPL
}

#define EHB_SEARCH_ARGS                                                                                  \
  const GraphView &g, const WalkCfg &cfg, const float *queries, uint32_t nq, uint32_t k, uint32_t ef,    \
      const ResultSink &sink, uint32_t *out_counts, uint32_t *stats, uint32_t wpb, cudaStream_t s
#define EHB_SEARCH_PASS g, cfg, queries, nq, k, ef, sink, out_counts, stats, wpb, s

cudaError_t launch_search_d32(EHB_SEARCH_ARGS);
cudaError_t launch_search_d64(EHB_SEARCH_ARGS);
cudaError_t launch_search_d128(EHB_SEARCH_ARGS);
cudaError_t launch_search_d256(EHB_SEARCH_ARGS);
cudaError_t launch_search_d384(EHB_SEARCH_ARGS);
cudaError_t launch_search_d512(EHB_SEARCH_ARGS);
cudaError_t launch_search_d768(EHB_SEARCH_ARGS);
cudaError_t launch_search_d1024(EHB_SEARCH_ARGS);
cudaError_t launch_search_d1536(EHB_SEARCH_ARGS);
cudaError_t launch_search_d2048(EHB_SEARCH_ARGS);

}  // namespace ehb
