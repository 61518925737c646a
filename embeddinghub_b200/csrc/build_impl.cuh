// K5 — batched HNSW construction on the GPU.
//
// Replaces the reference's build path: ANNIndex::set -> hnswlib addPoint
// (embeddinghub/embeddingstore/index.cc:20-37), driven one row at a time by
// Version::create_ann_index (version.cc:64-74).  Same algorithm per point —
// greedy descent through the upper layers, an ef_construction beam search per
// layer, hnswlib's getNeighborsByHeuristic2 neighbour selection, mutual
// connection with re-pruning of full rows — but a whole wave of points is
// linked per launch:
//   phase A (build_search_kernel, one warp per new point): search the already
//     linked graph, select <= M neighbours per layer, write the point's own
//     rows, emit one (target row, source, distance) record per selected edge;
//   phase B (count / alloc / scatter kernels): bucket the edge records by target
//     row with atomics (no global sort);
//   phase C (merge_rows_kernel, one warp per touched row): append the incoming
//     links, or, when the row would overflow, re-select the row with the same
//     heuristic over (existing + incoming) — what mutuallyConnectNewElement
//     does one edge at a time.
// Results are deterministic: candidates are ordered by (distance, id) before
// any selection, so atomic arrival order never matters.
#pragma once
#include "kernels.h"

namespace ehb {

struct Aux {
  uint32_t* sel_id;   // [32]
  float* sel_dist;    // [32]
};

// Aux arrays live in the tail of the 128 B mbarrier region + an extra 256 B.
__device__ __forceinline__ Aux aux_of(const WarpCtx& c) {
  Aux a;
  // cand_id (128 B) | cand_dist (128 B) | mbar (128 B) | stage ...
  // the build kernels reserve 256 extra bytes in front of keys (see build_warp_smem)
  a.sel_id = (uint32_t*)((unsigned char*)c.keys - 256);
  a.sel_dist = (float*)((unsigned char*)c.keys - 128);
  return a;
}
__host__ __device__ inline uint32_t build_warp_smem(const WalkCfg& cfg, uint32_t dpad) {
  return 256u + warp_smem_bytes(cfg, dpad);
}

// hnswlib getNeighborsByHeuristic2 over keys[0..cnt) (ascending distance to
// the point being linked): keep a candidate iff it is not closer to an already
// kept neighbour than to the point.  Returns the number kept (<= Msel).
template <int LPV, int NQ>
__device__ __forceinline__ uint32_t heuristic_select(WarpCtx& c, const GraphView& g, uint32_t Msel, const Aux& a) {
  if (c.cnt < Msel) {
    for (uint32_t i = c.lane; i < c.cnt; i += 32) {
      a.sel_id[i] = key_id(c.keys[i]);
      a.sel_dist[i] = key_dist(c.keys[i]);
    }
    __syncwarp();
    return c.cnt;
  }
  uint32_t nsel = 0;
  for (uint32_t i = 0; i < c.cnt && nsel < Msel; ++i) {
    uint64_t key = c.keys[i];
    uint32_t cid = key_id(key);
    float dq = key_dist(key);
    bool good = true;
    if (nsel > 0) {
      float4 cr[NQ];
      load_vec_regs<LPV, NQ>(cr, g.vecs + (size_t)cid * g.dpad, c.lane);
      if (c.lane < nsel) c.cand_id[c.lane] = a.sel_id[c.lane];
      __syncwarp();
      eval_candidates<LPV, NQ>(c, g.vecs, cr, nsel, g.metric);
      bool bad = c.lane < nsel && c.cand_dist[c.lane] < dq;
      good = !__any_sync(0xffffffffu, bad);
    }
    if (good) {
      if (c.lane == 0) {
        a.sel_id[nsel] = cid;
        a.sel_dist[nsel] = dq;
      }
      nsel++;
    }
    __syncwarp();
  }
  return nsel;
}

template <int LPV, int NQ, int KPL, bool HASDEL>
__global__ void __launch_bounds__(128) build_search_kernel(BuildGraph bg, WalkCfg cfg, const uint32_t* __restrict__ ids,
                                                           uint32_t first, uint32_t b, int is_update, BuildBuffers bb,
                                                           uint32_t warp_smem) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GraphView& g = bg.g;
  const uint32_t w = threadIdx.x >> 5;
  const uint32_t pi = blockIdx.x * (blockDim.x >> 5) + w;
  if (pi >= b) return;
  const uint32_t p = ids ? ids[pi] : first + pi;
  WarpCtx c;
  ctx_init(c, smem + (size_t)w * warp_smem + 256, cfg, g.dpad);
  Aux a = aux_of(c);
  float4 qr[NQ];
  load_vec_regs<LPV, NQ>(qr, g.vecs + (size_t)p * g.dpad, c.lane);
  UList<KPL> ul;
  WalkCounters wc = {0, 0, 0, 0};
  const int level_p = bg.levels[p];
  const int top = g.max_level;
  uint32_t cur = g.entry;
  if (c.lane == 0) c.cand_id[0] = cur;
  __syncwarp();
  eval_candidates<LPV, NQ>(c, g.vecs, qr, 1, g.metric);
  float curdist = c.cand_dist[0];
  __syncwarp();
  if (level_p < top) greedy_descent<LPV, NQ>(c, g, qr, cur, curdist, top, level_p, wc);
  uint32_t* links0 = const_cast<uint32_t*>(g.links0);
  uint32_t* links_up = const_cast<uint32_t*>(g.links_up);
  for (int level = min(level_p, top); level >= 0; --level) {
    beam_search<LPV, NQ, KPL, false, HASDEL>(c, g, qr, ul, cur, curdist, level, bg.efc, is_update ? p : kInvalid, wc);
    // ascending dump into the shared-memory list the selection heuristic walks
    c.cnt = 0;
    for (;;) {
      uint64_t key = ul_extract_min<KPL>(ul, c.lane);
      if (key == kMaxKey) break;
      if (c.lane == 0) c.keys[c.cnt] = key;
      c.cnt++;
    }
    __syncwarp();
    if (c.cnt == 0) continue;
    uint32_t nsel = heuristic_select<LPV, NQ>(c, g, g.M, a);
    uint32_t width = level == 0 ? g.M0 : g.M;
    uint32_t* row = level == 0 ? links0 + (size_t)p * g.M0 : links_up + (size_t)(g.up_off[p] + level - 1) * g.M;
    if (c.lane < width) row[c.lane] = c.lane < nsel ? a.sel_id[c.lane] : kInvalid;
    uint32_t base = 0;
    if (c.lane == 0) base = atomicAdd(bb.edge_count, nsel);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + nsel > bb.edge_cap) {
      if (c.lane == 0) atomicExch(bb.error_flag, 1u);
    } else if (c.lane < nsel) {
      uint32_t t = a.sel_id[c.lane];
      bb.edge_row[base + c.lane] = level == 0 ? t : bg.cap + g.up_off[t] + (uint32_t)(level - 1);
      bb.edge_src[base + c.lane] = p;
      bb.edge_dist[base + c.lane] = a.sel_dist[c.lane];
    }
    cur = key_id(c.keys[0]);
    curdist = key_dist(c.keys[0]);
    __syncwarp();
  }
}

// hnswlib updatePoint, first half (the part before repairConnectionsForUpdate): when the vector of an
// already linked point p changes, every one-hop neighbour nb of p (per layer) gets its adjacency row
// re-selected by the heuristic over the closest ef_construction members of
//   sCand = {p} U one-hop(p) U two-hop(p)   (minus nb itself),
// distances measured from nb.  One warp per updated point; sCand (<= 1 + 2M + 2M*2M ids) is gathered
// into `upd_cand` with the visited table deduplicating (a probe-budget overflow falls back to a linear
// scan, so the set is exact).  Rows are written under a per-row spin lock (bb.row_fill, idle in this
// phase) because two updated points of one wave may share a neighbour.
template <int LPV, int NQ, int KPL>
__global__ void __launch_bounds__(128) update_neighbors_kernel(BuildGraph bg, WalkCfg cfg,
                                                               const uint32_t* __restrict__ ids, uint32_t b,
                                                               BuildBuffers bb, uint32_t warp_smem) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GraphView& g = bg.g;
  const uint32_t w = threadIdx.x >> 5;
  const uint32_t pi = blockIdx.x * (blockDim.x >> 5) + w;
  if (pi >= b) return;
  const uint32_t p = ids[pi];
  WarpCtx c;
  ctx_init(c, smem + (size_t)w * warp_smem + 256, cfg, g.dpad);
  Aux a = aux_of(c);
  uint32_t* cand = bb.upd_cand + (size_t)pi * kUpdCandCap;
  uint32_t* links0 = const_cast<uint32_t*>(g.links0);
  uint32_t* links_up = const_cast<uint32_t*>(g.links_up);
  const int level_p = min((int)bg.levels[p], g.max_level);
  for (int layer = 0; layer <= level_p; ++layer) {
    const uint32_t one = load_row(g, p, layer, c.lane);
    const uint32_t n1 = __popc(__ballot_sync(0xffffffffu, one != kInvalid));
    if (n1 == 0) continue;
    // ---- sCand ----------------------------------------------------------------------------------
    hash_clear(c);
    uint32_t ncand = 0;
    auto add_ids = [&](uint32_t id) {  // one id per lane (kInvalid = none; ids of one call are distinct)
      uint32_t o = 0;
      bool is_new = id != kInvalid && hash_insert(c, id, o);
      __syncwarp();
      if (__any_sync(0xffffffffu, o != 0)) {  // probe budget exhausted: decide by scanning what is stored
        if (o)
          for (uint32_t i = 0; i < ncand && is_new; ++i) is_new = cand[i] != id;
        __syncwarp();
      }
      const uint32_t mask = __ballot_sync(0xffffffffu, is_new);
      if (is_new) cand[ncand + __popc(mask & lanemask_lt())] = id;
      ncand += __popc(mask);
      __syncwarp();
    };
    add_ids(c.lane == 0 ? p : kInvalid);
    add_ids(one);
    for (uint32_t j = 0; j < n1; ++j) {
      const uint32_t e1 = __shfl_sync(0xffffffffu, one, j);
      add_ids(load_row(g, e1, layer, c.lane));
    }
    // ---- re-select the row of every one-hop neighbour ----------------------------------------------
    const uint32_t Mmax = layer == 0 ? g.M0 : g.M;
    for (uint32_t j = 0; j < n1; ++j) {
      const uint32_t nbid = __shfl_sync(0xffffffffu, one, j);
      float4 qr[NQ];
      load_vec_regs<LPV, NQ>(qr, g.vecs + (size_t)nbid * g.dpad, c.lane);
      const uint32_t keep = min(bg.efc, ncand - 1u);  // nb is always a member of sCand
      UList<KPL> u;
      ul_clear<KPL>(u, keep, c.lane);
      uint32_t cnt = 0, worst_hi = 0xFFFFFFFFu;
      for (uint32_t b0 = 0; b0 < ncand; b0 += 32) {
        uint32_t id = b0 + c.lane < ncand ? cand[b0 + c.lane] : kInvalid;
        if (id == nbid) id = kInvalid;
        uint32_t mask = __ballot_sync(0xffffffffu, id != kInvalid);
        uint32_t m = __popc(mask);
        if (!m) continue;
        if (id != kInvalid) c.cand_id[__popc(mask & lanemask_lt())] = id;
        __syncwarp();
        eval_candidates<LPV, NQ>(c, g.vecs, qr, m, g.metric);
        uint32_t myhi = 0xFFFFFFFFu, myid = kInvalid;
        if (c.lane < m) myhi = f2ord(c.cand_dist[c.lane]), myid = c.cand_id[c.lane];
        __syncwarp();
        uint32_t qual = __ballot_sync(0xffffffffu, c.lane < m && (cnt < keep || myhi < worst_hi));
        while (qual) {
          int l = __ffs(qual) - 1;
          qual &= qual - 1;
          uint32_t hj = __shfl_sync(0xffffffffu, myhi, l);
          uint32_t ij = __shfl_sync(0xffffffffu, myid, l);
          if (cnt >= keep && hj >= worst_hi) continue;
          ul_insert<KPL>(u, hj, ij, keep, cnt, worst_hi, c.lane);
        }
      }
      c.cnt = 0;
      for (;;) {
        uint64_t key = ul_extract_min<KPL>(u, c.lane);
        if (key == kMaxKey) break;
        if (c.lane == 0) c.keys[c.cnt] = key;
        c.cnt++;
      }
      __syncwarp();
      const uint32_t nsel = heuristic_select<LPV, NQ>(c, g, Mmax, a);
      const uint32_t rid = layer == 0 ? nbid : bg.cap + g.up_off[nbid] + (uint32_t)(layer - 1);
      uint32_t* row = layer == 0 ? links0 + (size_t)nbid * g.M0
                                 : links_up + (size_t)(g.up_off[nbid] + (uint32_t)(layer - 1)) * g.M;
      if (c.lane == 0)
        while (atomicCAS(&bb.row_fill[rid], 0u, 1u) != 0u) {
        }
      __syncwarp();
      if (c.lane < Mmax) row[c.lane] = c.lane < nsel ? a.sel_id[c.lane] : kInvalid;
      __threadfence();
      __syncwarp();
      if (c.lane == 0) atomicExch(&bb.row_fill[rid], 0u);
    }
  }
}

static __global__ void edge_count_kernel(BuildBuffers bb) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = min(*bb.edge_count, bb.edge_cap);
  if (e >= n) return;
  uint32_t r = bb.edge_row[e];
  if (atomicAdd(&bb.row_cnt[r], 1u) == 0u) bb.touched[atomicAdd(bb.touched_count, 1u)] = r;
}
static __global__ void edge_alloc_kernel(BuildBuffers bb) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *bb.touched_count) return;
  uint32_t r = bb.touched[t];
  bb.row_start[r] = atomicAdd(bb.seg_cursor, bb.row_cnt[r]);
}
static __global__ void edge_scatter_kernel(BuildBuffers bb) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = min(*bb.edge_count, bb.edge_cap);
  if (e >= n) return;
  uint32_t r = bb.edge_row[e];
  uint32_t pos = bb.row_start[r] + atomicAdd(&bb.row_fill[r], 1u);
  bb.seg_src[pos] = bb.edge_src[e];
  bb.seg_dist[pos] = bb.edge_dist[e];
}

// Phase C.  Incoming links are applied in source-id (= insertion) order, one at
// a time, exactly like hnswlib's mutuallyConnectNewElement: append while the row
// has room; when it is full re-select the row with the heuristic over
// (existing + the one new link).  Only after kMaxSeqPrunes such re-selections
// in one wave (hub rows) is the remainder folded into a single re-selection.
constexpr uint32_t kMaxSeqPrunes = 3;

template <int LPV, int NQ>
__global__ void __launch_bounds__(128) merge_rows_kernel(BuildGraph bg, WalkCfg cfg, BuildBuffers bb,
                                                         uint32_t warp_smem) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GraphView& g = bg.g;
  const uint32_t w = threadIdx.x >> 5;
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + w;
  if (t >= *bb.touched_count) return;
  const uint32_t r = bb.touched[t];
  const uint32_t ninc_all = bb.row_cnt[r], start = bb.row_start[r];
  WarpCtx c;
  ctx_init(c, smem + (size_t)w * warp_smem + 256, cfg, g.dpad);
  Aux a = aux_of(c);
  uint32_t node, W;
  uint32_t* row;
  if (r < bg.cap) {
    node = r;
    W = g.M0;
    row = const_cast<uint32_t*>(g.links0) + (size_t)r * g.M0;
  } else {
    uint32_t ur = r - bg.cap;
    node = bg.up_owner[ur];
    W = g.M;
    row = const_cast<uint32_t*>(g.links_up) + (size_t)ur * g.M;
  }
  const uint32_t limit = cfg.lcap;  // 128
  // 1. order the incoming links by source id (deterministic, = insertion order)
  c.cnt = 0;
  for (uint32_t j = 0; j < ninc_all; ++j)
    list_insert(c, ((uint64_t)bb.seg_src[start + j] << 32) | (j & kIdMask), limit);
  const uint32_t ninc = c.cnt;
  uint32_t in_src[4], in_j[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    uint32_t i = (uint32_t)s * 32u + c.lane;
    uint64_t key = i < ninc ? c.keys[i] : kMaxKey;
    in_src[s] = (uint32_t)(key >> 32);
    in_j[s] = (uint32_t)key & kIdMask;
  }
  __syncwarp();
  // 2. the row, one entry per lane (compact: valid ids first)
  uint32_t e = c.lane < W ? row[c.lane] : kInvalid;
  float ed = 0.f;
  uint32_t ne = __popc(__ballot_sync(0xffffffffu, e != kInvalid));
  bool have_dists = false;
  uint32_t prunes = 0;
  float4 qr[NQ];
  for (uint32_t i = 0; i < ninc; ++i) {
    uint32_t sv = in_src[0], jv = in_j[0];
#pragma unroll
    for (int s = 1; s < 4; ++s)
      if ((i >> 5) == (uint32_t)s) sv = in_src[s], jv = in_j[s];
    const uint32_t sid = __shfl_sync(0xffffffffu, sv, i & 31);
    const uint32_t sj = __shfl_sync(0xffffffffu, jv, i & 31);
    if (__any_sync(0xffffffffu, e == sid)) continue;  // already linked (update path)
    const float sd = bb.seg_dist[start + sj];
    if (ne < W) {
      if (c.lane == ne) e = sid, ed = sd;
      ne++;
      continue;
    }
    // row full: distances of the existing entries to the row's owner are needed once
    if (!have_dists) {
      load_vec_regs<LPV, NQ>(qr, g.vecs + (size_t)node * g.dpad, c.lane);
      if (c.lane < ne) c.cand_id[c.lane] = e;
      __syncwarp();
      eval_candidates<LPV, NQ>(c, g.vecs, qr, ne, g.metric);
      if (c.lane < ne) ed = c.cand_dist[c.lane];
      __syncwarp();
      have_dists = true;
    }
    c.cnt = 0;
    uint64_t mykey = c.lane < ne ? make_key(ed, e) : kMaxKey;
    for (uint32_t j = 0; j < ne; ++j) list_insert(c, __shfl_sync(0xffffffffu, mykey, j), limit);
    list_insert(c, make_key(sd, sid), limit);
    bool fold_rest = ++prunes > kMaxSeqPrunes;
    if (fold_rest) {
      for (uint32_t i2 = i + 1; i2 < ninc; ++i2) {
        uint32_t sv2 = in_src[0], jv2 = in_j[0];
#pragma unroll
        for (int s = 1; s < 4; ++s)
          if ((i2 >> 5) == (uint32_t)s) sv2 = in_src[s], jv2 = in_j[s];
        uint32_t sid2 = __shfl_sync(0xffffffffu, sv2, i2 & 31);
        uint32_t sj2 = __shfl_sync(0xffffffffu, jv2, i2 & 31);
        list_insert(c, make_key(bb.seg_dist[start + sj2], sid2), limit);  // list_insert drops duplicate ids
      }
    }
    uint32_t nsel = heuristic_select<LPV, NQ>(c, g, W, a);
    e = c.lane < nsel ? a.sel_id[c.lane] : kInvalid;
    ed = c.lane < nsel ? a.sel_dist[c.lane] : 0.f;
    ne = nsel;
    __syncwarp();
    if (fold_rest) break;
  }
  if (c.lane < W) row[c.lane] = e;
  if (c.lane == 0) {
    bb.row_cnt[r] = 0;
    bb.row_fill[r] = 0;
  }
}

template <int LPV, int NQ>
cudaError_t launch_build_t(const BuildGraph& bg, const WalkCfg& cfg, const uint32_t* ids, uint32_t first, uint32_t b,
                           bool is_update, BuildBuffers& bb, uint32_t wpb, cudaStream_t s) {
  constexpr int KPL = 8;  // ef_construction <= 256
  cudaError_t e;
  if ((e = cudaMemsetAsync(bb.edge_count, 0, 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(bb.touched_count, 0, 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(bb.seg_cursor, 0, 4, s)) != cudaSuccess) return e;
  uint32_t wsm = build_warp_smem(cfg, bg.g.dpad);
  size_t smem = (size_t)wsm * wpb;
  // the merge pass needs no visited table
  WalkCfg mcfg = cfg;
  mcfg.hash_size = 0;
  mcfg.lcap = 128;
  uint32_t mwsm = build_warp_smem(mcfg, bg.g.dpad);
  uint32_t mwpb = 4;
  size_t msmem = (size_t)mwsm * mwpb;
  while (msmem > 200 * 1024 && mwpb > 1) mwpb >>= 1, msmem = (size_t)mwsm * mwpb;
  dim3 grid((b + wpb - 1) / wpb), block(32 * wpb);
  uint32_t ethreads = bb.edge_cap;
  auto ks = bg.g.deleted ? build_search_kernel<LPV, NQ, KPL, true> : build_search_kernel<LPV, NQ, KPL, false>;
  auto km = merge_rows_kernel<LPV, NQ>;
  if (is_update) {
    // updatePoint's neighbour re-selection runs before the moved points are re-linked
    if (!ids || !bb.upd_cand) return cudaErrorInvalidValue;
    WalkCfg ucfg = cfg;
    if (ucfg.hash_size < 4096) ucfg.hash_size = 4096;
    uint32_t uwsm = build_warp_smem(ucfg, bg.g.dpad), uwpb = wpb;
    while (uwpb > 1 && (size_t)uwsm * uwpb > 200 * 1024) uwpb >>= 1;
    auto ku = update_neighbors_kernel<LPV, NQ, KPL>;
    if ((e = cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)uwsm * uwpb))) !=
        cudaSuccess)
      return e;
    ku<<<(b + uwpb - 1) / uwpb, 32 * uwpb, (size_t)uwsm * uwpb, s>>>(bg, ucfg, ids, b, bb, uwsm);
  }
  if ((e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem)) != cudaSuccess) return e;
  ks<<<grid, block, smem, s>>>(bg, cfg, ids, first, b, is_update ? 1 : 0, bb, wsm);
  edge_count_kernel<<<(ethreads + 255) / 256, 256, 0, s>>>(bb);
  edge_alloc_kernel<<<(ethreads + 255) / 256, 256, 0, s>>>(bb);
  edge_scatter_kernel<<<(ethreads + 255) / 256, 256, 0, s>>>(bb);
  km<<<(ethreads + mwpb - 1) / mwpb, 32 * mwpb, msmem, s>>>(bg, mcfg, bb, mwsm);
  return cudaGetLastError();
}

#define EHB_BUILD_ARGS                                                                                     \
  const BuildGraph &bg, const WalkCfg &cfg, const uint32_t *ids, uint32_t first, uint32_t b, bool is_update, \
      BuildBuffers &bb, uint32_t wpb, cudaStream_t s
#define EHB_BUILD_PASS bg, cfg, ids, first, b, is_update, bb, wpb, s
cudaError_t launch_build_d32(EHB_BUILD_ARGS);
cudaError_t launch_build_d64(EHB_BUILD_ARGS);
cudaError_t launch_build_d128(EHB_BUILD_ARGS);
cudaError_t launch_build_d256(EHB_BUILD_ARGS);
cudaError_t launch_build_d384(EHB_BUILD_ARGS);
cudaError_t launch_build_d512(EHB_BUILD_ARGS);
cudaError_t launch_build_d768(EHB_BUILD_ARGS);
cudaError_t launch_build_d1024(EHB_BUILD_ARGS);
cudaError_t launch_build_d1536(EHB_BUILD_ARGS);
cudaError_t launch_build_d2048(EHB_BUILD_ARGS);

}  // namespace ehb
