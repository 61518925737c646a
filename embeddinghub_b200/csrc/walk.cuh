// Warp-per-query HNSW graph walk machinery (device side).
//
// Replaces hnswlib's searchKnn / searchBaseLayerST / searchBaseLayer hot loops
// (called from embeddinghub/embeddingstore/index.cc:36,41) with a design that
// fits the B200 memory system:
//   * one warp owns one query (or one point being inserted);
//   * an adjacency row is one 128 B line (2M = 32 u32, padded with kInvalid);
//   * hnswlib's ef-bounded result heap and its candidate heap are ONE unordered
//     array of (ordered distance, id | expanded flag) held in registers, KPL
//     entries per lane; the heap tops are warp reductions (redux.sync);
//   * the visited set is a per-warp open-addressing table in shared memory;
//   * distances of the unvisited neighbours of a node are evaluated together:
//       - rows up to 1 KB (LPV = 8 lanes per vector): every lane issues its
//         128-bit loads for up to 16 vectors before the first use, so a hop
//         has 8 KB..16 KB in flight per warp straight into registers;
//       - larger rows (LPV = 32): the TMA engine pulls whole rows HBM -> shared
//         memory (cp.async.bulk, one bulk copy per vector, completion counted
//         on an mbarrier) in a ring of groups, and the math on group r overlaps
//         the copies of the following groups.
//     (v1 staged every row through TMA; ncu showed the walk issue-bound with
//      17 % of all issued instructions in the per-lane UBLKCP serialisation
//      loops at d=128, see profiles/r01_walk_v1_summary.md.)
//   * rows are padded to an exact multiple of the per-lane tile, so the inner
//     loops carry no bounds checks.
#pragma once
#include "common.cuh"

namespace ehb {

struct GraphView {
  const float* vecs;         // [n][dpad] fp32, rows 16 B aligned, zero padded
  const uint32_t* links0;    // [n][M0]
  const uint32_t* up_off;    // [n] first upper row of node i, kInvalid if level 0
  const uint32_t* links_up;  // [rows][M]
  const uint64_t* labels;    // [n]
  const uint8_t* deleted;    // [n] tombstones (hnswlib markDelete), nullptr when the index has none
  uint32_t n, dim, dpad, M, M0, entry;
  int32_t max_level;
  int32_t metric;            // 0 = squared L2, 1 = 1 - dot (IP and cosine)
};

struct WalkCfg {
  uint32_t lcap;       // shared-memory key list capacity (0 = none; cold paths only)
  uint32_t hash_size;  // entries of the visited table (multiple of 4), 0 = none
  uint32_t G;          // vectors per TMA staging group (<= 32), LPV = 32 only
  uint32_t NG;         // staging groups (ring depth, <= 8)
  uint32_t staged;     // 1 when the TMA staging ring is allocated
  uint32_t dense;      // 1: launch the low-register form of the walk (more resident warps; rows <= 512 B only)
  uint32_t prefetch;   // 1: pull the speculated next hop's vectors towards L2 (rows <= 1 KB)
  uint32_t dcap;       // capacity of the side queue of admitted-but-deleted candidates (0 = index has no tombstones)
};

// Where a walk writes its [nq][k] results.  Destination 0 is local; in a sharded deployment the others are
// THIS rank's block inside every peer's receive buffer (CUDA-IPC / peer mappings): the walk's epilogue stores
// the results of each query to all of them with coalesced stores over NVLink, and the warp that completes a
// slice of `qs` queries raises that slice's flag on every peer (st.release.sys) — the transfer overlaps the
// rest of the walk, and the merge kernel (exchange.cu) only waits on flags.
constexpr uint32_t kMaxSinks = 16;
struct ResultSink {
  uint64_t* labels[kMaxSinks];
  float* dists[kMaxSinks];     // all null or none
  uint32_t* flags[kMaxSinks];  // [slices] of (parity, this rank) on destination t; unused for t = 0
  uint32_t* slice_count;       // local [slices], zero between steps
  uint32_t n, qs, epoch;       // destinations; queries per slice (0: no flags); value the flags take
};

__host__ __device__ inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// Supported padded row lengths: LPV=8 -> NQ in {1,2,4,8}; LPV=32 -> NQ in {3,4,6,8,12,16}.
__host__ __device__ inline uint32_t pad_dim(uint32_t dim) {
  const uint32_t sizes[10] = {32, 64, 128, 256, 384, 512, 768, 1024, 1536, 2048};
  for (int i = 0; i < 10; ++i)
    if (dim <= sizes[i]) return sizes[i];
  return 0;
}

// Per-warp shared-memory slice; every region offset is a multiple of 128 B.
__host__ __device__ inline uint32_t warp_smem_bytes(const WalkCfg& c, uint32_t dpad) {
  uint32_t b = 0;
  b += align_up(c.lcap * 8u, 128);
  b += align_up(c.hash_size * 4u, 128);
  b += 128;  // cand_id[32]
  b += 128;  // cand_dist[32]
  b += 128;  // mbarriers (<= 8) + spare
  b += align_up(c.dcap * 8u, 128);  // deleted-candidate queue: hi[dcap] | id[dcap]
  b += c.staged ? align_up(c.G * c.NG * dpad * 4u, 128) : 0u;
  return b;
}

struct WarpCtx {
  uint64_t* keys;
  uint32_t* hash;
  uint32_t* cand_id;
  float* cand_dist;
  uint64_t* mbar;
  uint32_t* dq_hi;   // deleted-candidate queue (unordered), see beam_search
  uint32_t* dq_id;
  float* stage;
  uint32_t lcap, hsize, G, NG, dpad, vbytes, dcap, prefetch;
  uint32_t phases;  // one parity bit per staging group
  uint32_t cnt;     // live entries in keys[] (shared-memory list only)
  uint32_t lane;
};

__device__ __forceinline__ void ctx_init(WarpCtx& c, unsigned char* base, const WalkCfg& cfg, uint32_t dpad) {
  c.lane = lane_id();
  c.lcap = cfg.lcap;
  c.G = cfg.G;
  c.NG = cfg.NG;
  c.dpad = dpad;
  c.vbytes = dpad * 4u;
  c.hsize = cfg.hash_size;
  unsigned char* p = base;
  c.keys = (uint64_t*)p;
  p += align_up(cfg.lcap * 8u, 128);
  c.hash = (uint32_t*)p;
  p += align_up(cfg.hash_size * 4u, 128);
  c.cand_id = (uint32_t*)p;
  p += 128;
  c.cand_dist = (float*)p;
  p += 128;
  c.mbar = (uint64_t*)p;
  p += 128;
  c.dcap = cfg.dcap;
  c.prefetch = cfg.prefetch;
  c.dq_hi = (uint32_t*)p;
  c.dq_id = c.dq_hi + cfg.dcap;
  p += align_up(cfg.dcap * 8u, 128);
  c.stage = (float*)p;
  c.phases = 0;
  c.cnt = 0;
  if (cfg.staged) {
    if (c.lane < cfg.NG) mbar_init(&c.mbar[c.lane], 1);
    fence_mbar_init();
  }
  __syncwarp();
}

__device__ __forceinline__ void hash_clear(WarpCtx& c) {
  uint4* h4 = (uint4*)c.hash;
  uint32_t n4 = c.hsize >> 2;
  for (uint32_t i = c.lane; i < n4; i += 32) h4[i] = make_uint4(kInvalid, kInvalid, kInvalid, kInvalid);
  __syncwarp();
}

// Lane-parallel "test and set".  Returns true when id was not in the table
// (and is now, unless the probe budget ran out -> overflow).
__device__ __forceinline__ bool hash_insert(WarpCtx& c, uint32_t id, uint32_t& overflow) {
  uint32_t h = __umulhi(id * 0x9E3779B1u, c.hsize);  // fast range reduction: any table size
#pragma unroll 1
  for (int probe = 0; probe < 8; ++probe) {  // bounded: a crowded table costs re-evaluations, never long probes
    uint32_t old = atomicCAS(&c.hash[h], kInvalid, id);
    if (old == kInvalid) return true;
    if (old == id) return false;
    h = h + 1u == c.hsize ? 0u : h + 1u;
  }
  overflow = 1;
  return true;
}

// ---------------------------------------------------------------------------
// Query / vector registers: lane `sub = lane % LPV` holds float4 chunks
// sub + LPV*t, t < NQ  (dpad == 4 * LPV * NQ exactly).
// ---------------------------------------------------------------------------
template <int LPV, int NQ>
__device__ __forceinline__ void load_query_regs(float4 (&qr)[NQ], const float* __restrict__ src, uint32_t dim,
                                                uint32_t lane) {
  uint32_t sub = lane % LPV;
#pragma unroll
  for (int t = 0; t < NQ; ++t) {
    uint32_t e = (sub + LPV * t) * 4u;
    float4 v;
    v.x = e + 0 < dim ? src[e + 0] : 0.f;
    v.y = e + 1 < dim ? src[e + 1] : 0.f;
    v.z = e + 2 < dim ? src[e + 2] : 0.f;
    v.w = e + 3 < dim ? src[e + 3] : 0.f;
    qr[t] = v;
  }
}
template <int LPV, int NQ>
__device__ __forceinline__ void load_vec_regs(float4 (&r)[NQ], const float* __restrict__ row, uint32_t lane) {
  const float4* r4 = (const float4*)row + (lane % LPV);
#pragma unroll
  for (int t = 0; t < NQ; ++t) r[t] = r4[LPV * t];
}

template <int NQ>
__device__ __forceinline__ float partial_dist(const float4 (&v)[NQ], const float4 (&qr)[NQ], int metric) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (metric == 0) {
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      float dx = qr[t].x - v[t].x, dy = qr[t].y - v[t].y, dz = qr[t].z - v[t].z, dw = qr[t].w - v[t].w;
      a0 = fmaf(dx, dx, a0);
      a1 = fmaf(dy, dy, a1);
      a2 = fmaf(dz, dz, a2);
      a3 = fmaf(dw, dw, a3);
    }
  } else {
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      a0 = fmaf(qr[t].x, v[t].x, a0);
      a1 = fmaf(qr[t].y, v[t].y, a1);
      a2 = fmaf(qr[t].z, v[t].z, a2);
      a3 = fmaf(qr[t].w, v[t].w, a3);
    }
  }
  return (a0 + a1) + (a2 + a3);
}
template <int LPV>
__device__ __forceinline__ float group_reduce(float acc) {
#pragma unroll
  for (int o = LPV / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// ---- LPV = 8: direct 128-bit loads, U steps (4 vectors each) in flight ------
// default U: 64 registers of loads in flight (16 vectors at d <= 128)
__host__ __device__ constexpr int eval_u(int NQ, int UDIV) {
  return ((NQ <= 2 ? 8 : (NQ <= 4 ? 4 : 2)) / UDIV) > 0 ? (NQ <= 2 ? 8 : (NQ <= 4 ? 4 : 2)) / UDIV : 1;
}
template <int NQ, int U = eval_u(NQ, 1)>
__device__ __forceinline__ void eval_direct(WarpCtx& c, const float* __restrict__ vecs, const float4 (&qr)[NQ],
                                            uint32_t m, int metric) {
  const uint32_t sub = c.lane & 7u, grp = c.lane >> 3;
  __syncwarp();
#pragma unroll 1
  for (uint32_t j0 = 0; j0 < m; j0 += 4 * U) {
    float4 v[U][NQ];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (j0 + 4u * u < m) {                          // warp-uniform
        uint32_t j = min(j0 + 4u * u + grp, m - 1u);  // clamped lanes re-read the last row (same lines)
        const float4* p = (const float4*)(vecs + (size_t)c.cand_id[j] * c.dpad) + sub;
#pragma unroll
        for (int t = 0; t < NQ; ++t) v[u][t] = ld_nc_f4(p + 8 * t);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (j0 + 4u * u < m) {
        uint32_t j = j0 + 4u * u + grp;
        float acc = group_reduce<8>(partial_dist<NQ>(v[u], qr, metric));
        if (sub == 0 && j < m) c.cand_dist[j] = metric == 0 ? acc : 1.0f - acc;
      }
    }
  }
  __syncwarp();
}

// ---- LPV = 32: TMA bulk staging ring ----------------------------------------
__device__ __forceinline__ void issue_group(WarpCtx& c, const float* __restrict__ vecs, uint32_t r, uint32_t m) {
  uint32_t buf = r % c.NG;
  uint32_t first = r * c.G;
  uint32_t cnt = min(c.G, m - first);
  if (c.lane == 0) mbar_arrive_expect_tx(&c.mbar[buf], cnt * c.vbytes);
  __syncwarp();
  if (c.lane < cnt) {
    uint32_t id = c.cand_id[first + c.lane];
    bulk_g2s(c.stage + (size_t)(buf * c.G + c.lane) * c.dpad, vecs + (size_t)id * c.dpad, c.vbytes, &c.mbar[buf]);
  }
}
template <int NQ>
__device__ __forceinline__ void eval_staged(WarpCtx& c, const float* __restrict__ vecs, const float4 (&qr)[NQ],
                                            uint32_t m, int metric) {
  const uint32_t rounds = (m + c.G - 1) / c.G;
  const uint32_t pre = min(rounds, c.NG);
  for (uint32_t r = 0; r < pre; ++r) issue_group(c, vecs, r, m);
#pragma unroll 1
  for (uint32_t r = 0; r < rounds; ++r) {
    uint32_t buf = r % c.NG;
    mbar_wait(&c.mbar[buf], (c.phases >> buf) & 1u);
    c.phases ^= (1u << buf);
    uint32_t first = r * c.G;
    uint32_t cnt = min(c.G, m - first);
    // four staged vectors per step: their shuffle reductions are independent and overlap
#pragma unroll 1
    for (uint32_t v0 = 0; v0 < cnt; v0 += 4) {
      float acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t v = min(v0 + (uint32_t)i, cnt - 1u);  // clamped repeats are discarded below
        const float4* s4 = (const float4*)(c.stage + (size_t)(buf * c.G + v) * c.dpad) + c.lane;
        float4 x[NQ];
#pragma unroll
        for (int t = 0; t < NQ; ++t) x[t] = s4[32 * t];
        acc[i] = partial_dist<NQ>(x, qr, metric);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
      }
      if (c.lane < 4 && v0 + c.lane < cnt) {
        float a = c.lane == 0 ? acc[0] : (c.lane == 1 ? acc[1] : (c.lane == 2 ? acc[2] : acc[3]));
        c.cand_dist[first + v0 + c.lane] = metric == 0 ? a : 1.0f - a;
      }
    }
    __syncwarp();
    if (r + c.NG < rounds) {
      fence_proxy_async();
      issue_group(c, vecs, r + c.NG, m);
    }
  }
  __syncwarp();
}

// cand_id[0..m) -> cand_dist[0..m): distances from the register-held query.  UDIV > 1 halves (…) the
// load batches kept in flight per warp: fewer registers, more resident warps (the "dense" walk).
template <int LPV, int NQ, int UDIV = 1>
__device__ __forceinline__ void eval_candidates(WarpCtx& c, const float* __restrict__ vecs, const float4 (&qr)[NQ],
                                                uint32_t m, int metric) {
  if (LPV == 8)
    eval_direct<NQ, eval_u(NQ, UDIV)>(c, vecs, qr, m, metric);
  else
    eval_staged<NQ>(c, vecs, qr, m, metric);
}

// ---------------------------------------------------------------------------
// Unsorted register-resident result set ("ulist"): position p = slot*32 + lane,
// valid iff p < ef.  hnswlib keeps a max-heap (results) and a min-heap
// (candidates); here both are one unordered array and the two heap tops are
// found with warp reductions (redux.sync min / max on the ordered 32-bit
// distance): insert-or-replace-worst and pop-closest-unexpanded cost ~10
// instructions each instead of a ~70-instruction sorted insert.
//   hi[s]: ordered distance; empty valid position = 0xFFFFFFFF; dead (p >= ef) = 0
//   id[s]: node id | expanded flag; empty / dead = kInvalid (flag set -> never popped)
// ---------------------------------------------------------------------------
template <int KPL>
struct UList {
  uint32_t hi[KPL];
  uint32_t id[KPL];
};
template <int KPL>
__device__ __forceinline__ void ul_clear(UList<KPL>& u, uint32_t ef, uint32_t lane) {
#pragma unroll
  for (int s = 0; s < KPL; ++s) {
    u.hi[s] = ((uint32_t)s * 32u + lane) < ef ? 0xFFFFFFFFu : 0u;
    u.id[s] = kInvalid;
  }
}
// worst (largest) ordered distance over the valid positions; meaningful when the set is full
template <int KPL>
__device__ __forceinline__ uint32_t ul_worst(const UList<KPL>& u) {
  uint32_t m = u.hi[0];
#pragma unroll
  for (int s = 1; s < KPL; ++s) m = max(m, u.hi[s]);
  return __reduce_max_sync(0xffffffffu, m);
}
// Insert (hi, id) [warp-uniform]; cnt/worst_hi are maintained by the caller's copies.
// Precondition when cnt == ef: hi < worst_hi.
// (Round 2 tried a per-lane form — every lane reduces over its own KPL slots, then ONE ballot elects the
//  lane, predicated writes select the slot — to cut the ~KPL dependent ballots per insert.  Measured on a
//  B200 it lost everywhere: C5 shape 22.8 ms vs 9.0 ms (KPL = 8, 136 vs 128 registers), C3 shape 20.9 vs
//  19.6 ms, C2 0.299 vs 0.293 ms; profiles/r02_ab_ulist.txt.  The slot-walking form below stays.)
template <int KPL>
__device__ __forceinline__ void ul_insert(UList<KPL>& u, uint32_t hi, uint32_t id, uint32_t ef, uint32_t& cnt,
                                          uint32_t& worst_hi, uint32_t lane) {
  bool done = false;
  if (cnt < ef) {
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      if (!done) {
        uint32_t b = __ballot_sync(0xffffffffu, u.id[s] == kInvalid && u.hi[s] == 0xFFFFFFFFu);
        if (b) {
          if ((int)lane == __ffs(b) - 1) u.hi[s] = hi, u.id[s] = id;
          done = true;
        }
      }
    }
    cnt++;
    if (cnt == ef) worst_hi = ul_worst<KPL>(u);
  } else {
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      if (!done) {
        uint32_t b = __ballot_sync(0xffffffffu, u.hi[s] == worst_hi && u.id[s] != kInvalid);
        if (b) {
          if ((int)lane == __ffs(b) - 1) u.hi[s] = hi, u.id[s] = id;
          done = true;
        }
      }
    }
    worst_hi = ul_worst<KPL>(u);
  }
}
// closest unexpanded entry: returns its id (flag clear) or kInvalid; mark=true sets its expanded flag
template <int KPL>
__device__ __forceinline__ uint32_t ul_min_unexpanded(UList<KPL>& u, bool mark, uint32_t lane) {
  uint32_t m = 0xFFFFFFFFu;
#pragma unroll
  for (int s = 0; s < KPL; ++s)
    if (!(u.id[s] & kExpandedFlag)) m = min(m, u.hi[s]);
  m = __reduce_min_sync(0xffffffffu, m);
  uint32_t node = kInvalid;
  bool done = false;
#pragma unroll
  for (int s = 0; s < KPL; ++s) {
    if (!done) {
      uint32_t b = __ballot_sync(0xffffffffu, !(u.id[s] & kExpandedFlag) && u.hi[s] == m);
      if (b) {
        int l = __ffs(b) - 1;
        node = __shfl_sync(0xffffffffu, u.id[s], l);
        if (mark && (int)lane == l) u.id[s] |= kExpandedFlag;
        done = true;
      }
    }
  }
  return node;
}
// ordered distance of the closest unexpanded entry (0xFFFFFFFF when there is none)
template <int KPL>
__device__ __forceinline__ uint32_t ul_min_unexpanded_hi(const UList<KPL>& u) {
  uint32_t m = 0xFFFFFFFFu;
#pragma unroll
  for (int s = 0; s < KPL; ++s)
    if (!(u.id[s] & kExpandedFlag)) m = min(m, u.hi[s]);
  return __reduce_min_sync(0xffffffffu, m);
}
template <int KPL>
__device__ __forceinline__ bool ul_contains(const UList<KPL>& u, uint32_t id) {
  bool hit = false;
#pragma unroll
  for (int s = 0; s < KPL; ++s) hit |= u.id[s] != kInvalid && (u.id[s] & kIdMask) == id;
  return __any_sync(0xffffffffu, hit);
}
// Destructive extraction in ascending order: returns the next (hi, id) or id == kInvalid when empty.
template <int KPL>
__device__ __forceinline__ uint64_t ul_extract_min(UList<KPL>& u, uint32_t lane) {
  uint32_t m = 0xFFFFFFFFu;
#pragma unroll
  for (int s = 0; s < KPL; ++s)
    if (u.id[s] != kInvalid) m = min(m, u.hi[s]);
  m = __reduce_min_sync(0xffffffffu, m);
  uint64_t out = kMaxKey;
  bool done = false;
#pragma unroll
  for (int s = 0; s < KPL; ++s) {
    if (!done) {
      uint32_t b = __ballot_sync(0xffffffffu, u.id[s] != kInvalid && u.hi[s] == m);
      if (b) {
        int l = __ffs(b) - 1;
        uint32_t id = __shfl_sync(0xffffffffu, u.id[s], l);
        out = ((uint64_t)m << 32) | (id & kIdMask);
        if ((int)lane == l) u.id[s] = kInvalid, u.hi[s] = 0xFFFFFFFFu;
        done = true;
      }
    }
  }
  return out;
}

// ---------------------------------------------------------------------------
// Shared-memory sorted key list (cold paths: brute-force select, row merge).
// Returns the insert position, or kInvalid when rejected (duplicate id or
// beyond the limit).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t list_insert(WarpCtx& c, uint64_t key, uint32_t limit) {
  const uint32_t id = key_id(key);
  uint32_t pos = 0;
  uint32_t dup = 0;
  for (uint32_t base = 0; base < c.cnt; base += 32) {
    uint32_t i = base + c.lane;
    uint64_t k = i < c.cnt ? c.keys[i] : kMaxKey;
    pos += __popc(__ballot_sync(0xffffffffu, k < key));
    dup |= __ballot_sync(0xffffffffu, i < c.cnt && key_id(k) == id);
  }
  if (dup || pos >= limit) return kInvalid;
  const uint32_t newcnt = min(c.cnt + 1u, limit);
  for (int base = (int)((newcnt - 1u) & ~31u); base >= (int)(pos & ~31u); base -= 32) {
    uint32_t i = (uint32_t)base + c.lane;
    bool in = i >= pos && i < newcnt;
    uint64_t v = key;
    if (in && i > pos) v = c.keys[i - 1];
    __syncwarp();
    if (in) c.keys[i] = v;
  }
  __syncwarp();
  c.cnt = newcnt;
  return pos;
}

struct WalkCounters {
  uint32_t hops_upper, hops_base, evals, overflow;
};

// Adjacency row of `node` at `level` -> one id per lane (kInvalid beyond the row).
// Branch-free (clamped index + select) so the warp never splits here.
__device__ __forceinline__ uint32_t load_row(const GraphView& g, uint32_t node, int level, uint32_t lane) {
  const uint32_t* row;
  uint32_t width;
  if (level == 0) {  // warp-uniform
    row = g.links0 + (size_t)node * g.M0;
    width = g.M0;
  } else {
    row = g.links_up + (size_t)(g.up_off[node] + (uint32_t)(level - 1)) * g.M;
    width = g.M;
  }
  uint32_t v = row[min(lane, width - 1u)];
  return lane < width ? v : kInvalid;
}

// hnswlib searchKnn's upper-layer descent: at each level move to the closest
// neighbour until no neighbour improves.
template <int LPV, int NQ, int UDIV = 1>
__device__ __forceinline__ void greedy_descent(WarpCtx& c, const GraphView& g, const float4 (&qr)[NQ], uint32_t& cur,
                                               float& curdist, int from_level, int to_level_excl,
                                               WalkCounters& wc) {
  for (int level = from_level; level > to_level_excl; --level) {
    bool changed = true;
    while (changed) {
      changed = false;
      __syncwarp();
      uint32_t nb = load_row(g, cur, level, c.lane);
      uint32_t mask = __ballot_sync(0xffffffffu, nb != kInvalid);
      uint32_t m = __popc(mask);
      wc.hops_upper++;
      if (!m) break;
      if (nb != kInvalid) c.cand_id[__popc(mask & lanemask_lt())] = nb;
      __syncwarp();
      wc.evals += m;
      eval_candidates<LPV, NQ, UDIV>(c, g.vecs, qr, m, g.metric);
      float bd = c.lane < m ? c.cand_dist[c.lane] : INFINITY;
      uint32_t bl = c.lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, bd, o);
        uint32_t ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (od < bd || (od == bd && ol < bl)) bd = od, bl = ol;
      }
      if (bd < curdist) {
        curdist = bd;
        cur = c.cand_id[bl];
        changed = true;
      }
      __syncwarp();
    }
  }
}

// hnswlib searchBaseLayer(ST): best-first beam search with an ef-bounded
// result set.  On return `u` holds the (<= ef) closest visited nodes (unordered).
//
// Why one array replaces hnswlib's two heaps (same expansion order, same result):
//   hnswlib keeps `top_candidates` (max-heap, <= ef results) and `candidate_set` (min-heap of everything
//   ever admitted).  A node enters both at the same moment (when top is not full or it beats the worst
//   result) and is only ever removed from top by eviction of the worst.  The loop pops the closest
//   unexpanded admitted node c and stops when dist(c) > worst result and top is full.  An admitted node
//   that is no longer in top was evicted, i.e. is not closer than the current worst; every node still in
//   top is.  Hence "closest unexpanded admitted node" is in top whenever top holds any unexpanded node, and
//   when it does not, the pop would hit the stop condition (or the queue is empty).  So the walk is:
//   repeatedly expand the closest entry of the result set whose expanded flag is clear, until none is
//   left — which needs only the result set itself plus one flag per entry.  Admission is the same test
//   (`cnt < ef || d < worst`); the tests check id-for-id equality with the oracle on identical graphs.
// `exclude` (kInvalid = none) is never admitted (used when re-linking an updated
// node).  The adjacency row of the likely next node (the closest unexpanded entry
// before this hop's candidates are known) is requested ahead of the distance
// evaluation, so its latency overlaps the vector loads.
// Side queue of admitted-but-deleted candidates (only when g.deleted != nullptr).  hnswlib's
// searchBaseLayerST<has_deletions=true> puts a tombstoned node into candidate_set (it is traversed) but never
// into top_candidates (it is not a result and does not move lowerBound).  Such nodes cannot live in the
// result set, so they wait here, unordered, until they are the closest unexpanded candidate.
__device__ __forceinline__ void dq_push(WarpCtx& c, uint32_t& dn, uint32_t hi, uint32_t id, uint32_t& overflow) {
  if (dn < c.dcap) {
    if (c.lane == 0) c.dq_hi[dn] = hi, c.dq_id[dn] = id;
    dn++;
  } else {  // full: keep the closest dcap entries (the farthest is the least likely to be expanded)
    uint32_t w = 0, wp = 0;
    for (uint32_t i = c.lane; i < dn; i += 32)
      if (c.dq_hi[i] >= w) w = c.dq_hi[i], wp = i;
    uint32_t wm = __reduce_max_sync(0xffffffffu, w);
    uint32_t b = __ballot_sync(0xffffffffu, w == wm);
    uint32_t pos = __shfl_sync(0xffffffffu, wp, __ffs(b) - 1);
    if (hi < wm && c.lane == 0) c.dq_hi[pos] = hi, c.dq_id[pos] = id;
    overflow = 1;
  }
  __syncwarp();
}
// closest queued entry: returns its position (kInvalid when empty) and ordered distance
__device__ __forceinline__ uint32_t dq_min(const WarpCtx& c, uint32_t dn, uint32_t& hi_out) {
  uint32_t m = 0xFFFFFFFFu, mp = kInvalid;
  for (uint32_t i = c.lane; i < dn; i += 32)
    if (c.dq_hi[i] < m) m = c.dq_hi[i], mp = i;
  uint32_t mm = __reduce_min_sync(0xffffffffu, m);
  uint32_t b = __ballot_sync(0xffffffffu, m == mm && mp != kInvalid);
  hi_out = mm;
  return b ? __shfl_sync(0xffffffffu, mp, __ffs(b) - 1) : kInvalid;
}

// HASDEL = false compiles every trace of the tombstone machinery out (an index without tombstones runs
// exactly the round-1 loop: the extra live registers cost the 16-vector load batches their overlap).
template <int LPV, int NQ, int KPL, bool PREFETCH, bool HASDEL, int UDIV = 1>
__device__ __forceinline__ void beam_search(WarpCtx& c, const GraphView& g, const float4 (&qr)[NQ], UList<KPL>& u,
                                            uint32_t ep, float epdist, int level, uint32_t ef, uint32_t exclude,
                                            WalkCounters& wc) {
  hash_clear(c);
  ul_clear<KPL>(u, ef, c.lane);
  const uint8_t* __restrict__ del = (HASDEL && c.dcap) ? g.deleted : nullptr;  // warp-uniform
  uint32_t dn = 0;                                                  // entries in the deleted-candidate queue
  uint32_t ovf = 0;
  if (c.lane == 0) {
    hash_insert(c, ep, ovf);
    if (exclude != kInvalid) hash_insert(c, exclude, ovf);
  }
  __syncwarp();
  uint32_t cnt = 0;
  uint32_t worst_hi = 0xFFFFFFFFu;  // ordered distance of the worst entry once the set is full
  bool ovf_any = false;
  // a tombstoned (or excluded) entry point is expanded once but never becomes a result
  const bool ep_result = ep != exclude && !(del && del[ep]);
  if (ep_result) ul_insert<KPL>(u, f2ord(epdist), ep, ef, cnt, worst_hi, c.lane);
  uint32_t node = ep;
  if (ep_result) node = ul_min_unexpanded<KPL>(u, true, c.lane);
  uint32_t nb = load_row(g, node, level, c.lane);
  for (;;) {
    if (level == 0) wc.hops_base++; else wc.hops_upper++;
    // a walk expands each admitted node once; the bound only guards against a corrupt graph
    if (HASDEL && wc.hops_base + wc.hops_upper > 64u * ef + 65536u) break;
    __syncwarp();
    // speculative: the row of the closest entry still unexpanded
    const uint32_t spec = ul_min_unexpanded<KPL>(u, false, c.lane);
    uint32_t spec_row = kInvalid;
    if (spec != kInvalid) spec_row = load_row(g, spec & kIdMask, level, c.lane);
    bool is_new = false;
    uint32_t o = 0;  // this insert ran out of probes: "new" is then only a guess
    if (nb != kInvalid && nb != exclude) is_new = hash_insert(c, nb, o);
    ovf |= o;
    __syncwarp();  // hash probing diverges; reconverge before the collective section
    uint32_t mask = __ballot_sync(0xffffffffu, is_new);
    uint32_t m = __popc(mask);
    if (m) {
      const uint32_t pos = __popc(mask & lanemask_lt());
      if (is_new) c.cand_id[pos] = nb;
      // tombstones only: candidates (compacted positions) whose visited status is a guess
      const uint32_t unsure = (HASDEL && del) ? __reduce_or_sync(0xffffffffu, (is_new && o) ? (1u << pos) : 0u) : 0u;
      __syncwarp();
      wc.evals += m;
      eval_candidates<LPV, NQ, UDIV>(c, g.vecs, qr, m, g.metric);
      // the speculative row has arrived by now: pull its neighbours' vectors towards L2 while this hop's
      // candidates are inserted (rows <= 1 KB only; a wrong guess costs bandwidth, not correctness)
      if (PREFETCH && LPV == 8 && c.prefetch && spec_row != kInvalid) {
        const char* pv = (const char*)(g.vecs + (size_t)spec_row * g.dpad);
#pragma unroll
        for (int b = 0; b < NQ * 8 * 16; b += 128) prefetch_l2(pv + b);
      }
      uint32_t myhi = 0xFFFFFFFFu, myid = kInvalid;
      if (c.lane < m) myhi = f2ord(c.cand_dist[c.lane]), myid = c.cand_id[c.lane];
      const bool mydel = HASDEL && del && c.lane < m && del[myid];
      __syncwarp();
      ovf_any = ovf_any || __any_sync(0xffffffffu, ovf);
      uint32_t qual = __ballot_sync(0xffffffffu, c.lane < m && (cnt < ef || myhi < worst_hi));
      const uint32_t delmask = (HASDEL && del) ? __ballot_sync(0xffffffffu, mydel) : 0u;
      while (qual) {
        int j = __ffs(qual) - 1;
        qual &= qual - 1;
        uint32_t hj = __shfl_sync(0xffffffffu, myhi, j);
        uint32_t ij = __shfl_sync(0xffffffffu, myid, j);
        if (cnt >= ef && hj >= worst_hi) continue;
        if (HASDEL && ((delmask >> j) & 1u)) {  // admitted like any candidate, but queued instead of becoming a result
          // (a tombstone whose visited status is only a guess is dropped: nothing else would keep it from
          //  being queued and expanded again and again once the visited table is full)
          if (!((unsure >> j) & 1u)) dq_push(c, dn, hj, ij, ovf);
          continue;
        }
        if (ovf_any && ul_contains<KPL>(u, ij)) continue;
        ul_insert<KPL>(u, hj, ij, ef, cnt, worst_hi, c.lane);
        if (PREFETCH && c.lane == 0) prefetch_l2(g.links0 + (size_t)ij * g.M0);
      }
    }
    if (HASDEL && dn) {
      // candidate_set.top(): the closer of (closest unexpanded result, closest queued tombstone)
      uint32_t dhi;
      const uint32_t dpos = dq_min(c, dn, dhi);
      if (dhi < ul_min_unexpanded_hi<KPL>(u)) {
        if (cnt >= ef && dhi > worst_hi) break;  // hnswlib: dist > lowerBound && top_candidates.size() == ef
        node = c.dq_id[dpos];
        __syncwarp();
        if (c.lane == 0) c.dq_hi[dpos] = c.dq_hi[dn - 1], c.dq_id[dpos] = c.dq_id[dn - 1];
        dn--;
        __syncwarp();
        nb = load_row(g, node, level, c.lane);
        continue;
      }
    }
    node = ul_min_unexpanded<KPL>(u, true, c.lane);
    if (node == kInvalid) break;
    nb = (node == spec) ? spec_row : load_row(g, node, level, c.lane);
  }
  wc.overflow |= ovf_any ? 1u : 0u;
}

}  // namespace ehb
