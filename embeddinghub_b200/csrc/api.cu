// Host side of the ehb200 C ABI (include/ehb200.h): index state in HBM, batched
// construction driver, search entry points.  Mirrors the responsibilities of
// featureform::embedding::ANNIndex + hnswlib::HierarchicalNSW as used in
// embeddinghub/embeddingstore/index.cc:10-52.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ehb200.h"
#include "kernels.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(expr)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(_e == cudaErrorMemoryAllocation ? EHB_ERR_OOM : EHB_ERR_CUDA,                          \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                                   \
  } while (0)
#define RET(expr)              \
  do {                         \
    int _r = (expr);           \
    if (_r != EHB_OK) return _r; \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  // grow to >= want elements, preserving the first `keep` elements; fill new tail with byte `fill` if fill >= 0
  cudaError_t grow(size_t want, size_t keep, int fill, cudaStream_t s) {
    if (want <= n) return cudaSuccess;
    T* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p) e = cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && fill >= 0) e = cudaMemsetAsync(np + keep, fill, (want - keep) * sizeof(T), s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      cudaFree(np);  // the old buffer stays valid
      return e;
    }
    if (p) cudaFree(p);
    p = np;
    n = want;
    return cudaSuccess;
  }
  size_t bytes() const { return n * sizeof(T); }
};

}  // namespace

struct ehb_index {
  ehb_params prm;
  uint32_t dim, dpad, M, M0;
  int metric;
  int device;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  std::mutex mu;

  uint64_t cap = 0;        // vector capacity
  uint64_t n = 0;          // stored vectors
  uint64_t n_linked = 0;   // vectors linked into the graph
  uint64_t up_rows = 0;    // used upper rows
  uint32_t entry = 0;
  int32_t max_level = -1;
  uint32_t ef;

  DevBuf<float> vecs;
  DevBuf<uint64_t> labels;
  DevBuf<uint8_t> levels;
  DevBuf<uint32_t> links0, up_off, links_up, up_owner;

  std::vector<uint8_t> h_levels;
  std::vector<uint64_t> h_labels;
  bool identity_labels = true;
  std::unordered_map<uint64_t, uint32_t> lookup;
  std::vector<uint32_t> pending_updates;

  // search scratch
  DevBuf<float> q_in, q_norm, o_dists;
  DevBuf<uint64_t> o_labels;
  DevBuf<uint32_t> o_counts, stats;
  DevBuf<unsigned long long> stat_sum;
  uint64_t last_nq = 0;
  unsigned long long last_sum[4] = {0, 0, 0, 0};
  bool last_sum_valid = false;

  // brute-force scratch
  DevBuf<float> bf_dist, bf_qpad;
  DevBuf<uint64_t> bf_part, bf_run;
  DevBuf<uint16_t> x_bf16, q_bf16;   // bf16 shadows for the tensor-core path
  DevBuf<float> x_norm, q_norm2, bf_thr;
  DevBuf<uint64_t> bf_cbuf;
  DevBuf<uint32_t> bf_ccount;
  uint64_t bf16_rows = 0;            // rows of x_bf16 that are current (0 = stale)

  // build scratch
  DevBuf<uint32_t> b_edge_row, b_edge_src, b_row_cnt, b_row_fill, b_row_start, b_touched, b_seg_src, b_counters, b_ids;
  DevBuf<float> b_edge_dist, b_seg_dist, b_stage_in;

  // tuning (0 = auto)
  uint32_t t_slots = 0, t_groups = 0, t_hash_bits = 0, t_wpb = 0, t_team = 0;
  // options (ehb_index_set_option)
  uint32_t o_build_frac = 0;     // a wave links at most n_linked / build_frac points (0 = 64)
  bool o_bf16_unfused = false;   // bf16 brute force: keep the distance tiles in HBM (A/B)
  bool o_gemm_2cta = false;      // bf16 brute force: cta_group::2 cluster form of the fused GEMM

  ~ehb_index() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }

  ehb::GraphView view() const {
    ehb::GraphView g;
    g.vecs = vecs.p;
    g.links0 = links0.p;
    g.up_off = up_off.p;
    g.links_up = links_up.p;
    g.labels = labels.p;
    g.n = (uint32_t)n_linked;
    g.dim = dim;
    g.dpad = dpad;
    g.M = M;
    g.M0 = M0;
    g.entry = entry;
    g.max_level = max_level;
    g.metric = metric == EHB_L2 ? 0 : 1;
    return g;
  }

  // ef_eff: beam width; smem_list: capacity of the shared-memory key list (0 for plain searches);
  // jobs: warps (queries or points) of the launch; team: warps sharing one visited table.
  ehb::WalkCfg walk_cfg(uint32_t ef_eff, uint32_t smem_list, uint64_t jobs, uint32_t team) const {
    ehb::WalkCfg c;
    c.lcap = smem_list;
    c.staged = dpad > 256 ? 1 : 0;  // rows above 1 KB go through the TMA staging ring
    uint32_t vbytes = dpad * 4;
    uint32_t slots = std::max(4u, std::min(32u, 24576u / vbytes));
    uint32_t ng = 2;                      // two groups: math on one overlaps the copies of the other
    uint32_t g = std::max(4u, slots / ng / 4 * 4);  // vectors per group, multiple of the 4-vector math step
    if (t_slots) g = std::min(32u, t_slots);
    if (t_groups) ng = std::min(8u, t_groups);
    c.G = std::max(1u, g);
    c.NG = std::max(1u, ng);
    // Visited table.  A hop admits at most 2M new ids and the walk makes about ef hops.  "roomy" keeps
    // the final load near 0.5 even on iid Gaussian data (~29 new ids per hop); but every KB of table
    // costs occupancy, and a crowded table only costs re-evaluations (probes are bounded; duplicates
    // are filtered against the result set): measured at d=768/ef=128, a table HALF the visited count
    // gave +1.3 % evaluations and 1.5x the throughput of the roomy one.  So: as roomy as the
    // occupancy target allows, never below a quarter of the worst case.
    const uint32_t roomy = 2u * M0 * ef_eff + 64u, tight = std::max(256u, M0 * ef_eff / 4u);
    uint32_t hs = roomy;
    if (t_hash_bits) {
      hs = 1u << t_hash_bits;
    } else {
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      uint64_t ctas = (jobs + team - 1) / std::max(team, 1u);
      uint32_t want = (uint32_t)std::min<uint64_t>((ctas + sms - 1) / sms, c.staged ? 5u : 16u / team);
      want = std::max(want, c.staged ? 4u : 4u);
      c.hash_size = 0;
      uint32_t fixed = ehb::warp_smem_bytes(c, dpad) * team + 1024u + (smem_list ? 256u : 0u);
      uint32_t per_cta = (227u * 1024u) / want;
      uint32_t avail = per_cta > fixed + 1024u ? (per_cta - fixed) / 4u : 256u;
      hs = std::min(roomy, std::max(tight, avail));
    }
    c.hash_size = ehb::align_up(std::max(hs, 256u), 32);
    // stay inside the 227 KB per-block limit
    while (ehb::warp_smem_bytes(c, dpad) + 256 > 200 * 1024 && c.hash_size > 512) c.hash_size = ehb::align_up(c.hash_size / 2, 32);
    return c;
  }
  uint32_t wpb_for(const ehb::WalkCfg& c, uint32_t extra) const {
    uint32_t w = t_wpb ? t_wpb : 1;
    while (w > 1 && (size_t)(ehb::warp_smem_bytes(c, dpad) + extra) * w > 220 * 1024) w >>= 1;
    return w;
  }

  int ensure_capacity(uint64_t want) {
    if (want <= cap) return EHB_OK;
    uint64_t nc = std::max<uint64_t>(cap ? cap : 1, 1);
    while (nc < want) nc *= 2;  // index.cc:29-32 doubles
    if (nc >= 0x7FFFFFFFull) return fail(EHB_ERR_INVALID, "capacity must stay below 2^31 vectors per index");
    CU(vecs.grow(nc * dpad, n * dpad, -1, stream));
    CU(labels.grow(nc, n, -1, stream));
    CU(levels.grow(nc, n, 0, stream));
    CU(links0.grow(nc * M0, n * M0, 0xFF, stream));
    CU(up_off.grow(nc, n, 0xFF, stream));
    cap = nc;
    return EHB_OK;
  }
  int ensure_upper(uint64_t want_rows) {
    if (want_rows <= links_up.n / M && links_up.n) return EHB_OK;
    uint64_t nr = std::max<uint64_t>(links_up.n / M, 64);
    while (nr < want_rows) nr *= 2;
    CU(links_up.grow(nr * M, up_rows * M, 0xFF, stream));
    CU(up_owner.grow(nr, up_rows, 0, stream));
    return EHB_OK;
  }

  // hnswlib getRandomLevel: (int)(-log(U(0,1)) * 1/ln(M)) drawn from
  // std::default_random_engine(seed), one draw per new point in insertion order —
  // the same generator classes upstream uses, so levels match an hnswlib built
  // against the same C++ standard library.
  std::default_random_engine level_rng;
  int draw_level() {
    std::uniform_real_distribution<double> u(0.0, 1.0);
    double r = -std::log(u(level_rng)) * (1.0 / std::log((double)M));
    return std::min((int)r, 15);
  }

  // ---- ingest --------------------------------------------------------------
  int add_rows(uint64_t cnt, const float* src, bool src_is_device, const uint64_t* lab) {
    if (cnt == 0) return EHB_OK;
    const uint64_t chunk = std::max<uint64_t>(1, (256ull << 20) / (dim * 4));
    for (uint64_t off = 0; off < cnt; off += chunk) {
      uint64_t m = std::min(chunk, cnt - off);
      // resolve destination ids
      std::vector<uint32_t> dst(m);
      uint64_t first_new = n;
      bool contiguous_new = true;
      uint64_t nn = n;
      for (uint64_t i = 0; i < m; ++i) {
        uint64_t l = lab ? lab[off + i] : nn;
        bool exists = false;
        uint32_t id = 0;
        if (identity_labels) {
          if (l < nn) exists = true, id = (uint32_t)l;
          else if (l != nn) {
            // leave identity mode: materialise the map
            lookup.reserve(std::max<uint64_t>(nn * 2, 1024));
            for (uint64_t j = 0; j < nn; ++j) lookup[j] = (uint32_t)j;
            identity_labels = false;
          }
        }
        if (!identity_labels && !exists) {
          auto it = lookup.find(l);
          if (it != lookup.end()) exists = true, id = it->second;
        }
        if (exists) {
          dst[i] = id;
          contiguous_new = false;
          if (id < n_linked) pending_updates.push_back(id);  // already in the graph: re-link at next build
        } else {
          dst[i] = (uint32_t)nn;
          if (!identity_labels) lookup[l] = (uint32_t)nn;
          h_labels.push_back(l);
          nn++;
        }
      }
      RET(ensure_capacity(nn));
      // levels + upper rows for the new ids
      uint64_t new_cnt = nn - first_new;
      std::vector<uint8_t> lv(new_cnt);
      std::vector<uint32_t> uo(new_cnt), owners;
      uint64_t rows = up_rows;
      for (uint64_t j = 0; j < new_cnt; ++j) {
        int l = draw_level();
        lv[j] = (uint8_t)l;
        uo[j] = l ? (uint32_t)rows : ehb::kInvalid;
        for (int t = 0; t < l; ++t) owners.push_back((uint32_t)(first_new + j));
        rows += l;
      }
      RET(ensure_upper(rows));
      if (new_cnt) {
        CU(cudaMemcpyAsync(levels.p + first_new, lv.data(), new_cnt, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(up_off.p + first_new, uo.data(), new_cnt * 4, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(labels.p + first_new, h_labels.data() + first_new, new_cnt * 8, cudaMemcpyHostToDevice,
                           stream));
        if (!owners.empty())
          CU(cudaMemcpyAsync(up_owner.p + up_rows, owners.data(), owners.size() * 4, cudaMemcpyHostToDevice, stream));
        h_levels.insert(h_levels.end(), lv.begin(), lv.end());
      }
      // stage the rows and scatter/pad/normalise them into place
      const float* dsrc;
      if (src_is_device) {
        dsrc = src + off * dim;
      } else {
        CU(b_stage_in.grow(m * dim, 0, -1, stream));
        CU(cudaMemcpyAsync(b_stage_in.p, src + off * dim, m * dim * 4, cudaMemcpyHostToDevice, stream));
        dsrc = b_stage_in.p;
      }
      if (contiguous_new) {
        CU(ehb::launch_pad_rows(dsrc, vecs.p + first_new * dpad, m, dim, dpad, metric == EHB_COSINE, stream));
      } else {
        // rows go to arbitrary ids: one launch per run of consecutive destinations
        uint64_t i = 0;
        while (i < m) {
          uint64_t j = i + 1;
          while (j < m && dst[j] == dst[j - 1] + 1) ++j;
          CU(ehb::launch_pad_rows(dsrc + i * dim, vecs.p + (uint64_t)dst[i] * dpad, j - i, dim, dpad,
                                  metric == EHB_COSINE, stream));
          i = j;
        }
      }
      CU(cudaStreamSynchronize(stream));  // host staging vectors go out of scope
      n = nn;
      up_rows = rows;
      bf16_rows = 0;
    }
    return EHB_OK;
  }

  // ---- construction ----------------------------------------------------------
  int ensure_build_scratch(uint32_t maxb) {
    uint64_t ecap = (uint64_t)maxb * M * 2 + 1024;
    CU(b_edge_row.grow(ecap, 0, -1, stream));
    CU(b_edge_src.grow(ecap, 0, -1, stream));
    CU(b_edge_dist.grow(ecap, 0, -1, stream));
    CU(b_touched.grow(ecap, 0, -1, stream));
    CU(b_seg_src.grow(ecap, 0, -1, stream));
    CU(b_seg_dist.grow(ecap, 0, -1, stream));
    CU(b_counters.grow(8, 0, 0, stream));
    uint64_t rowspace = cap + links_up.n / M;
    if (b_row_cnt.n < rowspace) {
      b_row_cnt.release();
      b_row_fill.release();
      b_row_start.release();
      CU(b_row_cnt.grow(rowspace, 0, 0, stream));
      CU(b_row_fill.grow(rowspace, 0, 0, stream));
      CU(b_row_start.grow(rowspace, 0, 0, stream));
    }
    return EHB_OK;
  }
  ehb::BuildBuffers build_buffers(uint32_t b) {
    ehb::BuildBuffers bb;
    bb.edge_row = b_edge_row.p;
    bb.edge_src = b_edge_src.p;
    bb.edge_dist = b_edge_dist.p;
    bb.edge_count = b_counters.p + 0;
    bb.edge_cap = (uint32_t)std::min<uint64_t>(b_edge_row.n, (uint64_t)b * M * 2 + 1024);
    bb.row_cnt = b_row_cnt.p;
    bb.row_fill = b_row_fill.p;
    bb.row_start = b_row_start.p;
    bb.touched = b_touched.p;
    bb.touched_count = b_counters.p + 1;
    bb.seg_cursor = b_counters.p + 2;
    bb.seg_src = b_seg_src.p;
    bb.seg_dist = b_seg_dist.p;
    bb.error_flag = b_counters.p + 3;
    return bb;
  }
  ehb::BuildGraph build_graph() const {
    ehb::BuildGraph bg;
    bg.g = view();
    bg.levels = levels.p;
    bg.up_owner = up_owner.p;
    bg.cap = (uint32_t)cap;
    bg.efc = std::max(prm.ef_construction, M);
    return bg;
  }

  int build() {
    if (n_linked == n && pending_updates.empty()) return EHB_OK;
    const uint32_t maxb = prm.build_batch ? prm.build_batch : 16384;
    RET(ensure_build_scratch(maxb));
    const uint32_t maxb0 = prm.build_batch ? prm.build_batch : 16384;
    ehb::WalkCfg cfg = walk_cfg(std::max(prm.ef_construction, M), 256, std::min<uint64_t>(maxb0, n), 1);
    uint32_t wpb = wpb_for(cfg, 256);
    while (n_linked < n) {
      if (n_linked == 0) {
        entry = 0;
        max_level = h_levels[0];
        n_linked = 1;
        continue;
      }
      // a wave never exceeds 1/64 of the linked graph: points of one wave cannot see each other
      // (measured: recall within sampling noise of the sequential build from 1/32 on)
      const uint64_t frac = o_build_frac ? o_build_frac : 64;
      uint64_t b = std::min<uint64_t>(maxb, std::max<uint64_t>(1, n_linked / frac));
      b = std::min<uint64_t>(b, n - n_linked);
      ehb::BuildGraph bg = build_graph();
      ehb::BuildBuffers bb = build_buffers((uint32_t)b);
      CU(ehb::launch_build_batch(bg, cfg, nullptr, (uint32_t)n_linked, (uint32_t)b, false, bb, wpb, stream));
      for (uint64_t i = n_linked; i < n_linked + b; ++i)
        if ((int)h_levels[i] > max_level) max_level = h_levels[i], entry = (uint32_t)i;
      n_linked += b;
    }
    if (!pending_updates.empty()) {
      std::sort(pending_updates.begin(), pending_updates.end());
      pending_updates.erase(std::unique(pending_updates.begin(), pending_updates.end()), pending_updates.end());
      if (n_linked > 1) {
        for (size_t off = 0; off < pending_updates.size(); off += maxb) {
          uint32_t b = (uint32_t)std::min<size_t>(maxb, pending_updates.size() - off);
          CU(b_ids.grow(b, 0, -1, stream));
          CU(cudaMemcpyAsync(b_ids.p, pending_updates.data() + off, (size_t)b * 4, cudaMemcpyHostToDevice, stream));
          ehb::BuildGraph bg = build_graph();
          ehb::BuildBuffers bb = build_buffers(b);
          CU(ehb::launch_build_batch(bg, cfg, b_ids.p, 0, b, true, bb, wpb, stream));
          CU(cudaStreamSynchronize(stream));
        }
      }
      pending_updates.clear();
    }
    uint32_t err = 0;
    CU(cudaMemcpyAsync(&err, b_counters.p + 3, 4, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    if (err) return fail(EHB_ERR_STATE, "build: edge buffer overflow");
    return EHB_OK;
  }

  // ---- search ------------------------------------------------------------------
  int search_dev(uint64_t nq, const float* dq, uint32_t k, uint32_t ef_in, uint64_t* dl, float* dd, uint32_t* dc,
                 cudaStream_t s) {
    if (k == 0 || nq == 0) return EHB_OK;
    uint32_t ef_eff = std::max(ef_in ? ef_in : ef, k);
    if (ef_eff > ehb::kMaxEf) return fail(EHB_ERR_INVALID, "max(ef, k) must be <= 512");
    RET(build());
    // Warps per query (rows <= 1 KB, ef <= 256): four while 3 CTAs of 128 threads per SM hold every query (small
    // online batches; Q=1: 135 us vs 252 us with one warp), two while 7 CTAs of 64 threads do (C2, Q=1000:
    // 0.288 ms vs 0.409 ms), else one warp per query (C5 shape, Q=10k: 9.2 ms vs 10.7 ms with two).
    uint32_t team = t_team;
    if (team == 0) {
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      team = nq <= (uint64_t)sms * 3 ? 4 : (nq <= (uint64_t)sms * 7 ? 2 : 1);
    }
    if (dpad > 256 || ef_eff > 256) team = 1;
    ehb::WalkCfg cfg = walk_cfg(ef_eff, 0, nq * team, team);
    const float* q = dq;
    if (metric == EHB_COSINE) {
      CU(q_norm.grow(nq * dim, 0, -1, s));
      CU(ehb::launch_pad_rows(dq, q_norm.p, nq, dim, dim, true, s));
      q = q_norm.p;
    }
    CU(stats.grow(nq * 4, 0, -1, s));
    CU(stat_sum.grow(4, 0, 0, s));
    uint32_t wpb = wpb_for(cfg, 0);
    CU(cudaEventRecord(ev0, s));
    if (team >= 2)
      CU(ehb::launch_search_team(team, view(), cfg.hash_size, q, (uint32_t)nq, k, ef_eff, dl, dd, dc, stats.p, s));
    else
      CU(ehb::launch_search(view(), cfg, q, (uint32_t)nq, k, ef_eff, dl, dd, dc, stats.p, wpb, s));
    CU(cudaEventRecord(ev1, s));
    timed = true;
    last_nq = nq;
    last_sum_valid = false;
    return EHB_OK;
  }

  int ensure_out(uint64_t nq, uint32_t k, cudaStream_t s) {
    CU(q_in.grow(nq * dim, 0, -1, s));
    CU(o_labels.grow(nq * k, 0, -1, s));
    CU(o_dists.grow(nq * k, 0, -1, s));
    CU(o_counts.grow(nq, 0, -1, s));
    return EHB_OK;
  }

  int bruteforce_dev(uint64_t nq, const float* dq, uint32_t k, int precision, uint64_t* dl, float* dd, uint32_t* dc,
                     cudaStream_t s) {
    if (k == 0 || nq == 0) return EHB_OK;
    if (precision != EHB_FP32 && precision != EHB_BF16) return fail(EHB_ERR_INVALID, "unknown precision");
    const bool bf16 = precision == EHB_BF16;
    if (bf16 && dpad % 64 != 0) return fail(EHB_ERR_INVALID, "bf16 brute force needs dim > 32 (64-wide k-blocks)");
    if (k > 2048) return fail(EHB_ERR_INVALID, "k must be <= 2048 for brute force");
    // candidates kept by the bf16 pass: bf16 rounding perturbs each dot product by ~|q||x| 2^-9 / sqrt(d),
    // comparable to the spacing of the best matches, so 4x (>= k + 64) of them go to the fp32 re-rank
    const uint32_t kc = bf16 ? (uint32_t)std::min<uint64_t>(std::min<uint64_t>(2048, std::max<uint64_t>(n, 1)),
                                                            std::max<uint64_t>(4ull * k, k + 64ull)) : k;
    ehb::BruteScratch sc;
    sc.qb = std::min<uint64_t>(nq, bf16 ? 2048 : 1024);
    sc.nc = std::min<uint64_t>(std::max<uint64_t>(n, 1), 131072);
    sc.slices = 32;
    CU(bf_dist.grow(sc.qb * sc.nc, 0, -1, s));
    CU(bf_part.grow(sc.qb * sc.slices * std::max(kc, k), 0, -1, s));
    CU(bf_run.grow(nq * std::max(kc, k), 0, -1, s));
    CU(bf_qpad.grow(nq * dpad, 0, -1, s));
    CU(ehb::launch_pad_rows(dq, bf_qpad.p, nq, dim, dpad, metric == EHB_COSINE, s));
    sc.dist = bf_dist.p;
    sc.part_keys = bf_part.p;
    sc.run_keys = bf_run.p;
    ehb::Bf16Ctx bctx;
    if (bf16) {
      // bf16 shadow of the base rows (+ squared norms), refreshed lazily after mutations
      if (bf16_rows != n) {
        CU(x_bf16.grow(std::max<uint64_t>(n, 1) * dpad, 0, -1, s));
        CU(x_norm.grow(std::max<uint64_t>(n, 1), 0, -1, s));
        CU(ehb::launch_to_bf16(vecs.p, dpad, x_bf16.p, x_norm.p, n, dpad, s));
        bf16_rows = n;
      }
      CU(q_bf16.grow(nq * dpad, 0, -1, s));
      CU(q_norm2.grow(nq, 0, -1, s));
      CU(ehb::launch_to_bf16(bf_qpad.p, dpad, q_bf16.p, q_norm2.p, nq, dpad, s));
      bctx.q_bf16 = q_bf16.p;
      bctx.x_bf16 = x_bf16.p;
      bctx.qnorm = q_norm2.p;
      bctx.xnorm = x_norm.p;
      bctx.kc = kc;
      // fused selection state (option "bf16_unfused" keeps the distance tiles in HBM: A/B switch)
      bctx.fused = !o_bf16_unfused;
      bctx.variant = o_gemm_2cta ? 1 : 0;
      bctx.ccap = 2 * kc + 64;
      CU(bf_thr.grow(nq, 0, -1, s));
      CU(bf_cbuf.grow(nq * bctx.ccap, 0, -1, s));
      CU(bf_ccount.grow(nq + 1, 0, 0, s));
      bctx.thr = bf_thr.p;
      bctx.cbuf = bf_cbuf.p;
      bctx.ccount = bf_ccount.p;
      bctx.overflow = bf_ccount.p + nq;
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      bctx.sms = sms;
    }
    CU(cudaEventRecord(ev0, s));
    CU(ehb::launch_bruteforce(vecs.p, dpad, dim, n, labels.p, metric == EHB_L2 ? 0 : 1, bf_qpad.p, nq, k, sc,
                              bf16 ? &bctx : nullptr, dl, dd, dc, s));
    CU(cudaEventRecord(ev1, s));
    timed = true;
    last_nq = 0;
    return EHB_OK;
  }
};

// ============================================================================
// C ABI
// ============================================================================
extern "C" {

const char* ehb_last_error(void) { return g_err.c_str(); }
uint32_t ehb_abi_version(void) { return 1; }

void ehb_params_default(ehb_params* p, uint32_t dim) {
  std::memset(p, 0, sizeof(*p));
  p->dim = dim;
  p->metric = EHB_L2;
  p->capacity = 128;        // index.h:21
  p->M = 16;                // hnswlib default used by index.cc:14-15
  p->ef_construction = 200;
  p->ef_search = 10;        // hnswlib ef_ default; the reference never calls setEf
  p->seed = 100;
  p->device = 0;
}

int ehb_index_create(const ehb_params* p, ehb_index** out) {
  if (!p || !out) return fail(EHB_ERR_INVALID, "null argument");
  if (p->dim == 0 || p->dim > ehb::kMaxDim) return fail(EHB_ERR_INVALID, "dim must be in 1..2048");
  if (p->M < 2 || p->M > 16) return fail(EHB_ERR_INVALID, "M must be in 2..16");
  if (p->ef_construction > 256) return fail(EHB_ERR_INVALID, "ef_construction must be <= 256");
  if (p->metric < 0 || p->metric > 2) return fail(EHB_ERR_INVALID, "unknown metric");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(EHB_ERR_CUDA, std::string("no CUDA device (ehb200 has no CPU fallback): ") + cudaGetErrorString(e));
  if (p->device < 0 || p->device >= ndev) return fail(EHB_ERR_INVALID, "bad device ordinal");
  CU(cudaSetDevice(p->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, p->device));
  if (prop.major < 10) return fail(EHB_ERR_CUDA, "ehb200 kernels are built for sm_100a only");
  ehb_index* ix = new (std::nothrow) ehb_index();
  if (!ix) return fail(EHB_ERR_OOM, "host allocation failed");
  ix->prm = *p;
  ix->dim = p->dim;
  ix->dpad = ehb::pad_dim(p->dim);
  ix->M = p->M;
  ix->M0 = 2 * p->M;
  ix->metric = p->metric;
  ix->device = p->device;
  ix->ef = p->ef_search ? p->ef_search : 10;
  ix->level_rng.seed((unsigned)p->seed);
  if (ix->prm.ef_construction == 0) ix->prm.ef_construction = 200;
  cudaError_t ce = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreate(&ix->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&ix->ev1);
  if (ce != cudaSuccess) {
    delete ix;
    return fail(EHB_ERR_CUDA, cudaGetErrorString(ce));
  }
  int rc = ix->ensure_capacity(std::max<uint64_t>(p->capacity, 1));
  if (rc != EHB_OK) {
    delete ix;
    return rc;
  }
  *out = ix;
  return EHB_OK;
}

int ehb_index_destroy(ehb_index* ix) {
  if (!ix) return EHB_OK;
  cudaSetDevice(ix->device);
  cudaStreamSynchronize(ix->stream);
  delete ix;
  return EHB_OK;
}

#define ENTER(ix)                                              \
  if (!(ix)) return fail(EHB_ERR_INVALID, "null index handle"); \
  std::lock_guard<std::mutex> _g((ix)->mu);                    \
  CU(cudaSetDevice((ix)->device))

int ehb_index_add(ehb_index* ix, uint64_t n, const float* vecs, const uint64_t* labels) {
  ENTER(ix);
  if (n && !vecs) return fail(EHB_ERR_INVALID, "null vectors");
  return ix->add_rows(n, vecs, false, labels);
}
int ehb_index_add_dev(ehb_index* ix, uint64_t n, const float* vecs_dev, const uint64_t* labels) {
  ENTER(ix);
  if (n && !vecs_dev) return fail(EHB_ERR_INVALID, "null vectors");
  return ix->add_rows(n, vecs_dev, true, labels);
}
int ehb_index_build(ehb_index* ix) {
  ENTER(ix);
  RET(ix->build());
  CU(cudaStreamSynchronize(ix->stream));
  return EHB_OK;
}
int ehb_index_set_ef(ehb_index* ix, uint32_t ef) {
  ENTER(ix);
  if (ef == 0) return fail(EHB_ERR_INVALID, "ef must be > 0");
  ix->ef = ef;
  return EHB_OK;
}
int ehb_index_size(ehb_index* ix, uint64_t* out) {
  ENTER(ix);
  *out = ix->n;
  return EHB_OK;
}

int ehb_index_get(ehb_index* ix, uint64_t label, float* out) {
  ENTER(ix);
  uint32_t id;
  if (ix->identity_labels) {
    if (label >= ix->n) return fail(EHB_ERR_NOT_FOUND, "label not found");
    id = (uint32_t)label;
  } else {
    auto it = ix->lookup.find(label);
    if (it == ix->lookup.end()) return fail(EHB_ERR_NOT_FOUND, "label not found");
    id = it->second;
  }
  CU(cudaMemcpyAsync(out, ix->vecs.p + (uint64_t)id * ix->dpad, ix->dim * 4, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return EHB_OK;
}

static int search_host(ehb_index* ix, bool brute, uint64_t nq, const float* q, uint32_t k, uint32_t ef, int precision,
                       uint64_t* ol, float* od, uint32_t* oc) {
  if (nq && (!q || !ol)) return fail(EHB_ERR_INVALID, "null buffer");
  if (k == 0 || nq == 0) return EHB_OK;
  cudaStream_t s = ix->stream;
  RET(ix->ensure_out(nq, k, s));
  CU(cudaMemcpyAsync(ix->q_in.p, q, nq * ix->dim * 4, cudaMemcpyHostToDevice, s));
  if (brute)
    RET(ix->bruteforce_dev(nq, ix->q_in.p, k, precision, ix->o_labels.p, ix->o_dists.p, ix->o_counts.p, s));
  else
    RET(ix->search_dev(nq, ix->q_in.p, k, ef, ix->o_labels.p, ix->o_dists.p, ix->o_counts.p, s));
  CU(cudaMemcpyAsync(ol, ix->o_labels.p, nq * k * 8, cudaMemcpyDeviceToHost, s));
  if (od) CU(cudaMemcpyAsync(od, ix->o_dists.p, nq * k * 4, cudaMemcpyDeviceToHost, s));
  if (oc) CU(cudaMemcpyAsync(oc, ix->o_counts.p, nq * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return EHB_OK;
}

int ehb_index_search(ehb_index* ix, uint64_t nq, const float* q, uint32_t k, uint32_t ef, uint64_t* ol, float* od,
                     uint32_t* oc) {
  ENTER(ix);
  return search_host(ix, false, nq, q, k, ef, 0, ol, od, oc);
}
int ehb_index_search_dev(ehb_index* ix, uint64_t nq, const float* dq, uint32_t k, uint32_t ef, uint64_t* dl, float* dd,
                         uint32_t* dc, void* stream) {
  ENTER(ix);
  if (nq && (!dq || !dl)) return fail(EHB_ERR_INVALID, "null buffer");
  return ix->search_dev(nq, dq, k, ef, dl, dd, dc, stream ? (cudaStream_t)stream : ix->stream);
}
int ehb_index_search_bruteforce(ehb_index* ix, uint64_t nq, const float* q, uint32_t k, int precision, uint64_t* ol,
                                float* od, uint32_t* oc) {
  ENTER(ix);
  return search_host(ix, true, nq, q, k, 0, precision, ol, od, oc);
}
int ehb_index_search_bruteforce_dev(ehb_index* ix, uint64_t nq, const float* dq, uint32_t k, int precision,
                                    uint64_t* dl, float* dd, uint32_t* dc, void* stream) {
  ENTER(ix);
  if (nq && (!dq || !dl)) return fail(EHB_ERR_INVALID, "null buffer");
  return ix->bruteforce_dev(nq, dq, k, precision, dl, dd, dc, stream ? (cudaStream_t)stream : ix->stream);
}

int ehb_index_stats(ehb_index* ix, ehb_stats* out) {
  ENTER(ix);
  if (!out) return fail(EHB_ERR_INVALID, "null out");
  std::memset(out, 0, sizeof(*out));
  if (ix->last_nq && !ix->last_sum_valid) {
    CU(cudaEventSynchronize(ix->ev1));
    CU(ehb::launch_sum_stats(ix->stats.p, (uint32_t)ix->last_nq, ix->stat_sum.p, ix->stream));
    CU(cudaMemcpyAsync(ix->last_sum, ix->stat_sum.p, 32, cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
    ix->last_sum_valid = true;
  }
  if (ix->last_nq) {
    out->queries = ix->last_nq;
    out->hops_upper = ix->last_sum[0];
    out->hops_base = ix->last_sum[1];
    out->dist_evals = ix->last_sum[2];
    out->visited_overflow = ix->last_sum[3];
    out->algorithmic_bytes = out->hops_upper * 4ull * ix->M + out->hops_base * 4ull * ix->M0 +
                             out->dist_evals * 4ull * ix->dim + out->queries * 4ull * ix->dim;
  }
  out->size = ix->n;
  out->capacity = ix->cap;
  out->upper_rows = ix->up_rows;
  out->dim = ix->dim;
  out->M = ix->M;
  out->max_level = ix->max_level < 0 ? 0 : (uint32_t)ix->max_level;
  out->entry_point = ix->entry;
  out->device_bytes = ix->vecs.bytes() + ix->labels.bytes() + ix->levels.bytes() + ix->links0.bytes() +
                      ix->up_off.bytes() + ix->links_up.bytes() + ix->up_owner.bytes();
  return EHB_OK;
}

int ehb_index_last_kernel_ms(ehb_index* ix, float* out_ms) {
  ENTER(ix);
  if (!ix->timed) return fail(EHB_ERR_STATE, "no search has been timed yet");
  CU(cudaEventSynchronize(ix->ev1));
  CU(cudaEventElapsedTime(out_ms, ix->ev0, ix->ev1));
  return EHB_OK;
}

int ehb_index_set_search_width(ehb_index* ix, uint32_t warps_per_query) {
  ENTER(ix);
  if (warps_per_query > 4) return fail(EHB_ERR_INVALID, "warps_per_query must be 0 (auto) or 1..4");
  ix->t_team = warps_per_query;
  return EHB_OK;
}

int ehb_index_set_option(ehb_index* ix, const char* name, int64_t value) {
  ENTER(ix);
  if (!name) return fail(EHB_ERR_INVALID, "null option name");
  const std::string o(name);
  if (o == "build_frac") {
    if (value < 0 || value > (1 << 20)) return fail(EHB_ERR_INVALID, "build_frac must be in 0..2^20");
    ix->o_build_frac = (uint32_t)value;
  } else if (o == "bf16_unfused") {
    ix->o_bf16_unfused = value != 0;
  } else if (o == "gemm_2cta") {
    ix->o_gemm_2cta = value != 0;
  } else {
    return fail(EHB_ERR_INVALID, "unknown option: " + o);
  }
  return EHB_OK;
}

int ehb_index_set_tuning(ehb_index* ix, uint32_t slots, uint32_t groups, uint32_t hash_bits, uint32_t wpb) {
  ENTER(ix);
  ix->t_slots = slots;
  ix->t_groups = groups;
  ix->t_hash_bits = hash_bits;
  ix->t_wpb = wpb;
  return EHB_OK;
}

int ehb_index_export_graph(ehb_index* ix, float* vectors, uint64_t* labels, uint8_t* levels, uint32_t* links0,
                           uint32_t* up_off, uint32_t* links_up, uint32_t* entry, int32_t* max_level) {
  ENTER(ix);
  RET(ix->build());
  cudaStream_t s = ix->stream;
  uint64_t n = ix->n;
  if (vectors && n)
    CU(cudaMemcpy2DAsync(vectors, ix->dim * 4, ix->vecs.p, ix->dpad * 4, ix->dim * 4, n, cudaMemcpyDeviceToHost, s));
  if (labels && n) CU(cudaMemcpyAsync(labels, ix->labels.p, n * 8, cudaMemcpyDeviceToHost, s));
  if (levels && n) CU(cudaMemcpyAsync(levels, ix->levels.p, n, cudaMemcpyDeviceToHost, s));
  if (links0 && n) CU(cudaMemcpyAsync(links0, ix->links0.p, n * ix->M0 * 4, cudaMemcpyDeviceToHost, s));
  if (up_off && n) CU(cudaMemcpyAsync(up_off, ix->up_off.p, n * 4, cudaMemcpyDeviceToHost, s));
  if (links_up && ix->up_rows)
    CU(cudaMemcpyAsync(links_up, ix->links_up.p, ix->up_rows * ix->M * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (entry) *entry = ix->entry;
  if (max_level) *max_level = ix->max_level;
  return EHB_OK;
}

int ehb_index_import_graph(ehb_index* ix, uint64_t n, const float* vectors, const uint64_t* labels,
                           const uint8_t* levels, const uint32_t* links0, const uint32_t* up_off, uint64_t upper_rows,
                           const uint32_t* links_up, uint32_t entry, int32_t max_level) {
  ENTER(ix);
  if (n && (!vectors || !labels || !levels || !links0 || !up_off)) return fail(EHB_ERR_INVALID, "null buffer");
  if (upper_rows && !links_up) return fail(EHB_ERR_INVALID, "null links_up");
  cudaStream_t s = ix->stream;
  ix->n = ix->n_linked = ix->up_rows = 0;
  ix->lookup.clear();
  ix->h_labels.clear();
  ix->h_levels.clear();
  ix->pending_updates.clear();
  ix->identity_labels = true;
  RET(ix->ensure_capacity(std::max<uint64_t>(n, 1)));
  RET(ix->ensure_upper(std::max<uint64_t>(upper_rows, 1)));
  if (n) {
    CU(ix->b_stage_in.grow(n * ix->dim, 0, -1, s));
    CU(cudaMemcpyAsync(ix->b_stage_in.p, vectors, n * ix->dim * 4, cudaMemcpyHostToDevice, s));
    CU(ehb::launch_pad_rows(ix->b_stage_in.p, ix->vecs.p, n, ix->dim, ix->dpad, false, s));
    CU(cudaMemcpyAsync(ix->labels.p, labels, n * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ix->levels.p, levels, n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ix->links0.p, links0, n * ix->M0 * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ix->up_off.p, up_off, n * 4, cudaMemcpyHostToDevice, s));
    if (upper_rows) {
      CU(cudaMemcpyAsync(ix->links_up.p, links_up, upper_rows * ix->M * 4, cudaMemcpyHostToDevice, s));
      std::vector<uint32_t> owners(upper_rows);
      for (uint64_t i = 0; i < n; ++i)
        for (int l = 0; l < levels[i]; ++l) owners[up_off[i] + l] = (uint32_t)i;
      CU(cudaMemcpyAsync(ix->up_owner.p, owners.data(), upper_rows * 4, cudaMemcpyHostToDevice, s));
      CU(cudaStreamSynchronize(s));
    }
    CU(cudaStreamSynchronize(s));
  }
  // rows past the imported range must read as empty for later inserts
  if (ix->cap > n) {
    CU(cudaMemsetAsync(ix->links0.p + n * ix->M0, 0xFF, (ix->cap - n) * ix->M0 * 4, s));
    CU(cudaMemsetAsync(ix->up_off.p + n, 0xFF, (ix->cap - n) * 4, s));
  }
  if (ix->links_up.n > upper_rows * ix->M)
    CU(cudaMemsetAsync(ix->links_up.p + upper_rows * ix->M, 0xFF, (ix->links_up.n - upper_rows * ix->M) * 4, s));
  CU(cudaStreamSynchronize(s));
  ix->h_labels.assign(labels, labels + n);
  ix->h_levels.assign(levels, levels + n);
  for (uint64_t i = 0; i < n; ++i)
    if (labels[i] != i) ix->identity_labels = false;
  if (!ix->identity_labels)
    for (uint64_t i = 0; i < n; ++i) ix->lookup[labels[i]] = (uint32_t)i;
  ix->n = ix->n_linked = n;
  ix->bf16_rows = 0;
  ix->up_rows = upper_rows;
  ix->entry = entry;
  ix->max_level = n ? max_level : -1;
  return EHB_OK;
}

// File format: "EHB200\0\1" header, params, counts, then the export arrays.
int ehb_index_save(ehb_index* ix, const char* path) {
  if (!ix || !path) return fail(EHB_ERR_INVALID, "null argument");
  uint64_t n, rows;
  uint32_t dim, M;
  {
    std::lock_guard<std::mutex> g(ix->mu);
    n = ix->n, rows = ix->up_rows, dim = ix->dim, M = ix->M;
  }
  std::vector<float> v(n * dim);
  std::vector<uint64_t> lab(n);
  std::vector<uint8_t> lev(n);
  std::vector<uint32_t> l0(n * 2 * M), uo(n), lu(std::max<uint64_t>(rows, 1) * M);
  uint32_t entry = 0;
  int32_t maxl = -1;
  RET(ehb_index_export_graph(ix, v.data(), lab.data(), lev.data(), l0.data(), uo.data(), lu.data(), &entry, &maxl));
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(EHB_ERR_IO, std::string("cannot open ") + path);
  const char magic[8] = {'E', 'H', 'B', '2', '0', '0', 0, 1};
  uint64_t hdr[4] = {n, rows, entry, (uint64_t)(int64_t)maxl};
  bool ok = std::fwrite(magic, 1, 8, f) == 8 && std::fwrite(&ix->prm, sizeof(ehb_params), 1, f) == 1 &&
            std::fwrite(hdr, 8, 4, f) == 4;
  auto wr = [&](const void* p, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes); };
  wr(v.data(), v.size() * 4);
  wr(lab.data(), n * 8);
  wr(lev.data(), n);
  wr(l0.data(), l0.size() * 4);
  wr(uo.data(), n * 4);
  wr(lu.data(), rows * M * 4);
  ok = (std::fclose(f) == 0) && ok;
  return ok ? EHB_OK : fail(EHB_ERR_IO, "short write");
}

int ehb_index_load(const char* path, int32_t device, ehb_index** out) {
  if (!path || !out) return fail(EHB_ERR_INVALID, "null argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(EHB_ERR_IO, std::string("cannot open ") + path);
  char magic[8];
  ehb_params p;
  uint64_t hdr[4];
  if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "EHB200", 6) != 0 ||
      std::fread(&p, sizeof(p), 1, f) != 1 || std::fread(hdr, 8, 4, f) != 4) {
    std::fclose(f);
    return fail(EHB_ERR_IO, "bad header");
  }
  uint64_t n = hdr[0], rows = hdr[1];
  p.device = device;
  p.capacity = std::max<uint64_t>(n, 1);
  std::vector<float> v(n * p.dim);
  std::vector<uint64_t> lab(n);
  std::vector<uint8_t> lev(n);
  std::vector<uint32_t> l0(n * 2 * p.M), uo(n), lu(std::max<uint64_t>(rows, 1) * p.M);
  bool ok = true;
  auto rd = [&](void* d, size_t bytes) { ok = ok && (bytes == 0 || std::fread(d, 1, bytes, f) == bytes); };
  rd(v.data(), v.size() * 4);
  rd(lab.data(), n * 8);
  rd(lev.data(), n);
  rd(l0.data(), l0.size() * 4);
  rd(uo.data(), n * 4);
  rd(lu.data(), rows * p.M * 4);
  std::fclose(f);
  if (!ok) return fail(EHB_ERR_IO, "short read");
  ehb_index* ix = nullptr;
  RET(ehb_index_create(&p, &ix));
  int rc = ehb_index_import_graph(ix, n, v.data(), lab.data(), lev.data(), l0.data(), uo.data(), rows, lu.data(),
                                  (uint32_t)hdr[2], (int32_t)(int64_t)hdr[3]);
  if (rc != EHB_OK) {
    ehb_index_destroy(ix);
    return rc;
  }
  *out = ix;
  return EHB_OK;
}

int ehb_merge_topk_dev(uint32_t G, uint64_t nq, uint32_t k, const float* dists, const uint64_t* labels,
                       float* out_dists, uint64_t* out_labels, uint32_t* out_counts, int32_t device, void* stream) {
  if (G == 0 || G > 32) return fail(EHB_ERR_INVALID, "G must be in 1..32");
  if (nq && k && (!dists || !labels || !out_labels)) return fail(EHB_ERR_INVALID, "null buffer");
  CU(cudaSetDevice(device));
  CU(ehb::launch_merge_topk(G, nq, k, dists, labels, nq * k * 4ull, nq * k * 8ull, out_dists, out_labels, out_counts,
                            (cudaStream_t)stream));
  return EHB_OK;
}

int ehb_merge_topk_packed_dev(uint32_t G, uint64_t nq, uint32_t k, const void* packed, uint64_t rank_stride_bytes,
                              float* out_dists, uint64_t* out_labels, uint32_t* out_counts, int32_t device,
                              void* stream) {
  if (G == 0 || G > 32) return fail(EHB_ERR_INVALID, "G must be in 1..32");
  if (nq && k && (!packed || !out_labels)) return fail(EHB_ERR_INVALID, "null buffer");
  if (rank_stride_bytes < nq * k * 12ull || (rank_stride_bytes & 7u)) return fail(EHB_ERR_INVALID, "bad rank stride");
  CU(cudaSetDevice(device));
  const unsigned char* base = (const unsigned char*)packed;
  CU(ehb::launch_merge_topk(G, nq, k, (const float*)(base + nq * k * 8ull), (const uint64_t*)base, rank_stride_bytes,
                            rank_stride_bytes, out_dists, out_labels, out_counts, (cudaStream_t)stream));
  return EHB_OK;
}

}  // extern "C"
