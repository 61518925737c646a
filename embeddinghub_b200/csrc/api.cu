// Host side of the ehb200 C ABI (include/ehb200.h): index state in HBM, batched construction driver,
// tombstones, re-entrant search entry points with a combining queue.  Mirrors the responsibilities of
// featureform::embedding::ANNIndex + hnswlib::HierarchicalNSW as used in
// embeddinghub/embeddingstore/index.cc:10-52.  (Persistence: io.cu; sharding / shard exchange: exchange.cu.)
#include "index_impl.h"

namespace ehb {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
const std::string& last_error_text() { return g_err; }
}  // namespace ehb
using ehb::fail;

// ============================================================================================
// index state
// ============================================================================================
ehb_index::~ehb_index() {
  slots.clear();
  if (bf_ev0) cudaEventDestroy(bf_ev0);
  if (bf_ev1) cudaEventDestroy(bf_ev1);
  if (stream) cudaStreamDestroy(stream);
}

ehb::GraphView ehb_index::view() const {
  ehb::GraphView g;
  g.vecs = vecs.p;
  g.links0 = links0.p;
  g.up_off = up_off.p;
  g.links_up = links_up.p;
  g.labels = labels.p;
  g.deleted = n_deleted ? deleted.p : nullptr;
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
ehb::WalkCfg ehb_index::walk_cfg(uint32_t ef_eff, uint32_t smem_list, uint64_t jobs, uint32_t team) const {
  ehb::WalkCfg c;
  c.lcap = smem_list;
  c.staged = dpad > 256 ? 1 : 0;  // rows above 1 KB go through the TMA staging ring
  c.dcap = n_deleted ? ehb::kDeletedQueue : 0;
  c.prefetch = o_walk_prefetch ? 1 : 0;
  // dense walk (search_impl.cuh): batches big enough to fill 20 warps per SM, rows <= 512 B, no tombstones
  c.dense = (!smem_list && team == 1 && !c.staged && dpad <= 128 && !n_deleted && jobs >= 20ull * (uint64_t)sms) ? 1 : 0;
  const uint32_t warp_target = c.dense ? 20u : 16u;  // resident warps per SM the visited-table sizing aims at
  uint32_t vbytes = dpad * 4;
  uint32_t nslots = std::max(4u, std::min(32u, 24576u / vbytes));
  uint32_t ng = 2;                      // two groups: math on one overlaps the copies of the other
  uint32_t g = std::max(4u, nslots / ng / 4 * 4);  // vectors per group, multiple of the 4-vector math step
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
    uint64_t ctas = (jobs + team - 1) / std::max(team, 1u);
    uint32_t want = (uint32_t)std::min<uint64_t>((ctas + sms - 1) / sms, c.staged ? 5u : warp_target / team);
    want = std::max(want, 4u);
    c.hash_size = 0;
    uint32_t fixed = ehb::warp_smem_bytes(c, dpad) * team + 1024u + (smem_list ? 256u : 0u);
    uint32_t per_cta = (227u * 1024u) / want;
    uint32_t avail = per_cta > fixed + 1024u ? (per_cta - fixed) / 4u : 256u;
    hs = std::min(roomy, std::max(tight, avail));
    if (n_deleted) hs = roomy;  // tombstoned candidates rely on the visited table alone (no result-set filter)
  }
  c.hash_size = ehb::align_up(std::max(hs, 256u), 32);
  // stay inside the 227 KB per-block limit
  while (ehb::warp_smem_bytes(c, dpad) + 256 > 200 * 1024 && c.hash_size > 512)
    c.hash_size = ehb::align_up(c.hash_size / 2, 32);
  return c;
}

uint32_t ehb_index::wpb_for(const ehb::WalkCfg& c, uint32_t extra) const {
  uint32_t w = t_wpb ? t_wpb : 1;
  while (w > 1 && (size_t)(ehb::warp_smem_bytes(c, dpad) + extra) * w > 220 * 1024) w >>= 1;
  return w;
}

int ehb_index::ensure_capacity(uint64_t want) {
  if (want <= cap) return EHB_OK;
  uint64_t nc = std::max<uint64_t>(cap ? cap : 1, 1);
  while (nc < want) nc *= 2;  // index.cc:29-32 doubles
  if (nc >= 0x7FFFFFFFull) return fail(EHB_ERR_INVALID, "capacity must stay below 2^31 vectors per index");
  CU(vecs.grow(nc * dpad, n * dpad, -1, stream));
  CU(labels.grow(nc, n, -1, stream));
  CU(levels.grow(nc, n, 0, stream));
  CU(deleted.grow(nc, n, 0, stream));
  CU(links0.grow(nc * M0, n * M0, 0xFF, stream));
  CU(up_off.grow(nc, n, 0xFF, stream));
  cap = nc;
  return EHB_OK;
}

int ehb_index::ensure_upper(uint64_t want_rows) {
  if (want_rows <= links_up.n / M && links_up.n) return EHB_OK;
  uint64_t nr = std::max<uint64_t>(links_up.n / M, 64);
  while (nr < want_rows) nr *= 2;
  CU(links_up.grow(nr * M, up_rows * M, 0xFF, stream));
  CU(up_owner.grow(nr, up_rows, 0, stream));
  return EHB_OK;
}

// hnswlib getRandomLevel: (int)(-log(U(0,1)) * 1/ln(M)) drawn from std::default_random_engine(seed), one
// draw per new point in insertion order — the same generator classes upstream uses, so levels match an
// hnswlib built against the same C++ standard library.  Levels are stored in a byte and upper rows are laid
// out per level, so a draw above 31 (probability M^-32) is clamped.
int ehb_index::draw_level() {
  std::uniform_real_distribution<double> u(0.0, 1.0);
  double r = -std::log(u(level_rng)) * (1.0 / std::log((double)M));
  return std::min((int)r, 31);
}

bool ehb_index::find_id(uint64_t label, uint32_t* id) const {
  if (identity_labels) {
    if (label >= n) return false;
    *id = (uint32_t)label;
    return true;
  }
  auto it = lookup.find(label);
  if (it == lookup.end()) return false;
  *id = it->second;
  return true;
}

void ehb_index::reset_content() {
  n = n_linked = up_rows = n_deleted = 0;
  entry = 0;
  max_level = -1;
  lookup.clear();
  h_labels.clear();
  h_levels.clear();
  h_deleted.clear();
  pending_updates.clear();
  identity_labels = true;
  bf16_rows = 0;
}

// ---- ingest ---------------------------------------------------------------------------------------
// Insert-or-update (ANNIndex::set, index.cc:20-37).  Host-side maps are committed only after every device
// operation of the chunk succeeded, so a failed add (OOM while doubling) leaves the index unchanged.
int ehb_index::add_rows(uint64_t cnt, const float* src, bool src_is_device, const uint64_t* lab) {
  if (cnt == 0) return EHB_OK;
  const uint64_t chunk = std::max<uint64_t>(1, (256ull << 20) / (dim * 4));
  for (uint64_t off = 0; off < cnt; off += chunk) {
    const uint64_t m = std::min(chunk, cnt - off);
    std::vector<uint32_t> dst(m);
    std::vector<uint64_t> new_labels;
    std::vector<uint32_t> relink, undelete;
    std::unordered_map<uint64_t, uint32_t> new_map;  // labels first seen in this chunk (non-identity mode)
    const uint64_t first_new = n;
    bool contiguous_new = true;
    uint64_t nn = n;
    for (uint64_t i = 0; i < m; ++i) {
      const uint64_t l = lab ? lab[off + i] : nn;
      bool exists = false;
      uint32_t id = 0;
      if (identity_labels) {
        if (l < nn) {
          exists = true, id = (uint32_t)l;
        } else if (l != nn) {
          // leave identity mode: materialise the map (a consistent state on its own)
          lookup.reserve(std::max<uint64_t>(nn * 2, 1024));
          for (uint64_t j = 0; j < n; ++j) lookup[j] = (uint32_t)j;
          for (uint64_t j = n; j < nn; ++j) new_map[j] = (uint32_t)j;
          identity_labels = false;
        }
      }
      if (!identity_labels && !exists) {
        auto it = lookup.find(l);
        if (it != lookup.end()) {
          exists = true, id = it->second;
        } else {
          auto it2 = new_map.find(l);
          if (it2 != new_map.end()) exists = true, id = it2->second;
        }
      }
      if (exists) {
        dst[i] = id;
        contiguous_new = false;
        if (id < n_linked) relink.push_back(id);  // already in the graph: updatePoint at the next build
        if (id < n && h_deleted[id]) undelete.push_back(id);  // hnswlib addPoint un-deletes a re-added label
      } else {
        dst[i] = (uint32_t)nn;
        if (!identity_labels) new_map[l] = (uint32_t)nn;
        new_labels.push_back(l);
        nn++;
      }
    }
    RET(ensure_capacity(nn));
    // levels + upper rows for the new ids (the generator only advances on commit)
    const uint64_t new_cnt = nn - first_new;
    std::default_random_engine rng_backup = level_rng;
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
    auto device_part = [&]() -> int {
      RET(ensure_upper(rows));
      if (new_cnt) {
        CU(cudaMemcpyAsync(levels.p + first_new, lv.data(), new_cnt, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(up_off.p + first_new, uo.data(), new_cnt * 4, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(labels.p + first_new, new_labels.data(), new_cnt * 8, cudaMemcpyHostToDevice, stream));
        CU(cudaMemsetAsync(deleted.p + first_new, 0, new_cnt, stream));
        if (!owners.empty())
          CU(cudaMemcpyAsync(up_owner.p + up_rows, owners.data(), owners.size() * 4, cudaMemcpyHostToDevice, stream));
      }
      for (uint32_t id : undelete) CU(cudaMemsetAsync(deleted.p + id, 0, 1, stream));
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
      return EHB_OK;
    };
    int rc = device_part();
    if (rc != EHB_OK) {
      level_rng = rng_backup;
      return rc;
    }
    // ---- commit ----
    if (!identity_labels)
      for (auto& kv : new_map) lookup[kv.first] = kv.second;
    h_labels.insert(h_labels.end(), new_labels.begin(), new_labels.end());
    h_levels.insert(h_levels.end(), lv.begin(), lv.end());
    h_deleted.resize(nn, 0);
    for (uint32_t id : undelete)
      if (h_deleted[id]) h_deleted[id] = 0, n_deleted--;
    pending_updates.insert(pending_updates.end(), relink.begin(), relink.end());
    n = nn;
    up_rows = rows;
    bf16_rows = 0;
  }
  return EHB_OK;
}

// hnswlib markDelete (promised by embeddinghub/docs/reading_and_writing_embeddings.md:49-66): the point
// stays in the graph as a tombstone — still traversed, never returned.
int ehb_index::remove_labels(uint64_t cnt, const uint64_t* lab) {
  std::vector<uint32_t> ids(cnt);
  for (uint64_t i = 0; i < cnt; ++i) {
    if (!find_id(lab[i], &ids[i])) return fail(EHB_ERR_NOT_FOUND, "label not found");
    if (h_deleted[ids[i]]) return fail(EHB_ERR_STATE, "the requested to delete element is already deleted");
  }
  {
    std::vector<uint32_t> s = ids;
    std::sort(s.begin(), s.end());
    if (std::adjacent_find(s.begin(), s.end()) != s.end()) return fail(EHB_ERR_INVALID, "label listed twice");
  }
  for (uint32_t id : ids) CU(cudaMemsetAsync(deleted.p + id, 1, 1, stream));
  CU(cudaStreamSynchronize(stream));
  for (uint32_t id : ids) h_deleted[id] = 1;
  n_deleted += cnt;
  bf16_rows = 0;
  return EHB_OK;
}

// ---- construction ------------------------------------------------------------------------------------
int ehb_index::ensure_build_scratch(uint64_t edges, uint32_t batch, bool updates) {
  const uint64_t ecap = edges + 1024;
  CU(b_edge_row.grow(ecap, 0, -1, stream));
  CU(b_edge_src.grow(ecap, 0, -1, stream));
  CU(b_edge_dist.grow(ecap, 0, -1, stream));
  CU(b_touched.grow(ecap, 0, -1, stream));
  CU(b_seg_src.grow(ecap, 0, -1, stream));
  CU(b_seg_dist.grow(ecap, 0, -1, stream));
  CU(b_counters.grow(8, 0, 0, stream));
  if (updates) CU(b_upd_cand.grow((uint64_t)batch * ehb::kUpdCandCap, 0, -1, stream));
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

ehb::BuildBuffers ehb_index::build_buffers(uint64_t edges) {
  ehb::BuildBuffers bb;
  bb.edge_row = b_edge_row.p;
  bb.edge_src = b_edge_src.p;
  bb.edge_dist = b_edge_dist.p;
  bb.edge_count = b_counters.p + 0;
  bb.edge_cap = (uint32_t)std::min<uint64_t>(b_edge_row.n, edges + 1024);
  bb.row_cnt = b_row_cnt.p;
  bb.row_fill = b_row_fill.p;
  bb.row_start = b_row_start.p;
  bb.touched = b_touched.p;
  bb.touched_count = b_counters.p + 1;
  bb.seg_cursor = b_counters.p + 2;
  bb.seg_src = b_seg_src.p;
  bb.seg_dist = b_seg_dist.p;
  bb.error_flag = b_counters.p + 3;
  bb.upd_cand = b_upd_cand.p;
  return bb;
}

ehb::BuildGraph ehb_index::build_graph() const {
  ehb::BuildGraph bg;
  bg.g = view();
  bg.levels = levels.p;
  bg.up_owner = up_owner.p;
  bg.cap = (uint32_t)cap;
  bg.efc = std::max(prm.ef_construction, M);
  return bg;
}

int ehb_index::build() {
  if (!needs_build()) return EHB_OK;
  const uint32_t maxb = prm.build_batch ? prm.build_batch : 16384;
  // a point of level l emits at most M reverse-edge records on each of its l+1 layers
  auto edges_of = [&](uint64_t lo, uint64_t hi) {
    uint64_t e = 0;
    for (uint64_t i = lo; i < hi; ++i) e += (uint64_t)M * (h_levels[i] + 1u);
    return e;
  };
  RET(ensure_build_scratch((uint64_t)std::min<uint64_t>(maxb, std::max<uint64_t>(n, 1)) * M * 2, 1, false));
  CU(cudaMemsetAsync(b_counters.p + 3, 0, 4, stream));  // error flag of earlier builds
  ehb::WalkCfg cfg = walk_cfg(std::max(prm.ef_construction, M), 256, std::min<uint64_t>(maxb, n), 1);
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
    const uint64_t edges = edges_of(n_linked, n_linked + b);
    RET(ensure_build_scratch(edges, 1, false));
    ehb::BuildGraph bg = build_graph();
    ehb::BuildBuffers bb = build_buffers(edges);
    CU(ehb::launch_build_batch(bg, cfg, nullptr, (uint32_t)n_linked, (uint32_t)b, false, bb, wpb, stream));
    for (uint64_t i = n_linked; i < n_linked + b; ++i)
      if ((int)h_levels[i] > max_level) max_level = h_levels[i], entry = (uint32_t)i;
    n_linked += b;
  }
  if (!pending_updates.empty()) {
    // hnswlib updatePoint: neighbour re-selection over the two-hop set, then repairConnectionsForUpdate.
    // Up to kSeqUpdates moved points are processed one at a time, which is exactly the reference's
    // sequential semantics (index_test.cc:39-49); larger bulks go in waves (moved points of one wave do
    // not see each other's new links, like inserted points of one wave).
    std::vector<uint32_t> ups;
    {
      std::vector<uint32_t> sorted_ids = pending_updates;
      std::sort(sorted_ids.begin(), sorted_ids.end());
      sorted_ids.erase(std::unique(sorted_ids.begin(), sorted_ids.end()), sorted_ids.end());
      std::vector<bool> done(sorted_ids.size(), false);  // keep first-occurrence (arrival) order
      for (uint32_t id : pending_updates) {
        size_t pos = std::lower_bound(sorted_ids.begin(), sorted_ids.end(), id) - sorted_ids.begin();
        if (!done[pos]) done[pos] = true, ups.push_back(id);
      }
    }
    constexpr size_t kSeqUpdates = 4096;
    const uint32_t ub = ups.size() <= kSeqUpdates ? 1u : (prm.build_batch ? prm.build_batch : 1024u);
    if (n_linked > 1) {
      for (size_t off = 0; off < ups.size(); off += ub) {
        uint32_t b = (uint32_t)std::min<size_t>(ub, ups.size() - off);
        uint64_t edges = 0;
        for (uint32_t i = 0; i < b; ++i) edges += (uint64_t)M * (h_levels[ups[off + i]] + 1u);
        RET(ensure_build_scratch(edges, b, true));
        CU(b_ids.grow(std::max<uint32_t>(b, 64), 0, -1, stream));
        CU(cudaMemcpyAsync(b_ids.p, ups.data() + off, (size_t)b * 4, cudaMemcpyHostToDevice, stream));
        ehb::BuildGraph bg = build_graph();
        ehb::BuildBuffers bb = build_buffers(edges);
        CU(ehb::launch_build_batch(bg, cfg, b_ids.p, 0, b, true, bb, wpb, stream));
      }
      CU(cudaStreamSynchronize(stream));
    }
    pending_updates.clear();
  }
  uint32_t err = 0;
  CU(cudaMemcpyAsync(&err, b_counters.p + 3, 4, cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  if (err) return fail(EHB_ERR_STATE, "build: edge buffer overflow");
  return EHB_OK;
}

// Searches link pending points lazily; that needs the writer side of the lock.
int ehb_index::ensure_built(std::shared_lock<ehb::RwLock>& lk) {
  while (needs_build()) {
    lk.unlock();
    int rc;
    {
      std::unique_lock<ehb::RwLock> x(rw);
      rc = build();
    }
    lk.lock();
    if (rc != EHB_OK) return rc;
  }
  return EHB_OK;
}

// ---- search slots ------------------------------------------------------------------------------------
int ehb_index::acquire_slot(ehb::SearchSlot** out) {
  std::unique_lock<std::mutex> g(slot_mu);
  for (;;) {
    if (!free_slots.empty()) {
      *out = free_slots.back();  // LIFO: a single-threaded caller keeps reusing one slot (stream-ordered scratch)
      free_slots.pop_back();
      return EHB_OK;
    }
    if (slots.size() < ehb::kMaxSlots) {
      std::unique_ptr<ehb::SearchSlot> sl(new (std::nothrow) ehb::SearchSlot());
      if (!sl) return fail(EHB_ERR_OOM, "host allocation failed");
      cudaError_t e = cudaStreamCreateWithFlags(&sl->stream, cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreate(&sl->ev0);
      if (e == cudaSuccess) e = cudaEventCreate(&sl->ev1);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl->busy, cudaEventDisableTiming);
      if (e != cudaSuccess) return fail(EHB_ERR_CUDA, cudaGetErrorString(e));
      *out = sl.get();
      slots.push_back(std::move(sl));
      return EHB_OK;
    }
    slot_cv.wait(g);
  }
}

void ehb_index::release_slot(ehb::SearchSlot* sl, cudaStream_t used) {
  // the scratch stays in use until the work queued on `used` is done: the next user orders itself after it
  sl->busy_valid = cudaEventRecord(sl->busy, used) == cudaSuccess;
  {
    std::lock_guard<std::mutex> g(slot_mu);
    free_slots.push_back(sl);
  }
  slot_cv.notify_one();
}

// ---- search --------------------------------------------------------------------------------------------
// Caller holds the shared lock, the graph is built, `sl` is acquired.
int ehb_index::search_dev(ehb::SearchSlot* sl, uint64_t nq, const float* dq, uint32_t k, uint32_t ef_in, uint64_t* dl,
                          float* dd, uint32_t* dc, cudaStream_t s, const ehb::ResultSink* sink, bool* pushed) {
  if (pushed) *pushed = false;
  if (k == 0 || nq == 0) return EHB_OK;
  uint32_t ef_eff = std::max(ef_in ? ef_in : ef, k);
  if (ef_eff > ehb::kMaxEf) return fail(EHB_ERR_INVALID, "max(ef, k) must be <= 512");
  // Warps per query (rows <= 1 KB, ef <= 256): four while 3 CTAs of 128 threads per SM hold every query (small
  // online batches; Q=1: 135 us vs 252 us with one warp), two while 7 CTAs of 64 threads do (C2, Q=1000:
  // 0.288 ms vs 0.409 ms), else one warp per query (C5 shape, Q=10k: 9.2 ms vs 10.7 ms with two).
  uint32_t team = t_team;
  if (team == 0) team = nq <= (uint64_t)sms * 3 ? 4 : (nq <= (uint64_t)sms * 7 ? 2 : 1);
  if (dpad > 256 || ef_eff > 256 || n_deleted) team = 1;  // tombstones: the one-warp walk carries the side queue
  ehb::WalkCfg cfg = walk_cfg(ef_eff, 0, nq * team, team);
  if (sl->busy_valid) CU(cudaStreamWaitEvent(s, sl->busy, 0));
  const float* q = dq;
  if (metric == EHB_COSINE) {
    CU(sl->q_norm.grow(nq * dim, 0, -1, s));
    CU(ehb::launch_pad_rows(dq, sl->q_norm.p, nq, dim, dim, true, s));
    q = sl->q_norm.p;
  }
  CU(sl->stats.grow(nq * 4, 0, -1, s));
  CU(sl->stat_sum.grow(4, 0, 0, s));
  uint32_t wpb = wpb_for(cfg, 0);
  CU(cudaEventRecord(sl->ev0, s));
  if (team >= 2)
    CU(ehb::launch_search_team(team, view(), cfg.hash_size, q, (uint32_t)nq, k, ef_eff, dl, dd, dc, sl->stats.p, s));
  else {
    ehb::ResultSink one;
    if (!sink) {
      std::memset(&one, 0, sizeof(one));
      one.labels[0] = dl;
      one.dists[0] = dd;
      one.n = 1;
      sink = &one;
    } else if (pushed) {
      *pushed = true;
    }
    CU(ehb::launch_search(view(), cfg, q, (uint32_t)nq, k, ef_eff, *sink, dc, sl->stats.p, wpb, s));
  }
  CU(cudaEventRecord(sl->ev1, s));
  sl->last_nq = nq;
  {
    const uint32_t kpl = ef_eff <= 64 ? 2 : (ef_eff <= 128 ? 4 : (ef_eff <= 256 ? 8 : 16));
    const uint32_t lpv = dpad > 256 ? 32 : 8;
    if (team >= 2)
      std::snprintf(sl->last_kernel, sizeof(sl->last_kernel), "hnsw_search_team_kernel<NQ=%u,KPL=%u,T=%u>",
                    dpad / (4 * lpv), kpl, team);
    else
      std::snprintf(sl->last_kernel, sizeof(sl->last_kernel), "%s<LPV=%u,NQ=%u,KPL=%u>",
                    cfg.dense ? "hnsw_search_dense_kernel" : "hnsw_search_kernel", lpv, dpad / (4 * lpv), kpl);
  }
  {
    std::lock_guard<std::mutex> g(last_mu);
    last_slot = sl;
    last_was_brute = false;
    timed = true;
    last_sum_valid = false;
  }
  return EHB_OK;
}

// Caller holds the shared lock and bf_mu.
int ehb_index::bruteforce_dev(uint64_t nq, const float* dq, uint32_t k, int precision, uint64_t* dl, float* dd,
                              uint32_t* dc, cudaStream_t s) {
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
  sc.deleted = n_deleted ? deleted.p : nullptr;
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
    // fused selection state (option "bf16_unfused" keeps the distance tiles in HBM: A/B switch); tombstones
    // are filtered where keys are formed, which only the unfused selection does
    bctx.fused = !o_bf16_unfused && !n_deleted;
    bctx.variant = o_gemm_2cta ? 1 : 0;
    bctx.ccap = 2 * kc + 64;
    CU(bf_thr.grow(nq, 0, -1, s));
    CU(bf_cbuf.grow(nq * bctx.ccap, 0, -1, s));
    CU(bf_ccount.grow(nq + 1, 0, 0, s));
    bctx.thr = bf_thr.p;
    bctx.cbuf = bf_cbuf.p;
    bctx.ccount = bf_ccount.p;
    bctx.overflow = bf_ccount.p + nq;
    bctx.sms = sms;
  }
  CU(cudaEventRecord(bf_ev0, s));
  CU(ehb::launch_bruteforce(vecs.p, dpad, dim, n, labels.p, metric == EHB_L2 ? 0 : 1, bf_qpad.p, nq, k, sc,
                            bf16 ? &bctx : nullptr, dl, dd, dc, s));
  CU(cudaEventRecord(bf_ev1, s));
  {
    std::lock_guard<std::mutex> g(last_mu);
    last_was_brute = true;
    timed = true;
  }
  return EHB_OK;
}

// ============================================================================================
// C ABI
// ============================================================================================
#define ENTER_X(ix)                                                    \
  if (!(ix)) return fail(EHB_ERR_INVALID, "null index handle");        \
  std::unique_lock<ehb::RwLock> _g((ix)->rw);                    \
  CU(cudaSetDevice((ix)->device))
#define ENTER_S(ix)                                                    \
  if (!(ix)) return fail(EHB_ERR_INVALID, "null index handle");        \
  std::shared_lock<ehb::RwLock> _g((ix)->rw);                    \
  CU(cudaSetDevice((ix)->device))

namespace {

// One host graph search on its own slot: H2D (from `q`, host), walk, D2H, synchronise.
int search_host_direct(ehb_index* ix, uint64_t nq, const float* q, uint32_t k, uint32_t ef, uint64_t* ol, float* od,
                       uint32_t* oc) {
  ehb::SearchSlot* sl = nullptr;
  RET(ix->acquire_slot(&sl));
  cudaStream_t s = sl->stream;
  auto run = [&]() -> int {
    if (sl->busy_valid) CU(cudaStreamWaitEvent(s, sl->busy, 0));
    CU(sl->q_in.grow(nq * ix->dim, 0, -1, s));
    CU(sl->o_labels.grow(nq * k, 0, -1, s));
    CU(sl->o_dists.grow(nq * k, 0, -1, s));
    CU(sl->o_counts.grow(nq, 0, -1, s));
    CU(cudaMemcpyAsync(sl->q_in.p, q, nq * ix->dim * 4, cudaMemcpyHostToDevice, s));
    RET(ix->search_dev(sl, nq, sl->q_in.p, k, ef, sl->o_labels.p, sl->o_dists.p, sl->o_counts.p, s));
    CU(cudaMemcpyAsync(ol, sl->o_labels.p, nq * k * 8, cudaMemcpyDeviceToHost, s));
    if (od) CU(cudaMemcpyAsync(od, sl->o_dists.p, nq * k * 4, cudaMemcpyDeviceToHost, s));
    if (oc) CU(cudaMemcpyAsync(oc, sl->o_counts.p, nq * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return EHB_OK;
  };
  int rc = run();
  ix->release_slot(sl, s);
  return rc;
}

// The leader's part of the combining queue: all taken requests share (k, ef) and go out as one launch.
// (Every queued caller holds the shared lock and has seen the graph built, so no mutation can intervene.)
int run_combined(ehb_index* ix, std::vector<ehb::CombineReq*>& batch) {
  uint64_t tot = 0;
  for (auto* r : batch) tot += r->nq;
  const uint32_t k = batch[0]->k, ef = batch[0]->ef, dim = ix->dim;
  ehb::SearchSlot* sl = nullptr;
  RET(ix->acquire_slot(&sl));
  cudaStream_t s = sl->stream;
  auto run = [&]() -> int {
    if (sl->busy_valid) CU(cudaStreamWaitEvent(s, sl->busy, 0));
    CU(sl->h_q.reserve(tot * dim * 4));
    CU(sl->h_l.reserve(tot * k * 8));
    CU(sl->h_d.reserve(tot * k * 4));
    CU(sl->h_c.reserve(tot * 4));
    CU(sl->q_in.grow(tot * dim, 0, -1, s));
    CU(sl->o_labels.grow(tot * k, 0, -1, s));
    CU(sl->o_dists.grow(tot * k, 0, -1, s));
    CU(sl->o_counts.grow(tot, 0, -1, s));
    uint64_t off = 0;
    for (auto* r : batch) {
      std::memcpy(sl->h_q.p + off * dim * 4, r->q, r->nq * dim * 4);
      off += r->nq;
    }
    CU(cudaMemcpyAsync(sl->q_in.p, sl->h_q.p, tot * dim * 4, cudaMemcpyHostToDevice, s));
    RET(ix->search_dev(sl, tot, sl->q_in.p, k, ef, sl->o_labels.p, sl->o_dists.p, sl->o_counts.p, s));
    CU(cudaMemcpyAsync(sl->h_l.p, sl->o_labels.p, tot * k * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(sl->h_d.p, sl->o_dists.p, tot * k * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(sl->h_c.p, sl->o_counts.p, tot * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    off = 0;
    for (auto* r : batch) {
      std::memcpy(r->ol, sl->h_l.p + off * k * 8, r->nq * k * 8);
      if (r->od) std::memcpy(r->od, sl->h_d.p + off * k * 4, r->nq * k * 4);
      if (r->oc) std::memcpy(r->oc, sl->h_c.p + off * 4, r->nq * 4);
      off += r->nq;
    }
    return EHB_OK;
  };
  int rc = run();
  ix->release_slot(sl, s);
  ix->combined_batches++;
  ix->combined_queries += tot;
  return rc;
}

// Combining queue ("group commit"): a caller queues its request; whoever finds a free leader seat takes
// every queued request with its own (k, ef) and runs them as ONE batched search.  Nobody ever waits on a
// timer: requests pile up only while earlier batches occupy the leader seats, which is exactly when
// batching pays.  A lone caller becomes its own leader at once.
int search_host_combined(ehb_index* ix, uint64_t nq, const float* q, uint32_t k, uint32_t ef, uint64_t* ol, float* od,
                         uint32_t* oc) {
  ehb::CombineReq me{q, nq, k, ef, ol, od, oc};
  std::unique_lock<std::mutex> g(ix->cq_mu);
  ix->cq.push_back(&me);
  for (;;) {
    if (me.done) break;
    if (!me.taken && ix->cq_leaders < ehb::kCombineLeaders) {
      std::vector<ehb::CombineReq*> batch;
      uint64_t tot = 0;
      for (auto it = ix->cq.begin(); it != ix->cq.end();) {
        ehb::CombineReq* r = *it;
        if (r->k == k && r->ef == ef && (r == &me || tot + r->nq <= ehb::kCombineMaxBatch)) {
          r->taken = true;
          tot += r->nq;
          batch.push_back(r);
          it = ix->cq.erase(it);
        } else {
          ++it;
        }
      }
      ix->cq_leaders++;
      g.unlock();
      int rc = run_combined(ix, batch);
      const std::string err = rc == EHB_OK ? std::string() : ehb::last_error_text();
      g.lock();
      ix->cq_leaders--;
      for (auto* r : batch) r->rc = rc, r->err = err, r->done = true;
      ix->cq_cv.notify_all();
      continue;
    }
    ix->cq_cv.wait(g);
  }
  if (me.rc != EHB_OK) return fail(me.rc, me.err);
  return EHB_OK;
}

}  // namespace

extern "C" {

const char* ehb_last_error(void) { return ehb::g_err.c_str(); }
uint32_t ehb_abi_version(void) { return 2; }

int ehb_device_count(int32_t* out) {
  if (!out) return fail(EHB_ERR_INVALID, "null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out = 0;
    return fail(EHB_ERR_CUDA, std::string("no CUDA device (ehb200 has no CPU fallback): ") + cudaGetErrorString(e));
  }
  *out = n;
  return EHB_OK;
}

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
  ix->sms = prop.multiProcessorCount;
  ix->ef = p->ef_search ? p->ef_search : 10;
  ix->level_rng.seed((unsigned)p->seed);
  if (ix->prm.ef_construction == 0) ix->prm.ef_construction = 200;
  cudaError_t ce = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreate(&ix->bf_ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&ix->bf_ev1);
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
  cudaDeviceSynchronize();
  delete ix;
  return EHB_OK;
}

int ehb_index_add(ehb_index* ix, uint64_t n, const float* vecs, const uint64_t* labels) {
  ENTER_X(ix);
  if (n && !vecs) return fail(EHB_ERR_INVALID, "null vectors");
  return ix->add_rows(n, vecs, false, labels);
}
int ehb_index_add_dev(ehb_index* ix, uint64_t n, const float* vecs_dev, const uint64_t* labels) {
  ENTER_X(ix);
  if (n && !vecs_dev) return fail(EHB_ERR_INVALID, "null vectors");
  return ix->add_rows(n, vecs_dev, true, labels);
}
int ehb_index_remove(ehb_index* ix, uint64_t n, const uint64_t* labels) {
  ENTER_X(ix);
  if (n && !labels) return fail(EHB_ERR_INVALID, "null labels");
  return ix->remove_labels(n, labels);
}
int ehb_index_build(ehb_index* ix) {
  ENTER_X(ix);
  RET(ix->build());
  CU(cudaStreamSynchronize(ix->stream));
  return EHB_OK;
}
int ehb_index_set_ef(ehb_index* ix, uint32_t ef) {
  ENTER_X(ix);
  if (ef == 0) return fail(EHB_ERR_INVALID, "ef must be > 0");
  ix->ef = ef;
  return EHB_OK;
}
int ehb_index_size(ehb_index* ix, uint64_t* out) {
  ENTER_S(ix);
  if (!out) return fail(EHB_ERR_INVALID, "null out");
  *out = ix->n;
  return EHB_OK;
}

int ehb_index_get(ehb_index* ix, uint64_t label, float* out) {
  ENTER_S(ix);
  if (!out) return fail(EHB_ERR_INVALID, "null out");
  uint32_t id;
  // hnswlib getDataByLabel: a tombstoned label reads as "Label not found"
  if (!ix->find_id(label, &id) || ix->h_deleted[id]) return fail(EHB_ERR_NOT_FOUND, "label not found");
  CU(cudaMemcpy(out, ix->vecs.p + (uint64_t)id * ix->dpad, ix->dim * 4, cudaMemcpyDeviceToHost));
  return EHB_OK;
}

int ehb_index_search(ehb_index* ix, uint64_t nq, const float* q, uint32_t k, uint32_t ef, uint64_t* ol, float* od,
                     uint32_t* oc) {
  ENTER_S(ix);
  if (nq && (!q || !ol)) return fail(EHB_ERR_INVALID, "null buffer");
  if (k == 0 || nq == 0) return EHB_OK;
  if (std::max(ef ? ef : ix->ef, k) > ehb::kMaxEf) return fail(EHB_ERR_INVALID, "max(ef, k) must be <= 512");
  RET(ix->ensure_built(_g));  // before queueing: a waiting follower must never block a writer the leader needs
  if (ix->o_combine && nq <= ehb::kCombineMaxCall)
    return search_host_combined(ix, nq, q, k, ef ? ef : ix->ef, ol, od, oc);
  return search_host_direct(ix, nq, q, k, ef, ol, od, oc);
}

int ehb_index_search_dev(ehb_index* ix, uint64_t nq, const float* dq, uint32_t k, uint32_t ef, uint64_t* dl, float* dd,
                         uint32_t* dc, void* stream) {
  ENTER_S(ix);
  if (nq && (!dq || !dl)) return fail(EHB_ERR_INVALID, "null buffer");
  if (k == 0 || nq == 0) return EHB_OK;
  RET(ix->ensure_built(_g));
  ehb::SearchSlot* sl = nullptr;
  RET(ix->acquire_slot(&sl));
  cudaStream_t s = stream ? (cudaStream_t)stream : sl->stream;
  int rc = ix->search_dev(sl, nq, dq, k, ef, dl, dd, dc, s);
  ix->release_slot(sl, s);
  return rc;
}

}  // extern "C"

int ehb_index_search_dev_sink(ehb_index* ix, uint64_t nq, const float* dq, uint32_t k, uint32_t ef,
                              const ehb::ResultSink* sink, uint32_t* dc, cudaStream_t stream, bool* pushed) {
  ENTER_S(ix);
  if (!dq || !sink || !sink->n || !sink->labels[0]) return fail(EHB_ERR_INVALID, "null buffer");
  if (k == 0 || nq == 0) return EHB_OK;
  RET(ix->ensure_built(_g));
  ehb::SearchSlot* sl = nullptr;
  RET(ix->acquire_slot(&sl));
  cudaStream_t s = stream ? stream : sl->stream;
  int rc = ix->search_dev(sl, nq, dq, k, ef, sink->labels[0], sink->dists[0], dc, s, sink, pushed);
  ix->release_slot(sl, s);
  return rc;
}

extern "C" {

int ehb_index_search_bruteforce(ehb_index* ix, uint64_t nq, const float* q, uint32_t k, int precision, uint64_t* ol,
                                float* od, uint32_t* oc) {
  ENTER_S(ix);
  if (nq && (!q || !ol)) return fail(EHB_ERR_INVALID, "null buffer");
  if (k == 0 || nq == 0) return EHB_OK;
  std::lock_guard<std::mutex> bg(ix->bf_mu);
  cudaStream_t s = ix->stream;
  CU(ix->bf_q_in.grow(nq * ix->dim, 0, -1, s));
  CU(ix->bf_o_labels.grow(nq * k, 0, -1, s));
  CU(ix->bf_o_dists.grow(nq * k, 0, -1, s));
  CU(ix->bf_o_counts.grow(nq, 0, -1, s));
  CU(cudaMemcpyAsync(ix->bf_q_in.p, q, nq * ix->dim * 4, cudaMemcpyHostToDevice, s));
  RET(ix->bruteforce_dev(nq, ix->bf_q_in.p, k, precision, ix->bf_o_labels.p, ix->bf_o_dists.p, ix->bf_o_counts.p, s));
  CU(cudaMemcpyAsync(ol, ix->bf_o_labels.p, nq * k * 8, cudaMemcpyDeviceToHost, s));
  if (od) CU(cudaMemcpyAsync(od, ix->bf_o_dists.p, nq * k * 4, cudaMemcpyDeviceToHost, s));
  if (oc) CU(cudaMemcpyAsync(oc, ix->bf_o_counts.p, nq * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return EHB_OK;
}

int ehb_index_search_bruteforce_dev(ehb_index* ix, uint64_t nq, const float* dq, uint32_t k, int precision,
                                    uint64_t* dl, float* dd, uint32_t* dc, void* stream) {
  ENTER_S(ix);
  if (nq && (!dq || !dl)) return fail(EHB_ERR_INVALID, "null buffer");
  std::lock_guard<std::mutex> bg(ix->bf_mu);
  return ix->bruteforce_dev(nq, dq, k, precision, dl, dd, dc, stream ? (cudaStream_t)stream : ix->stream);
}

int ehb_index_stats(ehb_index* ix, ehb_stats* out) {
  ENTER_S(ix);
  if (!out) return fail(EHB_ERR_INVALID, "null out");
  std::memset(out, 0, sizeof(*out));
  {
    std::lock_guard<std::mutex> g(ix->last_mu);
    ehb::SearchSlot* sl = ix->last_slot;
    if (sl && !ix->last_was_brute && sl->last_nq) {
      if (!ix->last_sum_valid) {
        CU(cudaEventSynchronize(sl->ev1));
        CU(ehb::launch_sum_stats(sl->stats.p, (uint32_t)sl->last_nq, sl->stat_sum.p, ix->stream));
        CU(cudaMemcpyAsync(ix->last_sum, sl->stat_sum.p, 32, cudaMemcpyDeviceToHost, ix->stream));
        CU(cudaStreamSynchronize(ix->stream));
        ix->last_sum_valid = true;
      }
      out->queries = sl->last_nq;
      out->hops_upper = ix->last_sum[0];
      out->hops_base = ix->last_sum[1];
      out->dist_evals = ix->last_sum[2];
      out->visited_overflow = ix->last_sum[3];
      out->algorithmic_bytes = out->hops_upper * 4ull * ix->M + out->hops_base * 4ull * ix->M0 +
                               out->dist_evals * 4ull * ix->dim + out->queries * 4ull * ix->dim;
    }
  }
  out->size = ix->n;
  out->capacity = ix->cap;
  out->upper_rows = ix->up_rows;
  out->dim = ix->dim;
  out->M = ix->M;
  out->max_level = ix->max_level < 0 ? 0 : (uint32_t)ix->max_level;
  out->entry_point = ix->entry;
  out->device_bytes = ix->vecs.bytes() + ix->labels.bytes() + ix->levels.bytes() + ix->deleted.bytes() +
                      ix->links0.bytes() + ix->up_off.bytes() + ix->links_up.bytes() + ix->up_owner.bytes();
  out->deleted = ix->n_deleted;
  out->combined_batches = ix->combined_batches.load();
  out->combined_queries = ix->combined_queries.load();
  out->metric = (uint32_t)ix->metric;
  return EHB_OK;
}

int ehb_index_last_kernel_ms(ehb_index* ix, float* out_ms) {
  ENTER_S(ix);
  if (!out_ms) return fail(EHB_ERR_INVALID, "null out");
  std::lock_guard<std::mutex> g(ix->last_mu);
  if (!ix->timed) return fail(EHB_ERR_STATE, "no search has been timed yet");
  cudaEvent_t e0 = ix->last_was_brute ? ix->bf_ev0 : ix->last_slot->ev0;
  cudaEvent_t e1 = ix->last_was_brute ? ix->bf_ev1 : ix->last_slot->ev1;
  CU(cudaEventSynchronize(e1));
  CU(cudaEventElapsedTime(out_ms, e0, e1));
  return EHB_OK;
}

int ehb_index_last_kernel_name(ehb_index* ix, char* out, uint32_t out_bytes) {
  ENTER_S(ix);
  if (!out || !out_bytes) return fail(EHB_ERR_INVALID, "null out");
  std::lock_guard<std::mutex> g(ix->last_mu);
  const char* name = !ix->timed ? "" : (ix->last_was_brute ? "bruteforce" : ix->last_slot->last_kernel);
  std::snprintf(out, out_bytes, "%s", name);
  return EHB_OK;
}

int ehb_index_set_search_width(ehb_index* ix, uint32_t warps_per_query) {
  ENTER_X(ix);
  if (warps_per_query > 4) return fail(EHB_ERR_INVALID, "warps_per_query must be 0 (auto) or 1..4");
  ix->t_team = warps_per_query;
  return EHB_OK;
}

int ehb_index_set_option(ehb_index* ix, const char* name, int64_t value) {
  ENTER_X(ix);
  if (!name) return fail(EHB_ERR_INVALID, "null option name");
  const std::string o(name);
  if (o == "build_frac") {
    if (value < 0 || value > (1 << 20)) return fail(EHB_ERR_INVALID, "build_frac must be in 0..2^20");
    ix->o_build_frac = (uint32_t)value;
  } else if (o == "bf16_unfused") {
    ix->o_bf16_unfused = value != 0;
  } else if (o == "gemm_2cta") {
    ix->o_gemm_2cta = value != 0;
  } else if (o == "walk_prefetch") {
    ix->o_walk_prefetch = value != 0;
  } else if (o == "combine") {
    ix->o_combine = value != 0;
  } else {
    return fail(EHB_ERR_INVALID, "unknown option: " + o);
  }
  return EHB_OK;
}

int ehb_index_set_tuning(ehb_index* ix, uint32_t slots, uint32_t groups, uint32_t hash_bits, uint32_t wpb) {
  ENTER_X(ix);
  ix->t_slots = slots;
  ix->t_groups = groups;
  ix->t_hash_bits = hash_bits;
  ix->t_wpb = wpb;
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
