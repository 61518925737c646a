// Internal state of an ehb_index (shared by api.cu, io.cu and exchange.cu).  Mirrors the
// responsibilities of featureform::embedding::ANNIndex + hnswlib::HierarchicalNSW as used in
// embeddinghub/embeddingstore/index.cc:10-52.
//
// Concurrency ("searches re-entrant ... mutations exclusive").  The reference
// serialises every RPC under one service mutex (embeddinghub/embeddingstore/server.cc:175); here
//   * `rw` is a reader/writer lock: searches share it, mutations (add / remove / build / import) own it;
//   * every in-flight graph search works on a SearchSlot (own stream, own device scratch, own pinned
//     staging) taken from a small pool, so host threads never share scratch;
//   * concurrent small host searches are coalesced into one batched launch by the combining queue
//     (api.cu) — the cgo pattern of serving/serving.go:744-771: one goroutine, one query, per request.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <random>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ehb200.h"
#include "kernels.h"

namespace ehb {

int fail(int code, const std::string& msg);  // sets the thread-local error text, returns code
const std::string& last_error_text();
extern thread_local std::string g_err;

#define CU(expr)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return ehb::fail(_e == cudaErrorMemoryAllocation ? EHB_ERR_OOM : EHB_ERR_CUDA,                     \
                       std::string(#expr) + ": " + cudaGetErrorString(_e));                              \
  } while (0)
#define RET(expr)                \
  do {                           \
    int _r = (expr);             \
    if (_r != EHB_OK) return _r; \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
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

// page-locked host staging
struct PinBuf {
  unsigned char* p = nullptr;
  size_t n = 0;
  PinBuf() = default;
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= n) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    size_t want = std::max<size_t>(bytes, 4096);
    cudaError_t e = cudaMallocHost((void**)&p, want);
    if (e == cudaSuccess) n = want;
    return e;
  }
};

// Reader/writer lock that prefers writers: once a mutation waits, new searches queue behind it, so a stream
// of overlapping searches can never starve an add (glibc's rwlock, which std::shared_mutex wraps, prefers
// readers by default).  Meets the SharedMutex requirements used by std::shared_lock / std::unique_lock.
class RwLock {
 public:
  void lock() {
    std::unique_lock<std::mutex> g(mu_);
    ++writers_waiting_;
    cv_.wait(g, [&] { return !writer_ && readers_ == 0; });
    --writers_waiting_;
    writer_ = true;
  }
  void unlock() {
    {
      std::lock_guard<std::mutex> g(mu_);
      writer_ = false;
    }
    cv_.notify_all();
  }
  void lock_shared() {
    std::unique_lock<std::mutex> g(mu_);
    cv_.wait(g, [&] { return !writer_ && writers_waiting_ == 0; });
    ++readers_;
  }
  void unlock_shared() {
    bool wake;
    {
      std::lock_guard<std::mutex> g(mu_);
      wake = --readers_ == 0;
    }
    if (wake) cv_.notify_all();
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  uint32_t readers_ = 0, writers_waiting_ = 0;
  bool writer_ = false;
};

// Everything one in-flight graph search needs.
struct SearchSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, busy = nullptr;
  bool busy_valid = false;
  DevBuf<float> q_in, q_norm, o_dists;
  DevBuf<uint64_t> o_labels;
  DevBuf<uint32_t> o_counts, stats;
  DevBuf<unsigned long long> stat_sum;
  PinBuf h_q, h_l, h_d, h_c;
  uint64_t last_nq = 0;
  char last_kernel[64] = {0};  // name of the graph-walk kernel of the most recent search on this slot
  ~SearchSlot() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (busy) cudaEventDestroy(busy);
    if (stream) cudaStreamDestroy(stream);
  }
};

// One pending host search in the combining queue.
struct CombineReq {
  const float* q;
  uint64_t nq;
  uint32_t k, ef;
  uint64_t* ol;
  float* od;
  uint32_t* oc;
  int rc = EHB_OK;
  std::string err;
  bool taken = false, done = false;
};

constexpr uint32_t kMaxSlots = 4;          // graph searches in flight per index
constexpr uint32_t kCombineMaxCall = 256;  // host searches up to this many queries go through the combining queue
constexpr uint32_t kCombineMaxBatch = 8192;
constexpr uint32_t kCombineLeaders = 2;    // batches in flight (copies of one overlap the walk of the other)
constexpr uint32_t kDeletedQueue = 64;     // side queue of tombstoned candidates per walking warp

}  // namespace ehb

// ehb_index_search_dev with a result sink (exchange.cu)
int ehb_index_search_dev_sink(ehb_index* ix, uint64_t nq, const float* dq, uint32_t k, uint32_t ef,
                              const ehb::ResultSink* sink, uint32_t* dc, cudaStream_t stream, bool* pushed);

struct ehb_index {
  ehb_params prm;
  uint32_t dim, dpad, M, M0;
  int metric;
  int device;
  int sms = 148;
  cudaStream_t stream = nullptr;  // mutation / construction / brute-force stream
  cudaEvent_t bf_ev0 = nullptr, bf_ev1 = nullptr;
  ehb::RwLock rw;

  uint64_t cap = 0;        // vector capacity
  uint64_t n = 0;          // stored vectors (tombstones included, like hnswlib cur_element_count)
  uint64_t n_linked = 0;   // vectors linked into the graph
  uint64_t up_rows = 0;    // used upper rows
  uint64_t n_deleted = 0;  // tombstones
  uint32_t entry = 0;
  int32_t max_level = -1;
  uint32_t ef;

  ehb::DevBuf<float> vecs;
  ehb::DevBuf<uint64_t> labels;
  ehb::DevBuf<uint8_t> levels, deleted;
  ehb::DevBuf<uint32_t> links0, up_off, links_up, up_owner;

  std::vector<uint8_t> h_levels, h_deleted;
  std::vector<uint64_t> h_labels;
  bool identity_labels = true;
  std::unordered_map<uint64_t, uint32_t> lookup;
  std::vector<uint32_t> pending_updates;

  // search slots
  std::mutex slot_mu;
  std::condition_variable slot_cv;
  std::vector<std::unique_ptr<ehb::SearchSlot>> slots;
  std::vector<ehb::SearchSlot*> free_slots;
  // what ehb_index_stats / ehb_index_last_kernel_ms report: the most recently issued search
  std::mutex last_mu;
  ehb::SearchSlot* last_slot = nullptr;  // graph search (counters + events)
  bool last_was_brute = false, timed = false;
  unsigned long long last_sum[4] = {0, 0, 0, 0};
  bool last_sum_valid = false;

  // combining queue of small host searches
  std::mutex cq_mu;
  std::condition_variable cq_cv;
  std::deque<ehb::CombineReq*> cq;
  uint32_t cq_leaders = 0;
  std::atomic<uint64_t> combined_batches{0}, combined_queries{0};

  // brute-force scratch (one brute-force search at a time: bf_mu)
  std::mutex bf_mu;
  ehb::DevBuf<float> bf_q_in, bf_o_dists, bf_dist, bf_qpad;
  ehb::DevBuf<uint64_t> bf_o_labels, bf_part, bf_run;
  ehb::DevBuf<uint32_t> bf_o_counts;
  ehb::DevBuf<uint16_t> x_bf16, q_bf16;   // bf16 shadows for the tensor-core path
  ehb::DevBuf<float> x_norm, q_norm2, bf_thr;
  ehb::DevBuf<uint64_t> bf_cbuf;
  ehb::DevBuf<uint32_t> bf_ccount;
  uint64_t bf16_rows = 0;            // rows of x_bf16 that are current (0 = stale)

  // build scratch
  ehb::DevBuf<uint32_t> b_edge_row, b_edge_src, b_row_cnt, b_row_fill, b_row_start, b_touched, b_seg_src, b_counters,
      b_ids, b_upd_cand;
  ehb::DevBuf<float> b_edge_dist, b_seg_dist, b_stage_in;

  // tuning (0 = auto)
  uint32_t t_slots = 0, t_groups = 0, t_hash_bits = 0, t_wpb = 0, t_team = 0;
  // options (ehb_index_set_option)
  uint32_t o_build_frac = 0;     // a wave links at most n_linked / build_frac points (0 = 64)
  bool o_bf16_unfused = false;   // bf16 brute force: keep the distance tiles in HBM (A/B)
  bool o_gemm_2cta = false;      // bf16 brute force: cta_group::2 cluster form of the fused GEMM
  bool o_combine = true;         // coalesce concurrent small host searches
  // L2 prefetch of the speculated next hop's vectors (rows <= 1 KB).  Off: on the full C5 shard ncu showed 52.7 GB
  // of DRAM traffic for 41.4 GB algorithmic (already-visited neighbours and wrong guesses are fetched too) and
  // the walk is 5.9 % faster without it (10.43 -> 9.82 ms); r1 had measured no gain at C2 either.
  bool o_walk_prefetch = false;

  std::default_random_engine level_rng;

  ~ehb_index();

  ehb::GraphView view() const;
  ehb::WalkCfg walk_cfg(uint32_t ef_eff, uint32_t smem_list, uint64_t jobs, uint32_t team) const;
  uint32_t wpb_for(const ehb::WalkCfg& c, uint32_t extra) const;
  int ensure_capacity(uint64_t want);
  int ensure_upper(uint64_t want_rows);
  int draw_level();
  bool find_id(uint64_t label, uint32_t* id) const;
  int add_rows(uint64_t cnt, const float* src, bool src_is_device, const uint64_t* lab);
  int remove_labels(uint64_t cnt, const uint64_t* lab);
  int ensure_build_scratch(uint64_t edges, uint32_t batch, bool updates);
  ehb::BuildBuffers build_buffers(uint64_t edges);
  ehb::BuildGraph build_graph() const;
  int build();
  bool needs_build() const { return n_linked != n || !pending_updates.empty(); }
  int ensure_built(std::shared_lock<ehb::RwLock>& lk);
  int acquire_slot(ehb::SearchSlot** out);
  void release_slot(ehb::SearchSlot* sl, cudaStream_t used);
  // sink (optional): extra destinations + slice flags for the sharded exchange; *pushed tells whether the
  // launched kernel honoured it (the one-warp walk does, the team walk does not)
  int search_dev(ehb::SearchSlot* sl, uint64_t nq, const float* dq, uint32_t k, uint32_t ef_in, uint64_t* dl, float* dd,
                 uint32_t* dc, cudaStream_t s, const ehb::ResultSink* sink = nullptr, bool* pushed = nullptr);
  int bruteforce_dev(uint64_t nq, const float* dq, uint32_t k, int precision, uint64_t* dl, float* dd, uint32_t* dc,
                     cudaStream_t s);
  void reset_content();
};
