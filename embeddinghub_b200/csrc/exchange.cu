// Range-sharded search behind the C ABI (SURVEY.md §8b B4 "device_ids[], n_dev", §8e).
//
// The index is partitioned by label range; every GPU owns an independent graph; every GPU searches all
// queries; the per-shard top-k lists meet in ONE exchange step and are merged.  Two deployment shapes:
//
//  * ehb_sharded  — one process drives n_dev GPUs (what a C++ ANNIndex or the cgo provider links against).
//    Peer access is enabled between the devices and every shard's search kernels write their top-k
//    STRAIGHT INTO DEVICE 0's gather buffer (stores over NVLink from inside the walk kernel); device 0's
//    merge kernel is ordered after them with events.  No collective, no staging copy.
//
//  * ehb_exchange — one process per GPU (torchrun / MPI style).  Each rank owns a receive buffer
//    [2 parities][world][block] + flags, exported with CUDA IPC and mapped by every peer.  A step is ONE
//    kernel per rank (exchange_merge_kernel): phase 1 pushes this rank's block, slice by slice, into every
//    peer's buffer with coalesced stores over NVLink and raises a per-(rank, slice) flag with a
//    system-scope release; phase 2 waits (acquire) for the flags of each slice and merges the G lists of
//    its queries.  This replaces ncclAllGather + merge kernel: no collective launch, the merge of early
//    slices overlaps the transfer of late ones, and flags are epoch-numbered with parity double
//    buffering so consecutive steps need no barrier.
#include <thread>

#include "index_impl.h"
#include "merge.cuh"

using ehb::fail;

namespace ehb {

constexpr uint32_t kMaxWorld = 16;
constexpr uint32_t kMaxSlices = 256;

struct ExchangeView {
  unsigned char* recv[kMaxWorld];  // recv buffer of every rank as mapped HERE ([2][world][stride])
  uint32_t* flags[kMaxWorld];      // flag array of every rank ([2][world][kMaxSlices])
  uint32_t world, rank;
  uint64_t stride;                 // bytes per rank block
};

// One persistent launch per rank and step; grid <= resident capacity so no CTA waits on an unscheduled one.
// 1024 threads per CTA: the push is a plain copy and the merge is one latency-bound warp per query, so both
// want as many warps per SM as one resident CTA can hold.
constexpr uint32_t kExchangeThreads = 1024;
__global__ void __launch_bounds__(kExchangeThreads) exchange_merge_kernel(ExchangeView ev, uint32_t parity, uint32_t epoch,
                                                             uint64_t nq, uint32_t k, uint32_t qs, uint32_t nslices,
                                                             float* __restrict__ out_dists,
                                                             uint64_t* __restrict__ out_labels,
                                                             uint32_t* __restrict__ out_counts,
                                                             uint32_t* __restrict__ timeout_flag, uint32_t skip_push) {
  const uint32_t W = ev.world, me = ev.rank;
  const uint64_t blk = ((uint64_t)parity * W + me) * ev.stride;  // my block inside ANY rank's buffer
  const unsigned char* mine = ev.recv[me] + blk;                 // written by my search kernels
  const uint64_t lab_bytes = nq * k * 8ull;
  // ---- phase 1: push my slices to every peer, then raise their flags ------------------------------------
  // (skipped when the producer was the one-warp walk: its epilogue already stored every query's results into
  //  the peers' buffers and raised the slice flags — search_impl.cuh)
  for (uint32_t s = blockIdx.x; s < nslices && !skip_push; s += gridDim.x) {
    const uint64_t q0 = (uint64_t)s * qs, q1 = min(nq, q0 + qs);
    const uint64_t e0 = q0 * k, e1 = q1 * k;  // element range of the slice
    const uint64_t* src_l = (const uint64_t*)mine;
    const float* src_d = (const float*)(mine + lab_bytes);
    for (uint32_t g = 0; g < W; ++g) {
      if (g == me) continue;
      uint64_t* dst_l = (uint64_t*)(ev.recv[g] + blk);
      float* dst_d = (float*)(ev.recv[g] + blk + lab_bytes);
      for (uint64_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) dst_l[i] = src_l[i];
      for (uint64_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) dst_d[i] = src_d[i];
    }
    __threadfence_system();  // every thread's stores are ordered before the flags below
    __syncthreads();
    if (threadIdx.x < W && threadIdx.x != me)
      st_release_sys(ev.flags[threadIdx.x] + ((uint64_t)parity * W + me) * kMaxSlices + s, epoch);
    __syncthreads();
  }
  // ---- phase 2: wait for each slice from every peer, merge its queries -------------------------------------
  const unsigned char* base = ev.recv[me] + (uint64_t)parity * W * ev.stride;
  const uint32_t warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t s = blockIdx.x; s < nslices; s += gridDim.x) {
    if (threadIdx.x < W && threadIdx.x != me) {
      const uint32_t* f = ev.flags[me] + ((uint64_t)parity * W + threadIdx.x) * kMaxSlices + s;
      // bounded (~20 s): a peer that never arrives must not wedge the GPU; the host reports the flag
      uint32_t spins = 0;
      while (ld_acquire_sys(f) != epoch) {
        __nanosleep(64);
        if (++spins > (1u << 28)) {
          atomicExch(timeout_flag, 1u);
          break;
        }
      }
    }
    __syncthreads();
    const uint64_t q0 = (uint64_t)s * qs, q1 = min(nq, q0 + qs);
    for (uint64_t q = q0 + w; q < q1; q += warps)
      merge_one_query(W, q, lane, k, (const float*)(base + lab_bytes), (const uint64_t*)base, ev.stride, ev.stride,
                      out_dists, out_labels, out_counts);
  }
}

}  // namespace ehb

// =====================================================================================================
// ehb_exchange: one process per GPU
// =====================================================================================================
struct ehb_exchange {
  int device = 0;
  uint32_t world = 1, rank = 0;
  uint64_t stride = 0, max_elems = 0;
  unsigned char* local = nullptr;   // [flags (+ timeout word at the end of the flag page) | recv]
  size_t flag_bytes = 0, total_bytes = 0;
  unsigned char* mapped[ehb::kMaxWorld] = {nullptr};  // base of every rank's allocation as mapped here
  bool opened[ehb::kMaxWorld] = {false};
  bool attached = false;
  uint32_t* slice_count = nullptr;  // [kMaxSlices], zero between steps (raised by the walk's epilogue)
  uint32_t same_device_ranks = 1;  // ranks (this one included) whose exchange kernels share this GPU (tests)
  uint32_t epoch = 0;
  uint64_t slot_nq = 0;
  uint32_t slot_k = 0;
  int sms = 148;
  std::mutex mu;
};

extern "C" {

int ehb_exchange_create(int32_t device, uint32_t world, uint32_t rank, uint64_t max_nq, uint32_t max_k,
                        ehb_exchange** out) {
  if (!out) return fail(EHB_ERR_INVALID, "null argument");
  if (world == 0 || world > ehb::kMaxWorld || rank >= world) return fail(EHB_ERR_INVALID, "bad world / rank");
  if (max_nq == 0 || max_k == 0) return fail(EHB_ERR_INVALID, "max_nq and max_k must be positive");
  CU(cudaSetDevice(device));
  ehb_exchange* ex = new (std::nothrow) ehb_exchange();
  if (!ex) return fail(EHB_ERR_OOM, "host allocation failed");
  ex->device = device;
  ex->world = world;
  ex->rank = rank;
  ex->max_elems = max_nq * max_k;
  ex->stride = (ex->max_elems * 12ull + 255) / 256 * 256;
  ex->flag_bytes = (2ull * world * ehb::kMaxSlices * 4 + 4 + 4095) / 4096 * 4096;
  ex->total_bytes = ex->flag_bytes + 2ull * world * ex->stride;
  cudaDeviceGetAttribute(&ex->sms, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e = cudaMalloc((void**)&ex->local, ex->total_bytes);
  if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->flag_bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ex->slice_count, ehb::kMaxSlices * 4);
  if (e == cudaSuccess) e = cudaMemset(ex->slice_count, 0, ehb::kMaxSlices * 4);
  if (e != cudaSuccess) {
    if (ex->slice_count) cudaFree(ex->slice_count);
    if (ex->local) cudaFree(ex->local);
    delete ex;
    return fail(e == cudaErrorMemoryAllocation ? EHB_ERR_OOM : EHB_ERR_CUDA, cudaGetErrorString(e));
  }
  ex->mapped[rank] = ex->local;
  ex->attached = world == 1;
  *out = ex;
  return EHB_OK;
}

int ehb_exchange_destroy(ehb_exchange* ex) {
  if (!ex) return EHB_OK;
  cudaSetDevice(ex->device);
  cudaDeviceSynchronize();
  for (uint32_t g = 0; g < ex->world; ++g)
    if (ex->opened[g]) cudaIpcCloseMemHandle(ex->mapped[g]);
  if (ex->slice_count) cudaFree(ex->slice_count);
  if (ex->local) cudaFree(ex->local);
  delete ex;
  return EHB_OK;
}

// 64 bytes (cudaIpcMemHandle_t) other ranks pass to ehb_exchange_open.
int ehb_exchange_ipc_handle(ehb_exchange* ex, void* out_handle) {
  if (!ex || !out_handle) return fail(EHB_ERR_INVALID, "null argument");
  CU(cudaSetDevice(ex->device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, ex->local));
  static_assert(sizeof(cudaIpcMemHandle_t) == EHB_IPC_HANDLE_BYTES, "handle size");
  std::memcpy(out_handle, &h, sizeof(h));
  return EHB_OK;
}

// handles: [world][64] in rank order (this rank's own entry is ignored).
int ehb_exchange_open(ehb_exchange* ex, const void* handles) {
  if (!ex || !handles) return fail(EHB_ERR_INVALID, "null argument");
  CU(cudaSetDevice(ex->device));
  for (uint32_t g = 0; g < ex->world; ++g) {
    if (g == ex->rank || ex->opened[g]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const unsigned char*)handles + (size_t)g * EHB_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ex->mapped[g] = (unsigned char*)p;
    ex->opened[g] = true;
  }
  ex->attached = true;
  return EHB_OK;
}

// Same process, different device: attach a peer exchange directly (peer access must be possible).
int ehb_exchange_attach_local(ehb_exchange* ex, uint32_t peer_rank, ehb_exchange* peer) {
  if (!ex || !peer || peer_rank >= ex->world || peer_rank == ex->rank) return fail(EHB_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(ex->device));
  if (peer->device != ex->device) {
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, ex->device, peer->device));
    if (!can) return fail(EHB_ERR_CUDA, "devices cannot access each other's memory");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
    cudaGetLastError();
  }
  if (!ex->mapped[peer_rank] && peer->device == ex->device) ex->same_device_ranks++;
  ex->mapped[peer_rank] = peer->local;
  bool all = true;
  for (uint32_t g = 0; g < ex->world; ++g) all = all && ex->mapped[g] != nullptr;
  ex->attached = all;
  return EHB_OK;
}

// Starts a step: returns where THIS rank's search must write its [nq][k] labels and distances.
int ehb_exchange_begin(ehb_exchange* ex, uint64_t nq, uint32_t k, uint64_t** labels_dev, float** dists_dev) {
  if (!ex || !labels_dev || !dists_dev) return fail(EHB_ERR_INVALID, "null argument");
  if (nq == 0 || k == 0 || nq * k > ex->max_elems) return fail(EHB_ERR_INVALID, "nq * k exceeds the exchange capacity");
  if (!ex->attached) return fail(EHB_ERR_STATE, "peers are not attached yet");
  std::lock_guard<std::mutex> g(ex->mu);
  ex->epoch++;
  ex->slot_nq = nq;
  ex->slot_k = k;
  const uint32_t parity = ex->epoch & 1u;
  unsigned char* blk = ex->local + ex->flag_bytes + ((uint64_t)parity * ex->world + ex->rank) * ex->stride;
  *labels_dev = (uint64_t*)blk;
  *dists_dev = (float*)(blk + nq * k * 8ull);
  return EHB_OK;
}

}  // extern "C"

namespace {

// slices of whole queries, a multiple of 4 queries so every slice boundary is 16 B aligned
void slice_plan(const ehb_exchange* ex, uint64_t nq, uint32_t* qs, uint32_t* nslices) {
  uint32_t target = std::min<uint32_t>(ehb::kMaxSlices, (uint32_t)ex->sms);
  uint32_t per = (uint32_t)((nq + target - 1) / target);
  per = (per + 3) / 4 * 4;
  *qs = per;
  *nslices = (uint32_t)((nq + per - 1) / per);
}

int launch_exchange_merge(ehb_exchange* ex, uint64_t nq, uint32_t k, float* out_dists_dev, uint64_t* out_labels_dev,
                          uint32_t* out_counts_dev, cudaStream_t stream, bool skip_push) {
  ehb::ExchangeView ev;
  std::memset(&ev, 0, sizeof(ev));
  for (uint32_t r = 0; r < ex->world; ++r) {
    ev.flags[r] = (uint32_t*)ex->mapped[r];
    ev.recv[r] = ex->mapped[r] + ex->flag_bytes;
  }
  ev.world = ex->world;
  ev.rank = ex->rank;
  ev.stride = ex->stride;
  uint32_t qs, nslices;
  slice_plan(ex, nq, &qs, &nslices);
  // every CTA must be resident (a waiting CTA may depend on a peer's CTA): one CTA per SM, and when several ranks
  // share this GPU (single-GPU tests) they split the SMs
  uint32_t grid = std::min<uint32_t>(nslices, std::max<uint32_t>(1, (uint32_t)ex->sms / ex->same_device_ranks));
  ehb::exchange_merge_kernel<<<grid, ehb::kExchangeThreads, 0, stream>>>(
      ev, ex->epoch & 1u, ex->epoch, nq, k, qs, nslices, out_dists_dev, out_labels_dev, out_counts_dev,
      (uint32_t*)(ex->local + ex->flag_bytes - 4), skip_push ? 1u : 0u);
  CU(cudaGetLastError());
  return EHB_OK;
}

}  // namespace

extern "C" {

// Finishes the step on `stream` (the stream the search was queued on): push + flags + wait + merge.
int ehb_exchange_merge_dev(ehb_exchange* ex, float* out_dists_dev, uint64_t* out_labels_dev, uint32_t* out_counts_dev,
                           void* stream) {
  if (!ex || !out_labels_dev) return fail(EHB_ERR_INVALID, "null argument");
  CU(cudaSetDevice(ex->device));
  std::lock_guard<std::mutex> g(ex->mu);
  if (!ex->slot_nq) return fail(EHB_ERR_STATE, "ehb_exchange_begin was not called");
  const uint64_t nq = ex->slot_nq;
  const uint32_t k = ex->slot_k;
  ex->slot_nq = 0;
  return launch_exchange_merge(ex, nq, k, out_dists_dev, out_labels_dev, out_counts_dev, (cudaStream_t)stream, false);
}

// One sharded graph search step, fused: this rank's walk stores every query's top-k straight into every peer's
// receive buffer from its epilogue (coalesced stores over NVLink while the other queries are still walking) and
// raises per-slice flags; then one kernel waits for the peers' flags and merges.  Falls back to push-after-walk
// when the batch is small enough for the team walk.  Call in lock step on every rank.
int ehb_exchange_search_dev(ehb_exchange* ex, ehb_index* ix, uint64_t nq, const float* queries_dev, uint32_t k,
                            uint32_t ef, float* out_dists_dev, uint64_t* out_labels_dev, uint32_t* out_counts_dev,
                            uint32_t* shard_counts_dev, void* stream) {
  if (!ex || !ix || !queries_dev || !out_labels_dev) return fail(EHB_ERR_INVALID, "null argument");
  if (nq == 0 || k == 0 || nq * k > ex->max_elems) return fail(EHB_ERR_INVALID, "nq * k exceeds the exchange capacity");
  if (!ex->attached) return fail(EHB_ERR_STATE, "peers are not attached yet");
  CU(cudaSetDevice(ex->device));
  std::lock_guard<std::mutex> g(ex->mu);
  ex->epoch++;
  const uint32_t parity = ex->epoch & 1u, W = ex->world, me = ex->rank;
  const uint64_t blk = ((uint64_t)parity * W + me) * ex->stride;
  uint32_t qs, nslices;
  slice_plan(ex, nq, &qs, &nslices);
  ehb::ResultSink sink;
  std::memset(&sink, 0, sizeof(sink));
  uint32_t t = 0;
  auto add = [&](uint32_t r) {
    unsigned char* base = ex->mapped[r] + ex->flag_bytes + blk;
    sink.labels[t] = (uint64_t*)base;
    sink.dists[t] = (float*)(base + nq * k * 8ull);
    sink.flags[t] = (uint32_t*)ex->mapped[r] + ((uint64_t)parity * W + me) * ehb::kMaxSlices;
    ++t;
  };
  add(me);  // destination 0 = my own block of my own buffer
  for (uint32_t r = 0; r < W; ++r)
    if (r != me) add(r);
  sink.n = t;
  sink.qs = qs;
  sink.epoch = ex->epoch;
  sink.slice_count = ex->slice_count;
  bool pushed = false;
  RET(ehb_index_search_dev_sink(ix, nq, queries_dev, k, ef, &sink, shard_counts_dev, (cudaStream_t)stream, &pushed));
  CU(cudaSetDevice(ex->device));
  return launch_exchange_merge(ex, nq, k, out_dists_dev, out_labels_dev, out_counts_dev, (cudaStream_t)stream, pushed);
}

// 1 when some exchange kernel of this rank gave up waiting for a peer (its results are then invalid).
int ehb_exchange_timed_out(ehb_exchange* ex, uint32_t* out) {
  if (!ex || !out) return fail(EHB_ERR_INVALID, "null argument");
  CU(cudaSetDevice(ex->device));
  CU(cudaMemcpy(out, ex->local + ex->flag_bytes - 4, 4, cudaMemcpyDeviceToHost));
  return EHB_OK;
}

}  // extern "C"

// =====================================================================================================
// ehb_sharded: one process, n_dev GPUs
// =====================================================================================================
struct ehb_sharded {
  std::vector<ehb_index*> shard;
  std::vector<int> dev;
  uint64_t span = 0;  // labels per shard range (0: label % n_dev)
  ehb_params prm;
  bool peer_direct = true;  // every device can store into device 0
  // device-0 gather + result buffers, per-device query staging
  ehb::DevBuf<unsigned char> gather;       // [n_dev][labels | dists]
  ehb::DevBuf<float> m_dists;
  ehb::DevBuf<uint64_t> m_labels;
  ehb::DevBuf<uint32_t> m_counts;
  std::vector<ehb::DevBuf<float>*> q_dev;   // per device
  std::vector<ehb::DevBuf<unsigned char>*> local_out;  // per device (no peer access): [labels | dists]
  std::vector<ehb::DevBuf<uint32_t>*> cnt_dev;
  std::vector<cudaStream_t> st;
  std::vector<cudaEvent_t> done;
  cudaEvent_t q_ready = nullptr;
  std::mutex mu;
  uint64_t next_label = 0;

  uint32_t owner(uint64_t label) const {
    const uint64_t G = shard.size();
    return (uint32_t)(span ? (label / span) % G : label % G);
  }
};

extern "C" {

int ehb_sharded_destroy(ehb_sharded* sh) {
  if (!sh) return EHB_OK;
  for (size_t g = 0; g < sh->shard.size(); ++g) {
    cudaSetDevice(sh->dev[g]);
    cudaDeviceSynchronize();
    if (g < sh->st.size() && sh->st[g]) cudaStreamDestroy(sh->st[g]);
    if (g < sh->done.size() && sh->done[g]) cudaEventDestroy(sh->done[g]);
    if (g < sh->q_dev.size()) delete sh->q_dev[g];
    if (g < sh->local_out.size()) delete sh->local_out[g];
    if (g < sh->cnt_dev.size()) delete sh->cnt_dev[g];
    ehb_index_destroy(sh->shard[g]);
  }
  if (!sh->dev.empty()) {
    cudaSetDevice(sh->dev[0]);
    if (sh->q_ready) cudaEventDestroy(sh->q_ready);
    sh->gather.release();
    sh->m_dists.release();
    sh->m_labels.release();
    sh->m_counts.release();
  }
  delete sh;
  return EHB_OK;
}

// ehb_index_create over device_ids[0..n_dev): p->device is ignored, p->capacity is per shard.
// shard_span: labels [i*span, (i+1)*span) live on shard i % n_dev (0 = label % n_dev).
int ehb_sharded_create(const ehb_params* p, const int32_t* device_ids, uint32_t n_dev, uint64_t shard_span,
                       ehb_sharded** out) {
  if (!p || !device_ids || !out) return fail(EHB_ERR_INVALID, "null argument");
  if (n_dev == 0 || n_dev > ehb::kMaxWorld) return fail(EHB_ERR_INVALID, "n_dev must be in 1..16");
  // (a device may be listed more than once: several shards then share it — useful on one GPU)
  ehb_sharded* sh = new (std::nothrow) ehb_sharded();
  if (!sh) return fail(EHB_ERR_OOM, "host allocation failed");
  sh->prm = *p;
  sh->span = shard_span;
  auto body = [&]() -> int {
    for (uint32_t g = 0; g < n_dev; ++g) {
      ehb_params pg = *p;
      pg.device = device_ids[g];
      ehb_index* ix = nullptr;
      RET(ehb_index_create(&pg, &ix));
      sh->shard.push_back(ix);
      sh->dev.push_back(device_ids[g]);
      cudaStream_t s = nullptr;
      cudaEvent_t e = nullptr;
      CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
      sh->st.push_back(s);
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      sh->done.push_back(e);
      sh->q_dev.push_back(new ehb::DevBuf<float>());
      sh->local_out.push_back(new ehb::DevBuf<unsigned char>());
      sh->cnt_dev.push_back(new ehb::DevBuf<uint32_t>());
      if (g > 0 && device_ids[g] != device_ids[0]) {  // device g must be able to store into device 0
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, device_ids[g], device_ids[0]));
        if (can) {
          cudaError_t pe = cudaDeviceEnablePeerAccess(device_ids[0], 0);
          if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CU(pe);
          cudaGetLastError();
        } else {
          sh->peer_direct = false;
        }
      }
    }
    CU(cudaSetDevice(device_ids[0]));
    CU(cudaEventCreateWithFlags(&sh->q_ready, cudaEventDisableTiming));
    return EHB_OK;
  };
  int rc = body();
  if (rc != EHB_OK) {
    const std::string msg = ehb::last_error_text();
    ehb_sharded_destroy(sh);
    return fail(rc, msg);
  }
  *out = sh;
  return EHB_OK;
}

int ehb_sharded_n_shards(ehb_sharded* sh, uint32_t* out) {
  if (!sh || !out) return fail(EHB_ERR_INVALID, "null argument");
  *out = (uint32_t)sh->shard.size();
  return EHB_OK;
}

// Borrow shard i (stats, tuning); owned by the sharded index.
int ehb_sharded_shard(ehb_sharded* sh, uint32_t i, ehb_index** out) {
  if (!sh || !out || i >= sh->shard.size()) return fail(EHB_ERR_INVALID, "bad argument");
  *out = sh->shard[i];
  return EHB_OK;
}

// Insert-or-update, routed by label range; no communication between shards.
int ehb_sharded_add(ehb_sharded* sh, uint64_t n, const float* vecs, const uint64_t* labels) {
  if (!sh) return fail(EHB_ERR_INVALID, "null handle");
  if (n && !vecs) return fail(EHB_ERR_INVALID, "null vectors");
  std::lock_guard<std::mutex> g(sh->mu);
  const size_t G = sh->shard.size(), dim = sh->prm.dim;
  std::vector<std::vector<float>> rows(G);
  std::vector<std::vector<uint64_t>> labs(G);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t l = labels ? labels[i] : sh->next_label + i;
    const uint32_t o = sh->owner(l);
    rows[o].insert(rows[o].end(), vecs + i * dim, vecs + (i + 1) * dim);
    labs[o].push_back(l);
  }
  for (size_t o = 0; o < G; ++o)
    if (!labs[o].empty()) RET(ehb_index_add(sh->shard[o], labs[o].size(), rows[o].data(), labs[o].data()));
  if (!labels) sh->next_label += n;
  return EHB_OK;
}

int ehb_sharded_remove(ehb_sharded* sh, uint64_t n, const uint64_t* labels) {
  if (!sh || (n && !labels)) return fail(EHB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(sh->mu);
  for (uint64_t i = 0; i < n; ++i) RET(ehb_index_remove(sh->shard[sh->owner(labels[i])], 1, labels + i));
  return EHB_OK;
}

int ehb_sharded_get(ehb_sharded* sh, uint64_t label, float* out) {
  if (!sh) return fail(EHB_ERR_INVALID, "null handle");
  return ehb_index_get(sh->shard[sh->owner(label)], label, out);
}

int ehb_sharded_size(ehb_sharded* sh, uint64_t* out) {
  if (!sh || !out) return fail(EHB_ERR_INVALID, "null argument");
  uint64_t tot = 0, v = 0;
  for (ehb_index* ix : sh->shard) {
    RET(ehb_index_size(ix, &v));
    tot += v;
  }
  *out = tot;
  return EHB_OK;
}

// Links every shard; the shards build concurrently (one host thread per device).
int ehb_sharded_build(ehb_sharded* sh) {
  if (!sh) return fail(EHB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> g(sh->mu);
  const size_t G = sh->shard.size();
  std::vector<int> rc(G, EHB_OK);
  std::vector<std::string> msg(G);
  std::vector<std::thread> th;
  for (size_t i = 0; i < G; ++i)
    th.emplace_back([&, i]() {
      rc[i] = ehb_index_build(sh->shard[i]);
      if (rc[i] != EHB_OK) msg[i] = ehb::last_error_text();
    });
  for (auto& t : th) t.join();
  for (size_t i = 0; i < G; ++i)
    if (rc[i] != EHB_OK) return fail(rc[i], msg[i]);
  return EHB_OK;
}

int ehb_sharded_set_ef(ehb_sharded* sh, uint32_t ef) {
  if (!sh) return fail(EHB_ERR_INVALID, "null handle");
  for (ehb_index* ix : sh->shard) RET(ehb_index_set_ef(ix, ef));
  return EHB_OK;
}

// Host queries in, merged host results out.  mode: 0 = graph walk, 1 = exact brute force, 2 = bf16 brute force.
static int sharded_search(ehb_sharded* sh, int mode, uint64_t nq, const float* q, uint32_t k, uint32_t ef, uint64_t* ol,
                          float* od, uint32_t* oc) {
  if (!sh) return fail(EHB_ERR_INVALID, "null handle");
  if (nq && (!q || !ol)) return fail(EHB_ERR_INVALID, "null buffer");
  if (nq == 0 || k == 0) return EHB_OK;
  std::lock_guard<std::mutex> g(sh->mu);
  const uint32_t G = (uint32_t)sh->shard.size();
  const uint32_t dim = sh->prm.dim;
  const uint64_t blk = (nq * k * 12ull + 255) / 256 * 256;
  CU(cudaSetDevice(sh->dev[0]));
  cudaStream_t s0 = sh->st[0];
  CU(sh->gather.grow(blk * G, 0, -1, s0));
  CU(sh->m_labels.grow(nq * k, 0, -1, s0));
  CU(sh->m_dists.grow(nq * k, 0, -1, s0));
  CU(sh->m_counts.grow(nq, 0, -1, s0));
  CU(cudaEventRecord(sh->q_ready, s0));  // orders the peers' stores after earlier merges on device 0
  for (uint32_t i = 0; i < G; ++i) {
    CU(cudaSetDevice(sh->dev[i]));
    cudaStream_t s = sh->st[i];
    CU(sh->q_dev[i]->grow(nq * dim, 0, -1, s));
    CU(sh->cnt_dev[i]->grow(nq, 0, -1, s));
    CU(cudaMemcpyAsync(sh->q_dev[i]->p, q, nq * dim * 4, cudaMemcpyHostToDevice, s));
    unsigned char* dst;
    if (i == 0 || sh->peer_direct || sh->dev[i] == sh->dev[0]) {
      dst = sh->gather.p + blk * i;  // device i's kernels store straight into device 0's gather block
      if (i) CU(cudaStreamWaitEvent(s, sh->q_ready, 0));
    } else {
      CU(sh->local_out[i]->grow(blk, 0, -1, s));
      dst = sh->local_out[i]->p;
    }
    uint64_t* dl = (uint64_t*)dst;
    float* dd = (float*)(dst + nq * k * 8ull);
    if (mode == 0)
      RET(ehb_index_search_dev(sh->shard[i], nq, sh->q_dev[i]->p, k, ef, dl, dd, sh->cnt_dev[i]->p, s));
    else
      RET(ehb_index_search_bruteforce_dev(sh->shard[i], nq, sh->q_dev[i]->p, k, mode == 2 ? EHB_BF16 : EHB_FP32, dl, dd,
                                          sh->cnt_dev[i]->p, s));
    CU(cudaSetDevice(sh->dev[i]));
    if (i && !sh->peer_direct && sh->dev[i] != sh->dev[0])
      CU(cudaMemcpyPeerAsync(sh->gather.p + blk * i, sh->dev[0], dst, sh->dev[i], nq * k * 12ull, s));
    CU(cudaEventRecord(sh->done[i], s));
  }
  CU(cudaSetDevice(sh->dev[0]));
  for (uint32_t i = 1; i < G; ++i) CU(cudaStreamWaitEvent(s0, sh->done[i], 0));
  CU(ehb::launch_merge_topk(G, nq, k, (const float*)(sh->gather.p + nq * k * 8ull), (const uint64_t*)sh->gather.p, blk,
                            blk, sh->m_dists.p, sh->m_labels.p, sh->m_counts.p, s0));
  CU(cudaMemcpyAsync(ol, sh->m_labels.p, nq * k * 8, cudaMemcpyDeviceToHost, s0));
  if (od) CU(cudaMemcpyAsync(od, sh->m_dists.p, nq * k * 4, cudaMemcpyDeviceToHost, s0));
  if (oc) CU(cudaMemcpyAsync(oc, sh->m_counts.p, nq * 4, cudaMemcpyDeviceToHost, s0));
  CU(cudaStreamSynchronize(s0));
  return EHB_OK;
}

int ehb_sharded_search(ehb_sharded* sh, uint64_t nq, const float* q, uint32_t k, uint32_t ef, uint64_t* ol, float* od,
                       uint32_t* oc) {
  return sharded_search(sh, 0, nq, q, k, ef, ol, od, oc);
}
int ehb_sharded_search_bruteforce(ehb_sharded* sh, uint64_t nq, const float* q, uint32_t k, int precision, uint64_t* ol,
                                  float* od, uint32_t* oc) {
  return sharded_search(sh, precision == EHB_BF16 ? 2 : 1, nq, q, k, 0, ol, od, oc);
}

}  // extern "C"
