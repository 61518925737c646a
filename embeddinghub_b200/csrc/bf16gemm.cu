// K3 — brute-force distances as a bf16 tensor-core GEMM (tcgen05 + TMEM + TMA).
//
// The batched exact scan is a genuine dense contraction: D[q][n] = <Q[q], X[n]> over
// d, Q x N x d multiply-adds (C4: 4096 x 10M x 768 = 3.1e13 MACs).  This kernel
// computes 128 x 256 tiles of it on the 5th-generation tensor cores:
//   warp 0  : TMA producer — cp.async.bulk.tensor.2d (SASS UTMALDG) of a 128 x 64 bf16
//             query box and a 256 x 64 bf16 base box per k-block, 128B-swizzled, into a
//             4-stage shared-memory ring, completion on mbarriers;
//   warp 1  : MMA issuer — one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//             (SASS UTCHMMA), M=128 N=256 K=16, fp32 accumulators in TMEM (256 columns);
//             tcgen05.commit releases the smem stage / signals the epilogue;
//   warps 2-5: epilogue — tcgen05.ld (SASS LDTM) 32 lanes x 32 columns at a time, turn the
//             dot products into distances (1 - dot, or |q|^2 + |x|^2 - 2 dot) and store
//             the fp32 tile.
// The candidate selection (top-k' per query over the tile rows), and the fp32
// re-rank that restores exact ids, reuse the K1 select / merge kernels
// (bruteforce.cu) and rerank_kernel below.
//
// Replaces: hnswlib::BruteforceSearch semantics on the "bf16 tensor-core GEMM path"
// named by BASELINE.json (configs[3]); the reference itself has no such path.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "kernels.h"

namespace ehb {

constexpr int GM = 128, GN = 256, GK = 64, GSTAGES = 4;
constexpr uint32_t kStageBytesA = GM * GK * 2, kStageBytesB = GN * GK * 2;
constexpr uint32_t kGemmSmem = GSTAGES * (kStageBytesA + kStageBytesB) + 1024 /*align*/ + 256 /*barriers*/;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row swizzle atoms of 1024 B):
// start address >> 4, SBO = 1024 B >> 4, descriptor version 1 (sm_100), layout type SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t make_smem_desc(const void* smem_tile) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem_tile) & 0x3FFFFu) >> 4);  // bits [0,14)
  d |= (uint64_t)0 << 16;                                   // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024u >> 4) << 32;                        // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                   // version
  d |= (uint64_t)2 << 61;                                   // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3, M >> 4.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// dist[(q - q0) * ldd + (n - n0)], q in [q0, q0 + qn), n in [n0, n0 + nn).
// metric 0: qnorm[q] + xnorm[n] - 2 dot;  metric 1: 1 - dot.
__global__ void __launch_bounds__(192, 1)
    bf16_dist_gemm_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                          uint32_t kblocks, int metric, const float* __restrict__ qnorm,
                          const float* __restrict__ xnorm, uint64_t q0, uint64_t qn, uint64_t n0, uint64_t nn,
                          float* __restrict__ dist, uint64_t ldd) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // 128B swizzle: 1 KB align
  unsigned char* sA = smem;
  unsigned char* sB = smem + GSTAGES * kStageBytesA;
  uint64_t* full = (uint64_t*)(smem + GSTAGES * (kStageBytesA + kStageBytesB));
  uint64_t* empty = full + GSTAGES;
  uint64_t* tmem_full = empty + GSTAGES;
  uint32_t* tmem_ptr = (uint32_t*)(tmem_full + 1);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // query tiles vary fastest: the CTAs that share one 256-row base tile run together, so the base set is
  // read from HBM once per chunk (ncu, base-tile-fastest order: 3.2 GB read per launch for a 0.2 GB chunk)
  const uint32_t tile_q = blockIdx.x, tile_n = blockIdx.y;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < GSTAGES; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    mbar_init(tmem_full, 1);
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, GN);  // 256 fp32 columns x 128 lanes
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (uint32_t kb = 0; kb < kblocks; ++kb) {
        uint32_t s = kb % GSTAGES, ph = (kb / GSTAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);  // first pass through the ring passes immediately
        mbar_arrive_expect_tx(&full[s], kStageBytesA + kStageBytesB);
        tma_load_2d(sA + s * kStageBytesA, &map_q, (int)(kb * GK), (int)(q0 + (uint64_t)tile_q * GM), &full[s]);
        tma_load_2d(sB + s * kStageBytesB, &map_x, (int)(kb * GK), (int)(n0 + (uint64_t)tile_n * GN), &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(GM, GN);
      for (uint32_t kb = 0; kb < kblocks; ++kb) {
        uint32_t s = kb % GSTAGES, ph = (kb / GSTAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        uint64_t da = make_smem_desc(sA + s * kStageBytesA), db = make_smem_desc(sB + s * kStageBytesB);
#pragma unroll
        for (uint32_t k4 = 0; k4 < GK / 16; ++k4)  // UMMA_K = 16 bf16 = 32 B = +2 in the (>>4) start address
          umma_bf16(tmem_base, da + 2 * k4, db + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full);    // accumulators complete
    }
  } else {
    // epilogue warp w covers TMEM lanes [32 * (warp % 4), +32) = query rows of the tile
    const uint32_t quarter = warp & 3u;
    const uint64_t qrow = (uint64_t)tile_q * GM + quarter * 32u + lane;  // relative to q0
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const float qn2 = (metric == 0 && qrow < qn) ? qnorm[q0 + qrow] : 0.f;
    for (uint32_t c0 = 0; c0 < GN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((quarter * 32u) << 16) + c0, r);
      uint64_t ncol = (uint64_t)tile_n * GN + c0;  // relative to n0
      if (qrow < qn) {
        float* out = dist + qrow * ldd + ncol;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float dot = __uint_as_float(r[j]);
          float xn = (metric == 0 && ncol + j < nn) ? xnorm[n0 + ncol + j] : 0.f;
          v[j] = metric == 0 ? fmaxf(qn2 + xn - 2.0f * dot, 0.f) : 1.0f - dot;
        }
        if (ncol + 32 <= nn && (ldd & 3u) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *(float4*)(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ncol + j < nn) out[j] = v[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, GN);
}

// ---------------------------------------------------------------------------------------------------
// Persistent variant with the selection fused into the epilogue: the Q x N distance tile never goes to HBM.
// One CTA per SM walks the (query tile, base tile) list (query tile fastest, so the CTAs that share a base
// tile run together); accumulators are double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i
// overlaps the MMAs of tile i+1.  The epilogue compares every distance with the query's current threshold
// thr[q] (its kc-th best bf16 distance so far) and appends the survivors (ordered distance | row index) to
// the query's candidate buffer with one global atomic each; compact_candidates_kernel folds the buffer into
// the running top-kc and tightens thr between chunks.  With chunk sizes that double, a chunk admits about kc
// candidates per query, so the buffer (capacity 2 kc + 64) practically never overflows; if it does the
// driver re-runs that chunk through the unfused path.
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kFusedSmem = GSTAGES * (kStageBytesA + kStageBytesB) + 1024 + 256;
static cudaError_t make_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t dpad, uint32_t box_rows);

__global__ void __launch_bounds__(192, 1)
    bf16_topk_gemm_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                          uint32_t kblocks, int metric, const float* __restrict__ qnorm,
                          const float* __restrict__ xnorm, uint64_t nq, uint64_t n_lo, uint64_t n_hi,
                          const float* __restrict__ thr, uint64_t* __restrict__ cbuf, uint32_t* __restrict__ ccount,
                          uint32_t ccap) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + GSTAGES * kStageBytesA;
  uint64_t* full = (uint64_t*)(smem + GSTAGES * (kStageBytesA + kStageBytesB));
  uint64_t* empty = full + GSTAGES;
  uint64_t* tmem_full = empty + GSTAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t q_tiles = (nq + GM - 1) / GM, n_tiles = (n_hi - n_lo + GN - 1) / GN;
  const uint64_t tiles = q_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < GSTAGES; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&tmem_full[a], 1), mbar_init(&tmem_empty[a], 4);
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 2 * GN);  // 512 columns: two accumulator stages
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint64_t tq = t % q_tiles, tn = t / q_tiles;
        for (uint32_t kb = 0; kb < kblocks; ++kb, ++it) {
          uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], kStageBytesA + kStageBytesB);
          tma_load_2d(sA + s * kStageBytesA, &map_q, (int)(kb * GK), (int)(tq * GM), &full[s]);
          tma_load_2d(sB + s * kStageBytesB, &map_x, (int)(kb * GK), (int)(n_lo + tn * GN), &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(GM, GN);
      uint32_t it = 0, ti = 0;
      for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++ti) {
        const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
        mbar_wait(&tmem_empty[acc], aph ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        for (uint32_t kb = 0; kb < kblocks; ++kb, ++it) {
          uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          uint64_t da = make_smem_desc(sA + s * kStageBytesA), db = make_smem_desc(sB + s * kStageBytesB);
#pragma unroll
          for (uint32_t k4 = 0; k4 < GK / 16; ++k4)
            umma_bf16(tmem_base + acc * GN, da + 2 * k4, db + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const uint32_t quarter = warp & 3u;
    uint32_t ti = 0;
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++ti) {
      const uint64_t tq = t % q_tiles, tn = t / q_tiles;
      const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
      const uint64_t q = tq * GM + quarter * 32u + lane;
      const bool qok = q < nq;
      const float tau = qok ? thr[q] : -INFINITY;
      const float qn2 = (metric == 0 && qok) ? qnorm[q] : 0.f;
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      for (uint32_t c0 = 0; c0 < GN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + acc * GN + ((quarter * 32u) << 16) + c0, r);
        const uint64_t nbase = n_lo + tn * GN + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float dot = __uint_as_float(r[j]);
          float d;
          if (metric == 0) {
            float xn = nbase + j < n_hi ? xnorm[nbase + j] : 0.f;
            d = fmaxf(qn2 + xn - 2.0f * dot, 0.f);
          } else {
            d = 1.0f - dot;
          }
          if (d < tau && nbase + j < n_hi) {
            uint32_t pos = atomicAdd(&ccount[q], 1u);
            if (pos < ccap) cbuf[q * ccap + pos] = make_key(d, (uint32_t)(nbase + j));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * GN);
}

// ---------------------------------------------------------------------------------------------------
// 2-CTA form of the fused kernel (cta_group::2): a cluster of two SMs owns a 256 x 256 tile.  Each CTA
// stages its own 128 query rows and HALF of the base tile (128 rows) per k-block — 32 KB instead of 48 KB for
// the same tensor-pipe time — so the six-stage ring covers the L2 -> SM latency that bounds the 1-CTA kernel.
// The leader CTA's MMA thread issues tcgen05.mma.cta_group::2 (M = 256); both CTAs' TMA loads complete on the
// leader's `full` barrier; tcgen05.commit multicasts the `empty` / `tmem_full` arrivals to both CTAs; the
// non-leader's epilogue warps release the accumulator stage on the leader's `tmem_empty` barrier remotely.
// Opt-in (ehb_index_set_option "gemm_2cta"): measured on C4-shaped chunks it reaches 765 TFLOP/s in-kernel vs 826 TFLOP/s
// for the 1-CTA kernel — both sit on the L2 -> SM operand traffic (92 resp. 61 B/clk/SM requested against
// a chip-wide LTS cap of ~6.3 KB/clk), so the next step is TMA multicast across a larger cluster, not this.
// ---------------------------------------------------------------------------------------------------
constexpr int G2STAGES = 6;
constexpr uint32_t kStage2 = (GM * GK * 2) + (128 * GK * 2);  // A 128x64 + B-half 128x64 bf16 = 32 KB
constexpr uint32_t kFused2Smem = G2STAGES * kStage2 + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  // both CTAs execute this; clearing the peer bit makes the transaction bytes land on CTA 0's barrier
  uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "l"(0x1000000000000000ull)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}

__global__ void __launch_bounds__(192, 1)
    bf16_topk_gemm2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                           uint32_t kblocks, int metric, const float* __restrict__ qnorm,
                           const float* __restrict__ xnorm, uint64_t nq, uint64_t n_lo, uint64_t n_hi,
                           const float* __restrict__ thr, uint64_t* __restrict__ cbuf, uint32_t* __restrict__ ccount,
                           uint32_t ccap) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                   // [stage][128 x 64]
  unsigned char* sB = smem + G2STAGES * (GM * GK * 2);        // [stage][128 x 64] (this CTA's half of the base tile)
  uint64_t* full = (uint64_t*)(smem + G2STAGES * kStage2);
  uint64_t* empty = full + G2STAGES;
  uint64_t* tmem_full = empty + G2STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2] (used in the leader)
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const uint64_t q_tiles = (nq + 2 * GM - 1) / (2 * GM), n_tiles = (n_hi - n_lo + GN - 1) / GN;
  const uint64_t tiles = q_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < G2STAGES; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&tmem_full[a], 1), mbar_init(&tmem_empty[a], 8);
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(2u * GN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (uint64_t t = cluster_id; t < tiles; t += n_clusters) {
        const uint64_t tq = t % q_tiles, tn = t / q_tiles;
        for (uint32_t kb = 0; kb < kblocks; ++kb, ++it) {
          uint32_t s = it % G2STAGES, ph = (it / G2STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * kStage2);  // both CTAs' bytes land here
          tma_load_2d_2sm(sA + s * (GM * GK * 2), &map_q, (int)(kb * GK), (int)(tq * 2 * GM + rank * GM), &full[s]);
          tma_load_2d_2sm(sB + s * (128 * GK * 2), &map_x, (int)(kb * GK), (int)(n_lo + tn * GN + rank * 128),
                          &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc(2 * GM, GN);
      uint32_t it = 0, ti = 0;
      for (uint64_t t = cluster_id; t < tiles; t += n_clusters, ++ti) {
        const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
        mbar_wait(&tmem_empty[acc], aph ^ 1u);
        tc_fence_after();
        for (uint32_t kb = 0; kb < kblocks; ++kb, ++it) {
          uint32_t s = it % G2STAGES, ph = (it / G2STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          uint64_t da = make_smem_desc(sA + s * (GM * GK * 2)), db = make_smem_desc(sB + s * (128 * GK * 2));
#pragma unroll
          for (uint32_t k4 = 0; k4 < GK / 16; ++k4)
            umma_bf16_2sm(tmem_base + acc * GN, da + 2 * k4, db + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty[s]);
        }
        umma_commit_2sm(&tmem_full[acc]);
      }
    }
  } else {
    const uint32_t quarter = warp & 3u;
    uint32_t ti = 0;
    for (uint64_t t = cluster_id; t < tiles; t += n_clusters, ++ti) {
      const uint64_t tq = t % q_tiles, tn = t / q_tiles;
      const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
      const uint64_t q = tq * 2 * GM + rank * GM + quarter * 32u + lane;
      const bool qok = q < nq;
      const float tau = qok ? thr[q] : -INFINITY;
      const float qn2 = (metric == 0 && qok) ? qnorm[q] : 0.f;
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      for (uint32_t c0 = 0; c0 < GN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + acc * GN + ((quarter * 32u) << 16) + c0, r);
        const uint64_t nbase = n_lo + tn * GN + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float dot = __uint_as_float(r[j]);
          float d;
          if (metric == 0) {
            float xn = nbase + j < n_hi ? xnorm[nbase + j] : 0.f;
            d = fmaxf(qn2 + xn - 2.0f * dot, 0.f);
          } else {
            d = 1.0f - dot;
          }
          if (d < tau && nbase + j < n_hi) {
            uint32_t pos = atomicAdd(&ccount[q], 1u);
            if (pos < ccap) cbuf[q * ccap + pos] = make_key(d, (uint32_t)(nbase + j));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tmem_empty[acc], 0);  // the leader's MMA thread owns the accumulator hand-off
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * GN) : "memory");
}

// One warp per query: fold the candidate buffer into the running top-kc, publish the new threshold, reset
// the counter; overflow[0] != 0 tells the driver the buffer was too small.  The (<= kc + ccap) keys are
// sorted with a warp-wide bitonic network in shared memory (P = next power of two; ~20 k instructions per
// query for P = 2048, 4x cheaper than ~kc sorted inserts into a kc-long list).
__global__ void compact_candidates_kernel(uint64_t* __restrict__ run_keys, uint64_t* __restrict__ cbuf,
                                          uint32_t* __restrict__ ccount, uint32_t ccap, uint32_t kc, uint32_t P,
                                          uint64_t nq, float* __restrict__ thr, uint32_t* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t w = threadIdx.x >> 5, wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  const uint64_t q = (uint64_t)blockIdx.x * wpb + w;
  if (q >= nq) return;
  uint64_t* keys = (uint64_t*)smem + (size_t)w * P;
  uint64_t* run = run_keys + q * kc;
  uint32_t m = ccount[q];
  if (m > ccap) {
    if (lane == 0) atomicExch(overflow, 1u);
    m = ccap;
  }
  for (uint32_t i = lane; i < P; i += 32) {
    uint64_t key = kMaxKey;
    if (i < kc) key = run[i];
    else if (i - kc < m) key = cbuf[q * ccap + (i - kc)];
    keys[i] = key;
  }
  __syncwarp();
  for (uint32_t size = 2; size <= P; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = lane; t < (P >> 1); t += 32) {
        uint32_t i = 2 * t - (t & (stride - 1));     // lower index of the pair
        uint32_t j = i + stride;
        bool up = (i & size) == 0;                   // ascending blocks
        uint64_t a = keys[i], b = keys[j];
        if ((a > b) == up) keys[i] = b, keys[j] = a;
      }
      __syncwarp();
    }
  }
  for (uint32_t i = lane; i < kc; i += 32) run[i] = keys[i];
  if (lane == 0) {
    thr[q] = keys[kc - 1] != kMaxKey ? key_dist(keys[kc - 1]) : INFINITY;
    ccount[q] = 0;
  }
}

cudaError_t launch_bf16_topk_chunk(const void* q_bf16, uint64_t nq, const void* x_bf16, uint64_t x_rows, uint32_t dpad,
                                   int metric, const float* qnorm, const float* xnorm, uint64_t n_lo, uint64_t n_hi,
                                   float* thr, uint64_t* cbuf, uint32_t* ccount, uint32_t ccap, uint64_t* run_keys,
                                   uint32_t kc, uint32_t* overflow, int sms, int variant, cudaStream_t s) {
  if (dpad % GK != 0) return cudaErrorInvalidValue;
  CUtensorMap mq, mx;
  cudaError_t e;
  if ((e = make_map(&mq, q_bf16, nq, dpad, GM)) != cudaSuccess) return e;
  if ((e = make_map(&mx, x_bf16, x_rows, dpad, GN)) != cudaSuccess) return e;
  e = cudaFuncSetAttribute(bf16_topk_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem);
  if (e != cudaSuccess) return e;
  if (n_hi > n_lo) {
    const bool two_cta = variant == 1;
    if (two_cta) {
      if ((e = make_map(&mx, x_bf16, x_rows, dpad, 128)) != cudaSuccess) return e;  // each CTA stages half a base tile
      e = cudaFuncSetAttribute(bf16_topk_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFused2Smem);
      if (e != cudaSuccess) return e;
      uint64_t tiles = ((nq + 2 * GM - 1) / (2 * GM)) * ((n_hi - n_lo + GN - 1) / GN);
      unsigned clusters = (unsigned)(tiles < (uint64_t)(sms / 2) ? tiles : (uint64_t)(sms / 2));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(clusters * 2);
      cfg.blockDim = dim3(192);
      cfg.dynamicSmemBytes = kFused2Smem;
      cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, bf16_topk_gemm2_kernel, mq, mx, dpad / GK, metric == 0 ? 0 : 1, qnorm, xnorm,
                             (uint64_t)nq, n_lo, n_hi, (const float*)thr, cbuf, ccount, ccap);
      if (e != cudaSuccess) return e;
    } else {
      uint64_t tiles = ((nq + GM - 1) / GM) * ((n_hi - n_lo + GN - 1) / GN);
      unsigned grid = (unsigned)(tiles < (uint64_t)sms ? tiles : (uint64_t)sms);
      bf16_topk_gemm_kernel<<<grid, 192, kFusedSmem, s>>>(mq, mx, dpad / GK, metric == 0 ? 0 : 1, qnorm, xnorm, nq,
                                                        n_lo, n_hi, thr, cbuf, ccount, ccap);
    }
  }
  // fold the survivors into the running top-kc
  uint32_t P = 64;
  while (P < kc + ccap) P <<= 1;
  uint32_t wpb = (uint32_t)(65536u / (P * 8u));
  wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);
  size_t smem = (size_t)wpb * P * 8;
  e = cudaFuncSetAttribute(compact_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  compact_candidates_kernel<<<(unsigned)((nq + wpb - 1) / wpb), 32 * wpb, smem, s>>>(run_keys, cbuf, ccount, ccap, kc,
                                                                                     P, nq, thr, overflow);
  return cudaGetLastError();
}

// ---- fp32 -> bf16 rows (+ squared norms of the rounded values) ------------------------------------
__global__ void to_bf16_rows_kernel(const float* __restrict__ in, uint32_t in_stride, __nv_bfloat16* __restrict__ out,
                                    float* __restrict__ norms, uint64_t n, uint32_t dpad) {
  uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint32_t lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* src = in + row * in_stride;
  float acc = 0.f;
  for (uint32_t i = lane; i < dpad; i += 32) {
    __nv_bfloat16 b = __float2bfloat16_rn(src[i]);
    out[row * dpad + i] = b;
    float f = __bfloat162float(b);
    acc = fmaf(f, f, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && norms) norms[row] = acc;
}

cudaError_t launch_to_bf16(const float* in, uint32_t in_stride, void* out_bf16, float* norms, uint64_t n, uint32_t dpad,
                           cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  to_bf16_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(in, in_stride, (__nv_bfloat16*)out_bf16, norms, n, dpad);
  return cudaGetLastError();
}

// ---- tensor maps -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static cudaError_t make_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t dpad, uint32_t box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || !p) return cudaErrorNotSupported;
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {dpad, rows};
  cuuint64_t strides[1] = {(cuuint64_t)dpad * 2};
  cuuint32_t box[2] = {(cuuint32_t)GK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// distances of queries [q0, q0+qn) x base rows [n0, n0+nn) into dist (row stride ldd)
cudaError_t launch_bf16_dist_tile(const void* q_bf16, uint64_t q_rows, const void* x_bf16, uint64_t x_rows,
                                  uint32_t dpad, int metric, const float* qnorm, const float* xnorm, uint64_t q0,
                                  uint64_t qn, uint64_t n0, uint64_t nn, float* dist, uint64_t ldd, cudaStream_t s) {
  if (dpad % GK != 0) return cudaErrorInvalidValue;
  CUtensorMap mq, mx;
  cudaError_t e;
  if ((e = make_map(&mq, q_bf16, q_rows, dpad, GM)) != cudaSuccess) return e;
  if ((e = make_map(&mx, x_bf16, x_rows, dpad, GN)) != cudaSuccess) return e;
  e = cudaFuncSetAttribute(bf16_dist_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((qn + GM - 1) / GM), (unsigned)((nn + GN - 1) / GN));
  bf16_dist_gemm_kernel<<<grid, 192, kGemmSmem, s>>>(mq, mx, dpad / GK, metric == 0 ? 0 : 1, qnorm, xnorm, q0, qn, n0,
                                                    nn, dist, ldd);
  return cudaGetLastError();
}

// ---- fp32 re-rank of the bf16 candidates (canonical arithmetic, total order (dist, index)) ----------------
// One warp per query: candidates cand[q][kc] (keys from the select pass: low 32 bits = row index), exact
// distance per candidate by lane-strided loop over candidates (each lane runs the sequential FMA chain),
// then the k best by repeated warp arg-min.
__global__ void rerank_kernel(const uint64_t* __restrict__ cand, uint32_t kc, const float* __restrict__ qpad,
                              const float* __restrict__ vecs, uint32_t dpad, uint32_t dim, int metric,
                              const uint64_t* __restrict__ labels, uint64_t nq, uint32_t k,
                              uint64_t* __restrict__ out_labels, float* __restrict__ out_dists,
                              uint32_t* __restrict__ out_counts) {
  extern __shared__ uint64_t skeys[];  // [wpb][kc] keys, then [wpb][32][33] row tile, [wpb][32] query segment
  const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const uint64_t q = (uint64_t)blockIdx.x * wpb + w;
  if (q >= nq) return;
  uint64_t* keys = skeys + (size_t)w * kc;
  float* tile = (float*)(skeys + (size_t)wpb * kc) + (size_t)w * (32 * 33 + 32);
  float* qseg = tile + 32 * 33;
  const float* qv = qpad + q * dpad;
  // 32 candidates at a time: rows are staged 32 floats per row per step with coalesced loads, then every
  // lane advances the canonical (k ascending, single accumulator) chain of ITS candidate by 32 terms
  for (uint32_t c0 = 0; c0 < kc; c0 += 32) {
    uint32_t c = c0 + lane;
    uint64_t ck = c < kc ? cand[q * kc + c] : kMaxKey;
    uint32_t idx = ck != kMaxKey ? (uint32_t)ck : 0u;
    float acc = 0.f;
    for (uint32_t s0 = 0; s0 < dim; s0 += 32) {
      __syncwarp();
#pragma unroll 8
      for (uint32_t r = 0; r < 32; ++r) {
        uint32_t ridx = __shfl_sync(0xffffffffu, idx, r);
        tile[r * 33 + lane] = s0 + lane < dim ? vecs[(size_t)ridx * dpad + s0 + lane] : 0.f;
      }
      qseg[lane] = s0 + lane < dim ? qv[s0 + lane] : 0.f;
      __syncwarp();
      const uint32_t lim = min(32u, dim - s0);
      if (metric == 0) {
        for (uint32_t i = 0; i < lim; ++i) {
          float t = qseg[i] - tile[lane * 33 + i];
          acc = fmaf(t, t, acc);
        }
      } else {
        for (uint32_t i = 0; i < lim; ++i) acc = fmaf(qseg[i], tile[lane * 33 + i], acc);
      }
    }
    if (c < kc) keys[c] = ck != kMaxKey ? make_key(metric == 0 ? acc : 1.0f - acc, idx) : kMaxKey;
  }
  __syncwarp();
  uint32_t found = 0;
  for (uint32_t i = 0; i < k; ++i) {
    uint64_t best = kMaxKey;
    uint32_t bpos = 0;
    for (uint32_t c = lane; c < kc; c += 32)
      if (keys[c] < best) best = keys[c], bpos = c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      uint64_t ob = __shfl_xor_sync(0xffffffffu, best, o);
      uint32_t op = __shfl_xor_sync(0xffffffffu, bpos, o);
      if (ob < best) best = ob, bpos = op;
    }
    if (best == kMaxKey) break;
    if (lane == 0) {
      out_labels[q * k + i] = labels[(uint32_t)best];
      if (out_dists) out_dists[q * k + i] = key_dist(best);
      keys[bpos] = kMaxKey;
    }
    __syncwarp();
    found++;
  }
  for (uint32_t i = found + lane; i < k; i += 32) {
    out_labels[q * k + i] = 0xFFFFFFFFFFFFFFFFull;
    if (out_dists) out_dists[q * k + i] = INFINITY;
  }
  if (lane == 0 && out_counts) out_counts[q] = found;
}

cudaError_t launch_rerank(const uint64_t* cand, uint32_t kc, const float* qpad, const float* vecs, uint32_t dpad,
                          uint32_t dim, int metric, const uint64_t* labels, uint64_t nq, uint32_t k,
                          uint64_t* out_labels, float* out_dists, uint32_t* out_counts, cudaStream_t s) {
  if (nq == 0) return cudaSuccess;
  uint32_t wpb = 4;
  size_t smem = (size_t)wpb * kc * 8 + (size_t)wpb * (32 * 33 + 32) * 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  rerank_kernel<<<(unsigned)((nq + wpb - 1) / wpb), 32 * wpb, smem, s>>>(cand, kc, qpad, vecs, dpad, dim, metric,
                                                                         labels, nq, k, out_labels, out_dists,
                                                                         out_counts);
  return cudaGetLastError();
}

}  // namespace ehb
