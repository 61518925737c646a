// K4 device part: per query, G sorted lists of (dist, label) -> global top-k.  One warp per query, lane g
// walks list g; each step is a warp arg-min on (distance, label).  Shared by merge_topk_kernel
// (bruteforce.cu) and the peer-memory shard exchange (exchange.cu).
#pragma once
#include "common.cuh"

namespace ehb {

// rank g's lists start at (bytes) dists + g * stride_d and labels + g * stride_l, so separate [G][nq][k]
// arrays and one packed gather buffer [G][labels | dists] are both merged in place.  (The inputs carry no
// __restrict__/read-only hint: in the shard exchange they are written by peer GPUs while this kernel runs.)
__device__ __forceinline__ void merge_one_query(uint32_t G, uint64_t q, uint32_t lane, uint32_t k,
                                                const float* dists, const uint64_t* labels, uint64_t stride_d,
                                                uint64_t stride_l,
                                                float* __restrict__ out_dists,
                                                uint64_t* __restrict__ out_labels,
                                                uint32_t* __restrict__ out_counts) {
  uint32_t head = 0;
  const uint32_t g = lane < G ? lane : 0;
  const float* dl = (const float*)((const unsigned char*)dists + (uint64_t)g * stride_d) + q * k;
  const uint64_t* ll = (const uint64_t*)((const unsigned char*)labels + (uint64_t)g * stride_l) + q * k;
  uint32_t found = 0;
  for (uint32_t i = 0; i < k; ++i) {
    uint32_t od = 0xFFFFFFFFu;
    uint64_t lab = 0xFFFFFFFFFFFFFFFFull;
    if (lane < G && head < k) {
      lab = ll[head];
      if (lab != 0xFFFFFFFFFFFFFFFFull) od = f2ord(dl[head]);
    }
    uint32_t bd = od, bl = lane;
    uint64_t blab = lab;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      uint32_t xd = __shfl_xor_sync(0xffffffffu, bd, o);
      uint64_t xl = __shfl_xor_sync(0xffffffffu, blab, o);
      uint32_t xn = __shfl_xor_sync(0xffffffffu, bl, o);
      if (xd < bd || (xd == bd && (xl < blab || (xl == blab && xn < bl)))) bd = xd, blab = xl, bl = xn;
    }
    bool ok = blab != 0xFFFFFFFFFFFFFFFFull;
    if (lane == 0) {
      out_labels[q * k + i] = ok ? blab : 0xFFFFFFFFFFFFFFFFull;
      if (out_dists) out_dists[q * k + i] = ok ? ord2f(bd) : INFINITY;
    }
    if (!ok) {
      for (uint32_t j = i + 1 + lane; j < k; j += 32) {
        out_labels[q * k + j] = 0xFFFFFFFFFFFFFFFFull;
        if (out_dists) out_dists[q * k + j] = INFINITY;
      }
      break;
    }
    found++;
    if (lane == bl) head++;
  }
  if (lane == 0 && out_counts) out_counts[q] = found;
}


}  // namespace ehb
