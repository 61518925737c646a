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
#include "build_impl.cuh"

namespace ehb {

cudaError_t launch_build_batch(EHB_BUILD_ARGS) {
  if (b == 0) return cudaSuccess;
  switch (bg.g.dpad) {
    case 32: return launch_build_d32(EHB_BUILD_PASS);
    case 64: return launch_build_d64(EHB_BUILD_PASS);
    case 128: return launch_build_d128(EHB_BUILD_PASS);
    case 256: return launch_build_d256(EHB_BUILD_PASS);
    case 384: return launch_build_d384(EHB_BUILD_PASS);
    case 512: return launch_build_d512(EHB_BUILD_PASS);
    case 768: return launch_build_d768(EHB_BUILD_PASS);
    case 1024: return launch_build_d1024(EHB_BUILD_PASS);
    case 1536: return launch_build_d1536(EHB_BUILD_PASS);
    case 2048: return launch_build_d2048(EHB_BUILD_PASS);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ehb
