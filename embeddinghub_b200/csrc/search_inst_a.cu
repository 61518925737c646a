// K2 instantiations (generated list of row shapes; see search_impl.cuh)
#include "search_impl.cuh"
namespace ehb {
cudaError_t launch_search_d32(EHB_SEARCH_ARGS) { return launch_search_kpl<8, 1>(EHB_SEARCH_PASS); }
cudaError_t launch_search_d64(EHB_SEARCH_ARGS) { return launch_search_kpl<8, 2>(EHB_SEARCH_PASS); }
cudaError_t launch_search_d128(EHB_SEARCH_ARGS) { return launch_search_kpl<8, 4>(EHB_SEARCH_PASS); }
}  // namespace ehb
