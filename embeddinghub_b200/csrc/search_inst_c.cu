// K2 instantiations (generated list of row shapes; see search_impl.cuh)
#include "search_impl.cuh"
namespace ehb {
cudaError_t launch_search_d512(EHB_SEARCH_ARGS) { return launch_search_kpl<32, 4>(EHB_SEARCH_PASS); }
cudaError_t launch_search_d768(EHB_SEARCH_ARGS) { return launch_search_kpl<32, 6>(EHB_SEARCH_PASS); }
}  // namespace ehb
