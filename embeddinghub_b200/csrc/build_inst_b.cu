// K5 instantiations (generated list of row shapes; see build_impl.cuh)
#include "build_impl.cuh"
namespace ehb {
cudaError_t launch_build_d256(EHB_BUILD_ARGS) { return launch_build_t<8, 8>(EHB_BUILD_PASS); }
cudaError_t launch_build_d384(EHB_BUILD_ARGS) { return launch_build_t<32, 3>(EHB_BUILD_PASS); }
}  // namespace ehb
