// K5 instantiations (generated list of row shapes; see build_impl.cuh)
#include "build_impl.cuh"
namespace ehb {
cudaError_t launch_build_d1024(EHB_BUILD_ARGS) { return launch_build_t<32, 8>(EHB_BUILD_PASS); }
cudaError_t launch_build_d1536(EHB_BUILD_ARGS) { return launch_build_t<32, 12>(EHB_BUILD_PASS); }
cudaError_t launch_build_d2048(EHB_BUILD_ARGS) { return launch_build_t<32, 16>(EHB_BUILD_PASS); }
}  // namespace ehb
