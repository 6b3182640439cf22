// flow_tc.cu -- tensor-core (tcgen05) conditioner path.  Placeholder until the kernel lands.
#include "rnf_common.cuh"
namespace rnf {
bool flow_tc_supported(const rnf_flow*) { return false; }
cudaError_t launch_flow_tc(const FlowArgs&, bool, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace rnf
