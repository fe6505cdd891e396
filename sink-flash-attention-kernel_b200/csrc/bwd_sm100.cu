// placeholder until the tcgen05 backward lands (next commit)
#include "common.cuh"
namespace sfa {
bool tc_bwd_supported(const AttnParams&, int) { return false; }
cudaError_t tc_bwd(const AttnParams&, int, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace sfa
