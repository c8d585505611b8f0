// tcgen05 tensor-core path for the dense node / edge transforms (placeholder until validated).
#include "common.cuh"
#include "resgcn.cuh"

namespace gg {
int gemm_tc_prepare_weights(gg_context*, const std::vector<float>&) { return GG_OK; }
bool gemm_tc_supported(const gg_context*, int, int, int) { return false; }
int gemm_tc(gg_context*, cudaStream_t, int, const float*, const float*, float*, const int*, long long,
            int, int, int, int) {
  set_error("gemm_tc: not built");
  return GG_ERR_STATE;
}
}  // namespace gg
