#include "hv_tc_win16.cuh"
namespace hv {
size_t wattn_tc256_bwd_workspace_bytes(const Geom& g) { return 16; }
int wattn_tc256_bwd(const Geom& g, const void* qkv, const void* out, const void* dout, const float* stats,
                    const float* bias_table, const float* tau, void* dqkv, float* dbias_table, float* dtau, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  HV_FAIL(HV_ERR_SHAPE, "wattn_tc256_bwd: not built yet");
}
}  // namespace hv
