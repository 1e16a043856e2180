// placeholder: tensor-core kernel comes next
#include "hv_common.cuh"
namespace hv {
bool wattn_mma64_supported(const Geom& g, int dtype) { (void)g; (void)dtype; return false; }
size_t wattn_mma64_bwd_workspace_bytes(const Geom& g) { (void)g; return 16; }
int wattn_mma64_fwd(const Geom&, const void*, const float*, const float*, const float*, int, void*, float*, cudaStream_t) { HV_FAIL(HV_ERR_SHAPE, "not built"); }
int wattn_mma64_bwd(const Geom&, const void*, const void*, const void*, const float*, const float*, const float*, const float*, int, void*, float*, float*, void*, size_t, cudaStream_t) { HV_FAIL(HV_ERR_SHAPE, "not built"); }
}
