// placeholder -- replaced by the tcgen05 kernels
#include "zb200_common.cuh"
namespace zb200 {
int init_tensor_maps(zb200_plan*) { return ZB200_OK; }
bool tc_supported(const zb200_plan*) { return false; }
int project_tc(const zb200_plan*, const float*, int64_t, int, int, void*, void*, const float*, const uint8_t*, int, int,
               cudaStream_t) {
    set_error("tcgen05 projection not built");
    return ZB200_EUNSUP;
}
}  // namespace zb200
