// placeholder: backward of the fused path (implemented next)
#include "common.cuh"
#include "../../include/swnerf_b200.h"
using namespace swnerf;
extern "C" {
int64_t swnerf_tc_packed_t_bytes(void) { return 0; }
int swnerf_tc_pack_weights_t(const float* const*, void*, void*) { return set_err(SWNERF_ERR_UNSUPPORTED, "tc backward not built"); }
int swnerf_tc_mlp_bwd(const float*, int64_t, int, const void*, const void*, const float* const*, void*, float* const*, float, void*) { return set_err(SWNERF_ERR_UNSUPPORTED, "tc backward not built"); }
}
