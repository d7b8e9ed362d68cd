// placeholder until the tcgen05 kernels land (keeps every declared symbol exported)
#include "common.cuh"
#include "../../include/swnerf_b200.h"
using namespace swnerf;
extern "C" {
int64_t swnerf_tc_packed_bytes(void) { return 0; }
int64_t swnerf_tc_packed_t_bytes(void) { return 0; }
int64_t swnerf_tc_workspace_bytes(int64_t, int) { return 0; }
int swnerf_tc_pack_weights(const float* const*, void*, void*) { return set_err(SWNERF_ERR_UNSUPPORTED, "tc path not built"); }
int swnerf_tc_pack_weights_t(const float* const*, void*, void*) { return set_err(SWNERF_ERR_UNSUPPORTED, "tc path not built"); }
int swnerf_tc_mlp_fwd(const float*, int, int, const float*, int64_t, int, const void*, float*, void*, int, void*) { return set_err(SWNERF_ERR_UNSUPPORTED, "tc path not built"); }
int swnerf_tc_mlp_bwd(const float*, int64_t, int, const void*, const void*, const float* const*, void*, float* const*, float, void*) { return set_err(SWNERF_ERR_UNSUPPORTED, "tc path not built"); }
}
