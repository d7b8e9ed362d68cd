// Shared helpers for the swnerf_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SWNERF_OK 0
#define SWNERF_ERR_ARG 1
#define SWNERF_ERR_CUDA 2
#define SWNERF_ERR_UNSUPPORTED 3

namespace swnerf {

// thread-local last error message (include/swnerf_b200.h: swnerf_last_error)
char* err_buf();
int set_err(int code, const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

// Rotating copies of a per-network parameter block in constant memory (per device and per `family` of kernels).  A
// stream keeps the slot it used last (stream order makes re-staging safe); a stream without one takes the least
// recently assigned slot and first waits (event) for the work of the stream that owned it, so launches on different
// streams (coarse and fine network, or two callers) never share a block.  Returns the slot, or -1 without a device.
// (While a stream is being captured into a CUDA graph no cross-stream wait is inserted: captures use one stream.)
int const_slot_acquire(int family, int nslots, cudaStream_t s);
constexpr int CONST_FAMILY_FWD_F32 = 0, CONST_FAMILY_BWD_WRGB = 1;

// One-time, PER-DEVICE setup (cudaFuncSetAttribute opt-ins, __constant__ uploads): true exactly once for each
// (id, current device) pair, whatever thread asks; callers run their setup when it returns true.  Ids below.
bool once_per_device(int id);
enum OnceId {
  ONCE_FWD4 = 0, ONCE_FWD1_BASE, ONCE_FWD1_1, ONCE_FWD1_2, ONCE_FWD1_3, ONCE_FWD1_4, ONCE_FWD1_5, ONCE_BWD_BASE, ONCE_BWD_DATA_PAIR, ONCE_BWD_WEIGHT_PAIR,
  ONCE_HGEMM, ONCE_HGEMM_WGRAD, ONCE_RESAMPLE64Q, ONCE_RESAMPLE64Q_CHECK, ONCE_FWDX, ONCE_BWDX, ONCE_BWD_LW, ONCE_BWD_INPUT_BASE, ONCE_BWD_INPUT_1, ONCE_BWD_INPUT_2, ONCE_BWD_INPUT_3, ONCE_BWD_INPUT_4, ONCE_COUNT
};

// u8 tensor map (1..3 dims, no swizzle / interleave) through the driver entry point fetched at run time, so the
// library does not link libcuda.  `map` points to a 128-byte CUtensorMap.  strides[] has ndim - 1 entries (bytes).
int encode_u8_tensor_map(void* map, const void* base, int ndim, const unsigned long long* dims,
                         const unsigned long long* strides, const unsigned int* box);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define SW_REQUIRE(cond, ...)                                          \
  do {                                                                 \
    if (!(cond)) return swnerf::set_err(SWNERF_ERR_ARG, __VA_ARGS__);  \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive warp scan (product / sum)
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

}  // namespace swnerf
