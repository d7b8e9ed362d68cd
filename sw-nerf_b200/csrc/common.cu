// Error channel and device queries of the C-ABI library.
#include "common.cuh"
#include "../../include/swnerf_b200.h"
#include <stdarg.h>

namespace swnerf {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;   // process-wide: backward launches happen on autograd's engine thread

char* err_buf() { return g_err; }

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(SWNERF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SWNERF_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace swnerf

extern "C" {

const char* swnerf_last_error(void) { return swnerf::err_buf(); }

int64_t swnerf_launch_count(int reset) {
  int64_t n = (int64_t)__atomic_load_n(&swnerf::g_launches, __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&swnerf::g_launches, 0ULL, __ATOMIC_RELAXED);
  return n;
}

int swnerf_version(void) { return SWNERF_B200_VERSION; }

int swnerf_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return swnerf::set_err(SWNERF_ERR_CUDA, "no CUDA device");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) return swnerf::set_err(SWNERF_ERR_UNSUPPORTED, "swnerf_b200 needs sm_100a, found sm_%d%d", major, minor);
  return SWNERF_OK;
}

}  // extern "C"
