// Error channel and device queries of the C-ABI library.
#include "common.cuh"
#include "../../include/swnerf_b200.h"
#include <stdarg.h>
#include <cuda.h>
#include <mutex>

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

int encode_u8_tensor_map(void* map, const void* base, int ndim, const unsigned long long* dims,
                         const unsigned long long* strides, const unsigned int* box) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (!encode) return set_err(SWNERF_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
  cuuint64_t d[3], st[2];
  cuuint32_t b[3], es[3] = {1, 1, 1};
  for (int i = 0; i < ndim; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < ndim; ++i) st[i] = strides[i];
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)ndim,
                      const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(SWNERF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return SWNERF_OK;
}

int const_slot_acquire(int family, int nslots, cudaStream_t s) {
  constexpr int kMaxDev = 64, kFam = 2, kSlots = 8;
  struct Slot { cudaStream_t owner; bool used; unsigned long long stamp; };
  static std::mutex mu;
  static Slot slots[kFam][kMaxDev][kSlots] = {};
  static cudaEvent_t ev[kMaxDev] = {};
  static unsigned long long clock_ = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev || family < 0 || family >= kFam || nslots > kSlots) return -1;
  std::lock_guard<std::mutex> lock(mu);
  Slot* sl = slots[family][dev];
  int k = -1;
  for (int i = 0; i < nslots; ++i) if (sl[i].used && sl[i].owner == s) k = i;
  if (k < 0) {
    for (int i = 0; i < nslots; ++i) if (!sl[i].used) { k = i; break; }
    if (k < 0) {
      k = 0;
      for (int i = 1; i < nslots; ++i) if (sl[i].stamp < sl[k].stamp) k = i;
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(s, &cap);
      if (cap == cudaStreamCaptureStatusNone) {
        if (!ev[dev]) cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming);
        cudaEventRecord(ev[dev], sl[k].owner);
        cudaStreamWaitEvent(s, ev[dev], 0);
      }
    }
    sl[k].owner = s; sl[k].used = true;
  }
  sl[k].stamp = ++clock_;
  return k;
}

bool once_per_device(int id) {
  constexpr int kMaxDev = 64;
  static std::mutex mu;
  static bool done[ONCE_COUNT][kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev || id < 0 || id >= ONCE_COUNT) return true;
  std::lock_guard<std::mutex> lock(mu);
  if (done[id][dev]) return false;
  done[id][dev] = true;
  return true;
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
