// Single-tile tcgen05 GEMM used by the tests to pin the operand image, the UMMA descriptors (K-major
// and MN-major), the bulk-copy/mbarrier handshake and the TMEM read-back on real hardware, in
// isolation from the fused kernels that are built from the same pieces.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/swnerf_b200.h"

namespace swnerf {
using namespace tc;

// fp32 row-major [rows x cols] -> fp16 tile images, one [rows x 64] image per 64-column block
__global__ void selftest_pack_kernel(const float* __restrict__ src, int rows, int cols, uint8_t* __restrict__ dst) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int nblk = (cols + 63) / 64;
  if (idx >= rows * nblk * 64) return;
  int c = idx % (nblk * 64), r = idx / (nblk * 64);
  float v = (c < cols) ? src[(size_t)r * cols + c] : 0.f;
  int blk = c >> 6;
  *reinterpret_cast<__half*>(dst + (size_t)blk * rows * 128 + tile_off(r, c & 63)) = __float2half_rn(v);
}

// mode 0: D[128 x N] = A[128 x K] . B[N x K]^T     (both operands K-major; K <= 256)
// mode 1: D[128 x N] = P[K x 128]^T . Q[K x N]      (both operands MN-major; K = samples = 128)
__global__ void __launch_bounds__(128, 1)
selftest_gemm_kernel(int mode, const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, int a_bytes,
                     int b_bytes, int N, int K, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) & ~1023);
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_full, (uint32_t)(a_bytes + b_bytes));
    bulk_g2s(sa, a_img, (uint32_t)a_bytes, &bar_full);
    bulk_g2s(sb, b_img, (uint32_t)b_bytes, &bar_full);
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    uint32_t acc = 0;
    if (mode == 0) {
      const uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
      for (int kb = 0; kb < K / 64; ++kb) {
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t ad = umma_desc_kmajor(smem_u32(sa) + kb * 128 * 128 + ks * 32);
          uint64_t bd = umma_desc_kmajor(smem_u32(sb) + kb * N * 128 + ks * 32);
          umma_f16(tmem, ad, bd, idesc, acc);
          acc = 1;
        }
      }
    } else {
      const uint32_t idesc = umma_idesc_f16(128, N, 1, 1);
      for (int ks = 0; ks < K / 16; ++ks) {     // 16 samples = two 8-row atoms per k-step
        uint64_t ad = umma_desc_mnmajor(smem_u32(sa) + ks * 2048, (uint32_t)K * 128);
        uint64_t bd = umma_desc_mnmajor(smem_u32(sb) + ks * 2048, (uint32_t)K * 128);
        umma_f16(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    umma_commit(&bar_done);
  }
  __syncwarp();
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    int row = warp * 32 + lane;
    for (int i = 0; i < 32; ++i)
      if (c0 + i < N) D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace swnerf

using namespace swnerf;

extern "C" int swnerf_tc_selftest(int mode, const float* A, const float* B, float* D, int N, int K, void* scratch,
                                  void* stream) {
  SW_REQUIRE(A && B && D && scratch, "tc_selftest: null pointer");
  SW_REQUIRE(mode == 0 || mode == 1, "tc_selftest: mode must be 0 or 1");
  SW_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256, "tc_selftest: N must be a multiple of 16 in [16, 256]");
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  int a_bytes, b_bytes;
  if (mode == 0) {
    SW_REQUIRE(K % 64 == 0 && K >= 64 && K <= 256, "tc_selftest: mode 0 needs K in {64,128,192,256}");
    a_bytes = (K / 64) * 128 * 128;
    b_bytes = (K / 64) * N * 128;
    selftest_pack_kernel<<<(128 * K + 255) / 256, 256, 0, s>>>(A, 128, K, sc);
    selftest_pack_kernel<<<(N * K + 255) / 256, 256, 0, s>>>(B, N, K, sc + a_bytes);
  } else {
    SW_REQUIRE(K == 128, "tc_selftest: mode 1 needs K == 128 samples");
    a_bytes = 2 * K * 128;            // P[K x 128]: two 64-channel blocks of K rows
    b_bytes = ((N + 63) / 64) * K * 128;
    selftest_pack_kernel<<<(K * 128 + 255) / 256, 256, 0, s>>>(A, K, 128, sc);
    selftest_pack_kernel<<<(K * ((N + 63) / 64) * 64 + 255) / 256, 256, 0, s>>>(B, K, N, sc + a_bytes);
  }
  size_t smem = (size_t)((a_bytes + 1023) & ~1023) + b_bytes + 2048;
  cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  selftest_gemm_kernel<<<1, 128, smem, s>>>(mode, sc, sc + a_bytes, a_bytes, b_bytes, N, K, D);
  return check_launch("tc_selftest");
}
