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
// mode 2: as mode 0 with A written into tensor memory by the four warps (tcgen05.st) and read from there
__global__ void __launch_bounds__(128, 1)
selftest_gemm_kernel(int mode, const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, int a_bytes,
                     int b_bytes, int N, int K, float* __restrict__ D, const float* __restrict__ A_f32) {
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
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (mode == 2) {
    const int row = warp * 32 + lane;
    for (int k0 = 0; k0 < K; k0 += 32) {
      uint32_t pk[16];
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_half2(A_f32[(size_t)row * K + k0 + 2 * i], A_f32[(size_t)row * K + k0 + 2 * i + 1]);
      tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + k0 / 2, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_full, (uint32_t)(a_bytes + b_bytes));
    bulk_g2s(sa, a_img, (uint32_t)a_bytes, &bar_full);
    bulk_g2s(sb, b_img, (uint32_t)b_bytes, &bar_full);
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    uint32_t acc = 0;
    if (mode == 2) {
      const uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
      for (int kb = 0; kb < K / 64; ++kb) {
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t bd = umma_desc_kmajor(smem_u32(sb) + kb * N * 128 + ks * 32);
          umma_f16_ts(tmem, tmem + 256 + kb * 32 + ks * 8, bd, idesc, acc);
          acc = 1;
        }
      }
    } else if (mode == 0) {
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
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Issue-rate probe: one CTA per SM issues `iters` groups of four back-to-back K=16 MMAs into one accumulator and
// reports clock cycles per MMA.  variant & 1: A from tensor memory instead of shared memory; (variant >> 4) & 3:
// tcgen05.commit per group.
__global__ void __launch_bounds__(128, 1) probe_kernel(int variant, int N, int iters, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done, bar_x[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const int ncommit = (variant >> 4) & 3;
  const bool ts = variant & 1;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_x[0], 1); mbar_init(&bar_x[1], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (!ts) umma_f16(tmem, umma_desc_kmajor(sa + ks * 32), umma_desc_kmajor(sb + ks * 32), idesc, 1u);
          else umma_f16_ts(tmem, tmem + 256 + ks * 8, umma_desc_kmajor(sb + ks * 32), idesc, 1u);
        }
        if (ncommit > 0) umma_commit(&bar_x[0]);
        if (ncommit > 1) umma_commit(&bar_x[1]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar_done);
    __syncwarp();
    mbar_wait(&bar_done, 0);
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = (float)(t1 - t0) / (4.f * iters);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// CTA-pair GEMM (cta_group::2): D[256 x N] = A[256 x K] . B[N x K]^T.  CTA r of the pair holds rows 128r.. of A and
// rows (N/2)r.. of B; the leader issues M=256 MMAs, each CTA reads its 128 rows of D from its own tensor memory.
// iters > 1 repeats the K loop (accumulating) to measure cycles per MMA; iters >> 20 = multicast commits per group.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_gemm_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, int N, int K, float* __restrict__ D,
                 int iters, float* __restrict__ cycles, int mn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // K-major: per 64-wide K block an image of [rows x 64]; MN-major (mn): per 64-wide M/N block an image of [K x 64]
  const int kb_n = K / 64;
  const int a_bytes = mn ? 2 * K * 128 : kb_n * 128 * 128;
  const int b_bytes = mn ? (N / 128) * K * 128 : kb_n * (N / 2) * 128;
  uint8_t* sa = smem;
  uint8_t* sb = smem + a_bytes;
  __shared__ uint64_t bar_full, bar_peer, bar_done, bar_x[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int ncommit = iters >> 20;
  iters &= 0xfffff;

  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_peer, 1);
    mbar_init(&bar_done, 1);
    mbar_init(&bar_x[0], 1);
    mbar_init(&bar_x[1], 1);
    mbar_fence_init();
  }
  cluster_sync_all();
  if (warp == 0) tmem_alloc_pair<256>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 1) {
    if (lane == 0) {
      // a bulk copy can only signal an mbarrier of the CTA it writes to, so the peer relays "my operands have landed"
      mbar_expect_tx(&bar_full, (uint32_t)(a_bytes + b_bytes));
      bulk_g2s(sa, a_img + (size_t)rank * a_bytes, (uint32_t)a_bytes, &bar_full);
      bulk_g2s(sb, b_img + (size_t)rank * b_bytes, (uint32_t)b_bytes, &bar_full);
      mbar_wait(&bar_full, 0);
      if (rank == 1) mbar_arrive_remote(mapa_u32(smem_u32(&bar_peer), 0));
    }
    __syncwarp();
    if (rank == 0) {
      mbar_wait_cluster(&bar_peer, 0);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_f16(256, N, mn, mn);
      const uint32_t sa32 = smem_u32(sa), sb32 = smem_u32(sb);
      long long t0 = clock64();
      if (mn) {
        // each CTA: A = its 128 M-columns (two [K x 64] blocks, LBO = block stride), B = its N/2 columns likewise
        if (elect_one()) {
          for (int ks = 0; ks < K / 16; ++ks)
            umma_f16_pair(tmem, umma_desc_mnmajor(sa32 + ks * 2048, (uint32_t)K * 128),
                          umma_desc_mnmajor(sb32 + ks * 2048, (uint32_t)K * 128), idesc, ks ? 1u : 0u);
        }
        __syncwarp();
      } else
      for (int it = 0; it < iters; ++it) {
        for (int kb = 0; kb < kb_n; ++kb) {
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_pair(tmem, umma_desc_kmajor(sa32 + kb * 128 * 128 + ks * 32),
                            umma_desc_kmajor(sb32 + kb * (N / 2) * 128 + ks * 32), idesc, (it | kb | ks) ? 1u : 0u);
            if (ncommit > 0) umma_commit_pair(&bar_x[0], 3);
            if (ncommit > 1) umma_commit_pair(&bar_x[1], 3);
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit_pair(&bar_done, 3);
      __syncwarp();
      mbar_wait(&bar_done, 0);
      long long t1 = clock64();
      if (lane == 0 && cycles) cycles[pair] = (float)(t1 - t0) / (4.f * kb_n * iters);
    }
  }
  __syncwarp();
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  if (D && pair == 0) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      int row = rank * 128 + warp * 32 + lane;
      for (int i = 0; i < 32; ++i)
        if (c0 + i < N) D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<256>(tmem);
}

}  // namespace swnerf

using namespace swnerf;

extern "C" int swnerf_tc_probe(int variant, int N, int iters, float* cycles_per_mma, void* stream) {
  SW_REQUIRE(cycles_per_mma && variant >= 0 && N % 16 == 0 && N >= 16 && N <= 256 && iters > 0, "tc_probe: bad arguments");
  const int smem = 16384 + 32768 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<sm_count(), 128, smem, (cudaStream_t)stream>>>(variant, N, iters, cycles_per_mma);
  return check_launch("tc_probe");
}

extern "C" int swnerf_tc_selftest_pair(const float* A, const float* B, float* D, int N, int K, int iters, int n_pairs,
                                       float* cycles_per_mma, void* scratch, void* stream) {
  SW_REQUIRE(A && B && scratch, "tc_selftest_pair: null pointer");
  if (iters == 0) {
    // MN-major form (the weight-gradient shape): D[256 x N] = P[K x 256]^T . Q[K x N], A = P, B = Q, N in {128, 256}
    SW_REQUIRE((N == 128 || N == 256) && (K == 64 || K == 128) && D, "tc_selftest_pair: MN-major form needs N in {128,256}, K in {64,128}");
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
    // selftest_pack_kernel writes one [rows x 64] image per 64-column block: P -> 4 blocks, Q -> N/64 blocks; CTA r
    // takes blocks 2r, 2r+1 of P and blocks r*(N/128).. of Q, which are contiguous in that order
    const int a_all = 4 * K * 128, b_all = (N / 64) * K * 128;
    selftest_pack_kernel<<<(K * 256 + 255) / 256, 256, 0, s>>>(A, K, 256, sc);
    selftest_pack_kernel<<<(K * N + 255) / 256, 256, 0, s>>>(B, K, N, sc + a_all);
    size_t smem = (size_t)a_all / 2 + b_all / 2 + 2048;
    cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pair_gemm_kernel<<<2, 128, smem, s>>>(sc, sc + a_all, N, K, D, 1, nullptr, 1);
    return check_launch("tc_selftest_pair");
  }
  SW_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256, "tc_selftest_pair: N must be a multiple of 32 in [32, 256]");
  SW_REQUIRE(K % 64 == 0 && K >= 64 && K <= 256, "tc_selftest_pair: K in {64,128,192,256}");
  SW_REQUIRE((iters & 0xfffff) >= 1 && n_pairs >= 1, "tc_selftest_pair: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  const int a_bytes = (K / 64) * 128 * 128, b_half = (K / 64) * (N / 2) * 128;
  // A: two [128 x K] row tiles (every pair reads the same two); B: two [N/2 x K] halves
  for (int r = 0; r < 2; ++r) {
    selftest_pack_kernel<<<(128 * K + 255) / 256, 256, 0, s>>>(A + (size_t)r * 128 * K, 128, K, sc + (size_t)r * a_bytes);
    selftest_pack_kernel<<<((N / 2) * K + 255) / 256, 256, 0, s>>>(B + (size_t)r * (N / 2) * K, N / 2, K,
                                                                   sc + 2 * (size_t)a_bytes + (size_t)r * b_half);
  }
  size_t smem = (size_t)a_bytes + b_half + 2048;
  cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pair_gemm_kernel<<<2 * n_pairs, 128, smem, s>>>(sc, sc + 2 * (size_t)a_bytes, N, K, D, iters, cycles_per_mma, 0);
  return check_launch("tc_selftest_pair");
}

extern "C" int swnerf_tc_selftest(int mode, const float* A, const float* B, float* D, int N, int K, void* scratch,
                                  void* stream) {
  SW_REQUIRE(A && B && D && scratch, "tc_selftest: null pointer");
  SW_REQUIRE(mode >= 0 && mode <= 2, "tc_selftest: mode must be 0, 1 or 2");
  SW_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256, "tc_selftest: N must be a multiple of 16 in [16, 256]");
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  int a_bytes, b_bytes;
  if (mode == 0 || mode == 2) {
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
  selftest_gemm_kernel<<<1, 128, smem, s>>>(mode, sc, sc + a_bytes, a_bytes, b_bytes, N, K, D, A);
  return check_launch("tc_selftest");
}
