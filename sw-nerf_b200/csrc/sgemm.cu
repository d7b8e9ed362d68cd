// fp32 SIMT GEMM used by the fp32-accumulate CHECK path of the MLP (the <=1e-5 parity mode) and by
// network shapes the fused tcgen05 kernels are not instantiated for (D-NeRF deformation net,
// MultiRes PE widths, use_viewdirs=False).  True fp32 FMA, like the reference's cuBLAS SGEMM with
// allow_tf32=False (model.py:43-57 nn.Linear).  Three operand layouts cover forward (x W^T),
// data-gradient (dy W) and weight-gradient (dy^T x, split over samples with red.add).
#include "common.cuh"
#include "../../include/swnerf_b200.h"

namespace swnerf {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NT = 256;

struct GemmArgs {
  const float* A; int64_t sam, sak;   // A(m,k) = A[m*sam + k*sak]
  const float* B; int64_t sbk, sbn;   // B(k,n) = B[k*sbk + n*sbn]
  float* C; int64_t ldc;
  int64_t M, N, K;
  const float* bias;                  // [N] or null
  const float* mask; int64_t ldmask;  // C *= (mask[m,n] > 0) or null
  int accumulate;
  int act;                            // epilogue activation: 0 none, 1 ReLU, 2 ELU(alpha=1)
  int mask_kind;                      // 0: C *= (mask > 0) (ReLU'), 1: C *= (mask > 0 ? 1 : mask + 1) (ELU' from its output)
  int64_t k_per_split;                // split-K slab (atomics when gridDim.z > 1)
};

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(NT) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = min(g.K, kbeg + g.k_per_split);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int it = 0; it < (BM * BK) / NT; ++it) {
      int e = it * NT + tid;
      int mm, kk;
      if (A_KCONTIG) { mm = e / BK; kk = e % BK; } else { kk = e / BM; mm = e % BM; }
      int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < kend) ? __ldg(g.A + m * g.sam + k * g.sak) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < (BN * BK) / NT; ++it) {
      int e = it * NT + tid;
      int nn, kk;
      if (B_KCONTIG) { nn = e / BK; kk = e % BK; } else { kk = e / BN; nn = e % BN; }
      int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < kend) ? __ldg(g.B + k * g.sbk + n * g.sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t n = n0 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = acc[i][j];
      float* c = g.C + m * g.ldc + n;
      if (split) { atomicAdd(c, v); continue; }
      if (g.bias) v += __ldg(g.bias + n);
      if (g.accumulate) v += *c;
      if (g.act == 1) v = fmaxf(v, 0.f);
      else if (g.act == 2) v = v > 0.f ? v : expm1f(v);
      if (g.mask) {
        const float y = __ldg(g.mask + m * g.ldmask + n);
        v = (y > 0.f) ? v : (g.mask_kind == 1 ? v * (y + 1.f) : 0.f);
      }
      *c = v;
    }
  }
}

// d_out[m,n] = d[m,n] * act'(y[m,n]) from the activation's OUTPUT y: kind 0 ReLU (y > 0), kind 1 ELU (y > 0 ? 1 : y + 1)
__global__ void act_bwd_kernel(const float* __restrict__ d, int64_t ldd, const float* __restrict__ y, int64_t ldy,
                               int64_t rows, int cols, int kind, float* __restrict__ out, int64_t ldo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t m = idx / cols;
  const int n = (int)(idx - m * cols);
  const float yv = __ldg(y + m * ldy + n), dv = __ldg(d + m * ldd + n);
  out[m * ldo + n] = (yv > 0.f) ? dv : (kind == 1 ? dv * (yv + 1.f) : 0.f);
}

// out[n] (+)= sum_m x[m*ld + n]
__global__ void colsum_kernel(const float* __restrict__ x, int64_t ld, int64_t M, int N, float* __restrict__ out,
                              int64_t rows_per_block) {
  __shared__ float red[8][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  int64_t mb = (int64_t)blockIdx.y * rows_per_block;
  int64_t me = min(M, mb + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int64_t m = mb + threadIdx.y; m < me; m += 8) s += __ldg(x + m * ld + n);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

}  // namespace swnerf

using namespace swnerf;

extern "C" {

int swnerf_sgemm(int op, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                 int64_t N, int64_t K, const float* bias, int accumulate, int act_flags, const float* mask,
                 int64_t ldmask, void* stream) {
  const int relu = act_flags & 3, mask_kind = (act_flags >> 4) & 3;
  SW_REQUIRE(A && B && C, "sgemm: null pointer");
  SW_REQUIRE(relu <= 2 && mask_kind <= 1 && (act_flags & ~0x33) == 0, "sgemm: bad act_flags");
  SW_REQUIRE(op >= 0 && op <= 2, "sgemm: op must be 0 (x W^T), 1 (dy W) or 2 (dy^T x)");
  SW_REQUIRE(M >= 0 && N >= 0 && K >= 0, "sgemm: negative size");
  if (M == 0 || N == 0) return SWNERF_OK;
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.mask = mask; g.ldmask = ldmask; g.accumulate = accumulate; g.act = relu; g.mask_kind = mask_kind;
  g.k_per_split = K > 0 ? K : 1;
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), 1);
  cudaStream_t s = (cudaStream_t)stream;
  if (op == 0) {            // C[M,N] = A[M,K] . B[N,K]^T
    g.sam = lda; g.sak = 1; g.sbk = 1; g.sbn = ldb;
    sgemm_kernel<true, true><<<grid, NT, 0, s>>>(g);
  } else if (op == 1) {     // C[M,N] = A[M,K] . B[K,N]
    g.sam = lda; g.sak = 1; g.sbk = ldb; g.sbn = 1;
    sgemm_kernel<true, false><<<grid, NT, 0, s>>>(g);
  } else {                  // C[M,N] (+)= A[K,M]^T . B[K,N], K = samples: split over K with red.add
    SW_REQUIRE(!bias && !relu && !mask, "sgemm: op 2 has no epilogue");
    g.sam = 1; g.sak = lda; g.sbk = ldb; g.sbn = 1;
    int64_t tiles = (int64_t)grid.x * grid.y;
    int64_t want = (4LL * sm_count() + tiles - 1) / tiles;
    int64_t max_split = (K + 4 * BK - 1) / (4 * BK);
    int64_t split = want < max_split ? want : max_split;
    if (split < 1) split = 1;
    int64_t kps = (K + split - 1) / split;
    kps = (kps + BK - 1) / BK * BK;
    split = (K + kps - 1) / kps;
    g.k_per_split = kps;
    grid.z = (unsigned)split;
    if (split > 1 && !accumulate) {
      SW_REQUIRE(ldc == N, "sgemm: op 2 without accumulate needs a dense C");
      cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * (size_t)N, s);
    } else if (split > 1) {
      g.accumulate = 1;
    }
    sgemm_kernel<false, false><<<grid, NT, 0, s>>>(g);
  }
  return check_launch("sgemm");
}

int swnerf_act_bwd(const float* d, int64_t ldd, const float* y, int64_t ldy, int64_t rows, int cols, int kind,
                   float* out, int64_t ldo, void* stream) {
  SW_REQUIRE(d && y && out, "act_bwd: null pointer");
  SW_REQUIRE(kind == 0 || kind == 1, "act_bwd: kind must be 0 (ReLU) or 1 (ELU)");
  SW_REQUIRE(rows >= 0 && cols >= 0, "act_bwd: negative size");
  if (rows == 0 || cols == 0) return SWNERF_OK;
  const int64_t n = rows * cols;
  act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, ldd, y, ldy, rows, cols, kind, out, ldo);
  return check_launch("act_bwd");
}

int swnerf_colsum(const float* x, int64_t ld, int64_t rows, int cols, float* out, int accumulate, void* stream) {
  SW_REQUIRE(x && out, "colsum: null pointer");
  if (cols == 0) return SWNERF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * (size_t)cols, s);
  if (rows == 0) return SWNERF_OK;
  int64_t rpb = 2048;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + rpb - 1) / rpb));
  colsum_kernel<<<grid, dim3(32, 8), 0, s>>>(x, ld, rows, cols, out, rpb);
  return check_launch("colsum");
}

}  // extern "C"
