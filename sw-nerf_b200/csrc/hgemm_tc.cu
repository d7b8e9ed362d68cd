// Layer-at-a-time tcgen05 GEMM with the nn.Linear epilogue, for the network shapes the fused kernels (mlp_tc*.cu) are
// not instantiated for: T-NeRF (W = 128, ELU, model.py:152-210), MultiRes D-NeRF levels with other encoding widths
// (multires_dnerf.py:665), networks without view directions.  Same contract as swnerf_sgemm ops 0 / 1 (sgemm.cu), but
// the contraction runs on the tensor cores with fp16 operands and fp32 accumulation - the precision of the fused path:
//     C[M,N] = act( A[M,K] . B^T + bias ) (+ C) (* act'(mask)),     B = W[N,K]  (forward)  or  W[K,N] read transposed (dgrad)
// A and C stay fp32 row-major in HBM (any row stride, any alignment): the kernel converts a 128-row tile of A to the
// fp16 operand image itself (tc_common.cuh: rows of 128 B, 8-row atoms, 16-B units XOR-swizzled), and builds the image
// of the small weight matrix once per CTA.  Persistent CTAs, 512 threads: warps 8-15 load / convert the next A tile and
// one of their lanes issues the K/16 `tcgen05.mma.cta_group::1.kind::f16` of the tile; warps 0-7 read the fp32
// accumulator (two 256-column TMEM buffers alternate) and run the epilogue, so a tile's stores overlap the next tile's
// loads and MMAs.  HBM-bound by construction (fp32 activations in and out: 4 (K + N) bytes per row).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/swnerf_b200.h"

namespace swnerf {
using namespace tc;

struct HgArgs {
  const float* A; int64_t lda;
  const float* W; int64_t ldw; int trans_w;      // trans_w: B(n,k) = W[k*ldw + n] instead of W[n*ldw + k]
  float* C; int64_t ldc;
  int64_t M; int N, K;
  const float* bias;
  int accumulate, act, mask_kind;
  const float* mask; int64_t ldmask;
  int n_pad, k_chunks;                            // N rounded up to 16; number of 64-column K chunks
  float a_scale;                                  // A is multiplied by a_scale before rounding, the product by 1 / a_scale
  const float* a_scale_dev;                       // if set: the scale is read from device memory (swnerf_pow2_scale)
  int vec_a, vec_c;                               // float4 access allowed (alignment checked on the host)
  int a_bufs;                                     // 2 when two A images fit next to the weight image, else 1
  int stage;                                      // epilogue stores go through a per-warp staging slice (20 KB)
};

constexpr int HG_A_CHUNK = 128 * 128;             // one [128 x 64] fp16 image

__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// 128 rows m0 .. m0+127 of a row-major fp32 matrix (K columns, padded with zeros to kpad; rows beyond M are zero),
// times `scale`, into fp16 operand images [128 x 64] (one per 64 columns) at dst.  NT threads, t = 0 .. NT-1.
// When the column count divides the thread count a thread keeps its column and walks down the rows with eight
// independent loads in flight (the loop is latency-bound otherwise).
// a thread's column c of rows r0, r0 + rstep, ...: UN independent 16-byte loads in flight, then convert and store
template <int UN>
__device__ __forceinline__ void hg_walk_vec(const float* __restrict__ A, int64_t lda, int64_t m0, int64_t M, int c,
                                            bool c_ok, int r0, int rstep, int iters, float scale, uint8_t* dst_chunk) {
  for (int j0 = 0; j0 < iters; j0 += UN) {
    float4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t m = m0 + r0 + (j0 + u) * rstep;
      v[u] = (c_ok && m < M) ? __ldg(reinterpret_cast<const float4*>(A + m * lda + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int r = r0 + (j0 + u) * rstep;
      uint2 p;
      p.x = pack_half2(v[u].x * scale, v[u].y * scale);
      p.y = pack_half2(v[u].z * scale, v[u].w * scale);
      *reinterpret_cast<uint2*>(dst_chunk + tile_off((uint32_t)r, (uint32_t)(c & 63))) = p;
    }
  }
}

template <int NT>
__device__ __forceinline__ void hg_load_tile(const float* __restrict__ A, int64_t lda, int64_t m0, int64_t M, int K,
                                             int kpad, int vec, float scale, uint8_t* dst, int t) {
  if (vec && (NT % (kpad >> 2)) == 0) {
    const int units = kpad >> 2, rstep = NT / units, iters = 128 / rstep;
    const int c = (t % units) << 2, r0 = t / units;
    const bool c_ok = c < K;                               // K is a multiple of 4 on this path
    const uint32_t cbase = (uint32_t)(c >> 6) * HG_A_CHUNK;
    if (iters % 16 == 0) hg_walk_vec<16>(A, lda, m0, M, c, c_ok, r0, rstep, iters, scale, dst + cbase);
    else hg_walk_vec<8>(A, lda, m0, M, c, c_ok, r0, rstep, iters, scale, dst + cbase);
  } else if (vec) {
    const int units = kpad >> 2;
    for (int idx = t; idx < 128 * units; idx += NT) {
      const int r = idx / units, c = (idx - r * units) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int64_t m = m0 + r;
      if (m < M && c < K) v = __ldg(reinterpret_cast<const float4*>(A + m * lda + c));
      uint2 p; p.x = pack_half2(v.x * scale, v.y * scale); p.y = pack_half2(v.z * scale, v.w * scale);
      *reinterpret_cast<uint2*>(dst + (uint32_t)(c >> 6) * HG_A_CHUNK + tile_off((uint32_t)r, (uint32_t)(c & 63))) = p;
    }
  } else if ((NT % (kpad >> 1)) == 0) {
    // unaligned / odd-width rows (first layers: encodings inside a wider row): same walk, two floats per step
    const int pairs = kpad >> 1, rstep = NT / pairs, iters = 128 / rstep;
    const int c = (t % pairs) << 1, r0 = t / pairs;
    const bool x_ok = c < K, y_ok = c + 1 < K;
    const uint32_t cbase = (uint32_t)(c >> 6) * HG_A_CHUNK;
    for (int j0 = 0; j0 < iters; j0 += 8) {
      float x[8], y[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t m = m0 + r0 + (j0 + u) * rstep;
        const bool ok = m < M;
        x[u] = (ok && x_ok) ? __ldg(A + m * lda + c) : 0.f;
        y[u] = (ok && y_ok) ? __ldg(A + m * lda + c + 1) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + (j0 + u) * rstep;
        *reinterpret_cast<uint32_t*>(dst + cbase + tile_off((uint32_t)r, (uint32_t)(c & 63))) =
            pack_half2(x[u] * scale, y[u] * scale);
      }
    }
  } else {
    const int pairs = kpad >> 1;
    for (int idx = t; idx < 128 * pairs; idx += NT) {
      const int r = idx / pairs, c = (idx - r * pairs) << 1;
      const int64_t m = m0 + r;
      float x = 0.f, y = 0.f;
      if (m < M) {
        if (c < K) x = __ldg(A + m * lda + c) * scale;
        if (c + 1 < K) y = __ldg(A + m * lda + c + 1) * scale;
      }
      *reinterpret_cast<uint32_t*>(dst + (uint32_t)(c >> 6) * HG_A_CHUNK + tile_off((uint32_t)r, (uint32_t)(c & 63))) =
          pack_half2(x, y);
    }
  }
}

__global__ void __launch_bounds__(512, 1) hgemm_tc_kernel(const HgArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t w_chunk = (uint32_t)g.n_pad * 128u;
  uint8_t* sw = smem;
  uint8_t* sa = smem + (((uint32_t)g.k_chunks * w_chunk + 1023u) & ~1023u);
  float* s_stage = reinterpret_cast<float*>(sa + (size_t)g.a_bufs * g.k_chunks * HG_A_CHUNK);   // 8 warps x 32 x 20 floats
  __shared__ uint64_t a_free[2], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[256];               // bias padded with zeros: unconditional vector reads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kpad = g.k_chunks * 64;
  if (threadIdx.x < 256) s_bias[threadIdx.x] = (g.bias && (int)threadIdx.x < g.N) ? __ldg(g.bias + threadIdx.x) : 0.f;
  const float a_scale_ = g.a_scale_dev ? __ldg(g.a_scale_dev) : g.a_scale, c_scale = 1.f / a_scale_;

  if (threadIdx.x == 0) {
    mbar_init(&a_free[0], 1); mbar_init(&a_free[1], 1);
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_free[0], 8); mbar_init(&acc_free[1], 8);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  // weight image: B(n, k) for n < n_pad, k < kpad, zero outside [N x K]
  // (eight independent loads in flight per thread: every CTA rebuilds the image, a dependent load per element was a
  // fifth of the kernel's time)
  for (int base = threadIdx.x; base < g.n_pad * kpad; base += 8 * blockDim.x) {
    float v[8];
    uint32_t off[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * blockDim.x;
      int n, k;
      if (g.trans_w) { n = idx % g.n_pad; k = idx / g.n_pad; } else { k = idx % kpad; n = idx / kpad; }
      const bool in = idx < g.n_pad * kpad;
      v[u] = 0.f;
      if (in && n < g.N && k < g.K) v[u] = g.trans_w ? __ldg(g.W + (int64_t)k * g.ldw + n) : __ldg(g.W + (int64_t)n * g.ldw + k);
      off[u] = in ? (uint32_t)(k >> 6) * w_chunk + tile_off((uint32_t)n, (uint32_t)(k & 63)) : 0xffffffffu;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (off[u] != 0xffffffffu) *reinterpret_cast<__half*>(sw + off[u]) = __float2half_rn(v[u]);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t n_tiles = (g.M + 127) / 128;

  if (warp >= 8) {
    // ===================== loaders (+ the MMA issuer) =====================
    const int t = threadIdx.x - 256;                       // 0..255
    const uint32_t idesc = umma_idesc_f16(128, g.n_pad, 0, 0);
    const int ksteps = (g.K + 15) / 16;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int64_t m0 = tile * 128;
      // image buffer ab: with two buffers the load of tile it overlaps the MMAs of tile it-1
      const uint32_t ab = (g.a_bufs == 2) ? (it & 1) : 0u, use = (g.a_bufs == 2) ? (it >> 1) : it;
      if (use > 0) mbar_wait(&a_free[ab], (use - 1) & 1);  // the MMAs that read this buffer last have completed
      uint8_t* sab = sa + (size_t)ab * g.k_chunks * HG_A_CHUNK;
      hg_load_tile<256>(g.A, g.lda, m0, g.M, g.K, kpad, g.vec_a, a_scale_, sab, t);
      fence_async_smem();
      named_bar_sync(1, 256);
      if (warp == 8) {
        const uint32_t buf = it & 1;
        if (it >= 2) mbar_wait(&acc_free[buf], ((it >> 1) - 1) & 1);      // the epilogue has drained this accumulator
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = smem_u32(sab), b0 = smem_u32(sw);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t kb = ks >> 2, ko = (ks & 3) * 32;
            umma_f16(tmem + buf * 256, umma_desc_kmajor(a0 + kb * HG_A_CHUNK + ko),
                     umma_desc_kmajor(b0 + kb * w_chunk + ko), idesc, ks > 0 ? 1u : 0u);
          }
          umma_commit(&a_free[ab]);
          umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: bias, accumulate, activation, mask =====================
    // warp w reads TMEM lanes 32 (w & 3) ..: warps 0-3 take the lower half of the 32-column blocks, warps 4-7 the upper
    const int q = warp & 3, row = q * 32 + lane;
    const int n_blk = (g.n_pad + 31) / 32, blk_lo = (warp < 4) ? 0 : (n_blk + 1) / 2,
              blk_hi = (warp < 4) ? (n_blk + 1) / 2 : n_blk;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      const int64_t m = tile * 128 + row;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      for (int c0 = 32 * blk_lo; c0 < 32 * blk_hi; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + c0, v);
        tmem_ld_wait();
        const bool row_ok = m < g.M;
        float o[32];
        const bool full = c0 + 32 <= g.N, vec = g.vec_c && full;
        float* crow = g.C + m * g.ldc + c0;
        if (row_ok) {
          // every option is tested once per 32-column block; the element loops inside are branch-free
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = fmaf(__uint_as_float(v[i]), c_scale, s_bias[c0 + i]);
          if (g.accumulate) {
            if (vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 c4 = *reinterpret_cast<const float4*>(crow + i);
                o[i] += c4.x; o[i + 1] += c4.y; o[i + 2] += c4.z; o[i + 3] += c4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < g.N) o[i] += crow[i];
            }
          }
          if (g.act == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = fmaxf(o[i], 0.f);
          } else if (g.act == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = o[i] > 0.f ? o[i] : __expf(o[i]) - 1.f;   // abs. error 1e-7: below fp16 rounding
          }
          if (g.mask) {
            const float* mrow = g.mask + m * g.ldmask + c0;
            const bool mvec = full && aligned16_dev(mrow);
            float y[32];
            if (mvec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mrow + i));
                y[i] = m4.x; y[i + 1] = m4.y; y[i + 2] = m4.z; y[i + 3] = m4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] = (c0 + i < g.N) ? __ldg(mrow + i) : 1.f;
            }
            if (g.mask_kind == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = y[i] > 0.f ? o[i] : o[i] * (y[i] + 1.f);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = y[i] > 0.f ? o[i] : 0.f;
            }
          }
        }
        if (vec && g.stage) {
          // a lane owns a ROW of the accumulator, so its own stores would touch 32 different 128-byte lines per
          // instruction.  The block goes through a 32 x 16 staging slice per warp (row stride 20 floats: conflict-free
          // 16-byte accesses) and leaves as 64-byte runs, eight rows per instruction
          float* stg = s_stage + warp * (32 * 20);
          const int pr = lane >> 2, pc = (lane & 3) * 4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(stg + lane * 20 + i) =
                  make_float4(o[16 * h + i], o[16 * h + i + 1], o[16 * h + i + 2], o[16 * h + i + 3]);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = pr + 8 * j;
              const int64_t mr = tile * 128 + q * 32 + r;
              if (mr < g.M)
                *reinterpret_cast<float4*>(g.C + mr * g.ldc + c0 + 16 * h + pc) = *reinterpret_cast<const float4*>(stg + r * 20 + pc);
            }
            __syncwarp();
          }
        } else if (row_ok) {
          if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(crow + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < g.N) crow[i] = o[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Weight gradient G[N_out, K_in] += sum over samples of dY[m, :]^T X[m, :] (swnerf_sgemm op 2) on the tensor cores.
// Both operands are read with K = the sample index, i.e. as MN-major operands over the same [128 samples x 64 columns]
// images the forward builds (the saved-image trick of mlp_tc_bwd.cu): per 128-sample tile the CTA converts its dY tile
// (times the gradient scale) and its X tile, issues 8 k-steps of M = 128 (one or two blocks of 128 output channels) x
// N = K_in (padded to 64) MMAs, and keeps accumulating into tensor memory over ALL its tiles; one red.add flush at the end.
struct HgWgArgs {
  const float* dY; int64_t ldy;
  const float* X; int64_t ldx;
  float* G; int64_t ldg;
  int64_t M; int n_out, k_in;
  int m_blocks, n_pad;                            // blocks of 128 output channels (1 or 2); K_in rounded up to 64
  float a_scale; const float* a_scale_dev;
  int vec_y, vec_x;
};

template <int TMEM_COLS>
__global__ void __launch_bounds__(256, (TMEM_COLS <= 256) ? 2 : 1) hgemm_tc_wgrad_kernel(const HgWgArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sy = smem;                                          // dY image: 2 m_blocks chunks of [128 x 64]
  uint8_t* sx = smem + (size_t)g.m_blocks * 2 * HG_A_CHUNK;    // X image: n_pad / 64 chunks
  __shared__ uint64_t mma_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float a_scale_ = g.a_scale_dev ? __ldg(g.a_scale_dev) : g.a_scale, c_scale = 1.f / a_scale_;
  if (threadIdx.x == 0) { mbar_init(&mma_done, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t n_tiles = (g.M + 127) / 128;
  const uint32_t idesc = umma_idesc_f16(128, g.n_pad, 1, 1);
  uint32_t it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    if (it > 0) mbar_wait(&mma_done, (it - 1) & 1);            // the previous tile's MMAs have read both images
    hg_load_tile<256>(g.dY, g.ldy, tile * 128, g.M, g.n_out, g.m_blocks * 128, g.vec_y, a_scale_, sy, threadIdx.x);
    hg_load_tile<256>(g.X, g.ldx, tile * 128, g.M, g.k_in, g.n_pad, g.vec_x, 1.f, sx, threadIdx.x);
    fence_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t y0 = smem_u32(sy), x0 = smem_u32(sx);
        for (int ks = 0; ks < 8; ++ks) {                       // 16 samples per k-step = two 8-row atoms
          const uint64_t bd = umma_desc_mnmajor(x0 + ks * 2048, HG_A_CHUNK);
          for (int mb = 0; mb < g.m_blocks; ++mb)
            umma_f16(tmem + mb * g.n_pad, umma_desc_mnmajor(y0 + mb * 2 * HG_A_CHUNK + ks * 2048, HG_A_CHUNK), bd, idesc,
                     (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&mma_done);
      }
      __syncwarp();
    }
  }
  mbar_wait(&mma_done, (it - 1) & 1);                          // it >= 1: every CTA owns at least one tile
  tc_fence_after();
  {
    // all eight warps flush: warp w reads TMEM lanes 32 (w & 3) .., warps 0-3 the lower half of the 32-column blocks,
    // warps 4-7 the upper half
    const int q = warp & 3, n_blk = g.n_pad / 32, blk_lo = (warp < 4) ? 0 : (n_blk + 1) / 2,
              blk_hi = (warp < 4) ? (n_blk + 1) / 2 : n_blk;
    for (int mb = 0; mb < g.m_blocks; ++mb) {
      const int row = mb * 128 + q * 32 + lane;
      for (int c0 = 32 * blk_lo; c0 < 32 * blk_hi; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mb * g.n_pad + c0, v);
        tmem_ld_wait();
        if (row < g.n_out) {
          float* grow = g.G + (int64_t)row * g.ldg + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < g.k_in) atomicAdd(grow + i, __uint_as_float(v[i]) * c_scale);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
}

// out[0] = 2^floor(log2(target / max|x|)) (1 if x is all zero): the power of two that lifts a gradient tensor into
// fp16's normal range with `target` as its largest magnitude (the fused backward does the same, mlp_tc_bwd.cu)
__global__ void __launch_bounds__(1024) pow2_scale_kernel(const float* __restrict__ x, int64_t n, float target,
                                                          float* __restrict__ out) {
  __shared__ float red[32];
  float mx = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, fabsf(__ldg(x + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) {
      float sc = 1.f;
      if (mx > 0.f && mx < 3.0e38f) sc = exp2f(fminf(fmaxf(floorf(log2f(target / mx)), -40.f), 40.f));
      out[0] = sc;
    }
  }
}

}  // namespace swnerf

using namespace swnerf;

extern "C" {

int swnerf_pow2_scale(const float* x, int64_t n, float target, float* out, void* stream) {
  SW_REQUIRE(x && out && n >= 0 && target > 0.f, "pow2_scale: bad argument");
  swnerf::pow2_scale_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, target, out);
  return swnerf::check_launch("pow2_scale");
}

int swnerf_hgemm_tc_wgrad(const float* dY, int64_t ldy, const float* X, int64_t ldx, float* G, int64_t ldg, int64_t M,
                          int64_t n_out, int64_t k_in, float a_scale, const float* a_scale_dev, void* stream) {
  SW_REQUIRE(dY && X && G, "hgemm_tc_wgrad: null pointer");
  SW_REQUIRE(M >= 0 && n_out >= 32 && n_out <= 256 && k_in >= 1 && k_in <= 256,
             "hgemm_tc_wgrad: needs 32 <= n_out <= 256 and 1 <= k_in <= 256");
  SW_REQUIRE(a_scale > 0.f, "hgemm_tc_wgrad: a_scale must be positive");
  if (M == 0) return SWNERF_OK;
  HgWgArgs g;
  g.dY = dY; g.ldy = ldy; g.X = X; g.ldx = ldx; g.G = G; g.ldg = ldg; g.M = M; g.n_out = (int)n_out; g.k_in = (int)k_in;
  g.m_blocks = n_out > 128 ? 2 : 1;
  g.n_pad = (int)((k_in + 63) / 64 * 64);
  g.a_scale = a_scale; g.a_scale_dev = a_scale_dev;
  g.vec_y = (aligned16(dY) && ldy % 4 == 0 && n_out % 4 == 0) ? 1 : 0;
  g.vec_x = (aligned16(X) && ldx % 4 == 0 && k_in % 4 == 0) ? 1 : 0;
  const size_t smem = (size_t)(g.m_blocks * 2 + g.n_pad / 64) * HG_A_CHUNK + 1024;
  if (once_per_device(ONCE_HGEMM_WGRAD)) {
    cudaFuncSetAttribute(hgemm_tc_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * HG_A_CHUNK + 1024);
    cudaFuncSetAttribute(hgemm_tc_wgrad_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * HG_A_CHUNK + 1024);
    cudaFuncSetAttribute(hgemm_tc_wgrad_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * HG_A_CHUNK + 1024);
  }
  // accumulator columns = m_blocks x n_pad: a CTA that needs at most half of tensor memory shares its SM with a second
  // one (their loads and MMAs overlap; the kernel has no other latency hiding)
  const int cols = g.m_blocks * g.n_pad;
  const int64_t tiles = (M + 127) / 128;
  const int per_sm = (cols <= 256 && smem <= 100 * 1024) ? 2 : 1;
  const int64_t max_ctas = (int64_t)per_sm * sm_count();
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  cudaStream_t st = (cudaStream_t)stream;
  if (cols <= 128) hgemm_tc_wgrad_kernel<128><<<grid, 256, smem, st>>>(g);
  else if (cols <= 256) hgemm_tc_wgrad_kernel<256><<<grid, 256, smem, st>>>(g);
  else hgemm_tc_wgrad_kernel<512><<<grid, 256, smem, st>>>(g);
  return check_launch("hgemm_tc_wgrad");
}

int swnerf_hgemm_tc_supported(int64_t N, int64_t K) { return (N >= 16 && N <= 256 && K >= 1 && K <= 256) ? 1 : 0; }

int swnerf_hgemm_tc(int op, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                    int64_t N, int64_t K, const float* bias, int accumulate, int act_flags, const float* mask,
                    int64_t ldmask, float a_scale, const float* a_scale_dev, void* stream) {
  SW_REQUIRE(A && W && C, "hgemm_tc: null pointer");
  SW_REQUIRE(op == 0 || op == 1, "hgemm_tc: op must be 0 (x W^T) or 1 (dy W)");
  SW_REQUIRE(M >= 0 && swnerf_hgemm_tc_supported(N, K), "hgemm_tc: needs 16 <= N <= 256 and 1 <= K <= 256");
  const int act = act_flags & 3, mask_kind = (act_flags >> 4) & 3;
  SW_REQUIRE(act <= 2 && mask_kind <= 1 && (act_flags & ~0x33) == 0, "hgemm_tc: bad act_flags");
  SW_REQUIRE(a_scale > 0.f, "hgemm_tc: a_scale must be positive");
  if (M == 0) return SWNERF_OK;
  HgArgs g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.trans_w = op; g.C = C; g.ldc = ldc; g.M = M; g.N = (int)N; g.K = (int)K;
  g.bias = bias; g.accumulate = accumulate; g.act = act; g.mask_kind = mask_kind; g.mask = mask; g.ldmask = ldmask;
  g.n_pad = (int)((N + 15) / 16 * 16);
  g.k_chunks = (int)((K + 63) / 64);
  g.a_scale = a_scale; g.a_scale_dev = a_scale_dev;
  g.vec_a = (aligned16(A) && lda % 4 == 0 && K % 4 == 0) ? 1 : 0;
  g.vec_c = (aligned16(C) && ldc % 4 == 0) ? 1 : 0;
  const size_t w_bytes = ((size_t)g.k_chunks * g.n_pad * 128 + 1023) & ~(size_t)1023, a_bytes = (size_t)g.k_chunks * HG_A_CHUNK;
  const size_t stage_bytes = 8 * 32 * 20 * sizeof(float);
  g.a_bufs = (w_bytes + 2 * a_bytes + stage_bytes + 1024 <= 220 * 1024) ? 2 : 1;
  g.stage = (w_bytes + g.a_bufs * a_bytes + stage_bytes + 1024 <= 220 * 1024) ? 1 : 0;
  const size_t smem = w_bytes + g.a_bufs * a_bytes + (g.stage ? stage_bytes : 0) + 1024;
  if (once_per_device(ONCE_HGEMM))
    cudaFuncSetAttribute(hgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);   // W <= 128 KB, A <= 64 KB, staging 20 KB
  const int64_t tiles = (M + 127) / 128;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  hgemm_tc_kernel<<<grid, 512, smem, (cudaStream_t)stream>>>(g);
  return check_launch("hgemm_tc");
}

}  // extern "C"
