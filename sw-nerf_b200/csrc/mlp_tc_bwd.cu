// Backward of the fused 8x256 MLP on tcgen05 tensor cores (autograd of model.py:39-62 at the sample
// positions of nerf/run.py:385): d_raw[N,S,4] -> fp32 gradients of the 24 parameter tensors.
//
//  1. bwd_data kernel (per 128-sample tile, same skeleton as the forward): the chain
//        d_raw -> dy9 = (d_rgb W_rgb) * mask9 -> dh7 = dy9 W_fv + d_sigma w_alpha -> dy7 = dh7 * mask7 -> dh6 = dy7 W7 ...
//     runs on tensor cores with transposed fp16 weight chunks; ReLU masks are the sign bits the forward
//     saved (32 B / row / layer).  Every dy_l tile is written back as an operand image for step 2.
//  2. bwd_weight kernel: dW_l = dy_l^T x_l, K = all samples.  The saved forward activations and the dy
//     tiles are the SAME 128B-swizzled images read as MN-major operands.  Each CTA owns one layer ("job")
//     and a slice of the tiles, accumulates in TMEM (up to 512 columns = a full 256x256 fp32 gradient) and
//     flushes once with red.add into the caller's (flat) gradient buffers.  HBM-bound: 2 x 64 KB per tile-layer.
//     The otherwise idle CUDA cores of the same CTAs form the bias gradients (column sums of the dy images
//     sitting in shared memory).
//  3. un-fold: d W_fv, d b_fv -> feature_linear / views_linears gradients (fp32 SIMT GEMMs, tiny).
// Gradients are scaled by a power of two (from max|d_raw|, on device) before the fp16 conversion and
// unscaled in the fp32 flush.
#include <cuda.h>
#include <mutex>
#include <cstdlib>
#include <atomic>
#include "common.cuh"
#include "tc_common.cuh"
#include "mlp_tc_layout.cuh"
#include "../../include/swnerf_b200.h"

namespace swnerf {
using namespace tc;
using namespace tcl;

struct ParamPtrsB {
  const float* p[24];
  int kind;            // 0: canonical net (24 tensors); 1: deformation net (_time.0..7, _time_out), see mlp_tc.cu
  Enc enc;
};

// ------------------------------------------------------------------------------------------------
// transposed weight image
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bwd_weight(const ParamPtrsB& P, const float* fold, int c, int n, int k) {
  // B[n][k]: n = input channel of the layer (output of the data-gradient GEMM), k = its output unit
  if (c < 2) return fold[(c * 64 + k) * 257 + n];                 // W_fv[u][n]
  if (c == 2) {
    if (P.kind == 1) return k < 3 ? P.p[16][k * 256 + n] : 0.f;   // _time_out rows
    return k == 0 ? P.p[20][n] : 0.f;                             // alpha_linear row
  }
  int t = (c - 3) / 4, kc = (c - 3) % 4;
  int l = 7 - t;                                                  // 7,6,5,4,3,2,1
  int ld = (l == 5) ? 256 + P.enc.pc : 256, off = (l == 5) ? P.enc.pc : 0;      // pts_linears.5 is [256, pc + 256]
  return P.p[2 * l][(size_t)(kc * 64 + k) * ld + off + n];
}

__global__ void pack_bwd_kernel(ParamPtrsB P, const uint8_t* __restrict__ packed_fwd, uint8_t* __restrict__ packed_t) {
  const float* fold = reinterpret_cast<const float*>(packed_fwd + PK_FOLD_OFF);
  int unit = blockIdx.x * blockDim.x + threadIdx.x;
  if (unit >= PKT_TOTAL_BYTES / 16) return;
  int byte = unit * 16;
  __align__(16) __half h[8];
  if (byte < PKT_DPE_OFF) {
    int c = byte / CHUNK_B, in = byte % CHUNK_B;
    int n = (in >> 10) * 8 + ((in >> 7) & 7);
    int pu = (in >> 4) & 7;
    int k0 = (pu ^ (n & 7)) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = __float2half_rn(bwd_weight(P, fold, c, n, k0 + i));
  } else {
    // input-gradient weights, one group of eight [64 x 64] images per 64-column chunk of the position encoding:
    // chunk (layer 0 | 5, kc): image row = pe column c (inside the chunk), image column = unit kc*64 + k
    int b2 = byte - PKT_DPE_OFF;
    int grp = b2 / (8 * DPE_CHUNK_B);
    int ch = (b2 / DPE_CHUNK_B) & 7, in = b2 % DPE_CHUNK_B;
    int cc = (in >> 10) * 8 + ((in >> 7) & 7);
    int pu = (in >> 4) & 7;
    int k0 = (pu ^ (cc & 7)) * 8;
    const float* Wl = (ch < 4) ? P.p[0] : P.p[10];
    const int pc = P.enc.pc;
    int ld = (ch < 4) ? pc + (P.kind == 1 ? P.enc.tw : 0) : 256 + pc;
    const int col = grp * 64 + cc;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int nn = (ch & 3) * 64 + k0 + i;
      h[i] = __float2half_rn(col < pc ? Wl[(size_t)nn * ld + col] : 0.f);
    }
  }
  *reinterpret_cast<uint4*>(packed_t + byte) = *reinterpret_cast<const uint4*>(h);
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < __int_as_float(0x7f800000)) atomicMax(out, __float_as_uint(m));
}

// four consecutive fp32 reductions as ONE 16-byte request (the flush of an accumulator row is otherwise 32 separate
// 4-byte reductions into 32 different lines per warp instruction); p must be 16-byte aligned
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float grad_scale_from(const uint32_t* absmax, float fixed) {
  if (fixed > 0.f) return fixed;
  float mx = __uint_as_float(*absmax);
  if (!(mx > 0.f)) return 1.f;
  float e = floorf(log2f(32.f / mx));      // max|d_raw| -> [16, 32): 2000x headroom below fp16's 65504 for growth in the chain
  e = fminf(fmaxf(e, -40.f), 40.f);
  return exp2f(e);
}

// ------------------------------------------------------------------------------------------------
// 1. backward data
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
  const float* d_raw;          // kind 0: d_raw[P,4];  kind 1: d_dx[P,3]
  int kind;
  int64_t P; int64_t num_tiles;
  const uint8_t* packed; const uint8_t* packed_t;
  uint8_t* ws;
  float* grads[24];
  float* unfold;
  const uint32_t* absmax; float fixed_scale;
};

__device__ __forceinline__ float4 load_dout(const BwdArgs& g, int64_t idx) {
  if (idx >= g.P) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (g.kind == 0) return __ldg(reinterpret_cast<const float4*>(g.d_raw) + idx);
  // deformation net: the three d_dx components ride where d_sigma does (rows 128..130 of the head)
  return make_float4(__ldg(g.d_raw + idx * 3), __ldg(g.d_raw + idx * 3 + 1), __ldg(g.d_raw + idx * 3 + 2), 0.f);
}

__global__ void __launch_bounds__(384, 1) mlp_bwd_data_kernel(BwdArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_act = smem + SMB_ACT;
  uint8_t* s_ring = smem + SMB_RING;
  float* s_wrgb = reinterpret_cast<float*>(smem + SMB_WRGB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMB_BAR);
  uint64_t* w_full = bars;            // [3]
  uint64_t* w_empty = bars + 3;       // [3]
  uint64_t* act_full = bars + 6;      // [4]
  uint64_t* d_full = bars + 10;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  uint64_t* blk_ready = bars + 13;    // [4] block j of the current dy image is complete (for the store warp)
  uint64_t* st_done = bars + 17;      // [4] block j's bulk store has finished reading shared memory

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int j = 0; j < 4; ++j) mbar_init(&act_full[j], 8);     // one arrival per epilogue warp
    mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
    for (int j = 0; j < 4; ++j) { mbar_init(&blk_ready[j], 8); mbar_init(&st_done[j], 1); }
    mbar_fence_init();
  }
  if (warp == 10) tmem_alloc<512>(tmem_slot);
  {
    const float* src = reinterpret_cast<const float*>(g.packed + PK_F32_OFF) + F32_WRGB;
    for (int i = threadIdx.x; i < 384; i += blockDim.x) s_wrgb[i] = __ldg(src + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float scale = grad_scale_from(g.absmax, g.fixed_scale);
  const int64_t mask_base = g.num_tiles * WS_TILE_BYTES;
  const int64_t dy_base = mask_base + g.num_tiles * WS_MASK_BYTES;

  // warp roles: 0-7 epilogue | 8 producer | 9 MMA issuer (highest warp id of its sub-partition) | 10 TMEM alloc
  if (warp == 8) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        for (int c = 0; c < NT_CHUNKS; ++c, ++cnt) {
          uint32_t stage = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1;
          mbar_wait(&w_empty[stage], ph ^ 1);
          mbar_expect_tx(&w_full[stage], CHUNK_B);
          bulk_g2s(s_ring + stage * CHUNK_B, g.packed_t + (size_t)c * CHUNK_B, CHUNK_B, &w_full[stage]);
        }
      }
    }
  } else if (warp == 9) {
    // whole warp converged, one elected lane issues (keeps the operands in uniform registers)
    const uint32_t idesc = umma_idesc_f16(128, 256, 0, 0);
    const uint32_t act_u32 = smem_u32(s_act), ring_u32 = smem_u32(s_ring);
    uint32_t cnt = 0, dcnt = 0, alayer = 0;
    for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      for (int t = 0; t < 8; ++t, ++dcnt, ++alayer) {
        const uint32_t d_tmem = tmem + (dcnt & 1) * 256;
        const int nch = (t == 0) ? 3 : 4;
        for (int ci = 0; ci < nch; ++ci, ++cnt) {
          mbar_wait(&act_full[ci], alayer & 1);
          const uint32_t stage = cnt % NSTAGE;
          mbar_wait(&w_full[stage], (cnt / NSTAGE) & 1);
          tc_fence_after();
          const uint32_t a_base = act_u32 + ci * ACT_BLK;
          const uint32_t b_base = ring_u32 + stage * CHUNK_B;
          const int nks = (t == 0 && ci == 2) ? 1 : 4;           // sigma block: only the first 16 columns
          if (elect_one()) {
            for (int ks = 0; ks < nks; ++ks)
              umma_f16(d_tmem, umma_desc_kmajor(a_base + ks * 32), umma_desc_kmajor(b_base + ks * 32), idesc,
                       (ci > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&w_empty[stage]);
            if (ci == nch - 1) umma_commit(&d_full[dcnt & 1]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t dcnt = 0, wstep = 0;         // wstep: dy images written so far (9 per tile)
    float4 pre_dr = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t pre_hm[2] = {0u, 0u};
    if ((int64_t)blockIdx.x < g.num_tiles) {
      const int64_t idx0 = (int64_t)blockIdx.x * TILE + row;
      pre_dr = load_dout(g, idx0);
      const uint32_t* m0 = reinterpret_cast<const uint32_t*>(g.ws + mask_base) + (int64_t)blockIdx.x * (9 * 8 * 128);
      pre_hm[0] = __ldg(m0 + (8 * 8 + hh * 2 + 0) * 128 + row);
      pre_hm[1] = __ldg(m0 + (8 * 8 + hh * 2 + 1) * 128 + row);
    }
    for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      const uint32_t* ws_mask = reinterpret_cast<const uint32_t*>(g.ws + mask_base) + tile * (9 * 8 * 128);
      // ---- head prep: d_raw -> [dy9 | d_sigma | d_rgb] operand image
      // (d_raw and the sign words of this tile were prefetched during the previous tile's last layer)
      {
        if (wstep > 0) {   // the store warp may still be reading the previous image
#pragma unroll
          for (int j = 0; j < 4; ++j) mbar_wait(&st_done[j], (wstep - 1) & 1);
        }
        ++wstep;
        float4 dr = pre_dr;
        dr.x *= scale; dr.y *= scale; dr.z *= scale; dr.w *= scale;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int c0 = hh * 64 + jj * 32;
          const uint32_t m = pre_hm[jj];
          float val[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = dr.x * s_wrgb[c0 + i] + dr.y * s_wrgb[128 + c0 + i] + dr.z * s_wrgb[256 + c0 + i];
            val[i] = (g.kind == 0 && ((m >> i) & 1u)) ? v : 0.f;      // the deformation net has no view branch
          }
          uint8_t* blk = s_act + hh * ACT_BLK;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(blk + tile_unit_off(row, jj * 4 + u)) =
                make_uint4(pack_half2(val[8 * u], val[8 * u + 1]), pack_half2(val[8 * u + 2], val[8 * u + 3]),
                           pack_half2(val[8 * u + 4], val[8 * u + 5]), pack_half2(val[8 * u + 6], val[8 * u + 7]));
        }
        uint8_t* blk = s_act + (2 + hh) * ACT_BLK;
        uint4 u0 = make_uint4(0u, 0u, 0u, 0u);
        if (g.kind == 1) {     // [d_dx0, d_dx1, d_dx2, 0..] in the sigma block; nothing in the rgb block
          if (hh == 0) { u0.x = pack_half2(dr.x, dr.y); u0.y = pack_half2(dr.z, 0.f); }
        } else if (hh == 0) u0.x = pack_half2(dr.w, 0.f);
        else { u0.x = pack_half2(dr.x, dr.y); u0.y = pack_half2(dr.z, 0.f); }
        *reinterpret_cast<uint4*>(blk + tile_unit_off(row, 0)) = u0;
        *reinterpret_cast<uint4*>(blk + tile_unit_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { mbar_arrive(&act_full[j]); mbar_arrive(&blk_ready[j]); }
        }
      }
      // ---- layers: dh_l (TMEM) * mask_l -> dy_l.  The sign words of layer l-1 (and, during the last layer,
      // d_raw + head sign words of the NEXT tile) are fetched one layer ahead so that no global-load latency
      // sits between the TMEM read and the MMA hand-off.
      uint32_t cur_m[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) cur_m[j] = __ldg(ws_mask + (7 * 8 + j * 2 + hh) * 128 + row);
      for (int t = 0; t < 8; ++t, ++dcnt) {
        const int l = 7 - t;
        const uint32_t dcol = (dcnt & 1) * 256;
        uint32_t nxt_m[4] = {0u, 0u, 0u, 0u};
        if (t < 7) {
#pragma unroll
          for (int j = 0; j < 4; ++j) nxt_m[j] = __ldg(ws_mask + ((l - 1) * 8 + j * 2 + hh) * 128 + row);
        } else {
          const int64_t ntile = tile + gridDim.x;
          if (ntile < g.num_tiles) {
            const int64_t nidx = ntile * TILE + row;
            pre_dr = load_dout(g, nidx);
            const uint32_t* nmask = reinterpret_cast<const uint32_t*>(g.ws + mask_base) + ntile * (9 * 8 * 128);
            pre_hm[0] = __ldg(nmask + (8 * 8 + hh * 2 + 0) * 128 + row);
            pre_hm[1] = __ldg(nmask + (8 * 8 + hh * 2 + 1) * 128 + row);
          }
        }
        mbar_wait(&d_full[dcnt & 1], (dcnt >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t v[32];
          const int c0 = j * 64 + hh * 32;
          tmem_ld32(tmem + lane_addr + dcol + c0, v);
          tmem_ld_wait();
          const uint32_t m = cur_m[j];
          uint8_t* blk = s_act + j * ACT_BLK;
          mbar_wait(&st_done[j], (wstep - 1) & 1);       // previous image's block j has been read out
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 8 * u + 2 * e;
              float a = ((m >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
              float b = ((m >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : 0.f;
              pk[e] = pack_half2(a, b);
            }
            *reinterpret_cast<uint4*>(blk + tile_unit_off(row, hh * 4 + u)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (t < 7) mbar_arrive(&act_full[j]);
            mbar_arrive(&blk_ready[j]);
          }
        }
        ++wstep;
#pragma unroll
        for (int j = 0; j < 4; ++j) cur_m[j] = nxt_m[j];
      }
    }
  } else if (warp == 11) {
    // dy-image store warp: saves every block as soon as its eight epilogue warps have written it
    if (lane == 0) {
      uint32_t step = 0;
      for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        uint8_t* ws_dy = g.ws + dy_base + tile * WS_DY_BYTES;
        for (int sidx = 0; sidx < 9; ++sidx, ++step) {        // head image, then layers 7..0
          uint8_t* dst = ws_dy + (sidx == 0 ? (size_t)WS_DYH_OFF : (size_t)(8 - sidx) * ACT_BYTES);
          for (int j = 0; j < 4; ++j) {
            mbar_wait(&blk_ready[j], step & 1);
            bulk_s2g(dst + j * ACT_BLK, s_act + j * ACT_BLK, ACT_BLK);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(&st_done[j]);
          }
        }
      }
      bulk_wait_all0();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// 1b. backward data on CTA pairs (cta_group::2), two tile slots per CTA - the structure of mlp_fwd4_kernel:
// the leader issues M=256 MMAs over both CTAs' dy images and the two halves of each transposed weight chunk
// (tensor-map loads that report to the leader's barrier), the two slots' layers alternate so one slot's epilogue
// (TMEM -> ReLU gate -> fp16 -> dy image in place, one proxy fence and one arrival per layer) runs under the other
// slot's MMAs, completion is multicast.  Shared memory per CTA: 2 x 64 KB dy images + 5 x 16 KB weight ring.
// ------------------------------------------------------------------------------------------------
constexpr int NST5 = 5;
constexpr int S5_ACT = 0;
constexpr int S5_RING = S5_ACT + 2 * ACT_BYTES;
constexpr int S5_WRGB = S5_RING + NST5 * (CHUNK_B / 2);
constexpr int S5_BAR = S5_WRGB + 384 * 4;
constexpr int S5_TOTAL = S5_BAR + 512 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
mlp_bwd_data_pair_kernel(const __grid_constant__ BwdArgs g, const __grid_constant__ CUtensorMap tm_wt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_act = smem + S5_ACT;          // [2 slots][4 blocks]
  uint8_t* s_ring = smem + S5_RING;
  float* s_wrgb = reinterpret_cast<float*>(smem + S5_WRGB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S5_BAR);
  uint64_t* w_full = bars;                 // [NST5] leader: both halves of the stage have landed
  uint64_t* w_empty = bars + 8;            // [NST5] multicast
  uint64_t* act_full = bars + 16;          // [2] leader, 16 arrivals: slot's dy image written in both CTAs
  uint64_t* d_full = bars + 18;            // [2] multicast: slot's accumulator complete
  uint64_t* img_ready = bars + 20;         // [2] local: image written (store warp)
  uint64_t* st_done = bars + 22;           // [2] local: image's bulk stores have read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t num_quads = (g.num_tiles + 3) >> 2;
  const int64_t quad0 = blockIdx.x >> 1, quad_step = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST5; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&act_full[t], 16); mbar_init(&d_full[t], 1); mbar_init(&img_ready[t], 8); mbar_init(&st_done[t], 1);
    }
    mbar_fence_init();
  }
  {
    const float* src = reinterpret_cast<const float*>(g.packed + PK_F32_OFF) + F32_WRGB;
    for (int i = threadIdx.x; i < 384; i += blockDim.x) s_wrgb[i] = __ldg(src + i);
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 10) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float scale = grad_scale_from(g.absmax, g.fixed_scale);
  const int64_t mask_base = g.num_tiles * WS_TILE_BYTES;
  const int64_t dy_base = mask_base + g.num_tiles * WS_MASK_BYTES;
  const uint32_t act_full_l = mapa_u32(smem_u32(act_full), 0);

  // warp roles: 0-7 epilogue | 8 producer | 9 MMA issuer (leader) | 10 TMEM alloc | 11 store
  if (warp == 8) {
    if (lane == 0) {
      const uint32_t w_full_l = mapa_u32(smem_u32(w_full), 0);
      uint32_t cnt = 0;
      for (int64_t quad = quad0; quad < num_quads; quad += quad_step) {
        int cbase = 0;
        for (int t = 0; t < 8; ++t) {
          const int nch = (t == 0) ? 3 : 4;
          for (int sl = 0; sl < 2; ++sl)
            for (int ci = 0; ci < nch; ++ci, ++cnt) {
              const uint32_t stage = cnt % NST5, ph = (cnt / NST5) & 1;
              mbar_wait(&w_empty[stage], ph ^ 1);
              if (leader) mbar_expect_tx(&w_full[stage], CHUNK_B);
              tma_load_2d_pair(s_ring + stage * (CHUNK_B / 2), &tm_wt, 0, (cbase + ci) * 256 + (int)rank * 128,
                               w_full_l + stage * 8);
            }
          cbase += nch;
        }
      }
    }
  } else if (warp == 9 && leader) {
    const uint32_t idesc = umma_idesc_f16(256, 256, 0, 0);
    const uint32_t act_u32 = smem_u32(s_act), ring_u32 = smem_u32(s_ring);
    uint32_t cnt = 0;
    for (int64_t quad = quad0; quad < num_quads; quad += quad_step) {
      for (int t = 0; t < 8; ++t) {
        const int nch = (t == 0) ? 3 : 4;
        for (int sl = 0; sl < 2; ++sl) {
          // the slot's image (head image for t = 0, dy of the layer above otherwise) is written and its accumulator read
          mbar_wait_cluster(&act_full[sl], t & 1);
          const uint32_t d_tmem = tmem + sl * 256;
          for (int ci = 0; ci < nch; ++ci, ++cnt) {
            const uint32_t stage = cnt % NST5;
            mbar_wait(&w_full[stage], (cnt / NST5) & 1);
            tc_fence_after();
            const uint32_t a_base = act_u32 + sl * ACT_BYTES + ci * ACT_BLK;
            const uint32_t b_base = ring_u32 + stage * (CHUNK_B / 2);
            const int nks = (t == 0 && ci == 2) ? 1 : 4;           // sigma block: only the first 16 columns
            if (elect_one()) {
              for (int ks = 0; ks < nks; ++ks)
                umma_f16_pair(d_tmem, umma_desc_kmajor(a_base + ks * 32), umma_desc_kmajor(b_base + ks * 32), idesc,
                              (ci > 0 || ks > 0) ? 1u : 0u);
              umma_commit_pair(&w_empty[stage], 3);
              if (ci == nch - 1) umma_commit_pair(&d_full[sl], 3);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t dcnt = 0, wstep = 0;         // d_full completions per slot so far; images written per slot so far
    const uint32_t* mask0 = reinterpret_cast<const uint32_t*>(g.ws + mask_base);
    for (int64_t quad = quad0; quad < num_quads; quad += quad_step) {
      // ---- head prep of both slots: d_raw -> [dy9 | d_sigma | d_rgb] operand image
      uint32_t cur_m[2][4];
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const int64_t tile = quad * 4 + sl * 2 + rank;
        const bool tvalid = tile < g.num_tiles;
        const uint32_t* ws_mask = mask0 + tile * (9 * 8 * 128);
        float4 dr = load_dout(g, tile * TILE + row);
        dr.x *= scale; dr.y *= scale; dr.z *= scale; dr.w *= scale;
        uint32_t hm[2] = {0u, 0u};
        if (tvalid) {
          hm[0] = __ldg(ws_mask + (8 * 8 + hh * 2 + 0) * 128 + row);
          hm[1] = __ldg(ws_mask + (8 * 8 + hh * 2 + 1) * 128 + row);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cur_m[sl][j] = tvalid ? __ldg(ws_mask + (7 * 8 + j * 2 + hh) * 128 + row) : 0u;
        if (wstep > 0) mbar_wait(&st_done[sl], (wstep - 1) & 1);     // the previous image's stores have read it
        uint8_t* img = s_act + sl * ACT_BYTES;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int c0 = hh * 64 + jj * 32;
          const uint32_t m = hm[jj];
          float val[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = dr.x * s_wrgb[c0 + i] + dr.y * s_wrgb[128 + c0 + i] + dr.z * s_wrgb[256 + c0 + i];
            val[i] = (g.kind == 0 && ((m >> i) & 1u)) ? v : 0.f;      // the deformation net has no view branch
          }
          uint8_t* blk = img + hh * ACT_BLK;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(blk + tile_unit_off(row, jj * 4 + u)) =
                make_uint4(pack_half2(val[8 * u], val[8 * u + 1]), pack_half2(val[8 * u + 2], val[8 * u + 3]),
                           pack_half2(val[8 * u + 4], val[8 * u + 5]), pack_half2(val[8 * u + 6], val[8 * u + 7]));
        }
        uint8_t* blk = img + (2 + hh) * ACT_BLK;
        uint4 u0 = make_uint4(0u, 0u, 0u, 0u);
        if (g.kind == 1) {     // [d_dx0, d_dx1, d_dx2, 0..] in the sigma block; nothing in the rgb block
          if (hh == 0) { u0.x = pack_half2(dr.x, dr.y); u0.y = pack_half2(dr.z, 0.f); }
        } else if (hh == 0) u0.x = pack_half2(dr.w, 0.f);
        else { u0.x = pack_half2(dr.x, dr.y); u0.y = pack_half2(dr.z, 0.f); }
        *reinterpret_cast<uint4*>(blk + tile_unit_off(row, 0)) = u0;
        *reinterpret_cast<uint4*>(blk + tile_unit_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive_remote(act_full_l + sl * 8); mbar_arrive(&img_ready[sl]); }
      }
      ++wstep;
      // ---- layers 7..0: dh_l (TMEM) gated by the layer's sign bits -> dy_l, in place
      for (int t = 0; t < 8; ++t, ++dcnt) {
        const int l = 7 - t;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int64_t tile = quad * 4 + sl * 2 + rank;
          const bool tvalid = tile < g.num_tiles;
          const uint32_t* ws_mask = mask0 + tile * (9 * 8 * 128);
          uint32_t nxt_m[4] = {0u, 0u, 0u, 0u};
          if (t < 7 && tvalid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) nxt_m[j] = __ldg(ws_mask + ((l - 1) * 8 + j * 2 + hh) * 128 + row);
          }
          mbar_wait(&d_full[sl], dcnt & 1);
          tc_fence_after();
          mbar_wait(&st_done[sl], (wstep - 1) & 1);       // the previous image's stores have read it
          uint8_t* img = s_act + sl * ACT_BYTES;
          const uint32_t acc = tmem + lane_addr + sl * 256 + hh * 32;
          uint32_t va[32], vb[32];
          tmem_ld32(acc, va);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t (&v)[32] = (j & 1) ? vb : va;
            uint32_t (&vn)[32] = (j & 1) ? va : vb;
            tmem_ld_wait_on(v);
            if (j < 3) tmem_ld32(acc + (j + 1) * 64, vn);
            const uint32_t m = cur_m[sl][j];
            uint8_t* blk = img + j * ACT_BLK;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = 8 * u + 2 * e;
                float a = ((m >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
                float b = ((m >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : 0.f;
                pk[e] = pack_half2(a, b);
              }
              *reinterpret_cast<uint4*>(blk + tile_unit_off(row, hh * 4 + u)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (t < 7) mbar_arrive_remote(act_full_l + sl * 8);
            mbar_arrive(&img_ready[sl]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) cur_m[sl][j] = nxt_m[j];
        }
        ++wstep;
      }
    }
  } else if (warp == 11) {
    // dy-image store warp: one image per (slot, step), in 16-KB pieces
    if (lane == 0) {
      uint32_t step = 0;
      for (int64_t quad = quad0; quad < num_quads; quad += quad_step) {
        for (int sidx = 0; sidx < 9; ++sidx, ++step) {        // head image, then layers 7..0
          for (int sl = 0; sl < 2; ++sl) {
            const int64_t tile = quad * 4 + sl * 2 + rank;
            uint8_t* dst = g.ws + dy_base + tile * WS_DY_BYTES + (sidx == 0 ? (size_t)WS_DYH_OFF : (size_t)(8 - sidx) * ACT_BYTES);
            uint8_t* img = s_act + sl * ACT_BYTES;
            mbar_wait(&img_ready[sl], step & 1);
            for (int j = 0; j < 4; ++j) {
              if (tile < g.num_tiles) bulk_s2g(dst + j * ACT_BLK, img + j * ACT_BLK, ACT_BLK);
              bulk_commit();
              bulk_wait_read0();
            }
            mbar_arrive(&st_done[sl]);
          }
        }
      }
      bulk_wait_all0();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 10) tmem_dealloc_pair<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// 2. backward weights
// ------------------------------------------------------------------------------------------------
constexpr int WG_MAX_PIECES = 5, WG_MAX_MMA = 5, WG_JOBS = 9;

struct WgPiece {      // one operand of a job: nblk 64-column blocks of a saved tile image
  int from_dy;        // 0: forward workspace tile, 1: dy workspace tile, 2: the tile's record of second encoding blocks
  int tile_off;       // byte offset of block 0 inside the tile record
  int nblk;
  int smem_off;       // inside the stage (each block contributes 8 KB = 64 samples x 128 B)
};
struct WgMma {
  int a_off, b_off;   // stage-relative offsets of the M=128 (two blocks) A operand and of the B operand
  int N, dcol;
  int out_param;      // index into grads[], or -1 for the un-fold scratch
  int out_off;        // element offset added to the base
  int row_stride, col_stride, ncols;
};
struct WgBias {       // column sums of a dy operand sitting in the stage: bias gradients
  int smem_off;       // first block
  int ncols;          // columns summed (in units of 8)
  int nvalid;         // columns written out
  int out_param;      // grads[] index; -1: folded head bias -> un-fold scratch gb AND grads[17]
  int out_off;
};
constexpr int WG_MAX_BIAS = 3;
// Bias gradients on the tensor pipe instead: one more MMA of the dy operand (M = 128 channels: two blocks) against a
// block of ones (N = 16), column 0 of its accumulator is the column sum.  Up to two row ranges of the 128 go to
// gradient tensors (out_param as WgBias: -1 / -2 are the special targets).
struct WgBiasSeg { int row0, nrow, out_param, out_off; };
struct WgBiasMma { int a_off, dcol; WgBiasSeg seg[2]; };
constexpr int WG_MAX_BMMA = 4;
struct WgJob {
  int npieces, nmma, nbias, stage_bytes, nstage, nbmma;
  WgPiece pc[WG_MAX_PIECES];
  WgMma mm[WG_MAX_MMA];
  WgBias bs[WG_MAX_BIAS];
  WgBiasMma bm[WG_MAX_BMMA];
};
struct WgArgs {
  uint8_t* ws; int64_t num_tiles;
  uint8_t* ws_ext;                           // two-chunk encodings (WS_EXT_BYTES per tile), else unused
  float* grads[24]; float* unfold;
  const uint32_t* absmax; float fixed_scale;
  int job_first_cta[WG_JOBS + 1];
  int kind;
};
constexpr int WG_JOBS_ALL = WG_JOBS + 2;     // + the two halves of job 0 as separate jobs (layer-pipelined kernel)
// The job table depends on the network kind and on the encoding widths (leading dimensions, column counts, one or two
// encoding blocks): the host builds it per call and hands it to the kernels as a parameter (3.5 KB of the constant
// bank).  The layer-pipelined kernel serves the default encoding only and keeps its copy in c_jobs.
struct WgJobTable { WgJob j[WG_JOBS_ALL]; };
__constant__ WgJob c_jobs[2][WG_JOBS_ALL];   // [kind], default encoding (mlp_bwd_lw_kernel)

constexpr int HALF_BLK = 64 * 128;     // 64 samples of one 64-column block

__global__ void __launch_bounds__(256, 1) mlp_bwd_weight_kernel(const __grid_constant__ WgArgs g,
                                                                const __grid_constant__ WgJobTable T) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t s_full[3], s_empty[3], s_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int job = 0;
  while (job + 1 < WG_JOBS && (int)blockIdx.x >= g.job_first_cta[job + 1]) ++job;
  const WgJob& J = T.j[job];
  const int ncta = g.job_first_cta[job + 1] - g.job_first_cta[job];
  const int cta = blockIdx.x - g.job_first_cta[job];
  const int64_t t_begin = g.num_tiles * cta / ncta, t_end = g.num_tiles * (cta + 1) / ncta;
  const int64_t nhalf = (t_end - t_begin) * 2;

  if (threadIdx.x == 0) {
    // a stage is released by the MMA commit and, when the job sums bias columns on the CUDA cores, by the 4 bias warps
    for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], J.nbias > 0 ? 1 + 4 : 1); }
    mbar_init(&s_done, 1);
    mbar_fence_init();
  }
  uint8_t* s_ones = smem + J.nstage * J.stage_bytes;          // [64 samples x 64] fp16 ones: B operand of the bias MMAs
  if (J.nbmma > 0) {
    for (int i = threadIdx.x; i < HALF_BLK / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3c003c00u;
    fence_async_smem();
  }
  if (warp == 2) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t mask_base = g.num_tiles * WS_TILE_BYTES;
  const int64_t dy_base = mask_base + g.num_tiles * WS_MASK_BYTES;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t bytes = 0;
      for (int p = 0; p < J.npieces; ++p) bytes += J.pc[p].nblk * HALF_BLK;
      for (int64_t h = 0; h < nhalf; ++h) {
        uint32_t stage = h % J.nstage, ph = (h / J.nstage) & 1;
        mbar_wait(&s_empty[stage], ph ^ 1);
        mbar_expect_tx(&s_full[stage], bytes);
        const int64_t tile = t_begin + (h >> 1);
        const int half = h & 1;
        uint8_t* dst = smem + stage * J.stage_bytes;
        for (int p = 0; p < J.npieces; ++p) {
          const WgPiece& pc = J.pc[p];
          const uint8_t* src = (pc.from_dy == 2 ? g.ws_ext + tile * WS_EXT_BYTES
                                : g.ws + (pc.from_dy ? dy_base + tile * WS_DY_BYTES : tile * WS_TILE_BYTES)) +
                               pc.tile_off + half * HALF_BLK;
          for (int b = 0; b < pc.nblk; ++b)
            bulk_g2s(dst + pc.smem_off + b * HALF_BLK, src + (size_t)b * ACT_BLK, HALF_BLK, &s_full[stage]);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t smem0 = smem_u32(smem), ones = smem_u32(s_ones);
    const uint32_t idesc_b = umma_idesc_f16(128, 16, 1, 1);
    for (int64_t h = 0; h < nhalf; ++h) {
      const uint32_t stage = h % J.nstage;
      mbar_wait(&s_full[stage], (h / J.nstage) & 1);
      tc_fence_after();
      const uint32_t base = smem0 + stage * J.stage_bytes;
      if (elect_one()) {
        for (int m = 0; m < J.nmma; ++m) {
          const WgMma& mm = J.mm[m];
          const uint32_t idesc = umma_idesc_f16(128, mm.N, 1, 1);
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tmem + mm.dcol, umma_desc_mnmajor(base + mm.a_off + ks * 2048, HALF_BLK),
                     umma_desc_mnmajor(base + mm.b_off + ks * 2048, HALF_BLK), idesc, (h > 0 || ks > 0) ? 1u : 0u);
        }
        for (int m = 0; m < J.nbmma; ++m) {                     // column sums of the dy operand: bias gradients
          const WgBiasMma& bm = J.bm[m];
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tmem + bm.dcol, umma_desc_mnmajor(base + bm.a_off + ks * 2048, HALF_BLK),
                     umma_desc_mnmajor(ones + ks * 2048, HALF_BLK), idesc_b, (h > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&s_empty[stage]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&s_done);
    __syncwarp();
  } else if (warp >= 4) {
    // bias gradients on the CUDA cores while the tensor pipe / HBM stream run: column sums of the dy images.
    // Thread t of the 128: 64-column block (t >> 5), 16-byte unit (t & 7) = 8 columns, row group ((t >> 3) & 3)
    // = 16 of the stage's 64 samples; one LDS.128 per row, conflict-free (a quarter warp reads one 128-byte row).
    const int bt = threadIdx.x - 128;
    const int b_blk = bt >> 5, b_u = bt & 7, b_rg = (bt >> 3) & 3;
    float bacc[WG_MAX_BIAS][8];
#pragma unroll
    for (int b = 0; b < WG_MAX_BIAS; ++b)
#pragma unroll
      for (int e = 0; e < 8; ++e) bacc[b][e] = 0.f;
    for (int64_t h = 0; J.nbias > 0 && h < nhalf; ++h) {
      uint32_t stage = h % J.nstage;
      mbar_wait(&s_full[stage], (h / J.nstage) & 1);
      const uint8_t* st = smem + stage * J.stage_bytes;
#pragma unroll
      for (int b = 0; b < WG_MAX_BIAS; ++b) {
        if (b < J.nbias && b_blk * 64 + b_u * 8 < J.bs[b].ncols) {
          const uint8_t* blk = st + J.bs[b].smem_off + b_blk * HALF_BLK;
#pragma unroll 4
          for (int rr = 0; rr < 16; ++rr) {
            const uint4 q4 = *reinterpret_cast<const uint4*>(blk + tile_unit_off(b_rg * 16 + rr, b_u));
            const __half2* h2 = reinterpret_cast<const __half2*>(&q4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __half22float2(h2[e]);
              bacc[b][2 * e] += f.x; bacc[b][2 * e + 1] += f.y;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[stage]);
    }
    // combine the four row groups (lanes l, l^8, l^16, l^24 hold the same columns)
#pragma unroll
    for (int b = 0; b < WG_MAX_BIAS; ++b)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        bacc[b][e] += __shfl_xor_sync(0xffffffffu, bacc[b][e], 8);
        bacc[b][e] += __shfl_xor_sync(0xffffffffu, bacc[b][e], 16);
      }
    // flush: TMEM -> scaled red.add into the gradient buffers
    mbar_wait(&s_done, 0);
    tc_fence_after();
    if (nhalf > 0) {
      const float inv = 1.f / grad_scale_from(g.absmax, g.fixed_scale);
      const int q = warp & 3;
      const int r = q * 32 + lane;
#pragma unroll
      for (int b = 0; b < WG_MAX_BIAS; ++b) {
        if (b < J.nbias && b_rg == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = b_blk * 64 + b_u * 8 + e;
            if (col < J.bs[b].nvalid) {
              const float v = bacc[b][e] * inv;
              if (J.bs[b].out_param >= 0) atomicAdd(g.grads[J.bs[b].out_param] + J.bs[b].out_off + col, v);
              else if (J.bs[b].out_param == -1) { atomicAdd(g.unfold + 128 * 256 + col, v); atomicAdd(g.grads[17] + col, v); }
              else { atomicAdd(g.unfold + col, v); atomicAdd(g.grads[1] + col, v); }   // -2: d b0 of THIS call + grads
            }
          }
        }
      }
      for (int m = 0; m < J.nbmma; ++m) {                       // bias MMAs: column 0 of the N = 16 accumulator, row r
        const WgBiasMma& bm = J.bm[m];
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + bm.dcol, v);
        tmem_ld_wait();
        const float val = __uint_as_float(v[0]) * inv;
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
          const WgBiasSeg& S = bm.seg[sg];
          const int col = r - S.row0;
          if (col >= 0 && col < S.nrow) {
            if (S.out_param >= 0) atomicAdd(g.grads[S.out_param] + S.out_off + col, val);
            else if (S.out_param == -1) { atomicAdd(g.unfold + 128 * 256 + S.out_off + col, val); atomicAdd(g.grads[17] + S.out_off + col, val); }
            else { atomicAdd(g.unfold + S.out_off + col, val); atomicAdd(g.grads[1] + S.out_off + col, val); }   // -2: d b0 of THIS call + grads
          }
        }
      }
      for (int m = 0; m < J.nmma; ++m) {
        const WgMma& mm = J.mm[m];
        float* base = (mm.out_param >= 0 ? g.grads[mm.out_param] : g.unfold) + mm.out_off + (size_t)r * mm.row_stride;
        // rows of 16-byte-aligned, contiguous outputs (the 256-wide blocks) go out as four-wide reductions
        const bool vec = mm.col_stride == 1 && (mm.ncols & 31) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
        for (int c0 = 0; c0 < mm.ncols; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mm.dcol + c0, v);
          tmem_ld_wait();
          if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              red_add_v4(base + c0 + i, __uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv,
                         __uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < mm.ncols) atomicAdd(base + (size_t)(c0 + i) * mm.col_stride, __uint_as_float(v[i]) * inv);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// 2b. backward weights of the seven 256 x 256 trunk layers on CTA pairs (cta_group::2)
//
// dW_l = dy_l^T x_l with M = 256 output channels split over the two CTAs of a cluster: CTA r streams only ITS half of
// dy_l (blocks 2r, 2r+1 = its 128 output channels, the A operand) and ITS half of x_l (blocks 2r, 2r+1 = half of the
// N = 256 input channels, the B operand the pair shares), so a 64-sample stage is 32 KB per SM instead of 64 KB, the
// ring is 6 stages deep instead of 3, a stage's MMAs take 4 x 128 cycles instead of 8 x 161, and each CTA's
// accumulator is 256 TMEM columns - which leaves room for the bias gradient as one more MMA against a block of ones
// (column sums on the tensor pipe; the CUDA-core sums out of shared memory cost 0.25 ms per step).  Work is the
// concatenation of the seven layers' tile lists cut into equal slices, one per pair (at most two layers per pair).
// Loads are tensor-map copies that report to the leader's barrier; completion is a multicast commit.
struct WgPairArgs {
  int64_t num_tiles;
  float* grads[24];
  const uint32_t* absmax; float fixed_scale;
  int kind;
};
struct WgPairMaps { CUtensorMap fwd, dy; };      // workspace as [16-KB block][128 rows][128 B], box = 2 blocks x 64 rows
constexpr int WGP_STAGE = 4 * HALF_BLK;           // dy half (2 blocks) | x half (2 blocks), 64 samples each
constexpr int WGP_NST = 6;
constexpr int WGP_SMEM = WGP_NST * WGP_STAGE + HALF_BLK + 1024;      // + ones block + alignment slack
constexpr int WGP_JOBS = 7;                       // jobs 1..7 of the table

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
mlp_bwd_weight_pair_kernel(const __grid_constant__ WgPairArgs g, const __grid_constant__ WgPairMaps tm,
                           const __grid_constant__ WgJobTable T) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ones = smem + WGP_NST * WGP_STAGE;
  __shared__ uint64_t s_full[WGP_NST], s_empty[WGP_NST], s_done, flush_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  // this pair's slice of the work list (layer-major tiles)
  const int64_t W = (int64_t)WGP_JOBS * g.num_tiles;
  const int64_t w_begin = W * pair / n_pairs, w_end = W * (pair + 1) / n_pairs;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WGP_NST; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
    mbar_init(&s_done, 1);
    mbar_init(&flush_done, 8);               // leader: the four flush warps of both CTAs have read the accumulators
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < HALF_BLK / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3c003c00u;   // fp16 1.0
  fence_async_smem();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_alloc_pair<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ===================== producer: this CTA's halves of dy_l and x_l, 64 samples per stage =====================
    if (lane == 0) {
      const uint32_t s_full_l = mapa_u32(smem_u32(s_full), 0);
      uint32_t cnt = 0;
      for (int64_t w = w_begin; w < w_end; ++w) {
        const int job = 1 + (int)(w / g.num_tiles);
        const int64_t tile = w % g.num_tiles;
        const WgJob& J = T.j[job];
        const int dy_blk = (int)(tile * 36) + J.pc[0].tile_off / ACT_BLK + 2 * (int)rank;
        const int x_blk = (int)(tile * 36) + J.pc[1].tile_off / ACT_BLK + 2 * (int)rank;
        for (int half = 0; half < 2; ++half, ++cnt) {
          const uint32_t stage = cnt % WGP_NST, ph = (cnt / WGP_NST) & 1;
          mbar_wait(&s_empty[stage], ph ^ 1);
          if (leader) mbar_expect_tx(&s_full[stage], 2u * WGP_STAGE);
          uint8_t* dst = smem + stage * WGP_STAGE;
          tma_load_3d_pair(dst, &tm.dy, 0, half * 64, dy_blk, s_full_l + stage * 8);
          tma_load_3d_pair(dst + 2 * HALF_BLK, &tm.fwd, 0, half * 64, x_blk, s_full_l + stage * 8);
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader) =====================
    const uint32_t idesc = umma_idesc_f16(256, 256, 1, 1), idesc_b = umma_idesc_f16(256, 32, 1, 1);
    const uint32_t smem0 = smem_u32(smem), ones = smem_u32(s_ones);
    uint32_t cnt = 0, seg = 0;
    for (int64_t w = w_begin; w < w_end; ++w) {
      const bool first = (w == w_begin) || (w % g.num_tiles == 0);     // a new layer starts: fresh accumulators
      const bool last = (w + 1 == w_end) || ((w + 1) % g.num_tiles == 0);
      if (first && seg > 0) mbar_wait_cluster(&flush_done, (seg - 1) & 1);   // the previous layer's accumulators were read
      for (int half = 0; half < 2; ++half, ++cnt) {
        const uint32_t stage = cnt % WGP_NST;
        mbar_wait(&s_full[stage], (cnt / WGP_NST) & 1);
        tc_fence_after();
        const uint32_t base = smem0 + stage * WGP_STAGE;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (first && half == 0 && ks == 0) ? 0u : 1u;
            const uint64_t ad = umma_desc_mnmajor(base + ks * 2048, HALF_BLK);
            umma_f16_pair(tmem, ad, umma_desc_mnmajor(base + 2 * HALF_BLK + ks * 2048, HALF_BLK), idesc, acc);
            umma_f16_pair(tmem + 256, ad, umma_desc_mnmajor(ones + ks * 2048, HALF_BLK), idesc_b, acc);   // column sums
          }
          umma_commit_pair(&s_empty[stage], 3);
          if (last && half == 1) umma_commit_pair(&s_done, 3);
        }
        __syncwarp();
      }
      if (last) ++seg;
    }
  }
  // (flush code below runs for warps 4-7 of both CTAs)
  if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const float inv = 1.f / grad_scale_from(g.absmax, g.fixed_scale);
    const uint32_t flush_done_l = mapa_u32(smem_u32(&flush_done), 0);
    uint32_t seg = 0;
    for (int64_t w = w_begin; w < w_end;) {
      const int job = 1 + (int)(w / g.num_tiles);
      int64_t seg_end = (int64_t)job * g.num_tiles;            // first work item of the next layer
      if (seg_end > w_end) seg_end = w_end;
      const WgJob& J = T.j[job];
      const WgMma& mm = J.mm[rank];
      mbar_wait(&s_done, seg & 1);
      tc_fence_after();
      float* base = g.grads[mm.out_param] + mm.out_off + (size_t)r * mm.row_stride;
      const bool vec = (reinterpret_cast<uintptr_t>(base) & 15) == 0;     // every trunk layer but the skip layer (ld 319)
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (vec) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            red_add_v4(base + c0 + i, __uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv,
                       __uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(base + c0 + i, __uint_as_float(v[i]) * inv);
        }
      }
      if (J.nbias > 0) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256, v);
        tmem_ld_wait();
        atomicAdd(g.grads[J.bs[0].out_param] + J.bs[0].out_off + rank * 128 + r, __uint_as_float(v[0]) * inv);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(flush_done_l);
      w = seg_end;
      ++seg;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// 3. un-fold of the head (model.py:50-55):  G = d W_fv [128,256], gb = d b_fv [128]
//      dW_f[c][k]  += sum_u W_v[u][c] G[u][k]
//      dW_v[u][c]  += sum_k G[u][k] W_f[c][k] + gb[u] b_f[c]
//      db_f[c]     += sum_u W_v[u][c] gb[u]
// ------------------------------------------------------------------------------------------------
// 64 blocks: four rows c of dW_f each (thread k streams column k of G, coalesced; the rows' W_v columns sit in shared
// memory) | 32 blocks: four rows u of dW_v[:, :256] each (an NT product: W_f goes through a transposed 256 x 32 shared-
// memory tile per 32 k) | 1 block: db_f.  Every element has one writer; atomicAdd because the buffers accumulate.
constexpr int UNFOLD_BLOCKS = 64 + 32 + 1;
__global__ void __launch_bounds__(256) unfold_head_kernel(const float* __restrict__ Wv, const float* __restrict__ Wf,
                                                           const float* __restrict__ bf, const float* __restrict__ G,
                                                           const float* __restrict__ gb, float* __restrict__ dWf,
                                                           float* __restrict__ dbf, float* __restrict__ dWv,
                                                           int ldv /* 256 + view columns */) {
  __shared__ float sa[4][256];
  __shared__ float tile[256][33];
  const int t = threadIdx.x;
  int blk = blockIdx.x;
  if (blk < 64) {                     // dW_f[c0 + r][k] += sum_u W_v[u][c0 + r] G[u][k]
    const int c0 = blk * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {     // 4 x 128 entries of W_v's columns c0 .. c0 + 3
      const int e = t + 256 * i, r = e & 3, u = e >> 2;
      sa[r][u] = Wv[(size_t)u * ldv + c0 + r];
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int u0 = 0; u0 < 128; u0 += 32) {                      // 32 independent loads in flight per thread
      float g[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) g[i] = __ldg(G + (size_t)(u0 + i) * 256 + t);
#pragma unroll
      for (int i = 0; i < 32; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = fmaf(sa[r][u0 + i], g[i], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) atomicAdd(dWf + (size_t)(c0 + r) * 256 + t, acc[r]);
  } else if (blk < 96) {              // dW_v[u0 + r][c] += sum_k G[u0 + r][k] W_f[c][k] + gb[u0 + r] b_f[c]
    const int u0 = (blk - 64) * 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) sa[r][t] = G[(size_t)(u0 + r) * 256 + t];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int lane = t & 31, w = t >> 5;
    for (int k0 = 0; k0 < 256; k0 += 32) {
      __syncthreads();                // (also orders the sa[] writes before their first use)
      {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __ldg(Wf + (size_t)(w + 8 * i) * 256 + k0 + lane);      // W_f[c][k0 + lane]
#pragma unroll
        for (int i = 0; i < 32; ++i) tile[w + 8 * i][lane] = v[i];
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const float wv = tile[t][kk];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = fmaf(sa[r][k0 + kk], wv, acc[r]);
      }
    }
    const float b = bf[t];
#pragma unroll
    for (int r = 0; r < 4; ++r) atomicAdd(dWv + (size_t)(u0 + r) * ldv + t, fmaf(gb[u0 + r], b, acc[r]));
  } else {                            // db_f[c] += sum_u W_v[u][c] gb[u]
    if (t < 128) sa[0][t] = gb[t];
    __syncthreads();
    float acc = 0.f;
    for (int u0 = 0; u0 < 128; u0 += 32) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __ldg(Wv + (size_t)(u0 + i) * ldv + t);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc = fmaf(v[i], sa[0][u0 + i], acc);
    }
    atomicAdd(dbf + t, acc);
  }
}

// deformation net: d _time.0.weight[:, pc : pc + tw] += d b0 (this call's, 256 values) x PE(t)   (K = 1 outer product)
__global__ void __launch_bounds__(256) time_outer_kernel(const float* __restrict__ db0, const float* __restrict__ tpe,
                                                          float* __restrict__ dW0, int ld, int pc, int tw) {
  const int m = threadIdx.x;
  const float a = db0[m];
  float* row = dW0 + (size_t)m * ld + pc;
  for (int n = 0; n < tw; ++n) row[n] += a * __ldg(tpe + n);
}

// ------------------------------------------------------------------------------------------------
// 4. input gradient of the canonical net (D-NeRF: PE sits inside the graph, model.py:148-149)
//      dPE[s, c] = sum_n dy0[s, n] W0[n, c] + dy5[s, n] W5[n, c]   (c < 63)        tcgen05, 128 x 64 x 512 per tile
//      dx[s, j]  = dPE[j] + sum_k 2^k (cos(2^k x_j) dPE[3+6k+j] - sin(2^k x_j) dPE[6+6k+j])      (d embed / d x)
// ------------------------------------------------------------------------------------------------
struct DpeArgs {
  uint8_t* ws; int64_t num_tiles; int64_t P;
  const uint8_t* packed_t; const float* pts; float* d_pts;
  const uint32_t* absmax; float fixed_scale;
};

// One launch per 64-column chunk GRP of the position encoding (L = 20: two launches, the second accumulates).
template <int L, int GRP>
__global__ void __launch_bounds__(192, 1) mlp_bwd_input_kernel(DpeArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                         // 8 x [64 x 64] weights, resident
  uint8_t* s_dy = smem + 8 * DPE_CHUNK_B;      // dy0 | dy5 images of the current tile (2 x 64 KB)
  __shared__ uint64_t b_w, b_full, b_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&b_w, 1); mbar_init(&b_full, 1); mbar_init(&b_done, 1); mbar_fence_init(); }
  if (warp == 5) tmem_alloc<64>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t dy_base = g.num_tiles * (WS_TILE_BYTES + WS_MASK_BYTES);
  const float inv = 1.f / grad_scale_from(g.absmax, g.fixed_scale);
  if (threadIdx.x == 128) {
    mbar_expect_tx(&b_w, 8 * DPE_CHUNK_B);
    bulk_g2s(s_w, g.packed_t + PKT_DPE_OFF + GRP * 8 * DPE_CHUNK_B, 8 * DPE_CHUNK_B, &b_w);
  }
  uint32_t it = 0;
  for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
    if (warp == 4) {
      if (lane == 0) {
        const uint8_t* src = g.ws + dy_base + tile * WS_DY_BYTES;
        mbar_expect_tx(&b_full, 2 * ACT_BYTES);
        bulk_g2s(s_dy, src + 0 * ACT_BYTES, ACT_BYTES, &b_full);
        bulk_g2s(s_dy + ACT_BYTES, src + 5 * (size_t)ACT_BYTES, ACT_BYTES, &b_full);
      }
      if (it == 0) mbar_wait(&b_w, 0);
      mbar_wait(&b_full, it & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_f16(128, 64, 0, 0);
        const uint32_t a0 = smem_u32(s_dy), w0 = smem_u32(s_w);
        for (int ch = 0; ch < 8; ++ch)
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tmem, umma_desc_kmajor(a0 + ch * ACT_BLK + ks * 32), umma_desc_kmajor(w0 + ch * DPE_CHUNK_B + ks * 32),
                     idesc, (ch > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&b_done);
      }
      __syncwarp();
    } else if (warp < 4) {
      mbar_wait(&b_done, it & 1);
      tc_fence_after();
      const int row = warp * 32 + lane;
      const int64_t idx = tile * TILE + row;
      uint32_t va[32], vb[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), va);
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 32, vb);
      tmem_ld_wait();
      if (idx < g.P) {
        float d[64];
#pragma unroll
        for (int i = 0; i < 32; ++i) { d[i] = __uint_as_float(va[i]); d[32 + i] = __uint_as_float(vb[i]); }
        constexpr int lo = 64 * GRP, hi = lo + 64;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float x = __ldg(g.pts + idx * 3 + j);
          const float h1 = 0.15915494f, l1 = 6.4206e-9f;   // 1/(2 pi) = h1 + l1
          const float th = x * h1;
          const float tl = fmaf(x, h1, -th) + x * l1;
          float acc = GRP == 0 ? d[j] : 0.f;
#pragma unroll
          for (int k = 0; k < L; ++k) {
            const int cs = 3 + 6 * k + j, cc = 6 + 6 * k + j;       // columns of sin / cos (2^k x_j): compile-time once unrolled
            const bool s_in = cs >= lo && cs < hi, c_in = cc >= lo && cc < hi;
            if (s_in || c_in) {
              float f = (float)(1 << k), sn, cs_;
              sincos_turns(th, tl, f, sn, cs_);             // exact range reduction in turns, as the forward's encoder
              if (s_in) acc += f * cs_ * d[s_in ? cs - lo : 0];
              if (c_in) acc -= f * sn * d[c_in ? cc - lo : 0];
            }
          }
          if (GRP == 0) g.d_pts[idx * 3 + j] = acc * inv;
          else g.d_pts[idx * 3 + j] += acc * inv;
        }
      }
      tc_fence_before();
    }
    // the next tile's images may only land once every consumer of this tile is done
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<64>(tmem);
}

// ------------------------------------------------------------------------------------------------
// 5. Layer-pipelined backward ("LW"): dy never reaches HBM
//
// The two-kernel backward above is HBM-bound by construction: the data-gradient kernel writes every dy_l image
// (4.5 KB / sample) and the weight-gradient kernels read them back next to the saved activations (9.7 KB / sample).
// Here ONE persistent kernel runs both as a pipeline of layer workers.  Every CTA owns one role for its whole life:
//     D(H), D(7) .. D(1)   data gradient through ONE layer: dy_l -> dy_{l-1} = (dy_l W_l) * relu'(layer l-1), with
//                          the layer's transposed weights RESIDENT in shared memory (no weight streaming at all)
//     W(H), W(7) .. W(0)   weight / bias gradient of ONE layer group, accumulated in tensor memory over ALL of the
//                          role's tiles (the jobs of mlp_bwd_weight_kernel), one red.add flush at the end
// A tile's dy_l image travels from D(l+1) to its two consumers D(l) and W(l) through a small ring of 64-KB slots in
// global memory (9 rings x 32 slots = 19 MB, rewritten continuously, so it lives in the 126 MB L2) guarded by two
// words per slot: `ready` (tile index + 1, st.release by the producer after its stores) and `done` (a running count
// of consumers that have pulled the slot into shared memory).  Only the saved forward activations are read from HBM
// (4.7 KB / sample).  CTAs of a role take tiles round-robin (tile = j, j + P, ...), every role walks its tiles in
// increasing order and all CTAs are co-resident (grid <= SM count), so the pipeline cannot deadlock: the smallest
// unproduced tile of any ring always has a free slot once its predecessors were consumed, and the last stage (the W
// roles) has no output ring.
// ------------------------------------------------------------------------------------------------
constexpr int LW_ROLES = 18;         // 0: D(H) | 1..7: D(7)..D(1) | 8: W(H) (job 8) | 9..15: W(7)..W(1) (jobs 7..1) | 16: W(0) (job 9) | 17: W(5,PE) (job 10)
constexpr int LW_RINGS = 9;          // 0: dyH | 1 + (7 - l): dy_l, l = 7..0
constexpr int LW_R = 32;             // slots per ring
constexpr int LW_IN_BLKS = 6;        // D roles: input ring of 16-KB blocks in shared memory
constexpr int64_t LW_RING_BYTES = (int64_t)LW_RINGS * LW_R * ACT_BYTES;
constexpr int LW_FLAG_WORDS = 2 * LW_RINGS * LW_R;
constexpr int LW_FLAGS_OFF = 200 * 1024;            // inside the workspace tail (WS_TAIL_BYTES = 256 KB)
static_assert(256 + UNFOLD_FLOATS * 4 <= LW_FLAGS_OFF && LW_FLAGS_OFF + LW_FLAG_WORDS * 4 <= WS_TAIL_BYTES, "workspace tail layout");

struct LwArgs {
  const float* d_raw; int64_t P; int64_t num_tiles;
  const uint8_t* packed; const uint8_t* packed_t;
  uint8_t* ws;                 // forward workspace: saved activation images and sign masks
  uint8_t* ring;               // [LW_RINGS][LW_R] slots of 64 KB (lives in the dy region of the workspace)
  uint32_t* flags;             // ready[LW_RINGS][LW_R] | done[LW_RINGS][LW_R], zeroed before the launch
  float* grads[24]; float* unfold;
  const uint32_t* absmax; float fixed_scale;
  int role_first[LW_ROLES + 1];
  int wrgb_slot;               // which copy of rgb_linear.weight in constant memory (c_wrgb)
  unsigned long long* dbg;     // -DSWNERF_LW_DEBUG builds: 12 cycle counters per CTA (tools/lw_profile.py), else unused
};
#ifdef SWNERF_LW_DEBUG
#define LW_T0() const long long lw_t0_ = clock64()
#define LW_ACC(slot) do { if (g.dbg) atomicAdd(g.dbg + (size_t)blockIdx.x * 12 + (slot), (unsigned long long)(clock64() - lw_t0_)); } while (0)
#else
#define LW_T0() do {} while (0)
#define LW_ACC(slot) do {} while (0)
#endif

__device__ __forceinline__ int lw_ring_of_layer(int l) { return 1 + (7 - l); }            // dy_l
__device__ __forceinline__ int lw_ring_consumers(int ring) { return (ring == 0 || ring == 8) ? 1 : (ring == 3 ? 3 : 2); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// the producer of tile t has published slot t % LW_R of `ring`
__device__ __forceinline__ void lw_wait_ready(const LwArgs& g, int ring, int64_t t) {
  const uint32_t* f = g.flags + ring * LW_R + (int)(t % LW_R);
  while (ld_acquire_gpu(f) != (uint32_t)(t + 1)) __nanosleep(40);
}
// every consumer of the slot's previous occupant (tile t - LW_R) has pulled it out
__device__ __forceinline__ void lw_wait_free(const LwArgs& g, int ring, int64_t t) {
  if (t < LW_R) return;
  const uint32_t* f = g.flags + (LW_RINGS + ring) * LW_R + (int)(t % LW_R);
  const uint32_t need = (uint32_t)(lw_ring_consumers(ring) * (t / LW_R));
  while (ld_acquire_gpu(f) < need) __nanosleep(40);
}
__device__ __forceinline__ void lw_mark_done(const LwArgs& g, int ring, int64_t t) {
  red_release_gpu_add(g.flags + (LW_RINGS + ring) * LW_R + (int)(t % LW_R), 1u);
}
__device__ __forceinline__ uint8_t* lw_slot(const LwArgs& g, int ring, int64_t t) {
  return g.ring + ((int64_t)ring * LW_R + (t % LW_R)) * ACT_BYTES;
}

// rgb_linear.weight [3][128] of the network whose backward is running, one slot per stream (const_slot_acquire):
// every access is warp-uniform, so the values are FFMA operands straight from the constant bank.
constexpr int WRGB_SLOTS = 4;
__constant__ float c_wrgb[WRGB_SLOTS][384];

__device__ __forceinline__ void bulk_wait_group1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
// publish slot t % LW_R of `ring` after the bulk stores that filled it have completed
__device__ __forceinline__ void lw_publish(const LwArgs& g, int ring, int64_t t) {
  fence_proxy_async_all();
  __threadfence();
  st_release_gpu(g.flags + ring * LW_R + (int)(t % LW_R), (uint32_t)(t + 1));
}

__global__ void __launch_bounds__(384, 1) mlp_bwd_lw_kernel(const __grid_constant__ LwArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[32];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int role = 0;
  while (role + 1 < LW_ROLES && (int)blockIdx.x >= g.role_first[role + 1]) ++role;
  const int P = g.role_first[role + 1] - g.role_first[role];
  const int j = blockIdx.x - g.role_first[role];
  const int64_t my_tiles = ((int64_t)j < g.num_tiles) ? (g.num_tiles - j + P - 1) / P : 0;
  const int64_t mask_base = g.num_tiles * WS_TILE_BYTES;
  const uint32_t* mask0 = reinterpret_cast<const uint32_t*>(g.ws + mask_base);
#ifdef SWNERF_LW_DEBUG
  const long long lw_start = clock64();
  if (threadIdx.x == 0 && g.dbg) { g.dbg[(size_t)blockIdx.x * 12 + 6] = (unsigned long long)my_tiles; g.dbg[(size_t)blockIdx.x * 12 + 7] = (unsigned long long)role; }
#endif

  if (role <= 7) {
    // =========================================================================================== D roles
    // D(l): the layer's transposed weights are RESIDENT in shared memory (4 chunks, 128 KB: no weight streaming); the
    // tiles' dy_l images arrive block by block (16 KB = one K-chunk) through a 6-block ring, so the next tile loads while
    // this one multiplies; two accumulators alternate, and the epilogue stores the gated fp16 result STRAIGHT to the
    // output ring slot in global memory - transposed across lanes first, so that eight lanes write one full 128-byte line.
    // D(H): three resident head chunks (96 KB) and two head-image buffers (written by the head prep; operand of the MMAs
    // and source of the bulk store to ring 0).
    const bool head = role == 0;
    const int l = 8 - role;                       // D(l): consumes dy_l, produces dy_{l-1}     (head: produces dyH and dy7)
    const int in_ring = head ? -1 : lw_ring_of_layer(l);
    const int out_ring = head ? lw_ring_of_layer(7) : lw_ring_of_layer(l - 1);
    const int mask_layer = head ? 7 : l - 1;
    const int nch = head ? 3 : 4;
    const int cbase = head ? 0 : 3 + (7 - l) * 4;
    uint8_t* s_w = smem;                          // resident weights: nch x [256 x 64] K-major images
    uint8_t* s_in = smem + nch * CHUNK_B;         // D(l): 6 x 16 KB input blocks | D(H): 2 x 64 KB head images
    uint64_t* w_bar = bars;                       // weights landed
    uint64_t* blk_full = bars + 1;                // [6]  (D(H): img_full[2], 8 arrivals)
    uint64_t* blk_empty = bars + 7;               // [6]  (D(H): img_empty[2]: the MMAs have read the image)
    uint64_t* d_full = bars + 13;                 // [2]
    uint64_t* d_empty = bars + 15;                // [2]  8 arrivals (epilogue warps)
    uint64_t* hs_done = bars + 17;                // [2]  D(H): the head image's bulk store has read the buffer
    uint64_t* pub_full = bars + 19;               // [2]  8 arrivals: the epilogue warps have stored the tile's output image
    if (threadIdx.x == 0) {
      mbar_init(w_bar, 1);
      for (int i = 0; i < LW_IN_BLKS; ++i) { mbar_init(&blk_full[i], head ? 8 : 1); mbar_init(&blk_empty[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 8); mbar_init(&hs_done[i], 1); mbar_init(&pub_full[i], 8);
      }
      mbar_fence_init();
    }
    if (warp == 10) tmem_alloc<512>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 8 && lane == 0) {                 // resident weights, once
      mbar_expect_tx(w_bar, (uint32_t)nch * CHUNK_B);
      for (int c = 0; c < nch; ++c)
        bulk_g2s(s_w + c * CHUNK_B, g.packed_t + (size_t)(cbase + c) * CHUNK_B, CHUNK_B, w_bar);
    }
    if (warp == 8 && !head) {
      // ---------------- loader (D(l)): the tiles' dy_l images block by block
      if (lane == 0) {
        {
          uint32_t cnt = 0;
          for (int64_t t = j; t < g.num_tiles; t += P) {
            { LW_T0(); lw_wait_ready(g, in_ring, t); LW_ACC(1); }
            fence_proxy_async_all();              // the image was written with generic stores by another SM
            const uint8_t* src = lw_slot(g, in_ring, t);
            for (int ci = 0; ci < 4; ++ci, ++cnt) {
              const uint32_t pos = cnt % LW_IN_BLKS, ph = (cnt / LW_IN_BLKS) & 1;
              { LW_T0(); mbar_wait(&blk_empty[pos], ph ^ 1); LW_ACC(2); }
              mbar_expect_tx(&blk_full[pos], ACT_BLK);
              bulk_g2s(s_in + pos * ACT_BLK, src + ci * ACT_BLK, ACT_BLK, &blk_full[pos]);
            }
          }
        }
      }
    } else if (warp == 9) {
      // ---------------- MMA issuer (converged warp, one elected lane issues)
      const uint32_t idesc = umma_idesc_f16(128, 256, 0, 0);
      const uint32_t w_u32 = smem_u32(s_w), in_u32 = smem_u32(s_in);
      mbar_wait(w_bar, 0);
      uint32_t cnt = 0, dcnt = 0;
      for (int64_t t = j; t < g.num_tiles; t += P, ++dcnt) {
        const uint32_t a = dcnt & 1;
        if (dcnt >= 2) { LW_T0(); mbar_wait(&d_empty[a], ((dcnt >> 1) - 1) & 1); if (lane == 0) LW_ACC(5); }
        const uint32_t d_tmem = tmem + a * 256;
        if (head) {
          { LW_T0(); mbar_wait(&blk_full[a], (dcnt >> 1) & 1); if (lane == 0) LW_ACC(4); }     // head image `a` written
          tc_fence_after();
          const uint32_t img = in_u32 + a * ACT_BYTES;
          if (elect_one()) {
            for (int ci = 0; ci < 3; ++ci) {
              const int nks = (ci == 2) ? 1 : 4;                   // sigma block: only the first 16 columns
              for (int ks = 0; ks < nks; ++ks)
                umma_f16(d_tmem, umma_desc_kmajor(img + ci * ACT_BLK + ks * 32), umma_desc_kmajor(w_u32 + ci * CHUNK_B + ks * 32),
                         idesc, (ci > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(&blk_empty[a]);
            umma_commit(&d_full[a]);
          }
          __syncwarp();
        } else {
          for (int ci = 0; ci < 4; ++ci, ++cnt) {
            const uint32_t pos = cnt % LW_IN_BLKS;
            { LW_T0(); mbar_wait(&blk_full[pos], (cnt / LW_IN_BLKS) & 1); if (lane == 0) LW_ACC(4); }
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_f16(d_tmem, umma_desc_kmajor(in_u32 + pos * ACT_BLK + ks * 32), umma_desc_kmajor(w_u32 + ci * CHUNK_B + ks * 32),
                         idesc, (ci > 0 || ks > 0) ? 1u : 0u);
              umma_commit(&blk_empty[pos]);
              if (ci == 3) umma_commit(&d_full[a]);
            }
            __syncwarp();
          }
          // all four blocks of the tile have landed in shared memory: the ring slot may be overwritten
          if (elect_one()) lw_mark_done(g, in_ring, t);
          __syncwarp();
        }
      }
    } else if (warp == 10) {
      // ---------------- publisher: the epilogue warps only STORE the output image (generic stores) and arrive on
      // pub_full; this lane makes those stores visible at GPU scope (the fence is cumulative over what it has observed
      // through the barrier) and publishes the slot, so the 1.5 - 3 us a fence waits for L2's write acknowledgements
      // never sit in the epilogue's path.
      if (lane == 0) {
        uint32_t dcnt = 0;
        for (int64_t t = j; t < g.num_tiles; t += P, ++dcnt) {
          mbar_wait(&pub_full[dcnt & 1], (dcnt >> 1) & 1);
          __threadfence();
          st_release_gpu(g.flags + out_ring * LW_R + (int)(t % LW_R), (uint32_t)(t + 1));
        }
      }
      __syncwarp();
    } else if (head && (warp == 8 || warp == 11)) {
      // ---------------- D(H): head images -> ring 0 (bulk stores from the image buffers; one warp per buffer, because a
      // 64-KB store takes ~2.8 us from issue to completion).  A slot is published as soon as ITS stores have completed -
      // never later: a producer that sat on an unpublished tile while waiting for a free slot further down its list
      // would close a cycle with its consumers.
      if (lane == 0) {
        const uint32_t b = warp == 11 ? 1u : 0u;
        const uint8_t* img = s_in + b * ACT_BYTES;
        uint32_t use = 0;
        for (int64_t t = j + (int64_t)b * P; t < g.num_tiles; t += 2 * (int64_t)P, ++use) {
          mbar_wait(&blk_full[b], use & 1);
          LW_T0();
          lw_wait_free(g, 0, t);
          uint8_t* dst = lw_slot(g, 0, t);
          for (int c = 0; c < 4; ++c) bulk_s2g(dst + c * ACT_BLK, img + c * ACT_BLK, ACT_BLK);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(&hs_done[b]);                         // the buffer may be rewritten (once the MMAs have read it too)
          bulk_wait_all0();
          lw_publish(g, 0, t);
          LW_ACC(11);
        }
      }
      __syncwarp();
    } else if (warp < 8) {
      // ---------------- epilogue warps (and, for D(H), the head image).  Warp w: TMEM lane quarter q = w & 3 (rows
      // 32q .. 32q+31), blocks 2 (w >> 2) and 2 (w >> 2) + 1 - all 64 columns of each, i.e. whole 128-byte image rows.
      const int q = warp & 3, bh = warp >> 2;
      const int row = q * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      const float scale = grad_scale_from(g.absmax, g.fixed_scale);
      const float* wrgb = c_wrgb[g.wrgb_slot];
      // D(H): head image of tile t = [dy9 0..63 | dy9 64..127 | (d_sigma, 0..) | (d_rgb, 0..)] into buffer it & 1
      // (thread: row `row`, dy9 block bh, and unit 0/1 of block 2 + bh)
      auto head_prep = [&](int64_t t, uint32_t it) {
        LW_T0();
        const uint32_t ib = it & 1;
        const uint32_t* ws_mask = mask0 + t * (9 * 8 * 128);
        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t * TILE + row < g.P) dr = __ldg(reinterpret_cast<const float4*>(g.d_raw) + t * TILE + row);
        dr.x *= scale; dr.y *= scale; dr.z *= scale; dr.w *= scale;
        const uint32_t hm0 = __ldg(ws_mask + (8 * 8 + bh * 2 + 0) * 128 + row), hm1 = __ldg(ws_mask + (8 * 8 + bh * 2 + 1) * 128 + row);
        if (it >= 2) {                                         // the buffer's previous image: MMAs and bulk store have read it
          mbar_wait(&blk_empty[ib], ((it >> 1) - 1) & 1);
          mbar_wait(&hs_done[ib], ((it >> 1) - 1) & 1);
        }
        if (threadIdx.x == 0) LW_ACC(2);
        uint8_t* img = s_in + ib * ACT_BYTES;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int c0 = bh * 64 + jj * 32;
          const uint32_t m = jj ? hm1 : hm0;
          float val[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = dr.x * wrgb[c0 + i] + dr.y * wrgb[128 + c0 + i] + dr.z * wrgb[256 + c0 + i];
            val[i] = ((m >> i) & 1u) ? v : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(img + bh * ACT_BLK + tile_unit_off(row, jj * 4 + u)) =
                make_uint4(pack_half2(val[8 * u], val[8 * u + 1]), pack_half2(val[8 * u + 2], val[8 * u + 3]),
                           pack_half2(val[8 * u + 4], val[8 * u + 5]), pack_half2(val[8 * u + 6], val[8 * u + 7]));
        }
        uint4 u0 = make_uint4(0u, 0u, 0u, 0u);
        if (bh == 0) u0.x = pack_half2(dr.w, 0.f);
        else { u0.x = pack_half2(dr.x, dr.y); u0.y = pack_half2(dr.z, 0.f); }
        *reinterpret_cast<uint4*>(img + (2 + bh) * ACT_BLK + tile_unit_off(row, 0)) = u0;
        *reinterpret_cast<uint4*>(img + (2 + bh) * ACT_BLK + tile_unit_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&blk_full[ib]);
        if (threadIdx.x == 0) LW_ACC(8);
      };
      uint32_t dcnt = 0;
      if (head && my_tiles > 0) head_prep(j, 0);
      const int lg = lane & 7, lq = lane >> 3;               // position in the 8-lane transpose group, group in the warp
      for (int64_t t = j; t < g.num_tiles; t += P, ++dcnt) {
        const uint32_t a = dcnt & 1;
        const uint32_t* ws_mask = mask0 + t * (9 * 8 * 128);
        uint32_t m[2][2];                                      // sign words of this row: [block][column half]
#pragma unroll
        for (int jb = 0; jb < 2; ++jb)
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) m[jb][h2] = __ldg(ws_mask + (mask_layer * 8 + (2 * bh + jb) * 2 + h2) * 128 + row);
        if (head && t + P < g.num_tiles) head_prep(t + P, dcnt + 1);      // next tile's image while this tile's MMAs run
        if (threadIdx.x == 0) { LW_T0(); lw_wait_free(g, out_ring, t); LW_ACC(10); }
        named_bar_sync(1, 256);
        { LW_T0(); mbar_wait(&d_full[a], (dcnt >> 1) & 1); if (threadIdx.x == 0) LW_ACC(3); }
        tc_fence_after();
        LW_T0();
        uint8_t* dst = lw_slot(g, out_ring, t);
        const uint32_t acc = tmem + lane_addr + a * 256 + bh * 128;
#pragma unroll
        for (int jb = 0; jb < 2; ++jb) {
          uint32_t va[32], vb[32];
          tmem_ld32(acc + jb * 64, va);
          tmem_ld32(acc + jb * 64 + 32, vb);
          tmem_ld_wait();
          if (jb == 1) {                         // every accumulator read of this thread is complete: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[a]);
          }
          uint4 x[8];                            // x[u] = 16-byte unit u (columns 8u .. 8u+7) of this thread's row
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t (&v)[32] = (u < 4) ? va : vb;
            const uint32_t mk = m[jb][u >> 2];
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 8 * (u & 3) + 2 * e;
              float f0 = ((mk >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
              float f1 = ((mk >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : 0.f;
              pk[e] = pack_half2(f0, f1);
            }
            x[u] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          // 8 x 8 transpose of the units inside each group of eight lanes: afterwards lane i of a group holds unit i of
          // the group's eight rows, so one store instruction writes four complete 128-byte lines (32 separate lines
          // before: the LSU spent 4096 line-cycles per tile on them)
#pragma unroll
          for (int sft = 4; sft >= 1; sft >>= 1) {
            const bool up = (lg & sft) != 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if ((u & sft) == 0) {
                const uint4 snd = up ? x[u] : x[u | sft];
                uint4 rcv;
                rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, sft); rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, sft);
                rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, sft); rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, sft);
                if (up) x[u] = rcv; else x[u | sft] = rcv;
              }
            }
          }
          uint8_t* blk = dst + (2 * bh + jb) * ACT_BLK;
#pragma unroll
          for (int k = 0; k < 8; ++k) {          // x[k] = unit lg of row q*32 + 8*lq + k; it lives at 16-byte slot lg ^ k of that row
            const int r = q * 32 + 8 * lq + k;
            *reinterpret_cast<uint4*>(blk + (r >> 3) * 1024 + k * 128 + ((lg ^ k) << 4)) = x[k];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pub_full[a]);             // (release at CTA scope: orders this warp's stores before it)
        if (threadIdx.x == 0) LW_ACC(9);
      }
    }
    tc_fence_before();
    __syncthreads();
#ifdef SWNERF_LW_DEBUG
    if (threadIdx.x == 0 && g.dbg) g.dbg[(size_t)blockIdx.x * 12] = (unsigned long long)(clock64() - lw_start);
#endif
    if (warp == 10) tmem_dealloc<512>(tmem);
    return;
  }

  // ============================================================================================= W roles
  // The jobs of mlp_bwd_weight_kernel with the dy operands pulled from the rings instead of the dy workspace.
  {
    const int job = (role == 16) ? 9 : (role == 17 ? 10 : (role == 8 ? 8 : 16 - role));      // roles 9..15 -> jobs 7..1
    const WgJob& J = c_jobs[0][job];
    uint64_t* s_full = bars;            // [3]
    uint64_t* s_empty = bars + 3;       // [3]
    uint64_t* s_done = bars + 6;
    const int64_t nhalf = my_tiles * 2;
    if (threadIdx.x == 0) {
      for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1 + 4); }
      mbar_init(s_done, 1);
      mbar_fence_init();
    }
    if (warp == 10) tmem_alloc<512>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // ring of a dy piece: its offset inside the dy tile record names the layer (8 = the head image)
    auto piece_ring = [](const WgPiece& pc) { const int L = pc.tile_off / ACT_BYTES; return L == 8 ? 0 : 1 + (7 - L); };

    if (warp == 0) {
      if (lane == 0) {
        uint32_t bytes = 0;
        for (int p = 0; p < J.npieces; ++p) bytes += J.pc[p].nblk * HALF_BLK;
        for (int64_t h = 0; h < nhalf; ++h) {
          const uint32_t stage = h % J.nstage, ph = (h / J.nstage) & 1;
          const int64_t tile = j + (h >> 1) * P;
          const int half = h & 1;
          if (half == 0) {
            LW_T0();
            for (int p = 0; p < J.npieces; ++p)
              if (J.pc[p].from_dy) lw_wait_ready(g, piece_ring(J.pc[p]), tile);
            LW_ACC(1);
            fence_proxy_async_all();
          }
          { LW_T0(); mbar_wait(&s_empty[stage], ph ^ 1); LW_ACC(2); }
          mbar_expect_tx(&s_full[stage], bytes);
          uint8_t* dst = smem + stage * J.stage_bytes;
          for (int p = 0; p < J.npieces; ++p) {
            const WgPiece& pc = J.pc[p];
            const uint8_t* src = (pc.from_dy ? lw_slot(g, piece_ring(pc), tile) + pc.tile_off % ACT_BYTES
                                             : g.ws + tile * WS_TILE_BYTES + pc.tile_off) + half * HALF_BLK;
            for (int b = 0; b < pc.nblk; ++b)
              bulk_g2s(dst + pc.smem_off + b * HALF_BLK, src + (size_t)b * ACT_BLK, HALF_BLK, &s_full[stage]);
          }
        }
      }
    } else if (warp == 1) {
      const uint32_t smem0 = smem_u32(smem);
      for (int64_t h = 0; h < nhalf; ++h) {
        const uint32_t stage = h % J.nstage;
        { LW_T0(); mbar_wait(&s_full[stage], (h / J.nstage) & 1); if (lane == 0) LW_ACC(4); }
        tc_fence_after();
        const uint32_t base = smem0 + stage * J.stage_bytes;
        if (elect_one()) {
          for (int m = 0; m < J.nmma; ++m) {
            const WgMma& mm = J.mm[m];
            const uint32_t idesc = umma_idesc_f16(128, mm.N, 1, 1);
            for (int ks = 0; ks < 4; ++ks)
              umma_f16(tmem + mm.dcol, umma_desc_mnmajor(base + mm.a_off + ks * 2048, HALF_BLK),
                       umma_desc_mnmajor(base + mm.b_off + ks * 2048, HALF_BLK), idesc, (h > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(&s_empty[stage]);
          if (h & 1) {                  // both halves of the tile are in shared memory: release its ring slots
            const int64_t tile = j + (h >> 1) * P;
            for (int p = 0; p < J.npieces; ++p)
              if (J.pc[p].from_dy) lw_mark_done(g, piece_ring(J.pc[p]), tile);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(s_done);
      __syncwarp();
    } else if (warp >= 4 && warp < 8) {
      // bias gradients on the CUDA cores + the final flush: as mlp_bwd_weight_kernel
      const int bt = threadIdx.x - 128;
      const int b_blk = bt >> 5, b_u = bt & 7, b_rg = (bt >> 3) & 3;
      float bacc[WG_MAX_BIAS][8];
#pragma unroll
      for (int b = 0; b < WG_MAX_BIAS; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) bacc[b][e] = 0.f;
      for (int64_t h = 0; h < nhalf; ++h) {
        uint32_t stage = h % J.nstage;
        mbar_wait(&s_full[stage], (h / J.nstage) & 1);
        const uint8_t* st = smem + stage * J.stage_bytes;
#pragma unroll
        for (int b = 0; b < WG_MAX_BIAS; ++b) {
          if (b < J.nbias && b_blk * 64 + b_u * 8 < J.bs[b].ncols) {
            const uint8_t* blk = st + J.bs[b].smem_off + b_blk * HALF_BLK;
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
              const uint4 q4 = *reinterpret_cast<const uint4*>(blk + tile_unit_off(b_rg * 16 + rr, b_u));
              const __half2* h2 = reinterpret_cast<const __half2*>(&q4);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(h2[e]);
                bacc[b][2 * e] += f.x; bacc[b][2 * e + 1] += f.y;
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);
      }
#pragma unroll
      for (int b = 0; b < WG_MAX_BIAS; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          bacc[b][e] += __shfl_xor_sync(0xffffffffu, bacc[b][e], 8);
          bacc[b][e] += __shfl_xor_sync(0xffffffffu, bacc[b][e], 16);
        }
      mbar_wait(s_done, 0);
      tc_fence_after();
      if (nhalf > 0) {
        const float inv = 1.f / grad_scale_from(g.absmax, g.fixed_scale);
        const int q = warp & 3;
        const int r = q * 32 + lane;
#pragma unroll
        for (int b = 0; b < WG_MAX_BIAS; ++b) {
          if (b < J.nbias && b_rg == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = b_blk * 64 + b_u * 8 + e;
              if (col < J.bs[b].nvalid) {
                const float v = bacc[b][e] * inv;
                if (J.bs[b].out_param >= 0) atomicAdd(g.grads[J.bs[b].out_param] + J.bs[b].out_off + col, v);
                else if (J.bs[b].out_param == -1) { atomicAdd(g.unfold + 128 * 256 + col, v); atomicAdd(g.grads[17] + col, v); }
              }
            }
          }
        }
        for (int m = 0; m < J.nmma; ++m) {
          const WgMma& mm = J.mm[m];
          float* base = (mm.out_param >= 0 ? g.grads[mm.out_param] : g.unfold) + mm.out_off + (size_t)r * mm.row_stride;
          for (int c0 = 0; c0 < mm.ncols; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mm.dcol + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < mm.ncols) atomicAdd(base + (size_t)(c0 + i) * mm.col_stride, __uint_as_float(v[i]) * inv);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
#ifdef SWNERF_LW_DEBUG
    if (threadIdx.x == 0 && g.dbg) g.dbg[(size_t)blockIdx.x * 12] = (unsigned long long)(clock64() - lw_start);
#endif
    if (warp == 10) tmem_dealloc<512>(tmem);
  }
}

#ifdef SWNERF_LW_DEBUG
static unsigned long long* g_lw_dbg = nullptr;
extern "C" int swnerf_tc_lw_debug(unsigned long long* dev_buf) { g_lw_dbg = dev_buf; return 0; }
#endif

// CTAs per role.  Every role costs about one "full" unit per tile (16 MMAs of 128 x 256 x 16, or the load time of its
// operands), the head roles a little less; SWNERF_LW_ROLES="n0,n1,...,n16" overrides (tuning).
static void lw_assign_roles(int n_cta, int* first) {
  // measured (tools/lw_sweep.py, 786 k samples): 12,6x7,12,9x7,9,9 -> 3.64 ms; D(l) = 5 -> 4.2 ms
  static const int weight_default[LW_ROLES] = {130, 65, 65, 65, 65, 65, 65, 65, 130, 100, 100, 100, 100, 100, 100, 100, 100, 100};
  int cnt[LW_ROLES];
  const char* env = getenv("SWNERF_LW_ROLES");
  bool ok = false;
  if (env) {
    int n = 0, tot = 0;
    const char* p = env;
    while (*p && n < LW_ROLES) { cnt[n] = atoi(p); tot += cnt[n]; ++n; while (*p && *p != ',') ++p; if (*p == ',') ++p; }
    ok = (n == LW_ROLES && tot <= n_cta);
    for (int i = 0; ok && i < LW_ROLES; ++i) ok = cnt[i] >= 1;
  }
  if (!ok) {
    int tot = 0, used = 0;
    for (int i = 0; i < LW_ROLES; ++i) tot += weight_default[i];
    for (int i = 0; i < LW_ROLES; ++i) { cnt[i] = n_cta * weight_default[i] / tot; if (cnt[i] < 1) cnt[i] = 1; used += cnt[i]; }
    // left-over CTAs go to the full-cost roles first (largest remainder would give the same within one CTA)
    for (int i = 0; used < n_cta; i = (i + 1) % LW_ROLES) { ++cnt[i]; ++used; }
  }
  first[0] = 0;
  for (int i = 0; i < LW_ROLES; ++i) first[i + 1] = first[i] + cnt[i];
}

// job table ------------------------------------------------------------------------------------
static void wg_std_job(WgJob& J, int l, int x_off /* forward tile offset of the layer input */, int col_off, int ld) {
  J.npieces = 2; J.nmma = 2; J.nbias = 1; J.stage_bytes = 8 * HALF_BLK; J.nstage = 3;
  J.bs[0] = {0, 256, 256, 2 * l + 1, 0};
  J.pc[0] = {1, l * ACT_BYTES, 4, 0};
  J.pc[1] = {0, x_off, 4, 4 * HALF_BLK};
  for (int m = 0; m < 2; ++m)
    J.mm[m] = {m * 2 * HALF_BLK, 4 * HALF_BLK, 256, m * 256, 2 * l, m * 128 * ld + col_off, ld, 1, 256};
}

// bias_mma: the layer-0 / skip / head jobs form their bias gradients on the tensor pipe (WgBiasMma) when tensor memory
// has room for the 16-column accumulators, instead of summing the dy images on the CUDA cores out of shared memory
// (whose LDS traffic competes with the tensor pipe's operand reads).  The layer-pipelined kernel keeps the CUDA-core sums.
static void build_jobs(WgJob* jobs, int kind, const Enc& E, bool bias_mma) {
  memset(jobs, 0, sizeof(WgJob) * WG_JOBS_ALL);
  const int pc = E.pc, vc = E.vc, PC = E.PC, VC = E.VC;
  const int ld0 = pc + (kind == 1 ? E.tw : 0);       // pts_linears.0 is [256, pc]; _time.0 is [256, pc + tw]
  const int ld5 = 256 + pc;                          // pts_linears.5 is [256, pc + 256]
  // job 0: PE inputs of layer 0 and of the skip layer 5:  dW0[:, :pc], dW5[:, :pc].  stage: dy0 (4 blk) | dy5 (4) | PE (PC)
  {
    WgJob& J = jobs[0];
    J.npieces = 2 + PC; J.nmma = 4; J.nbias = 2; J.stage_bytes = (8 + PC) * HALF_BLK; J.nstage = PC > 1 ? 2 : 3;
    J.bs[0] = {0, 256, 256, kind == 1 ? -2 : 1, 0};  // db of pts_linears.0 (deformation net: also kept per call, see bwd_impl)
    J.bs[1] = {4 * HALF_BLK, 256, 256, 11, 0};       // db of pts_linears.5
    J.pc[0] = {1, 0 * ACT_BYTES, 4, 0};
    J.pc[1] = {1, 5 * ACT_BYTES, 4, 4 * HALF_BLK};
    J.pc[2] = {0, WS_PE_OFF, 1, 8 * HALF_BLK};
    if (PC > 1) J.pc[3] = {2, 0, 1, 9 * HALF_BLK};   // second 64 columns of the position encoding
    for (int a = 0; a < 2; ++a)
      for (int m = 0; m < 2; ++m)
        J.mm[a * 2 + m] = {(a * 4 + m * 2) * HALF_BLK, 8 * HALF_BLK, 64 * PC, (a * 2 + m) * 64 * PC,
                           a == 0 ? 0 : 10, m * 128 * (a == 0 ? ld0 : ld5), a == 0 ? ld0 : ld5, 1, pc};
    if (bias_mma && PC == 1) {                         // accumulators take 256 of the 512 columns: room for 4 x 16 more
      J.nbias = 0; J.nbmma = 4;
      for (int a = 0; a < 2; ++a)
        for (int m = 0; m < 2; ++m) {
          J.bm[a * 2 + m] = {(a * 4 + m * 2) * HALF_BLK, 256 + (a * 2 + m) * 16,
                             {{0, 128, a == 0 ? (kind == 1 ? -2 : 1) : 11, m * 128}, {0, 0, 0, 0}}};
        }
    }
  }
  // jobs 9 / 10: the two halves of job 0 on their own (dy0 with PE -> dW0[:, :pc], db0;  dy5 with PE -> dW5[:, :pc],
  // db5).  The layer-pipelined kernel gives each its own role: a role that needed dy5 AND dy0 of a tile would hold the
  // dy5 slot until the tile has travelled five more stages.  (One-chunk encodings only.)
  for (int a = 0; a < 2; ++a) {
    WgJob& J = jobs[9 + a];
    J.npieces = 2; J.nmma = 2; J.nbias = 1; J.stage_bytes = 5 * HALF_BLK; J.nstage = 3;
    J.bs[0] = {0, 256, 256, a == 0 ? (kind == 1 ? -2 : 1) : 11, 0};
    J.pc[0] = {1, (a == 0 ? 0 : 5) * ACT_BYTES, 4, 0};
    J.pc[1] = {0, WS_PE_OFF, 1, 4 * HALF_BLK};
    for (int m = 0; m < 2; ++m)
      J.mm[m] = {m * 2 * HALF_BLK, 4 * HALF_BLK, 64, m * 64, a == 0 ? 0 : 10, m * 128 * (a == 0 ? ld0 : ld5), a == 0 ? ld0 : ld5, 1,
                 pc < 64 ? pc : 64};
  }
  for (int l = 1; l <= 4; ++l) wg_std_job(jobs[l], l, WS_H_OFF + (l - 1) * ACT_BYTES, 0, 256);
  wg_std_job(jobs[5], 5, WS_H_OFF + 4 * ACT_BYTES, pc, ld5);
  jobs[5].nbias = 0;                                 // pts_linears.5.bias is summed by job 0
  wg_std_job(jobs[6], 6, WS_H_OFF + 5 * ACT_BYTES, 0, 256);
  wg_std_job(jobs[7], 7, WS_H_OFF + 6 * ACT_BYTES, 0, 256);
  // job 8: head.  stage: dyH (4 blk) | views (VC) | h7 (4) | h9 (2)
  {
    WgJob& J = jobs[8];
    const int V = 4, H7 = 4 + VC, H9 = 8 + VC;       // first stage block of views / h7 / h9
    J.npieces = 3 + VC; J.nmma = 5; J.nbias = 3; J.stage_bytes = (10 + VC) * HALF_BLK; J.nstage = 2;
    J.bs[0] = {0, 128, 128, -1, 0};                  // folded head bias -> gb (un-fold) and views_linears.0.bias
    J.bs[1] = {2 * HALF_BLK, 8, 1, 21, 0};           // alpha_linear.bias
    J.bs[2] = {3 * HALF_BLK, 8, 3, 23, 0};           // rgb_linear.bias
    J.pc[0] = {1, WS_DYH_OFF, 4, 0};
    J.pc[1] = {0, WS_VW_OFF, 1, V * HALF_BLK};
    J.pc[2] = {0, WS_H_OFF + 7 * ACT_BYTES, 4, H7 * HALF_BLK};
    J.pc[3] = {0, WS_H9_OFF, 2, H9 * HALF_BLK};
    if (VC > 1) J.pc[4] = {2, ACT_BLK, 1, (V + 1) * HALF_BLK};               // second 64 columns of the view encoding
    const int c0 = 64 * VC;                          // accumulator columns behind the view block
    J.mm[0] = {0, V * HALF_BLK, 64 * VC, 0, 16, 256, 256 + vc, 1, vc};       // dW_v[:, 256:256+vc] = dy9^T views
    J.mm[1] = {0, H7 * HALF_BLK, 256, c0, -1, 0, 256, 1, 256};               // G = dy9^T h7  (d W_fv)
    J.mm[2] = {H7 * HALF_BLK, 2 * HALF_BLK, 16, c0 + 256, 20, 0, 1, 0, 1};           // d w_alpha[0:128]   = h7^T d_sigma
    J.mm[3] = {(H7 + 2) * HALF_BLK, 2 * HALF_BLK, 16, c0 + 272, 20, 128, 1, 0, 1};   // d w_alpha[128:256]
    J.mm[4] = {H9 * HALF_BLK, 3 * HALF_BLK, 16, c0 + 288, 22, 0, 1, 128, 3};         // dW_rgb[j][c] = h9^T d_rgb
    if (bias_mma) {
      J.nbias = 0; J.nbmma = 2;
      J.bm[0] = {0, c0 + 304, {{0, 128, -1, 0}, {0, 0, 0, 0}}};                      // dy9 (blocks 0, 1): folded head bias
      J.bm[1] = {2 * HALF_BLK, c0 + 320, {{0, 1, 21, 0}, {64, 3, 23, 0}}};           // block 2 col 0 = d_sigma; block 3 cols 0..2 = d_rgb
    }
    if (kind == 1) {
      // deformation net: only its 256->3 output layer lives in the head.  stage: dyH block 2 | h7 (4 blocks)
      memset(&J, 0, sizeof(J));
      J.npieces = 2; J.nmma = 2; J.nbias = 1; J.stage_bytes = 5 * HALF_BLK; J.nstage = 3;
      J.pc[0] = {1, WS_DYH_OFF + 2 * ACT_BLK, 1, 0};
      J.pc[1] = {0, WS_H_OFF + 7 * ACT_BYTES, 4, 1 * HALF_BLK};
      J.mm[0] = {1 * HALF_BLK, 0, 16, 0, 16, 0, 1, 256, 3};                  // dW_out[j][c], c < 128  = h7^T d_dx
      J.mm[1] = {3 * HALF_BLK, 0, 16, 16, 16, 128, 1, 256, 3};               // c >= 128
      J.bs[0] = {0, 8, 3, 17, 0};                                            // _time_out.bias
    }
  }
}

static void assign_ctas(int n_cta, int* first, int kind, bool edge_only) {
  // CTAs per job in proportion to the bytes a job streams per tile; with edge_only the seven 256 x 256 trunk layers
  // (jobs 1..7) run on the CTA-pair kernel and get no CTAs here
  const int t = edge_only ? 0 : 128;
  const int w[WG_JOBS] = {144, t, t, t, t, t, t, t, kind == 1 ? 80 : 176};
  int tot = 0;
  for (int j = 0; j < WG_JOBS; ++j) tot += w[j];
  int cnt[WG_JOBS], used = 0;
  for (int j = 0; j < WG_JOBS; ++j) { cnt[j] = n_cta * w[j] / tot; if (cnt[j] < 1 && w[j] > 0) cnt[j] = 1; used += cnt[j]; }
  for (int j = 0; used < n_cta; j = (j + 1) % WG_JOBS) if (w[j] > 0) { ++cnt[j]; ++used; }
  first[0] = 0;
  for (int j = 0; j < WG_JOBS; ++j) first[j + 1] = first[j] + cnt[j];
}

}  // namespace swnerf

using namespace swnerf;

// optional per-kernel device timing of the backward (bench.py): events on the launching stream
// (process-wide, not thread-local: autograd runs the backward on its own engine thread)
static int g_prof = 0;
static cudaEvent_t g_ev[4] = {nullptr, nullptr, nullptr, nullptr};
static int g_prof_valid = 0;

// -1 / 0 = the two-kernel backward (default), 1 = the layer-pipelined kernel for launches of more than two tiles per SM.
// SWNERF_BWD_LW=0/1 presets it.
static std::atomic<int> g_bwd_variant{[] { const char* e = getenv("SWNERF_BWD_LW"); return e ? (atoi(e) ? 1 : 0) : -1; }()};

extern "C" {

int swnerf_tc_set_bwd_variant(int variant) {
  SW_REQUIRE(variant >= -1 && variant <= 1, "tc_set_bwd_variant: variant must be -1 (automatic), 0 or 1");
  g_bwd_variant.store(variant);
  return SWNERF_OK;
}

int swnerf_tc_set_profiling(int on) {
  g_prof = on;
  g_prof_valid = 0;
  if (on && !g_ev[0])
    for (int i = 0; i < 4; ++i)
      if (cudaEventCreate(&g_ev[i]) != cudaSuccess) return set_err(SWNERF_ERR_CUDA, "cudaEventCreate failed");
  return SWNERF_OK;
}

int swnerf_tc_last_bwd_ms(float* data_ms, float* weight_ms) {
  SW_REQUIRE(data_ms && weight_ms, "tc_last_bwd_ms: null pointer");
  SW_REQUIRE(g_prof && g_prof_valid, "tc_last_bwd_ms: profiling is off or no backward has run");
  if (cudaEventSynchronize(g_ev[2]) != cudaSuccess) return set_err(SWNERF_ERR_CUDA, "event sync failed");
  cudaEventElapsedTime(data_ms, g_ev[0], g_ev[1]);
  cudaEventElapsedTime(weight_ms, g_ev[1], g_ev[2]);
  return SWNERF_OK;
}

int64_t swnerf_tc_packed_t_bytes(void) { return PKT_TOTAL_BYTES; }

static int pack_t_impl(const float* const* params, int kind, int enc, const void* packed, void* packed_t, void* stream) {
  SW_REQUIRE(params && packed && packed_t, "tc_pack_weights_t: null pointer");
  SW_REQUIRE(aligned16(packed_t), "tc_pack_weights_t: packed_t must be 16-byte aligned");
  ParamPtrsB P;
  SW_REQUIRE(decode_enc(enc, &P.enc), "tc_pack_weights_t: unsupported encoding code 0x%x", enc);
  const int np = kind == 0 ? 24 : 18;
  for (int i = 0; i < 24; ++i) {
    SW_REQUIRE(i >= np || params[i], "tc_pack_weights_t: null parameter %d", i);
    P.p[i] = i < np ? params[i] : nullptr;
  }
  P.kind = kind;
  pack_bwd_kernel<<<(PKT_TOTAL_BYTES / 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      P, reinterpret_cast<const uint8_t*>(packed), reinterpret_cast<uint8_t*>(packed_t));
  return check_launch("tc_pack_weights_t");
}

int swnerf_tc_pack_weights_t(const float* const* params, int enc, const void* packed, void* packed_t, void* stream) {
  return pack_t_impl(params, 0, enc, packed, packed_t, stream);
}
int swnerf_tc_pack_weights_time_t(const float* const* params, int enc, const void* packed, void* packed_t, void* stream) {
  return pack_t_impl(params, 1, enc, packed, packed_t, stream);
}

}  // extern "C"

template <int L, int GRP>
static int launch_input_grad(const DpeArgs& d, int grid, int smem, cudaStream_t s) {
  if (once_per_device(ONCE_BWD_INPUT_BASE + (L == 0 ? 0 : L == 4 ? 1 : L == 10 ? 2 : 3 + GRP)))
    cudaFuncSetAttribute(mlp_bwd_input_kernel<L, GRP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mlp_bwd_input_kernel<L, GRP><<<grid, 192, smem, s>>>(d);
  return check_launch("tc_mlp_bwd_input");
}

static int bwd_impl(const float* d_out, int64_t n_rays, int n_samples, const void* packed, const void* packed_t, int enc,
                    const float* const* params, void* workspace, float* const* grads, float grad_scale, int kind,
                    const float* tpe_dev, const float* pts, float* d_pts, void* stream) {
  Enc E;
  SW_REQUIRE(decode_enc(enc, &E), "tc_mlp_bwd: unsupported encoding code 0x%x", enc);
  SW_REQUIRE(d_out && packed && packed_t && params && workspace && grads, "tc_mlp_bwd: null pointer");
  SW_REQUIRE(aligned16(workspace) && (kind == 1 || aligned16(d_out)), "tc_mlp_bwd: buffers must be 16-byte aligned");
  SW_REQUIRE(grad_scale >= 0.f, "tc_mlp_bwd: grad_scale must be >= 0 (0 = automatic)");
  SW_REQUIRE(kind == 0 || tpe_dev, "tc_mlp_bwd: the deformation net needs the time embedding on the device");
  SW_REQUIRE(!d_pts || (pts && kind == 0), "tc_mlp_bwd: the input gradient needs the sample positions (canonical net only)");
  if (n_rays == 0) return SWNERF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int np = kind == 0 ? 24 : 18;
  const int64_t P = n_rays * n_samples;
  const int64_t tiles = (P + TILE - 1) / TILE;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* tail = ws + tiles * (WS_TILE_BYTES + WS_MASK_BYTES + WS_DY_BYTES);
  uint32_t* absmax = reinterpret_cast<uint32_t*>(tail);
  float* unfold = reinterpret_cast<float*>(tail + 256);
  cudaMemsetAsync(tail, 0, 256 + UNFOLD_FLOATS * sizeof(float), s);
  if (grad_scale == 0.f) {
    absmax_kernel<<<sm_count() * 4, 256, 0, s>>>(d_out, P * (kind == 0 ? 4 : 3), absmax);
    int rc = check_launch("tc_absmax");
    if (rc) return rc;
  }
  for (int i = 0; i < np; ++i) SW_REQUIRE(grads[i] && params[i], "tc_mlp_bwd: null gradient / parameter %d", i);

  BwdArgs b;
  b.d_raw = d_out; b.kind = kind; b.P = P; b.num_tiles = tiles;
  b.packed = reinterpret_cast<const uint8_t*>(packed); b.packed_t = reinterpret_cast<const uint8_t*>(packed_t);
  b.ws = ws; b.unfold = unfold; b.absmax = absmax; b.fixed_scale = grad_scale;
  for (int i = 0; i < 24; ++i) b.grads[i] = i < np ? grads[i] : nullptr;
  // job table of the default encoding: built once per process (host), uploaded once per DEVICE (c_jobs, read by the
  // layer-pipelined kernel) together with the shared-memory opt-ins; this call's table (kind, encoding) is a parameter
  static WgJobTable jobs_def[2], jobs_lw[2];
  static int wg_smem = 0;
  static std::once_flag jobs_once;
  std::call_once(jobs_once, [] {
    Enc D, Wd;
    decode_enc(ENC_DEFAULT_CODE, &D);
    decode_enc(20 | (20 << 8) | (20 << 16), &Wd);
    WgJobTable wide[2];
    for (int k = 0; k < 2; ++k) {
      build_jobs(jobs_def[k].j, k, D, true); build_jobs(jobs_lw[k].j, k, D, false); build_jobs(wide[k].j, k, Wd, true);
    }
    for (int k = 0; k < 2; ++k)
      for (int j = 0; j < WG_JOBS_ALL; ++j) {          // ring + ones block + alignment slack
        int need = jobs_def[k].j[j].stage_bytes * jobs_def[k].j[j].nstage + HALF_BLK + 1024;
        int need_w = wide[k].j[j].stage_bytes * wide[k].j[j].nstage + HALF_BLK + 1024;
        if (need > wg_smem) wg_smem = need;
        if (need_w > wg_smem) wg_smem = need_w;
      }
  });
  WgJobTable table_local;
  const WgJobTable* table = &jobs_def[kind];
  if (enc != ENC_DEFAULT_CODE) { build_jobs(table_local.j, kind, E, true); table = &table_local; }
  const int dpe_smem = 8 * DPE_CHUNK_B + 2 * ACT_BYTES + 1024;
  if (once_per_device(ONCE_BWD_BASE)) {
    cudaFuncSetAttribute(mlp_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMB_TOTAL);
    cudaMemcpyToSymbol(c_jobs, jobs_lw, sizeof(jobs_lw));
    cudaFuncSetAttribute(mlp_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem);
  }
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  if (g_prof) cudaEventRecord(g_ev[0], s);
  int rc = SWNERF_OK;
  // The layer-pipelined kernel (section 5) is OPT-IN (swnerf_tc_set_bwd_variant(1) / SWNERF_BWD_LW=1): it cuts the
  // backward's DRAM traffic 3x (ncu: 3.86 GB against 11.5 GB for 786 k samples, dy never leaves L2) but is still slower
  // than the two-kernel backward (3.6 against 2.1 ms): its roles are bound by the latency of their operand loads, not
  // by HBM (profiles/r2_lw_backward.md).  The two-kernel backward stays the default.
  const int lw_variant = g_bwd_variant.load();
  const int lw_min_ctas = 2 * LW_ROLES;
  if (kind == 0 && enc == ENC_DEFAULT_CODE && !d_pts && sm_count() >= lw_min_ctas && lw_variant == 1 && tiles > 2 * (int64_t)sm_count() &&
      tiles * WS_DY_BYTES >= LW_RING_BYTES) {
    const int lw_d_smem = 4 * CHUNK_B + LW_IN_BLKS * ACT_BLK + 1024;
    const int lw_smem = lw_d_smem > wg_smem ? lw_d_smem : wg_smem;
    if (once_per_device(ONCE_BWD_LW))
      cudaFuncSetAttribute(mlp_bwd_lw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lw_smem);
    LwArgs a;
    a.d_raw = d_out; a.P = P; a.num_tiles = tiles;
    a.packed = reinterpret_cast<const uint8_t*>(packed); a.packed_t = reinterpret_cast<const uint8_t*>(packed_t);
    a.ws = ws; a.ring = ws + tiles * (WS_TILE_BYTES + WS_MASK_BYTES);
    a.flags = reinterpret_cast<uint32_t*>(tail + LW_FLAGS_OFF);
    a.unfold = unfold; a.absmax = absmax; a.fixed_scale = grad_scale;
    for (int i = 0; i < 24; ++i) a.grads[i] = grads[i];
#ifdef SWNERF_LW_DEBUG
    a.dbg = g_lw_dbg;
#else
    a.dbg = nullptr;
#endif
    cudaMemsetAsync(tail + LW_FLAGS_OFF, 0, LW_FLAG_WORDS * sizeof(uint32_t), s);
    a.wrgb_slot = const_slot_acquire(CONST_FAMILY_BWD_WRGB, WRGB_SLOTS, s);
    SW_REQUIRE(a.wrgb_slot >= 0, "tc_mlp_bwd: no current device");
    if (cudaMemcpyToSymbolAsync(c_wrgb, reinterpret_cast<const uint8_t*>(packed) + PK_F32_OFF + F32_WRGB * sizeof(float),
                                384 * sizeof(float), (size_t)a.wrgb_slot * 384 * sizeof(float), cudaMemcpyDeviceToDevice,
                                s) != cudaSuccess)
      return set_err(SWNERF_ERR_CUDA, "tc_mlp_bwd: staging rgb_linear.weight failed");
    lw_assign_roles(sm_count(), a.role_first);
    mlp_bwd_lw_kernel<<<a.role_first[LW_ROLES], 384, lw_smem, s>>>(a);
    if (g_prof) { cudaEventRecord(g_ev[1], s); cudaEventRecord(g_ev[2], s); g_prof_valid = 1; }
    rc = check_launch("tc_mlp_bwd_lw");
    if (rc) return rc;
    unfold_head_kernel<<<UNFOLD_BLOCKS, 256, 0, s>>>(params[16], params[18], params[19], unfold, unfold + 128 * 256, grads[18],
                                            grads[19], grads[16], 256 + E.vc);
    return check_launch("tc_unfold_head");
  }
  static const int dg_variant = [] { const char* e = getenv("SWNERF_BWD_PAIR"); return e ? atoi(e) : -1; }();
  if (dg_variant == 1 || (dg_variant < 0 && tiles > 2 * (int64_t)sm_count())) {
    if (once_per_device(ONCE_BWD_DATA_PAIR))
      cudaFuncSetAttribute(mlp_bwd_data_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S5_TOTAL);
    CUtensorMap tm_wt;
    const unsigned long long dims[2] = {128, (unsigned long long)NT_CHUNKS * 256};
    const unsigned long long strides[1] = {128};
    const unsigned int box[2] = {128, 128};
    rc = encode_u8_tensor_map(&tm_wt, packed_t, 2, dims, strides, box);
    if (rc) return rc;
    const int64_t quads = (tiles + 3) / 4;
    const int grid_p = 2 * (int)(quads < sm_count() / 2 ? quads : sm_count() / 2);
    mlp_bwd_data_pair_kernel<<<grid_p, 384, S5_TOTAL, s>>>(b, tm_wt);
  } else {
    mlp_bwd_data_kernel<<<grid, 384, SMB_TOTAL, s>>>(b);
  }
  rc = check_launch("tc_mlp_bwd_data");
  if (rc) return rc;

  if (d_pts) {
    DpeArgs d;
    d.ws = ws; d.num_tiles = tiles; d.P = P; d.packed_t = reinterpret_cast<const uint8_t*>(packed_t);
    d.pts = pts; d.d_pts = d_pts; d.absmax = absmax; d.fixed_scale = grad_scale;
    switch (E.Lp) {
      case 0: rc = launch_input_grad<0, 0>(d, grid, dpe_smem, s); break;
      case 4: rc = launch_input_grad<4, 0>(d, grid, dpe_smem, s); break;
      case 10: rc = launch_input_grad<10, 0>(d, grid, dpe_smem, s); break;
      default:
        rc = launch_input_grad<20, 0>(d, grid, dpe_smem, s);
        if (!rc) rc = launch_input_grad<20, 1>(d, grid, dpe_smem, s);        // columns 64..122 of the encoding: accumulates
        break;
    }
    if (rc) return rc;
  }

  WgArgs w;
  w.ws = ws; w.num_tiles = tiles; w.unfold = unfold; w.absmax = absmax; w.fixed_scale = grad_scale; w.kind = kind;
  w.ws_ext = E.wide() ? tail + WS_TAIL_BYTES : nullptr;
  for (int i = 0; i < 24; ++i) w.grads[i] = i < np ? grads[i] : nullptr;
  const bool wg_pair = tiles > 2 * (int64_t)sm_count();      // small launches: every job on the one-CTA kernel
  // Every CTA ends with a flush of its accumulators (tens of thousands of red.add into the same gradient tensors the
  // job's other CTAs flush into), whatever its share of the tiles.  With ~3 us per tile and ~0.65 us of flush per CTA
  // the time is ~ 3 tiles jobs / n + 0.65 n, smallest at n ~ 2.1 sqrt(tiles jobs): a small launch (a MultiRes level of
  // 16 or 64 rays) takes only that many CTAs, at least one per job
  const int jobs_active = wg_pair ? 2 : WG_JOBS;
  int want = (int)(2.1 * sqrt((double)tiles * jobs_active));
  int n_cta = want < sm_count() ? want : sm_count();
  if (n_cta < WG_JOBS) n_cta = WG_JOBS;
  assign_ctas(n_cta, w.job_first_cta, kind, wg_pair);
  if (g_prof) cudaEventRecord(g_ev[1], s);
  if (wg_pair) {
    // trunk layers: CTA pairs
    if (once_per_device(ONCE_BWD_WEIGHT_PAIR))
      cudaFuncSetAttribute(mlp_bwd_weight_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WGP_SMEM);
    WgPairArgs pw;
    pw.num_tiles = tiles; pw.absmax = absmax; pw.fixed_scale = grad_scale; pw.kind = kind;
    for (int i = 0; i < 24; ++i) pw.grads[i] = i < np ? grads[i] : nullptr;
    WgPairMaps maps;
    static_assert(WS_TILE_BYTES == 36 * ACT_BLK && WS_DY_BYTES == 36 * ACT_BLK, "tile records are 36 blocks of 16 KB");
    const unsigned long long dims[3] = {128, 128, (unsigned long long)tiles * 36};
    const unsigned long long strides[2] = {128, (unsigned long long)ACT_BLK};
    const unsigned int box[3] = {128, 64, 2};
    rc = encode_u8_tensor_map(&maps.fwd, ws, 3, dims, strides, box);
    if (rc) return rc;
    rc = encode_u8_tensor_map(&maps.dy, ws + tiles * (WS_TILE_BYTES + WS_MASK_BYTES), 3, dims, strides, box);
    if (rc) return rc;
    const int n_pair_ctas = (sm_count() / 2) * 2;
    mlp_bwd_weight_pair_kernel<<<n_pair_ctas, 256, WGP_SMEM, s>>>(pw, maps, *table);
    rc = check_launch("tc_mlp_bwd_weight_pair");
    if (rc) return rc;
    if (g_prof) cudaEventRecord(g_ev[3], s);
  }
  mlp_bwd_weight_kernel<<<w.job_first_cta[WG_JOBS], 256, wg_smem, s>>>(w, *table);
  if (g_prof) { cudaEventRecord(g_ev[2], s); g_prof_valid = 1; }
  rc = check_launch("tc_mlp_bwd_weight");
  if (rc) return rc;

  if (kind == 0) {
    // un-fold the head: G = d W_fv, gb = d b_fv -> feature_linear / views_linears gradients (db_v was added above)
    unfold_head_kernel<<<UNFOLD_BLOCKS, 256, 0, s>>>(params[16], params[18], params[19], unfold, unfold + 128 * 256, grads[18],
                                            grads[19], grads[16], 256 + E.vc);
    rc = check_launch("tc_unfold_head");
  } else {
    // _time.0.weight[:, pc:pc+tw]: the time embedding is the same for every sample, so this block of the gradient
    // is the outer product (sum_s dy0[s]) x PE(t) = d b0 x PE(t); this call's d b0 sits in the scratch.
    time_outer_kernel<<<1, 256, 0, s>>>(unfold, tpe_dev, grads[0], E.pc + E.tw, E.pc, E.tw);
    rc = check_launch("tc_time_outer");
  }
  return rc;
}

extern "C" {

int swnerf_tc_mlp_bwd(const float* d_raw, int64_t n_rays, int n_samples, const void* packed, const void* packed_t,
                      int enc, const float* const* params, void* workspace, float* const* grads, float grad_scale,
                      void* stream) {
  return bwd_impl(d_raw, n_rays, n_samples, packed, packed_t, enc, params, workspace, grads, grad_scale, 0, nullptr,
                  nullptr, nullptr, stream);
}

int swnerf_tc_mlp_bwd_points(const float* d_raw, int64_t n_rays, int n_samples, const void* packed,
                             const void* packed_t, int enc, const float* const* params, void* workspace,
                             float* const* grads, float grad_scale, const float* pts, float* d_pts, void* stream) {
  return bwd_impl(d_raw, n_rays, n_samples, packed, packed_t, enc, params, workspace, grads, grad_scale, 0, nullptr, pts,
                  d_pts, stream);
}

int swnerf_tc_time_bwd(const float* d_dx, int64_t n_rays, int n_samples, const void* packed_time,
                       const void* packed_time_t, int enc, const float* const* params, const float* time_embedding_dev,
                       void* workspace, float* const* grads, float grad_scale, void* stream) {
  return bwd_impl(d_dx, n_rays, n_samples, packed_time, packed_time_t, enc, params, workspace, grads, grad_scale, 1,
                  time_embedding_dev, nullptr, nullptr, stream);
}

}  // extern "C"
