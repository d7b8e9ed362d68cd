// Fused points + positional encoding + 8x256 skip MLP on tcgen05 tensor cores (forward).
//
// Reference restated: nerf/run.py:385 (pts = o + d z), embedder.py:33-42 (PE L=10 / L=4),
// nerf/run.py:76-83 (expand viewdirs, concat), model.py:39-62 (vallina_NeRF.forward).
//
// One persistent CTA per SM walks 128-sample tiles.  Per tile the whole network runs on chip:
//   PE warps      : points + encodings in registers -> fp16 operand images in shared memory
//   producer warp : streams the packed fp16 weight chunks ([N x 64] K-major images, 32 KB) through a
//                   3-stage ring with cp.async.bulk + mbarrier (weights stay L2 resident, 1.05 MB)
//   MMA thread    : tcgen05.mma 128 x N x 16 (fp16 in, fp32 accumulate in TMEM), 4 per chunk
//   epilogue warps: tcgen05.ld -> +bias, ReLU -> fp16 -> back into the activation image IN PLACE, one
//                   64-column block at a time, so the next layer's MMAs start on block 0 while blocks 1..3
//                   are still being drained (two TMEM accumulators of 256 columns ping-pong)
// The skip connection is an extra K=64 chunk on the retained PE image (column order [pts | h],
// model.py:46); feature_linear is folded into views_linears at pack time (no nonlinearity between them,
// model.py:50-55) and alpha_linear rides along as row 128 of that N=144 head; rgb_linear (3 x 128) is
// evaluated in fp32 in the head epilogue.  Algorithmic work: 593,408 MAC per sample.
#include <cuda.h>
#include <atomic>
#include <mutex>
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"
#include "mlp_tc_layout.cuh"
#include "../../include/swnerf_b200.h"

namespace swnerf {
using namespace tc;
using namespace tcl;

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
struct ParamPtrs {
  const float* p[24];
  // kind 0: vallina_NeRF / NeRFOriginal (24 tensors, order of include/swnerf_b200.h)
  // kind 1: D-NeRF deformation net (model.py:113-136): p[2i],p[2i+1] = _time.i, p[16],p[17] = _time_out.
  //         It runs through the SAME kernels: its time embedding is constant per call, so W0[:, pc:] PE(t) is
  //         folded into the layer-0 bias, and its 256->3 output layer rides in rows 128..130 of the head
  //         (where alpha_linear sits for kind 0) with the view branch zeroed.
  int kind;
  Enc enc;
  float tpe[ENC_MAX_TW + 3];       // PE(t), enc.tw values used
};

// W_fv = W_v[:, :256] W_f (128 x 256), b_fv = W_v[:, :256] b_f + b_v      -> fold[128][257] fp32
// 64 blocks of two rows: the block keeps its rows of W_v in shared memory, thread c streams column c of W_f (coalesced
// across the block, L2-resident) and forms both rows' dot products in ascending j; the bias column is a block reduction.
constexpr int FOLD_ROWS = 2;
__global__ void __launch_bounds__(256) fold_head_kernel(ParamPtrs P, float* __restrict__ fold) {
  __shared__ float sv[FOLD_ROWS][256];
  __shared__ float red[FOLD_ROWS][8];
  const int c = threadIdx.x, r0 = blockIdx.x * FOLD_ROWS;
  if (P.kind != 0) {                                                     // deformation net: no view branch to fold
#pragma unroll
    for (int r = 0; r < FOLD_ROWS; ++r) {
      fold[(r0 + r) * 257 + c] = 0.f;
      if (c == 0) fold[(r0 + r) * 257 + 256] = 0.f;
    }
    return;
  }
  const int ldv = 256 + P.enc.vc;                                        // views_linears.0.weight is [128, 256 + vc]
#pragma unroll
  for (int r = 0; r < FOLD_ROWS; ++r) sv[r][c] = P.p[16][(size_t)(r0 + r) * ldv + c];       // W_v[r, j]
  __syncthreads();
  float acc[FOLD_ROWS];
#pragma unroll
  for (int r = 0; r < FOLD_ROWS; ++r) acc[r] = 0.f;
  const float* Wf = P.p[18];
  for (int j0 = 0; j0 < 256; j0 += 32) {                                 // 32 independent loads in flight per thread
    float w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = __ldg(Wf + (size_t)(j0 + i) * 256 + c);      // W_f[j, c]
#pragma unroll
    for (int i = 0; i < 32; ++i)
#pragma unroll
      for (int r = 0; r < FOLD_ROWS; ++r) acc[r] = fmaf(sv[r][j0 + i], w[i], acc[r]);
  }
  const float bf = P.p[19][c];
#pragma unroll
  for (int r = 0; r < FOLD_ROWS; ++r) {
    fold[(r0 + r) * 257 + c] = acc[r];
    float pb = sv[r][c] * bf;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pb += __shfl_xor_sync(0xffffffffu, pb, o);
    if ((c & 31) == 0) red[r][c >> 5] = pb;
  }
  __syncthreads();
  if (c < FOLD_ROWS) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[c][w];
    fold[(r0 + c) * 257 + 256] = sum + P.p[17][r0 + c];
  }
}

// element (n, k) of forward chunk c: n = output unit (row of the K-major B image), k = input column inside the chunk
__device__ __forceinline__ float fwd_weight(const ParamPtrs& P, const float* fold, int c, int n, int k) {
  const Enc& E = P.enc;
  const int pc = E.pc, vc = E.vc;
  const int ld0 = pc + (P.kind == 1 ? E.tw : 0);                         // _time.0 is [256, pc + tw]
  if (c < E.PC) { const int kk = c * 64 + k; return kk < pc ? P.p[0][(size_t)n * ld0 + kk] : 0.f; }
  c -= E.PC;
  if (c < 16) { const int l = 1 + c / 4, kc = c % 4; return P.p[2 * l][n * 256 + kc * 64 + k]; }
  c -= 16;
  if (c < E.PC) { const int kk = c * 64 + k; return kk < pc ? P.p[10][(size_t)n * (256 + pc) + kk] : 0.f; }
  c -= E.PC;
  if (c < 4) return P.p[10][(size_t)n * (256 + pc) + pc + c * 64 + k];
  c -= 4;
  if (c < 8) { const int l = 6 + c / 4, kc = c % 4; return P.p[2 * l][n * 256 + kc * 64 + k]; }
  c -= 8;
  // head: VC view chunks, then 4 chunks on h7
  if (c < E.VC) {
    const int kk = c * 64 + k;
    return (P.kind == 0 && n < 128 && kk < vc) ? P.p[16][(size_t)n * (256 + vc) + 256 + kk] : 0.f;
  }
  c -= E.VC;
  const int kk = c * 64 + k;
  if (P.kind == 1) return (n >= 128 && n < 131) ? P.p[16][(n - 128) * 256 + kk] : 0.f;
  if (n < 128) return fold[n * 257 + kk];
  if (n == 128) return P.p[20][kk];
  return 0.f;
}

__global__ void pack_fwd_kernel(ParamPtrs P, uint8_t* __restrict__ packed) {
  const float* fold = reinterpret_cast<const float*>(packed + PK_FOLD_OFF);
  const int n_full = P.enc.n_full(), n_all = P.enc.n_chunks();
  int unit = blockIdx.x * blockDim.x + threadIdx.x;
  if (unit < chunk_off_n(n_all, n_full) / 16) {
    int byte = unit * 16;
    int c, in;
    if (byte < n_full * CHUNK_B) { c = byte / CHUNK_B; in = byte % CHUNK_B; }
    else { int b2 = byte - n_full * CHUNK_B; c = n_full + b2 / HCHUNK_B; in = b2 % HCHUNK_B; }
    int n = (in >> 10) * 8 + ((in >> 7) & 7);
    int pu = (in >> 4) & 7;
    int k0 = (pu ^ (n & 7)) * 8;
    __align__(16) __half h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = __float2half_rn(fwd_weight(P, fold, c, n, k0 + i));
    *reinterpret_cast<uint4*>(packed + byte) = *reinterpret_cast<const uint4*>(h);
  }
  // fp32 block: biases, head bias, rgb_linear
  float* f = reinterpret_cast<float*>(packed + PK_F32_OFF);
  if (unit < F32_COUNT) {
    float v = 0.f;
    if (unit < F32_BHEAD) {
      int l = unit / 256, n = unit % 256;
      v = P.p[2 * l + 1][n];
      if (P.kind == 1 && l == 0) {
        const int pc = P.enc.pc, tw = P.enc.tw;
        for (int k = 0; k < tw; ++k) v = fmaf(P.p[0][(size_t)n * (pc + tw) + pc + k], P.tpe[k], v);     // + W0[:, pc:] PE(t)
      }
    } else if (P.kind == 1) {
      int n = unit - F32_BHEAD;
      v = (unit < F32_WRGB && n >= 128 && n < 131) ? P.p[17][n - 128] : 0.f;
    }
    else if (unit < F32_WRGB) { int n = unit - F32_BHEAD; v = n < 128 ? fold[n * 257 + 256] : (n == 128 ? P.p[21][0] : 0.f); }
    else if (unit < F32_BRGB) { v = P.p[22][unit - F32_WRGB]; }
    else { int i = unit - F32_BRGB; v = i < 3 ? P.p[23][i] : 0.f; }
    f[unit] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// fused forward
// ------------------------------------------------------------------------------------------------
struct FwdArgs {
  const float* rays; int ray_stride; int view_col;
  const float* z; int S; int64_t P;           // P = total sample rows
  const float* pts;                           // optional explicit sample positions [P,3] (D-NeRF: x + dx)
  const uint8_t* packed; float* raw;          // kind 0: raw[P,4];  kind 1: dx[P,3]
  uint8_t* ws; int64_t num_tiles;
  uint8_t* ws_ext;                            // two-chunk encodings: second blocks of the saved PE / view images, [tile][PE1 | VW1]
  int kind;
  int Lp, Lv, PC, VC;                         // encoding (mlp_tc_layout.cuh: Enc)
  int f32_slot;                               // CTA-pair kernel: which copy of the bias block in constant memory
#ifdef SWNERF_EXPERIMENTS
  int ko;                                     // knock-out experiments (profiles/r1_knockout_experiments.md): bench builds only
#endif
};
// The knock-out branches exist only in builds made with -DSWNERF_EXPERIMENTS; the shipped library compiles them out.
#ifdef SWNERF_EXPERIMENTS
#define SW_KO(g, bit) ((g).ko & (bit))
#else
#define SW_KO(g, bit) 0
#endif

// columns [64 CH, 64 CH + 64) of [x, sin(2^k x), cos(2^k x)]_{k < L} (embedder.py:33-42); columns past 3 (1 + 2 L) stay zero
template <int L, int CH>
__device__ __forceinline__ void encode3(const float (&x)[3], float (&f)[64]) {
  constexpr int lo = 64 * CH, hi = lo + 64;
#pragma unroll
  for (int i = 0; i < 64; ++i) f[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (CH == 0) f[j] = x[j];
    const float h1 = 0.15915494f, l1 = 6.4206e-9f;      // 1/(2 pi) = h1 + l1
    float th = x[j] * h1;
    float tl = fmaf(x[j], h1, -th) + x[j] * l1;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const int cs = 3 + 6 * k + j, cc = 6 + 6 * k + j;         // compile-time once unrolled
      const bool s_in = cs >= lo && cs < hi, c_in = cc >= lo && cc < hi;
      if (s_in || c_in) {
        float sn, cn;
        sincos_turns(th, tl, (float)(1 << k), sn, cn);
        if (s_in) f[s_in ? cs - lo : 0] = sn;
        if (c_in) f[c_in ? cc - lo : 0] = cn;
      }
    }
  }
}
// run-time frequency count (the encodings the reference instantiates: identity, 4, 10, 20), chunk ch of the image
__device__ __forceinline__ void encode3_any(int L, int ch, const float (&x)[3], float (&f)[64]) {
  switch (L) {
    case 0: encode3<0, 0>(x, f); break;
    case 4: encode3<4, 0>(x, f); break;
    case 10: encode3<10, 0>(x, f); break;
    default:
      if (ch == 0) encode3<20, 0>(x, f); else encode3<20, 1>(x, f);
      break;
  }
}

__device__ __forceinline__ void store_row64(uint8_t* img, int row, const float (&f)[64]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    uint4 q;
    q.x = pack_half2(f[8 * u + 0], f[8 * u + 1]);
    q.y = pack_half2(f[8 * u + 2], f[8 * u + 3]);
    q.z = pack_half2(f[8 * u + 4], f[8 * u + 5]);
    q.w = pack_half2(f[8 * u + 6], f[8 * u + 7]);
    *reinterpret_cast<uint4*>(img + tile_unit_off(row, u)) = q;
  }
}

// fp32 block of the packed weights (trunk biases, folded head bias, rgb_linear) for the CTA-pair kernel: staged per
// call with a device-to-device copy on the launching stream.  Every access is warp-uniform, so constant-cache reads
// cost the same as the shared-memory broadcast they replace, and the 10 KB of shared memory buy a 4th ring stage.
// There are F32_SLOTS copies: a stream keeps the slot it used last, another stream takes another slot, so the coarse
// and the fine network (or two callers) can run concurrently on different streams without sharing a bias block.
constexpr int F32_SLOTS = 4;
constexpr int F32_PAD = (F32_COUNT + 3) & ~3;
__constant__ float c_f32s[F32_SLOTS][F32_PAD];

// MODE 0: the default encoding (PE L = 10 / 4) at compile time; 1: one-chunk encodings chosen at run time (g.Lp, g.Lv);
// 2: two-chunk encodings (L = 20: 123 columns): 32-KB position / view images and a two-stage weight ring
template <int MODE> struct Fwd1Smem {
  static constexpr int EB = MODE == 2 ? 2 : 1;                  // 16-KB blocks per encoding image
  static constexpr int NS = NSTAGE;                             // weight ring depth
  static constexpr int ACT = 0;
  static constexpr int PE = ACT + ACT_BYTES;
  static constexpr int VW = PE + EB * ACT_BLK;
  static constexpr int RING = VW + EB * ACT_BLK;
  static constexpr int F32 = RING + NS * CHUNK_B;
  // MODE 2 has no room for the 10-KB fp32 block next to two 32-KB encoding images and a three-stage ring: it reads the
  // biases from constant memory (the CTA-pair kernel's slots) and checks the alignment of the base instead of padding it
  static constexpr int SCR = F32 + (MODE == 2 ? 0 : ((F32_COUNT * 4 + 127) / 128) * 128);
  static constexpr int BAR = SCR + TILE * 16;
  static constexpr int TOTAL = BAR + 256 + (MODE == 2 ? 0 : 1024);      // + alignment slack
};
static_assert(Fwd1Smem<0>::TOTAL == SM_TOTAL && Fwd1Smem<2>::TOTAL <= 232448, "shared memory budget");

template <bool TRAIN, int MODE>
__global__ void __launch_bounds__(512, 1) mlp_fwd_kernel(FwdArgs g) {
  using SL = Fwd1Smem<MODE>;
  constexpr int NS = SL::NS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = MODE == 2 ? smem_raw
                            : reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (MODE == 2 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* s_act = smem + SL::ACT;
  uint8_t* s_pe = smem + SL::PE;
  uint8_t* s_vw = smem + SL::VW;
  uint8_t* s_ring = smem + SL::RING;
  float* s_f32w = reinterpret_cast<float*>(smem + SL::F32);              // MODE 0 / 1: staged copy of the fp32 block
  const float* s_f32 = MODE == 2 ? c_f32s[g.f32_slot] : s_f32w;          // MODE 2: constant memory (warp-uniform reads)
  float4* s_scr = reinterpret_cast<float4*>(smem + SL::SCR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL::BAR);
  uint64_t* w_full = bars;            // [3]
  uint64_t* w_empty = bars + 3;       // [3]
  uint64_t* pe_full = bars + 6;
  uint64_t* pe_empty = bars + 7;
  uint64_t* vw_full = bars + 8;
  uint64_t* vw_empty = bars + 9;
  uint64_t* act_full = bars + 10;     // [4]
  uint64_t* d_full = bars + 14;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  uint64_t* st_done = bars + 17;      // [4] training: block j's bulk store has finished reading shared memory
  uint64_t* h9_full = bars + 21;      //     training: head epilogue has written h9 into blocks 0,1
  const int PC = MODE == 2 ? g.PC : 1, VC = MODE == 2 ? g.VC : 1;       // K-chunks of the position / view image

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(pe_full, 4); mbar_init(pe_empty, 1);       // one arrival per PE warp
    mbar_init(vw_full, 4); mbar_init(vw_empty, 1);
    for (int j = 0; j < 4; ++j) mbar_init(&act_full[j], 8);   // one arrival per epilogue warp
    mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
    for (int j = 0; j < 4; ++j) mbar_init(&st_done[j], 1);
    mbar_init(h9_full, 8);
    mbar_fence_init();
  }
  if (warp == 14) tmem_alloc<512>(tmem_slot);
  {
    const float* src = reinterpret_cast<const float*>(g.packed + PK_F32_OFF);
    if (MODE != 2)
      for (int i = threadIdx.x; i < F32_COUNT; i += blockDim.x) s_f32w[i] = __ldg(src + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Warp roles.  The issue arbiter of an SM sub-partition favours the highest warp id, so the two single-
  // thread roles (weight producer, MMA issuer) take warps 12/13: they win the issue slot whenever eligible.
  //   0-7 epilogue (TMEM lane quarter = warp & 3, column half = warp >> 2) | 8-11 PE | 12 producer | 13 MMA | 14 TMEM alloc
  if (warp == 12) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t cnt = 0;
      const int n_full = 28 + 2 * PC, n_all = 32 + 2 * PC + VC;
      for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        for (int c = 0; c < n_all; ++c, ++cnt) {
          uint32_t stage = cnt % NS, ph = (cnt / NS) & 1;
          mbar_wait(&w_empty[stage], ph ^ 1);
          uint32_t bytes = c < n_full ? CHUNK_B : HCHUNK_B;
          mbar_expect_tx(&w_full[stage], bytes);
          bulk_g2s(s_ring + stage * CHUNK_B, g.packed + chunk_off_n(c, n_full), bytes, &w_full[stage]);
        }
      }
    }
  } else if (warp == 13) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) control flow and one elected lane issues: descriptors and
    // addresses then live in uniform registers and the issue loop stays a handful of instructions per MMA
    // (issuing from inside an `if (lane == 0)` region makes ptxas serialise every operand through a waterfall loop).
    const uint32_t idesc256 = umma_idesc_f16(128, 256, 0, 0);
    const uint32_t idesc144 = umma_idesc_f16(128, HEAD_N, 0, 0);
    const uint32_t act_u32 = smem_u32(s_act), pe_u32 = smem_u32(s_pe), vw_u32 = smem_u32(s_vw), ring_u32 = smem_u32(s_ring);
    uint32_t cnt = 0, dcnt = 0, alayer = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      for (int li = 0; li < 9; ++li, ++dcnt) {
        const uint32_t d_tmem = tmem + (dcnt & 1) * 256;
        const uint32_t idesc = (li == 8) ? idesc144 : idesc256;
        const int nenc = (li == 0 || li == 5) ? PC : (li == 8 ? VC : 0);    // encoding chunks in front of the layer's input
        const int nch = (li == 0) ? PC : nenc + 4;
        for (int ci = 0; ci < nch; ++ci) {
          // which A operand feeds this chunk
          const int aj = ci - nenc;                          // activation block index, < 0 = encoding image chunk ci
          uint32_t a_base;
          if (aj < 0) {
            if (li == 8) { if (ci == 0) mbar_wait(vw_full, it & 1); a_base = vw_u32 + ci * ACT_BLK; }
            else { if (li == 0 && ci == 0) mbar_wait(pe_full, it & 1); a_base = pe_u32 + ci * ACT_BLK; }
          } else {
            mbar_wait(&act_full[aj], alayer & 1);
            a_base = act_u32 + aj * ACT_BLK;
          }
          const uint32_t stage = cnt % NS;
          mbar_wait(&w_full[stage], (cnt / NS) & 1);
          tc_fence_after();
          const uint32_t b_base = ring_u32 + stage * CHUNK_B;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16(d_tmem, umma_desc_kmajor(a_base + ks * 32), umma_desc_kmajor(b_base + ks * 32), idesc,
                       (ci > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&w_empty[stage]);
            if (aj == -1 && li == 5) umma_commit(pe_empty);      // the image's last chunk has been read
            if (aj == -1 && li == 8) umma_commit(vw_empty);
            if (ci == nch - 1) umma_commit(&d_full[dcnt & 1]);
          }
          __syncwarp();
          ++cnt;
        }
        if (li >= 1) ++alayer;
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue: TMEM -> bias/ReLU -> fp16 activation image =====================
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t dcnt = 0, it = 0;
    uint32_t wcnt[4] = {0u, 0u, 0u, 0u};     // training: writes issued so far to activation block j
    for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      uint32_t* ws_mask = TRAIN ? reinterpret_cast<uint32_t*>(g.ws + g.num_tiles * WS_TILE_BYTES) + tile * (9 * 8 * 128)
                                : nullptr;
      for (int li = 0; li < 9; ++li, ++dcnt) {
        const uint32_t dcol = (dcnt & 1) * 256;
        mbar_wait(&d_full[dcnt & 1], (dcnt >> 1) & 1);
        tc_fence_after();
        if (li < 8) {
          const float* bias = s_f32 + li * 256;
          // software pipeline over the four 64-column blocks: the TMEM load of block j+1 is in flight while
          // block j is converted and handed to the MMA thread
          uint32_t va[32], vb[32];
          tmem_ld32(tmem + lane_addr + dcol + hh * 32, va);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t (&v)[32] = (j & 1) ? vb : va;
            uint32_t (&vn)[32] = (j & 1) ? va : vb;
            const int c0 = j * 64 + hh * 32;
            tmem_ld_wait_on(v);
            if (j < 3) tmem_ld32(tmem + lane_addr + dcol + c0 + 64, vn);
            uint32_t pk[16];
            uint32_t mask = 0;
            const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 bb = b4[i4];                         // broadcast LDS.128
              float a0 = __uint_as_float(v[4 * i4]) + bb.x, a1 = __uint_as_float(v[4 * i4 + 1]) + bb.y;
              float a2 = __uint_as_float(v[4 * i4 + 2]) + bb.z, a3 = __uint_as_float(v[4 * i4 + 3]) + bb.w;
              if (TRAIN)
                mask |= (a0 > 0.f ? 1u : 0u) << (4 * i4) | (a1 > 0.f ? 1u : 0u) << (4 * i4 + 1) |
                        (a2 > 0.f ? 1u : 0u) << (4 * i4 + 2) | (a3 > 0.f ? 1u : 0u) << (4 * i4 + 3);
              pk[2 * i4] = pack_half2_relu(a0, a1);
              pk[2 * i4 + 1] = pack_half2_relu(a2, a3);
            }
            uint8_t* blk = s_act + j * ACT_BLK;
            if (TRAIN) {   // the store warp may still be reading the previous contents of this block
              if (wcnt[j] > 0) mbar_wait(&st_done[j], (wcnt[j] - 1) & 1);
              ++wcnt[j];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *reinterpret_cast<uint4*>(blk + tile_unit_off(row, hh * 4 + u)) =
                  make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            if (TRAIN) ws_mask[(li * 8 + j * 2 + hh) * 128 + row] = mask;
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_full[j]);      // 256 per-thread arrivals would serialise on one smem word
          }
        } else {
          // head: cols 0..127 = relu -> h9, col 128 = sigma; rgb = W_rgb h9 + b_rgb in fp32
          const float* bh = s_f32 + F32_BHEAD;
          const float* wr = s_f32 + F32_WRGB;
          float pr = 0.f, pg = 0.f, pb = 0.f;
          if (TRAIN) {   // h9 goes into blocks 0 / 1 (this thread writes block hh)
            if (wcnt[hh] > 0) mbar_wait(&st_done[hh], (wcnt[hh] - 1) & 1);
            ++wcnt[0]; ++wcnt[1];
          }
#pragma unroll 1
          for (int jj = 0; jj < 2; ++jj) {
            uint32_t v[32];
            const int c0 = hh * 64 + jj * 32;
            tmem_ld32(tmem + lane_addr + dcol + c0, v);
            tmem_ld_wait();
            uint32_t pk[16];
            uint32_t mask = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float a = fmaxf(__uint_as_float(v[2 * i]) + bh[c0 + 2 * i], 0.f);
              float b = fmaxf(__uint_as_float(v[2 * i + 1]) + bh[c0 + 2 * i + 1], 0.f);
              pr = fmaf(wr[c0 + 2 * i], a, pr); pr = fmaf(wr[c0 + 2 * i + 1], b, pr);
              pg = fmaf(wr[128 + c0 + 2 * i], a, pg); pg = fmaf(wr[128 + c0 + 2 * i + 1], b, pg);
              pb = fmaf(wr[256 + c0 + 2 * i], a, pb); pb = fmaf(wr[256 + c0 + 2 * i + 1], b, pb);
              if (TRAIN) {
                mask |= (a > 0.f ? 1u : 0u) << (2 * i) | (b > 0.f ? 1u : 0u) << (2 * i + 1);
                pk[i] = pack_half2(a, b);
              }
            }
            if (TRAIN) {
              uint8_t* blk = s_act + hh * ACT_BLK;
#pragma unroll
              for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(blk + tile_unit_off(row, jj * 4 + u)) =
                    make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
              ws_mask[(8 * 8 + hh * 2 + jj) * 128 + row] = mask;
            }
          }
          if (hh == 1) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_addr + dcol + 128, v);
            tmem_ld_wait();
            if (g.kind == 1) {       // deformation net: dx = rows 128..130 of the head (model.py:136)
              int64_t idx = tile * TILE + row;
              if (idx < g.P) {
                g.raw[idx * 3 + 0] = __uint_as_float(v[0]) + bh[128];
                g.raw[idx * 3 + 1] = __uint_as_float(v[1]) + bh[129];
                g.raw[idx * 3 + 2] = __uint_as_float(v[2]) + bh[130];
              }
            }
            s_scr[row] = make_float4(pr, pg, pb, __uint_as_float(v[0]) + bh[128]);
          }
          if (TRAIN) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(h9_full);
          }
          tc_fence_before();
          named_bar_sync(1, 256);
          if (hh == 0 && g.kind == 0) {
            float4 o = s_scr[row];
            const float* br = s_f32 + F32_BRGB;
            int64_t idx = tile * TILE + row;
            if (idx < g.P)
              reinterpret_cast<float4*>(g.raw)[idx] = make_float4(pr + o.x + br[0], pg + o.y + br[1], pb + o.z + br[2], o.w);
          }
        }
      }
    }
  } else if (warp < 12) {
    // ===================== points + positional encoding =====================
    const int p = (warp - 8) * 32 + lane;
    const bool p0 = (p == 0);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      int64_t idx = tile * TILE + p;
      bool valid = idx < g.P;
      float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
      if (valid) {
        int64_t r = idx / g.S;
        const float* ray = g.rays + r * g.ray_stride;
        float zz = g.pts ? 0.f : __ldg(g.z + idx);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          pos[j] = g.pts ? __ldg(g.pts + idx * 3 + j)
                         : __fadd_rn(__ldg(ray + j), __fmul_rn(__ldg(ray + 3 + j), zz));      // run.py:385
          dir[j] = __ldg(ray + g.view_col + j);
        }
      }
      float f[64];
      for (int ch = 0; ch < PC; ++ch) {
        if (MODE == 0) encode3<10, 0>(pos, f); else encode3_any(g.Lp, ch, pos, f);
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 64; ++i) f[i] = 0.f;
        }
        if (ch == 0) {             // (after the first chunk's arithmetic: it overlaps the wait)
          if (it > 0) mbar_wait(pe_empty, (it - 1) & 1);
          if (TRAIN && it > 0) {
            if (p0) bulk_wait_read0();
            named_bar_sync(2, 128);
          }
        }
        store_row64(s_pe + ch * ACT_BLK, p, f);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pe_full);
      for (int ch = 0; ch < VC; ++ch) {
        if (MODE == 0) encode3<4, 0>(dir, f); else encode3_any(g.Lv, ch, dir, f);
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 64; ++i) f[i] = 0.f;
        }
        if (ch == 0 && it > 0) mbar_wait(vw_empty, (it - 1) & 1);
        store_row64(s_vw + ch * ACT_BLK, p, f);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(vw_full);
      if (TRAIN) {
        named_bar_sync(2, 128);
        if (p0) {
          uint8_t* ws_tile = g.ws + tile * WS_TILE_BYTES;
          bulk_s2g(ws_tile + WS_PE_OFF, s_pe, ACT_BLK);
          bulk_s2g(ws_tile + WS_VW_OFF, s_vw, ACT_BLK);
          if (MODE == 2) {       // second blocks of the images (zero when the encoding has one chunk)
            uint8_t* ext = g.ws_ext + tile * (2 * (int64_t)ACT_BLK);
            if (PC > 1) bulk_s2g(ext, s_pe + ACT_BLK, ACT_BLK);
            if (VC > 1) bulk_s2g(ext + ACT_BLK, s_vw + ACT_BLK, ACT_BLK);
          }
          bulk_commit();
        }
      }
    }
    if (TRAIN && p0) bulk_wait_all0();
  }

  else if (TRAIN && warp == 15) {
    // ===================== activation store warp (training) =====================
    // Saves every activation block for the backward as soon as its eight epilogue warps have written it
    // (same act_full barrier the MMA warp waits on), so no block-wide barrier sits in the epilogue.
    if (lane == 0) {
      uint32_t pl = 0, it = 0;
      for (int64_t tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
        uint8_t* ws_tile = g.ws + tile * WS_TILE_BYTES;
        for (int li = 0; li < 8; ++li, ++pl) {
          for (int j = 0; j < 4; ++j) {
            mbar_wait(&act_full[j], pl & 1);
            bulk_s2g(ws_tile + WS_H_OFF + li * ACT_BYTES + j * ACT_BLK, s_act + j * ACT_BLK, ACT_BLK);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(&st_done[j]);
          }
        }
        mbar_wait(h9_full, it & 1);
        bulk_s2g(ws_tile + WS_H9_OFF, s_act, 2 * ACT_BLK);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(&st_done[0]);
        mbar_arrive(&st_done[1]);
      }
      bulk_wait_all0();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 14) tmem_dealloc<512>(tmem);
}


// ======================================================================================================
// Forward kernel on CTA pairs (cta_group::2), two tile slots per CTA.
//
// Two CTAs of a cluster (two SMs of one TPC) each own TWO 128-sample tiles (slots X, Y).  The leader CTA issues
// M=256 MMAs that take 128 rows of A from each CTA's shared memory and HALF of the weight rows from each CTA's ring,
// so every SM streams half the weight bytes and its tensor pipe runs at the full 128x256x16-per-128-cycles rate (one
// CTA feeding both operands from its own shared memory tops out at 161 cycles, tools/probe_mma.py).  Layers of the
// two slots alternate (X.l, Y.l, X.l+1, ...): while the pipe runs one slot's layer, the epilogue warps drain the
// other slot's accumulator and write its next input in place, so the ~1000-cycle hand-off chain (commit -> barrier
// -> tcgen05.ld -> bias/ReLU/fp16 -> st.shared -> fence -> barrier -> issue) is off the tensor pipe's critical path.
// Barriers that gate MMA issue live in the leader CTA and collect arrivals from both CTAs (cluster-scope arrive);
// completion (tcgen05.commit) is multicast to the same barrier offset in both CTAs.
// Shared memory per CTA: 2 x 64 KB activation images, 2 x 16 KB encoding images (PE until the skip layer has read
// it, then the view encoding for the head), 3 x 16 KB weight ring.
// DEFENC: the default encoding (PE L = 10 / 4) at compile time; otherwise a one-chunk encoding chosen at run time.
template <bool TRAIN, bool DEFENC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) mlp_fwd4_kernel(const __grid_constant__ FwdArgs g, const __grid_constant__ CUtensorMap tm_trunk,
                                                                                   const __grid_constant__ CUtensorMap tm_head) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;                // no alignment slack (shared memory is full): checked instead
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* s_act = smem + S4_ACT;          // [2 slots][4 blocks]
  uint8_t* s_enc = smem + S4_ENC;          // [2 slots]
  uint8_t* s_ring = smem + S4_RING;
  float4* s_scr = reinterpret_cast<float4*>(smem + S4_SCR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S4_BAR);
  uint64_t* w_full = bars;                 // [NST4] leader: both CTAs' halves of the stage have landed (tensor-map loads)
  uint64_t* w_empty = bars + 8;            // [NST4] multicast: the MMAs reading the stage are complete
  uint64_t* e_full = bars + 12;            // [2] leader, 8 arrivals: encoding image of slot t written (PE, then views)
  uint64_t* e_free = bars + 14;            // [2] multicast: its readers are complete (after the skip layer, after the head)
  uint64_t* act_full = bars + 16;          // [2] leader, 16 arrivals: slot t's layer output written in both CTAs
  uint64_t* head_done = bars + 18;         // [2] leader, 16 arrivals: slot t's head accumulator has been read
  uint64_t* d_full = bars + 20;            // [2] multicast: slot t's accumulator complete
  uint64_t* img_ready = bars + 22;         // [2] local, training: slot t's image written (store warp)
  uint64_t* st_done = bars + 24;           // [2] local, training: the image's bulk store has read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const float* c_f32 = c_f32s[g.f32_slot];
  const int64_t num_quads = (g.num_tiles + 3) >> 2;          // a pair iteration covers 4 tiles: tile = 4 q + 2 slot + rank
  const int64_t quad0 = blockIdx.x >> 1, quad_step = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST4; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&e_full[t], 8); mbar_init(&e_free[t], 1);
      mbar_init(&head_done[t], 16); mbar_init(&d_full[t], 1);
      mbar_init(&act_full[t], 16); mbar_init(&img_ready[t], 8); mbar_init(&st_done[t], 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  cluster_sync_all();                      // both CTAs' barriers are initialised before anyone signals the other
  if (warp == 14) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // cluster addresses (in the leader) of the barriers both CTAs signal
  const uint32_t e_full_l = mapa_u32(smem_u32(e_full), 0), act_full_l = mapa_u32(smem_u32(act_full), 0);
  const uint32_t head_done_l = mapa_u32(smem_u32(head_done), 0);

  //   0-7 epilogue (TMEM lane quarter = warp & 3, column half = warp >> 2) | 8-11 PE | 12 producer | 13 MMA (leader) |
  //   14 TMEM alloc | 15 store (training)
  if (warp == 12) {
    // ===================== weight producer: this CTA's half of every chunk, once per slot =====================
    // Tensor-map loads with .cta_group::2 report their bytes to the LEADER's barrier, which expects both halves: no
    // relay hop between the peer's copy landing and the leader issuing.
    if (lane == 0 && !SW_KO(g, 32)) {
      const uint32_t w_full_l = mapa_u32(smem_u32(w_full), 0);
      uint32_t cnt = 0;
      for (int64_t quad = quad0; quad < num_quads; quad += quad_step) {
        int row = 0;                                   // row of the packed weight image (128 B per row)
        for (int li = 0; li < 9; ++li) {
          const int nch = (li == 0) ? 1 : ((li == 5 || li == 8) ? 5 : 4);
          const int rows = (li == 8) ? HEAD_N : 256;
          const uint32_t bytes = (uint32_t)rows * 64u;                             // this CTA's half: rows / 2 x 128 B
          for (int t = 0; t < 2; ++t)
            for (int ci = 0; ci < nch; ++ci, ++cnt) {
              const uint32_t stage = cnt % NST4, ph = (cnt / NST4) & 1;
              mbar_wait(&w_empty[stage], ph ^ 1);
              if (leader) mbar_expect_tx(&w_full[stage], 2u * bytes);
              tma_load_2d_pair(s_ring + stage * STG4_B, li == 8 ? &tm_head : &tm_trunk, 0,
                               row + ci * rows + (int)rank * (rows / 2), w_full_l + stage * 8);
            }
          row += nch * rows;
        }
      }
    }
  } else if (warp == 13 && leader) {
    // ===================== MMA issuer (leader CTA; converged warp, one elected lane issues) =====================
    const uint32_t idesc256 = umma_idesc_f16(256, 256, 0, 0);
    const uint32_t idesc144 = umma_idesc_f16(256, HEAD_N, 0, 0);
    const uint32_t act_u32 = smem_u32(s_act), enc_u32 = smem_u32(s_enc), ring_u32 = smem_u32(s_ring);
    uint32_t cnt = 0, it = 0;
    for (int64_t quad = quad0; quad < num_quads; quad += quad_step, ++it) {
      for (int li = 0; li < 9; ++li) {
        const bool head = (li == 8);
        const int nch = (li == 0) ? 1 : ((li == 5 || head) ? 5 : 4);
        const uint32_t idesc = head ? idesc144 : idesc256;
        for (int t = 0; t < 2; ++t) {
          const uint32_t d_tmem = tmem + t * 256;
          // the slot's accumulator must have been drained and (li >= 1) its input written: the epilogue of the slot's
          // previous layer ran while the pipe worked on the other slot
          if (li == 0) {
            if (it > 0) mbar_wait_cluster(&head_done[t], (it - 1) & 1);
            mbar_wait_cluster(&e_full[t], 0);                 // PE image (phase 2 it)
          } else {
            mbar_wait_cluster(&act_full[t], (li - 1) & 1);
            if (head) mbar_wait_cluster(&e_full[t], 1);       // view encoding (phase 2 it + 1)
          }
          for (int ci = 0; ci < nch; ++ci) {
            int aj = (li == 5 || head) ? ci - 1 : ci;       // activation block index, -1 = encoding image
            if (li == 0) aj = -1;
            const uint32_t a_base = aj < 0 ? enc_u32 + t * ACT_BLK : act_u32 + t * ACT_BYTES + aj * ACT_BLK;
            const uint32_t stage = cnt % NST4, wph = (cnt / NST4) & 1;
            if (!SW_KO(g, 32)) mbar_wait(&w_full[stage], wph);
            tc_fence_after();
            const uint32_t b_base = ring_u32 + stage * STG4_B;
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_f16_pair(d_tmem, umma_desc_kmajor(a_base + ks * 32), umma_desc_kmajor(b_base + ks * 32), idesc,
                              (ci > 0 || ks > 0) ? 1u : 0u);
              umma_commit_pair(&w_empty[stage], 3);
              if (aj < 0 && li >= 5) umma_commit_pair(&e_free[t], 3);
              if (ci == nch - 1) umma_commit_pair(&d_full[t], 3);
            }
            __syncwarp();
            ++cnt;
          }
        }
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue: accumulator -> bias/ReLU -> fp16 -> the slot's activation image (in place) =====================
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t dcnt = 0;                       // d_full[t] completes once per layer: both slots advance together
    uint32_t wcnt = 0;                       // training: writes issued so far to every block of a slot (same for all)
    for (int64_t quad = quad0; quad < num_quads; quad += quad_step) {
      for (int li = 0; li < 9; ++li, ++dcnt) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int64_t tile = quad * 4 + t * 2 + rank;
          const bool tvalid = tile < g.num_tiles;
          uint32_t* ws_mask = (TRAIN && tvalid)
              ? reinterpret_cast<uint32_t*>(g.ws + g.num_tiles * WS_TILE_BYTES) + tile * (9 * 8 * 128) : nullptr;
          const uint32_t acc = tmem + lane_addr + t * 256;
          uint8_t* img = s_act + t * ACT_BYTES;
          mbar_wait(&d_full[t], dcnt & 1);
          tc_fence_after();
          if (li < 8 && SW_KO(g, 16)) {     // experiment: synchronisation only
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(act_full_l + t * 8);
          } else if (li < 8) {
            const float* bias = c_f32 + li * 256 + hh * 32;
            // software pipeline over the four 64-column blocks: the TMEM load of block j+1 is in flight while
            // block j is converted and handed to the MMA thread
            if (TRAIN && wcnt > 0) mbar_wait(&st_done[t], (wcnt - 1) & 1);   // the image's bulk store still reads it
            uint32_t va[32], vb[32];
            uint32_t masks[4];
            tmem_ld32(acc + hh * 32, va);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t (&v)[32] = (j & 1) ? vb : va;
              uint32_t (&vn)[32] = (j & 1) ? va : vb;
              tmem_ld_wait_on(v);
              if (j < 3) tmem_ld32(acc + (j + 1) * 64 + hh * 32, vn);
              uint32_t pk[16];
              uint32_t mask = 0;
              const float4* b4 = reinterpret_cast<const float4*>(bias + j * 64);
              if (TRAIN) {
                // sign masks at two instructions per element: n = -(acc + bias) (negations ride on the FADD), its sign
                // bit is funnel-shifted into the mask, the fp16 pack takes -n.  Elements go last to first so bit k
                // ends up being column k.  (An exact +0 pre-activation counts as active: its activation is 0 either way.)
#pragma unroll
                for (int i4 = 7; i4 >= 0; --i4) {
                  const float4 bb = b4[i4];                       // warp-uniform constant-cache read
                  const float n0 = -__uint_as_float(v[4 * i4]) - bb.x, n1 = -__uint_as_float(v[4 * i4 + 1]) - bb.y;
                  const float n2 = -__uint_as_float(v[4 * i4 + 2]) - bb.z, n3 = -__uint_as_float(v[4 * i4 + 3]) - bb.w;
                  mask = __funnelshift_l(__float_as_uint(n3), mask, 1);
                  mask = __funnelshift_l(__float_as_uint(n2), mask, 1);
                  mask = __funnelshift_l(__float_as_uint(n1), mask, 1);
                  mask = __funnelshift_l(__float_as_uint(n0), mask, 1);
                  pk[2 * i4] = pack_half2_relu(-n0, -n1);
                  pk[2 * i4 + 1] = pack_half2_relu(-n2, -n3);
                }
              } else {
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                  const float4 bb = b4[i4];                       // warp-uniform constant-cache read
                  float a0 = __uint_as_float(v[4 * i4]) + bb.x, a1 = __uint_as_float(v[4 * i4 + 1]) + bb.y;
                  float a2 = __uint_as_float(v[4 * i4 + 2]) + bb.z, a3 = __uint_as_float(v[4 * i4 + 3]) + bb.w;
                  pk[2 * i4] = pack_half2_relu(a0, a1);
                  pk[2 * i4 + 1] = pack_half2_relu(a2, a3);
                }
              }
              uint8_t* blk = img + j * ACT_BLK;
#pragma unroll
              for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(blk + tile_unit_off(row, hh * 4 + u)) =
                    make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
              masks[j] = mask;
            }
            // one proxy fence and one arrival for the whole layer output: the next layer of this slot is issued after
            // the other slot's layer anyway, and the fence is the expensive part of the hand-off
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive_remote(act_full_l + t * 8);
              if (TRAIN) mbar_arrive(&img_ready[t]);
            }
            // the sign masks go out AFTER the hand-off: in front of it the proxy fence and the release-arrive would
            // wait for these global stores (ncu: 15 % of the kernel's stall samples sat on that ERRBAR / arrive)
            if (TRAIN && tvalid && !SW_KO(g, 2)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) ws_mask[(li * 8 + j * 2 + hh) * 128 + row] = masks[j];
            }
          } else {
            // head: cols 0..127 = relu -> h9, col 128 = sigma; rgb = W_rgb h9 + b_rgb in fp32
            const float* bh = c_f32 + F32_BHEAD;
            const float* wr = c_f32 + F32_WRGB;
            float pr = 0.f, pg = 0.f, pb = 0.f;
            // training: h9 goes into blocks 0 / 1 of the slot (this thread writes block hh); blocks 2,3 are not
            // rewritten by the head, so their store counters advance without a store (see the store warp)
            if (TRAIN && wcnt > 0) mbar_wait(&st_done[t], (wcnt - 1) & 1);
            uint32_t vs[32];
            if (hh == 1) tmem_ld32(acc + 128, vs);
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
              uint32_t v[32];
              const int c0 = hh * 64 + jj * 32;
              tmem_ld32(acc + c0, v);
              tmem_ld_wait();
              uint32_t hk[16];
              uint32_t mask = 0;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float a = fmaxf(__uint_as_float(v[2 * i]) + bh[c0 + 2 * i], 0.f);
                float b = fmaxf(__uint_as_float(v[2 * i + 1]) + bh[c0 + 2 * i + 1], 0.f);
                pr = fmaf(wr[c0 + 2 * i], a, pr); pr = fmaf(wr[c0 + 2 * i + 1], b, pr);
                pg = fmaf(wr[128 + c0 + 2 * i], a, pg); pg = fmaf(wr[128 + c0 + 2 * i + 1], b, pg);
                pb = fmaf(wr[256 + c0 + 2 * i], a, pb); pb = fmaf(wr[256 + c0 + 2 * i + 1], b, pb);
                if (TRAIN) {
                  mask |= (a > 0.f ? 1u : 0u) << (2 * i) | (b > 0.f ? 1u : 0u) << (2 * i + 1);
                  hk[i] = pack_half2(a, b);
                }
              }
              if (TRAIN) {
                uint8_t* blk = img + hh * ACT_BLK;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  *reinterpret_cast<uint4*>(blk + tile_unit_off(row, jj * 4 + u)) =
                      make_uint4(hk[4 * u], hk[4 * u + 1], hk[4 * u + 2], hk[4 * u + 3]);
                if (tvalid) ws_mask[(8 * 8 + hh * 2 + jj) * 128 + row] = mask;
              }
            }
            // every accumulator read of this slot's tile is complete (the first tcgen05.wait::ld covered `vs`)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(head_done_l + t * 8);
            const int64_t idx = tile * TILE + row;
            if (hh == 1) {
              if (g.kind == 1 && idx < g.P) {       // deformation net: dx = rows 128..130 of the head (model.py:136)
                g.raw[idx * 3 + 0] = __uint_as_float(vs[0]) + bh[128];
                g.raw[idx * 3 + 1] = __uint_as_float(vs[1]) + bh[129];
                g.raw[idx * 3 + 2] = __uint_as_float(vs[2]) + bh[130];
              }
              s_scr[row] = make_float4(pr, pg, pb, __uint_as_float(vs[0]) + bh[128]);
            }
            if (TRAIN) {
              fence_async_smem();
              __syncwarp();
              if (lane == 0) mbar_arrive(&img_ready[t]);
            }
            named_bar_sync(1, 256);
            if (hh == 0 && g.kind == 0) {
              float4 o = s_scr[row];
              const float* br = c_f32 + F32_BRGB;
              if (idx < g.P)
                reinterpret_cast<float4*>(g.raw)[idx] = make_float4(pr + o.x + br[0], pg + o.y + br[1], pb + o.z + br[2], o.w);
            }
            named_bar_sync(1, 256);       // s_scr is rewritten by the other slot's head right away
          }
        }
        if (TRAIN) ++wcnt;                 // both slots' images have had one more write this layer
      }
    }
  } else if (warp < 12) {
    // ===================== points + positional encoding, both slots =====================
    const int p = (warp - 8) * 32 + lane;
    const bool p0 = (p == 0);
    uint32_t it = 0;
    for (int64_t quad = quad0; quad < num_quads; quad += quad_step, ++it) {
      float dirs[2][3];
      // PE images first (both slots need them at layer 0), the view encodings after the skip layer released the images
      for (int t = 0; t < 2; ++t) {
        const int64_t tile = quad * 4 + t * 2 + rank;
        int64_t idx = tile * TILE + p;
        bool valid = idx < g.P;
        float pos[3] = {0.f, 0.f, 0.f};
        dirs[t][0] = dirs[t][1] = dirs[t][2] = 0.f;
        if (valid) {
          int64_t r = idx / g.S;
          const float* ray = g.rays + r * g.ray_stride;
          float zz = g.pts ? 0.f : __ldg(g.z + idx);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            pos[j] = g.pts ? __ldg(g.pts + idx * 3 + j)
                           : __fadd_rn(__ldg(ray + j), __fmul_rn(__ldg(ray + 3 + j), zz));      // run.py:385
            dirs[t][j] = __ldg(ray + g.view_col + j);
          }
        }
        float f[64];
        if (DEFENC) encode3<10, 0>(pos, f); else encode3_any(g.Lp, 0, pos, f);
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 64; ++i) f[i] = 0.f;
        }
        if (it > 0) mbar_wait(&e_free[t], 1);          // the previous tile's head has read the view encoding
        if (TRAIN && (it > 0 || t > 0)) {
          if (p0) bulk_wait_read0();
          named_bar_sync(2, 128);
        }
        store_row64(s_enc + t * ACT_BLK, p, f);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(e_full_l + t * 8);
        if (TRAIN) {
          named_bar_sync(2, 128);
          if (p0) {
            if (tile < g.num_tiles) bulk_s2g(g.ws + tile * WS_TILE_BYTES + WS_PE_OFF, s_enc + t * ACT_BLK, ACT_BLK);
            bulk_commit();
          }
        }
      }
      for (int t = 0; t < 2; ++t) {
        const int64_t tile = quad * 4 + t * 2 + rank;
        const bool valid = tile * TILE + p < g.P;
        float f[64];
        if (DEFENC) encode3<4, 0>(dirs[t], f); else encode3_any(g.Lv, 0, dirs[t], f);
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 64; ++i) f[i] = 0.f;
        }
        mbar_wait(&e_free[t], 0);                      // the skip layer has read the PE image
        if (TRAIN) {
          if (p0) bulk_wait_read0();
          named_bar_sync(2, 128);
        }
        store_row64(s_enc + t * ACT_BLK, p, f);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(e_full_l + t * 8);
        if (TRAIN) {
          named_bar_sync(2, 128);
          if (p0) {
            if (tile < g.num_tiles) bulk_s2g(g.ws + tile * WS_TILE_BYTES + WS_VW_OFF, s_enc + t * ACT_BLK, ACT_BLK);
            bulk_commit();
          }
        }
      }
    }
    if (TRAIN && p0) bulk_wait_all0();
  } else if (TRAIN && warp == 15) {
    // ===================== activation store warp (training) =====================
    if (lane == 0) {
      uint32_t rc = 0;
      for (int64_t quad = quad0; quad < num_quads; quad += quad_step) {
        for (int li = 0; li < 9; ++li, ++rc) {
          for (int t = 0; t < 2; ++t) {
            const int64_t tile = quad * 4 + t * 2 + rank;
            uint8_t* ws_tile = g.ws + tile * WS_TILE_BYTES;
            uint8_t* img = s_act + t * ACT_BYTES;
            mbar_wait(&img_ready[t], rc & 1);
            // one 16-KB bulk store at a time (0.959 -> 0.946 ms against one 64-KB store; 8-KB pieces: no further gain):
            // the TMA unit also carries the weight ring's loads
            const int nblk = (li < 8) ? 4 : 2;                                 // h9 sits in blocks 0,1
            uint8_t* dst = ws_tile + (li < 8 ? WS_H_OFF + li * ACT_BYTES : WS_H9_OFF);
            for (int j = 0; j < nblk; ++j) {
              if (tile < g.num_tiles && !SW_KO(g, 8)) bulk_s2g(dst + j * ACT_BLK, img + j * ACT_BLK, ACT_BLK);
              bulk_commit();
              bulk_wait_read0();
            }
            mbar_arrive(&st_done[t]);
          }
        }
      }
      bulk_wait_all0();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer may still be using this CTA's barriers / the pair's tensor memory
  if (warp == 14) tmem_dealloc_pair<512>(tmem);
}

}  // namespace swnerf

using namespace swnerf;

extern "C" {

int64_t swnerf_tc_packed_bytes(void) { return PK_TOTAL_BYTES; }

int64_t swnerf_tc_workspace_bytes(int64_t n_points, int training, int enc) {
  Enc E;
  if (!training || n_points <= 0 || !decode_enc(enc, &E)) return 0;
  int64_t tiles = (n_points + TILE - 1) / TILE;
  return tiles * (WS_TILE_BYTES + WS_MASK_BYTES + WS_DY_BYTES) + WS_TAIL_BYTES + (E.wide() ? tiles * WS_EXT_BYTES : 0);
}

static int pack_impl(const float* const* params, int kind, const float* tpe_host, int enc, void* packed, void* stream) {
  SW_REQUIRE(params && packed, "tc_pack_weights: null pointer");
  SW_REQUIRE(kind == 0 || kind == 1, "tc_pack_weights: kind must be 0 (canonical net) or 1 (deformation net)");
  SW_REQUIRE(kind == 0 || tpe_host, "tc_pack_weights: the deformation net needs the time embedding");
  SW_REQUIRE(aligned16(packed), "tc_pack_weights: packed must be 16-byte aligned");
  ParamPtrs P;
  SW_REQUIRE(decode_enc(enc, &P.enc), "tc_pack_weights: unsupported encoding code 0x%x (SWNERF_TC_ENC: L in {0, 4, 10, 20})", enc);
  const int np = kind == 0 ? 24 : 18;
  for (int i = 0; i < 24; ++i) {
    SW_REQUIRE(i >= np || params[i], "tc_pack_weights: null parameter %d", i);
    P.p[i] = i < np ? params[i] : nullptr;
  }
  P.kind = kind;
  for (int i = 0; i < ENC_MAX_TW + 3; ++i) P.tpe[i] = (kind == 1 && i < P.enc.tw) ? tpe_host[i] : 0.f;
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* pk = reinterpret_cast<uint8_t*>(packed);
  fold_head_kernel<<<128 / FOLD_ROWS, 256, 0, s>>>(P, reinterpret_cast<float*>(pk + PK_FOLD_OFF));
  int rc = check_launch("tc_fold_head");
  if (rc) return rc;
  pack_fwd_kernel<<<(PK_CHUNK_BYTES / 16 + 255) / 256, 256, 0, s>>>(P, pk);
  return check_launch("tc_pack_weights");
}

int swnerf_tc_pack_weights(const float* const* params, int enc, void* packed, void* stream) {
  return pack_impl(params, 0, nullptr, enc, packed, stream);
}

int swnerf_tc_pack_weights_time(const float* const* params, const float* time_embedding_host, int enc, void* packed,
                                void* stream) {
  return pack_impl(params, 1, time_embedding_host, enc, packed, stream);
}

// Tensor maps over the packed weight image seen as rows of 128 bytes: boxes of 128 rows (half a trunk chunk) and 72
// rows (half a head chunk), copied verbatim (the image is already in the UMMA swizzle, so the map itself uses none).
static int weight_tensor_maps(const void* packed, CUtensorMap* trunk, CUtensorMap* head) {
  const unsigned long long dims[2] = {128, (unsigned long long)(PK_CHUNK_BYTES / 128)};
  const unsigned long long strides[1] = {128};
  for (int h = 0; h < 2; ++h) {
    const unsigned int box[2] = {128, h ? (unsigned)(HEAD_N / 2) : 128u};
    int rc = encode_u8_tensor_map(h ? head : trunk, packed, 2, dims, strides, box);
    if (rc) return rc;
  }
  return SWNERF_OK;
}

// Copies a network's fp32 block into one of the constant-memory slots on the launching stream (const_slot_acquire).
static int stage_f32_block(const uint8_t* src, cudaStream_t s, int* slot_out) {
  const int k = const_slot_acquire(CONST_FAMILY_FWD_F32, F32_SLOTS, s);
  if (k < 0) return set_err(SWNERF_ERR_CUDA, "tc_mlp_fwd: no current device");
  cudaError_t e = cudaMemcpyToSymbolAsync(c_f32s, src, F32_COUNT * sizeof(float), (size_t)k * F32_PAD * sizeof(float),
                                          cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return set_err(SWNERF_ERR_CUDA, "tc_mlp_fwd: staging the bias block failed: %s", cudaGetErrorString(e));
  *slot_out = k;
  return SWNERF_OK;
}

// -1 = automatic (the CTA-pair kernel: 17 % faster in inference, 2-5 % in training),
// 0 = one CTA per tile (the first-generation kernel, kept as the bit-identity reference), 1 = CTA pairs.
// SWNERF_FWD_PAIR=0/1 presets it.
static std::atomic<int> g_fwd_variant{[] { const char* e = getenv("SWNERF_FWD_PAIR"); return e ? (atoi(e) ? 1 : 0) : -1; }()};

}  // extern "C"

template <bool TRAIN, int MODE>
static int launch_fwd1(const FwdArgs& g, int grid, cudaStream_t s) {
  if (once_per_device(ONCE_FWD1_BASE + (TRAIN ? 3 : 0) + MODE))
    cudaFuncSetAttribute(mlp_fwd_kernel<TRAIN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd1Smem<MODE>::TOTAL);
  mlp_fwd_kernel<TRAIN, MODE><<<grid, 512, Fwd1Smem<MODE>::TOTAL, s>>>(g);
  return check_launch("tc_mlp_fwd");
}

static int fwd_impl(const float* rays, int ray_stride, int view_col, const float* z_vals, const float* pts,
                    int64_t n_rays, int n_samples, const void* packed, int enc, float* out, void* workspace, int training,
                    int kind, void* stream) {
  SW_REQUIRE(rays && packed && out && (z_vals || pts), "tc_mlp_fwd: null pointer");
  SW_REQUIRE(view_col >= 0 && view_col + 3 <= ray_stride, "tc_mlp_fwd: the fused kernel needs viewdirs in the ray batch");
  SW_REQUIRE(!training || workspace, "tc_mlp_fwd: training needs a workspace");
  SW_REQUIRE(aligned16(out) || kind == 1, "tc_mlp_fwd: raw must be 16-byte aligned");
  SW_REQUIRE(aligned16(packed) && aligned16(workspace), "tc_mlp_fwd: buffers must be 16-byte aligned");
  SW_REQUIRE(n_samples >= 1 && n_rays >= 0, "tc_mlp_fwd: bad sizes");
  Enc E;
  SW_REQUIRE(decode_enc(enc, &E), "tc_mlp_fwd: unsupported encoding code 0x%x", enc);
  if (n_rays == 0) return SWNERF_OK;
  FwdArgs g;
  g.rays = rays; g.ray_stride = ray_stride; g.view_col = view_col; g.z = z_vals; g.S = n_samples;
  g.P = n_rays * n_samples; g.pts = pts; g.packed = reinterpret_cast<const uint8_t*>(packed); g.raw = out;
  g.ws = reinterpret_cast<uint8_t*>(workspace); g.num_tiles = (g.P + TILE - 1) / TILE; g.kind = kind; g.f32_slot = 0;
  g.Lp = E.Lp; g.Lv = E.Lv; g.PC = E.PC; g.VC = E.VC;
  g.ws_ext = (training && E.wide())
      ? g.ws + g.num_tiles * (WS_TILE_BYTES + WS_MASK_BYTES + WS_DY_BYTES) + WS_TAIL_BYTES : nullptr;
#ifdef SWNERF_EXPERIMENTS
  { static const char* ko = getenv("SWNERF_KO"); g.ko = ko ? atoi(ko) : 0; }
#endif
  int grid = (int)(g.num_tiles < sm_count() ? g.num_tiles : sm_count());
  cudaStream_t s = (cudaStream_t)stream;
  const int variant = g_fwd_variant.load();
  const bool defenc = enc == ENC_DEFAULT_CODE;
  // automatic: a launch that does not fill the GPU twice over runs one tile per CTA (a pair CTA works through its two
  // slots back to back: 0.055 vs 0.033 ms at 128 tiles; equal from ~500 tiles; 17 % faster at 6144).  Two-chunk
  // encodings always run one CTA per tile: the pair kernel's shared memory has no room for 32-KB encoding images.
  if (!E.wide() && (variant == 1 || (variant < 0 && g.num_tiles > 2 * (int64_t)sm_count()))) {        // CTA-pair kernel
    if (once_per_device(ONCE_FWD4)) {
      cudaFuncSetAttribute(mlp_fwd4_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S4_TOTAL);
      cudaFuncSetAttribute(mlp_fwd4_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S4_TOTAL);
      cudaFuncSetAttribute(mlp_fwd4_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S4_TOTAL);
      cudaFuncSetAttribute(mlp_fwd4_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S4_TOTAL);
    }
    const int64_t num_quads = (g.num_tiles + 3) / 4;
    const int grid4 = 2 * (int)(num_quads < sm_count() / 2 ? num_quads : sm_count() / 2);
    CUtensorMap tm_trunk, tm_head;
    int rc = weight_tensor_maps(packed, &tm_trunk, &tm_head);
    if (rc) return rc;
    rc = stage_f32_block(reinterpret_cast<const uint8_t*>(packed) + PK_F32_OFF, s, &g.f32_slot);
    if (rc) return rc;
    if (training) {
      if (defenc) mlp_fwd4_kernel<true, true><<<grid4, 512, S4_TOTAL, s>>>(g, tm_trunk, tm_head);
      else mlp_fwd4_kernel<true, false><<<grid4, 512, S4_TOTAL, s>>>(g, tm_trunk, tm_head);
    } else {
      if (defenc) mlp_fwd4_kernel<false, true><<<grid4, 512, S4_TOTAL, s>>>(g, tm_trunk, tm_head);
      else mlp_fwd4_kernel<false, false><<<grid4, 512, S4_TOTAL, s>>>(g, tm_trunk, tm_head);
    }
    return check_launch("tc_mlp_fwd");
  }
  const int mode = E.wide() ? 2 : (defenc ? 0 : 1);
  if (mode == 2) {
    int rc = stage_f32_block(reinterpret_cast<const uint8_t*>(packed) + PK_F32_OFF, s, &g.f32_slot);
    if (rc) return rc;
  }
  if (training) return mode == 0 ? launch_fwd1<true, 0>(g, grid, s) : mode == 1 ? launch_fwd1<true, 1>(g, grid, s) : launch_fwd1<true, 2>(g, grid, s);
  return mode == 0 ? launch_fwd1<false, 0>(g, grid, s) : mode == 1 ? launch_fwd1<false, 1>(g, grid, s) : launch_fwd1<false, 2>(g, grid, s);
}

extern "C" {

int swnerf_tc_set_fwd_variant(int variant) {
  SW_REQUIRE(variant >= -1 && variant <= 1, "tc_set_fwd_variant: variant must be -1 (automatic), 0 or 1");
  g_fwd_variant.store(variant);
  return SWNERF_OK;
}

int swnerf_tc_mlp_fwd(const float* rays, int ray_stride, int view_col, const float* z_vals, int64_t n_rays,
                      int n_samples, const void* packed, int enc, float* raw, void* workspace, int training,
                      void* stream) {
  SW_REQUIRE(z_vals, "tc_mlp_fwd: null pointer");
  return fwd_impl(rays, ray_stride, view_col, z_vals, nullptr, n_rays, n_samples, packed, enc, raw, workspace, training,
                  0, stream);
}

int swnerf_tc_mlp_fwd_points(const float* rays, int ray_stride, int view_col, const float* pts, int64_t n_rays,
                             int n_samples, const void* packed, int enc, float* raw, void* workspace, int training,
                             void* stream) {
  SW_REQUIRE(pts, "tc_mlp_fwd_points: null pointer");
  return fwd_impl(rays, ray_stride, view_col, nullptr, pts, n_rays, n_samples, packed, enc, raw, workspace, training, 0,
                  stream);
}

int swnerf_tc_time_fwd(const float* rays, int ray_stride, int view_col, const float* z_vals, int64_t n_rays,
                       int n_samples, const void* packed_time, int enc, float* dx, void* workspace, int training,
                       void* stream) {
  SW_REQUIRE(z_vals, "tc_time_fwd: null pointer");
  return fwd_impl(rays, ray_stride, view_col, z_vals, nullptr, n_rays, n_samples, packed_time, enc, dx, workspace,
                  training, 1, stream);
}

}  // extern "C"
