// Byte layouts shared by the fused tcgen05 kernels: packed weight images, shared memory carve-up,
// per-tile workspace (saved activations / masks / output gradients).
#pragma once
#include <stdint.h>

namespace swnerf {
namespace tcl {

constexpr int TILE = 128;                 // sample rows per tile (UMMA M)
constexpr int HEAD_N = 144;               // folded head: 128 view-branch units + sigma + pad to 16
constexpr int NSTAGE = 3;                 // weight ring depth
constexpr int CHUNK_B = 256 * 128;        // [256 x 64] fp16 K-major image
constexpr int HCHUNK_B = HEAD_N * 128;    // [144 x 64]
constexpr int ACT_BLK = TILE * 128;       // [128 x 64] fp16 image (16 KB)
constexpr int ACT_BYTES = 4 * ACT_BLK;    // [128 x 256]

// ---- encoding widths of a network family (decoded from the C-ABI's `enc` code, SWNERF_TC_ENC(pos_L, view_L, time_L)).
// The reference's configs use PE L=10/4 (vanilla; time L=10 for D-NeRF); the MultiRes pyramid (multires_dnerf.py:665)
// adds (pos, time, view) = (20, 8, 20), (10, 4, 10) and the identity.  A position / view image is one or two 64-column
// K-chunks (63 or 123 columns, zero padded); the time embedding is constant per call and is folded into the layer-0 bias.
struct Enc {
  int Lp, Lv, Lt;       // frequency counts (0 = identity: the raw 3 / 3 / 1 inputs)
  int pc, vc, tw;       // columns: 3 (1 + 2 Lp), 3 (1 + 2 Lv), 1 + 2 Lt
  int PC, VC;           // K-chunks of the position / view image
  int code;
  __host__ __device__ int n_full() const { return 28 + 2 * PC; }       // [256 x 64] chunks
  __host__ __device__ int n_chunks() const { return 32 + 2 * PC + VC; }
  __host__ __device__ bool wide() const { return PC > 1 || VC > 1; }
};
constexpr int ENC_DEFAULT_CODE = 10 | (4 << 8) | (10 << 16);
constexpr int ENC_MAX_TW = 41;            // 1 + 2 * 20
inline __host__ bool decode_enc(int code, Enc* e) {
  const int Lp = code & 255, Lv = (code >> 8) & 255, Lt = (code >> 16) & 255;
  auto ok = [](int L) { return L == 0 || L == 4 || L == 10 || L == 20; };     // the encoders the kernels instantiate
  if ((code >> 24) != 0 || !ok(Lp) || !ok(Lv) || Lt > 20) return false;
  e->Lp = Lp; e->Lv = Lv; e->Lt = Lt;
  e->pc = 3 * (1 + 2 * Lp); e->vc = 3 * (1 + 2 * Lv); e->tw = 1 + 2 * Lt;
  e->PC = (e->pc + 63) / 64; e->VC = (e->vc + 63) / 64;
  e->code = code;
  return true;
}

// ---- forward weight stream: chunk order per tile (see mlp_tc.cu); PC / VC = chunks of the position / view image
//  L0: PC (pe) | L1..L4: 16 | L5: PC (pe) + 4 (h) | L6, L7: 8 | head: VC (views) + 4 (h7)
//  default encoding (PC = VC = 1):  0: L0(pe) | 1-16: L1..L4 | 17: L5(pe) 18-21: L5(h) | 22-29: L6,L7 | 30: head(views) 31-34: head(h7)
constexpr int N_CHUNKS = 35;              // default encoding (the CTA-pair kernels serve one-chunk encodings only)
constexpr int N_FULL_CHUNKS = 30;
constexpr int PK_MAX_FULL = 32, PK_MAX_HEAD = 6;
constexpr int PK_CHUNK_BYTES = PK_MAX_FULL * CHUNK_B + PK_MAX_HEAD * HCHUNK_B;       // sized for two-chunk encodings
__host__ __device__ constexpr int chunk_off_n(int c, int n_full) {
  return c < n_full ? c * CHUNK_B : n_full * CHUNK_B + (c - n_full) * HCHUNK_B;
}
__host__ __device__ constexpr int chunk_off(int c) { return chunk_off_n(c, N_FULL_CHUNKS); }
// fp32 block: trunk biases [8][256] | head bias [160] (128 folded + sigma + pad) | rgb weight [3][128] | rgb bias [4]
constexpr int F32_BHEAD = 8 * 256;
constexpr int F32_WRGB = F32_BHEAD + 160;
constexpr int F32_BRGB = F32_WRGB + 384;
constexpr int F32_COUNT = F32_BRGB + 4;                                       // 2596 floats
constexpr int PK_F32_OFF = PK_CHUNK_BYTES;
constexpr int PK_FOLD_OFF = PK_F32_OFF + ((F32_COUNT * 4 + 255) / 256) * 256; // fold scratch [128][257] fp32
constexpr int PK_TOTAL_BYTES = PK_FOLD_OFF + 128 * 257 * 4;

// ---- shared memory of the CTA-pair forward kernel (per CTA; two tile slots; offsets from a 1024-aligned base)
constexpr int NST4 = 4;                   // weight ring depth; a stage holds this CTA's half of one chunk
constexpr int STG4_B = CHUNK_B / 2;       // [128 x 64] fp16 (head chunks use 72 rows of it)
constexpr int S4_ACT = 0;                                 // [2 slots] activation images [128 x 256] fp16, in place
constexpr int S4_ENC = S4_ACT + 2 * ACT_BYTES;            // [2 slots] encoding image: PE (layers 0, 5), then the view encoding (head)
constexpr int S4_RING = S4_ENC + 2 * ACT_BLK;
constexpr int S4_SCR = S4_RING + NST4 * STG4_B;
constexpr int S4_BAR = S4_SCR + TILE * 16;
constexpr int S4_TOTAL = S4_BAR + 512;    // biases live in constant memory and the base must be 1024-aligned (checked):
                                          // that is what makes room for the 4th ring stage
static_assert(S4_TOTAL <= 232448, "shared memory budget");

// ---- shared memory of the forward kernel v1 (offsets from a 1024-aligned base)
constexpr int SM_ACT = 0;
constexpr int SM_PE = SM_ACT + ACT_BYTES;
constexpr int SM_VW = SM_PE + ACT_BLK;
constexpr int SM_RING = SM_VW + ACT_BLK;
constexpr int SM_F32 = SM_RING + NSTAGE * CHUNK_B;
constexpr int SM_SCR = SM_F32 + ((F32_COUNT * 4 + 127) / 128) * 128;
constexpr int SM_BAR = SM_SCR + TILE * 16;
constexpr int SM_TOTAL = SM_BAR + 256 + 1024;                                  // + alignment slack

// ---- per-tile workspace written by the training forward
constexpr int WS_PE_OFF = 0;
constexpr int WS_VW_OFF = ACT_BLK;
constexpr int WS_H_OFF = 2 * ACT_BLK;                    // h0..h7, 64 KB each
constexpr int WS_H9_OFF = WS_H_OFF + 8 * ACT_BYTES;      // h9 [128 x 128]
constexpr int64_t WS_TILE_BYTES = WS_H9_OFF + 2 * ACT_BLK;       // 589,824
constexpr int64_t WS_MASK_BYTES = 9 * 8 * 128 * 4;               // ReLU sign bits, 36,864
// written by the backward-data kernel: dy0..dy7 [128 x 256] and the head gradient tile
// [dy9 0..63 | dy9 64..127 | (d_sigma, 0..) | (d_rgb, 0..)] as four 64-column blocks
constexpr int WS_DYH_OFF = 8 * ACT_BYTES;
constexpr int64_t WS_DY_BYTES = 9 * ACT_BYTES;                   // 589,824
// two-chunk encodings only: the second 64-column blocks of the saved position / view images, one record per tile
// [PE block 1 | view block 1], in a region of its own behind the workspace tail (every other offset stays as it is)
constexpr int64_t WS_EXT_BYTES = 2 * ACT_BLK;

// ---- transposed weight stream of the backward-data kernel: chunk order per tile
//  0,1: head^T (y9 units 0..127) 2: head^T (sigma row) | 3..30: L7^T, L6^T, L5^T(h part), L4^T .. L1^T, 4 chunks each
constexpr int NT_CHUNKS = 31;
// followed by the input-gradient weights (D-NeRF needs d/dx of the canonical net): W0[:, :63]^T and
// W5[:, :63]^T as eight [64 x 64] K-major images (pe column x unit)
constexpr int PKT_DPE_OFF = NT_CHUNKS * CHUNK_B;
constexpr int DPE_CHUNK_B = 64 * 128;
constexpr int PKT_TOTAL_BYTES = PKT_DPE_OFF + 2 * 8 * DPE_CHUNK_B;       // two groups: two-chunk position encodings

// shared memory of the backward-data kernel
constexpr int SMB_ACT = 0;
constexpr int SMB_RING = SMB_ACT + ACT_BYTES;
constexpr int SMB_WRGB = SMB_RING + NSTAGE * CHUNK_B;             // rgb_linear weight [3][128] fp32
constexpr int SMB_BAR = SMB_WRGB + 384 * 4;
constexpr int SMB_TOTAL = SMB_BAR + 256 + 1024;

// ---- fp32 scratch the backward produces before un-folding the head: G = d W_fv [128][256] | gb = d b_fv [128]
constexpr int UNFOLD_FLOATS = 128 * 256 + 128;
// workspace tail: [0,256) scalars (max|d_raw| bits) | un-fold scratch
constexpr int64_t WS_TAIL_BYTES = 256 * 1024;

}  // namespace tcl
}  // namespace swnerf
