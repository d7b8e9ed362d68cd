/* swnerf_b200 - C ABI of the B200-native SW-NeRF per-ray volumetric rendering path.
 *
 * Drop-in boundary.  The reference has no FFI of its own for this path: its seam is
 * the Python surface `embedder.py`, `model.py`, `ray.py` and the runner functions
 * `run_network / render_rays / create_nerf`; its only native interface is the (dead) torchsearchsorted
 * extension `void searchsorted_cuda_wrapper(at::Tensor a, at::Tensor v, at::Tensor res, bool side_left)`
 * (d_nerf/torchsearchsorted/src/cuda/searchsorted_cuda_wrapper.cpp:5-19).  This header is what a
 * ctypes / pybind binding on the reference side binds instead (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - plain pointers and sizes, no torch types.  Every pointer is a DEVICE pointer to fp32 (or int64 /
 *    fp16 where stated), row-major contiguous; the caller owns and allocates every buffer, exactly like
 *    the reference extension whose `out` is caller-allocated (searchsorted.py:32-38).
 *  - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises.
 *  - return 0 on success; otherwise an error code, with the message in swnerf_last_error()
 *    (thread-local).  Bad arguments are rejected, never silently fixed.
 *  - `rays` is the reference's flat ray batch `[N, ray_stride]` (nerf/run.py:152-158: o3 d3 near far
 *    [time] viewdir3); columns are addressed by `d_col`, `near_col`, `view_col`.
 */
#ifndef SWNERF_B200_H_
#define SWNERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWNERF_B200_VERSION 100

/* ---- library ---- */
int swnerf_version(void);
const char* swnerf_last_error(void);
/* 0 iff the current device is sm_100 (B200). */
int swnerf_device_ok(void);

/* ---- a13: torchsearchsorted.searchsorted (searchsorted.py:20-52, searchsorted_cuda_kernel.cu:83-107).
 * a[nrow_a, ncol_a] sorted rows, v[nrow_v, ncol_v]; out[max(nrow_a,nrow_v), ncol_v] int64 ==
 * np.searchsorted(a[row], v[row], side).  nrow_a / nrow_v may be 1 (row broadcast). */
int swnerf_searchsorted(const float* a, const float* v, int64_t* out, int64_t nrow_a, int64_t nrow_v,
                        int64_t ncol_a, int64_t ncol_v, int side_left, void* stream);

/* ---- a2: stratified z-values (nerf/run.py:361-383).  near/far are rays[:, near_col], rays[:, near_col+1].
 * perturb != 0 consumes t_rand[N, S] in [0,1) (the caller draws it, e.g. torch.rand, so the generator
 * stream matches the reference's torch.rand(z_vals.shape)). */
int swnerf_stratified_z(const float* rays, int ray_stride, int near_col, const float* t_rand, float* z_vals,
                        int64_t n_rays, int n_samples, int lindisp, int perturb, void* stream);

/* ---- a4: Embedder.embed (embedder.py:33-42) and its gradient.  y[rows, dims*(1+2L)]. */
int swnerf_embed_fwd(const float* x, float* y, int64_t rows, int dims, int L, void* stream);
int swnerf_embed_bwd(const float* x, const float* dy, float* dx, int64_t rows, int dims, int L, void* stream);

/* ---- a3+a4+a5: points o + d*z, PE of points and of the per-ray unit viewdir, concatenated
 * (nerf/run.py:385, :76-83) -> out[N*S, out_stride] fp32.  view_col < 0: no viewdirs; L < 0: identity. */
int swnerf_encode_points(const float* rays, int ray_stride, int view_col, const float* z_vals, float* out,
                         int64_t n_rays, int n_samples, int L_pos, int L_dir, int out_stride, void* stream);

/* ---- a8: raw2outputs (ray.py:155-198).  raw[N,S,raw_ch] with raw_ch >= 4 (create_nerf builds output_ch = 5 networks
 * when N_importance > 0, nerf/run.py:231; ray.py:175-186 reads channels 0..2 = rgb and 3 = sigma only), z_vals[N,S],
 * rays_d = rays[:, d_col:d_col+3], noise[N,S] (already scaled by raw_noise_std) or NULL. */
int swnerf_composite_fwd(const float* raw, int raw_ch, const float* z_vals, const float* rays, int ray_stride, int d_col,
                         const float* noise, int white_bkgd, int64_t n_rays, int n_samples, float* rgb_map,
                         float* disp_map, float* acc_map, float* weights, float* depth_map, void* stream);
/* autograd of the above: any of g_* may be NULL (= zero); g_disp needs the saved acc/depth maps.  d_raw[N,S,raw_ch]:
 * channels >= 4 receive zeros.  Same n_samples limit as the forward (<= 1024). */
int swnerf_composite_bwd(const float* raw, int raw_ch, const float* z_vals, const float* rays, int ray_stride, int d_col,
                         const float* noise, int white_bkgd, int64_t n_rays, int n_samples, const float* g_rgb,
                         const float* g_disp, const float* g_acc, const float* g_weights, const float* g_depth,
                         const float* acc_map, const float* depth_map, float* d_raw, void* stream);

/* ---- a9: sample_pdf (ray.py:96-153).  bins[N,n_bins]; give weights[N,n_bins-1] OR a ready cdf[N,n_bins]
 * (the bit-exact index test entry).  det: u = linspace(0,1,n_samples); else u[N,n_samples] from the caller.
 * inds (optional) receives torch.searchsorted(cdf, u, right=True) as int64. */
int swnerf_sample_pdf(const float* bins, const float* weights, const float* cdf, const float* u, int det,
                      int64_t n_rays, int n_bins, int n_samples, float* samples, int64_t* inds, void* stream);

/* ---- a9+a10+a11 fused (nerf/run.py:396-400, 416): z_mid bins, sample_pdf on weights[:,1:-1],
 * z_fine = sort(cat(z_vals, z_samples)), z_std = std(z_samples, unbiased=False).  z_samples, z_std optional;
 * z_samples is returned in ASCENDING order (the reference only uses it through std and the sort). */
int swnerf_resample(const float* z_vals, const float* weights, const float* u, int det, int64_t n_rays,
                    int n_samples, int n_importance, float* z_samples, float* z_fine, float* z_std, void* stream);
/* Check-mode / test entry of the same stage (same outputs), with
 *  - cdf_in[N, n_samples-1] (optional): sample from THIS cdf instead of building one from the weights - the
 *    north_star's "bin indices bit-exact given identical CDFs" test, on the production kernel;
 *  - inds_out[N, n_importance] int64 (optional): torch.searchsorted(cdf, u, right=True) (ray.py:136) per sample, in
 *    ascending-u order (the kernels sort the uniforms first); cdf_out[N, n_samples-1] (optional): the cdf used;
 *  - variant 0: the production kernel of the shape (64+128: the eight-lanes-per-ray kernel); variant 1: the
 *    reference-order routine - pdf sum in the order of torch.sum on the reference's CPU path (`ref_lanes` = SIMD lanes of
 *    the host that produced the comparison data: 16 for AVX512, 8 for AVX2), IEEE divisions, cdf accumulated
 *    sequentially in double like torch.cumsum on CPU (ray.py:111-114): z_fine is bit-identical to the reference's given
 *    identical weights.  render_rays uses variant 1 in precision='fp32' (the <= 1e-5 check mode). */
int swnerf_resample_check(const float* z_vals, const float* weights, const float* cdf_in, const float* u, int det,
                          int64_t n_rays, int n_samples, int n_importance, int variant, int ref_lanes, float* z_samples,
                          float* z_fine, float* z_std, int64_t* inds_out, float* cdf_out, void* stream);
/* Kernel used for the 64 + 128 shape of the reference configs: 1 (default) = eight lanes per ray, four rays per
 * warp; 0 = the first specialisation, one warp per ray.  Both are verified per ray and fall back to the generic
 * routine, so the results are identical; the switch exists for A/B timing (tools/bench_ray_kernels.py). */
int swnerf_set_resample_variant(int variant);
/* Diagnostic (host-synchronising, not for the hot path): number of rays the eight-lane kernel handed to the generic
 * routine since the last reset.  Well-formed inputs should keep this at a fraction of a percent. */
int swnerf_resample_fallbacks(unsigned long long* count, int reset, void* stream);

/* ---- a6 (check path): fp32 SIMT GEMM with the nn.Linear epilogue (model.py:43-57).
 *  op 0: C[M,N] = A[M,K] . B[N,K]^T  (+bias[N]) (+C if accumulate) (activation)     forward  x W^T
 *  op 1: C[M,N] = A[M,K] . B[K,N]    (+C if accumulate) (times act'(mask[m,n]))      dgrad    dy W
 *  op 2: C[M,N] (+)= A[K,M]^T . B[K,N]                                              wgrad    dy^T x
 * act_flags: bits 0-1 = epilogue activation (0 none, 1 ReLU as model.py:43, 2 ELU as TNeRF's layers,
 * model.py:163-171); bits 4-5 = how `mask` (the saved activation OUTPUT) scales the result: 0 ReLU' = (mask > 0),
 * 1 ELU' = (mask > 0 ? 1 : mask + 1).  The values 0 / 1 keep the earlier relu-flag meaning. */
int swnerf_sgemm(int op, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                 int64_t N, int64_t K, const float* bias, int accumulate, int act_flags, const float* mask,
                 int64_t ldmask, void* stream);
/* ---- a6 / a7 / f4 on the tensor cores for the shapes the fused kernels are not instantiated for (TNeRF, model.py:152-210;
 * MultiRes encoding widths, multires_dnerf.py:665): swnerf_sgemm's ops 0 and 1 with fp16 operands and fp32
 * accumulation (tcgen05), same epilogue and act_flags.  A is multiplied by a_scale (a power of two; gradients are
 * lifted into fp16's normal range) before rounding and the product by 1 / a_scale; a non-null a_scale_dev (one float in
 * device memory, see swnerf_pow2_scale) overrides a_scale.  Needs 16 <= N <= 256, K <= 256
 * (swnerf_hgemm_tc_supported); everything else stays on swnerf_sgemm. */
int swnerf_hgemm_tc_supported(int64_t N, int64_t K);
int swnerf_hgemm_tc(int op, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                    int64_t N, int64_t K, const float* bias, int accumulate, int act_flags, const float* mask,
                    int64_t ldmask, float a_scale, const float* a_scale_dev, void* stream);
/* swnerf_sgemm's op 2 on the tensor cores: G[n_out, k_in] += dY[M, n_out]^T . X[M, k_in] (fp16 operands read as MN-major
 * images, fp32 accumulation in tensor memory over all of a CTA's 128-sample tiles, one red.add flush).  dY is scaled
 * like A above.  Needs 32 <= n_out <= 256, k_in <= 256. */
int swnerf_hgemm_tc_wgrad(const float* dY, int64_t ldy, const float* X, int64_t ldx, float* G, int64_t ldg, int64_t M,
                          int64_t n_out, int64_t k_in, float a_scale, const float* a_scale_dev, void* stream);
/* out[0] = 2^floor(log2(target / max|x|)) on the device (1 for an all-zero x): the a_scale_dev of a gradient chain, so
 * that no host synchronisation is needed to choose it.  The fused backward scales its gradients the same way. */
int swnerf_pow2_scale(const float* x, int64_t n, float target, float* out, void* stream);
/* out[rows, cols] = d * act'(y) from the activation's output y; kind 0 ReLU, 1 ELU (TNeRF's colour head ends in a
 * ReLU, model.py:183-186). */
int swnerf_act_bwd(const float* d, int64_t ldd, const float* y, int64_t ldy, int64_t rows, int cols, int kind,
                   float* out, int64_t ldo, void* stream);
/* out[cols] (+)= sum over rows of x[rows, ld]  (bias gradients). */
int swnerf_colsum(const float* x, int64_t ld, int64_t rows, int cols, float* out, int accumulate, void* stream);

/* ---- a3+a4+a5+a6 fused, tcgen05 (the hot kernel): see the second half of this header. ---- */

/* Geometry of the fused 8x256 skip MLP (model.py:11-37 with D=8, W=256, skips=[4], use_viewdirs=True).  The positional
 * encodings are a per-network-family property passed to every call as one `enc` code:
 *   SWNERF_TC_ENC(pos_L, view_L, time_L)   frequency counts of get_embedder (embedder.py:45-59); L <= 0 = identity
 * configs/lego.txt and every D-NeRF config: (10, 4, 10) -> input_ch 63 / input_ch_views 27 / input_ch_time 21;
 * the MultiRes pyramid (multires_dnerf.py:665, channels = (pos, time, view)): SWNERF_TC_ENC(20, 20, 8),
 * SWNERF_TC_ENC(10, 10, 4) and SWNERF_TC_ENC(0, 0, 0).  Supported: pos_L, view_L in {0, 4, 10, 20}, time_L <= 20; a
 * call with another code returns SWNERF_ERR_ARG.  Other shapes run through swnerf_sgemm / swnerf_hgemm_tc. */
#define SWNERF_TC_TILE 128            /* sample rows per tile */
#define SWNERF_TC_W 256
#define SWNERF_TC_NPARAM 24           /* state_dict tensors, order below */
#define SWNERF_TC_ENC(pos_L, view_L, time_L) \
  (((pos_L) > 0 ? (pos_L) : 0) | (((view_L) > 0 ? (view_L) : 0) << 8) | (((time_L) > 0 ? (time_L) : 0) << 16))
#define SWNERF_TC_ENC_DEFAULT SWNERF_TC_ENC(10, 4, 10)

/* Parameter pointer order for `params` / `grads` (state_dict names, model.py:22-37):
 *   [2i], [2i+1]  pts_linears.i.weight / .bias   (i = 0..7)
 *   [16],[17]     views_linears.0.weight / .bias
 *   [18],[19]     feature_linear.weight / .bias
 *   [20],[21]     alpha_linear.weight / .bias
 *   [22],[23]     rgb_linear.weight / .bias                                                          */

/* Bytes of the packed fp16 weight image (UMMA-canonical, 128B-swizzled K-major chunks + fp32 biases)
 * the fused kernels stream; and of the transposed image used by the backward-data kernel (sized for every encoding). */
int64_t swnerf_tc_packed_bytes(void);
/* Repack fp32 master weights (24 device pointers in a HOST array) into the packed image.  Also folds
 * feature_linear into views_linears (no nonlinearity between them, model.py:50-55). */
int swnerf_tc_pack_weights(const float* const* params, int enc, void* packed, void* stream);

/* Workspace bytes per call for n_points sample rows: saved activations (training only). */
int64_t swnerf_tc_workspace_bytes(int64_t n_points, int training, int enc);

/* Fused forward: for every ray n and sample s: p = o + d*z[n,s]; x = [PE(p) | PE(viewdir)];
 * raw[n,s,:] = MLP(x).  rays[N, ray_stride] as in nerf/run.py:152-158.  If training != 0 the
 * post-ReLU activations needed by the backward are written to `workspace`. */
int swnerf_tc_mlp_fwd(const float* rays, int ray_stride, int view_col, const float* z_vals, int64_t n_rays,
                      int n_samples, const void* packed, int enc, float* raw, void* workspace, int training,
                      void* stream);

/* Fused backward: d_raw[N,S,4] -> fp32 gradients of the 24 parameter tensors, ACCUMULATED into
 * grads[i] (so coarse+fine passes and the flat all-reduce buffer need no extra copy).  `packed_t`
 * is the transposed weight image from swnerf_tc_pack_weights_t.  grad_scale multiplies d_raw before
 * the fp16 conversion and is divided out again in the fp32 flush; 0 = pick a power of two from
 * max|d_raw| on the device (no host sync). */
int64_t swnerf_tc_packed_t_bytes(void);
/* `packed` is the forward image of the same parameters (its folded head is reused). */
int swnerf_tc_pack_weights_t(const float* const* params, int enc, const void* packed, void* packed_t, void* stream);
int swnerf_tc_mlp_bwd(const float* d_raw, int64_t n_rays, int n_samples, const void* packed,
                      const void* packed_t, int enc, const float* const* params, void* workspace,
                      float* const* grads, float grad_scale, void* stream);

/* ---- D-NeRF (a1d, a5d, a6d, a7) on the same fused kernels -------------------------------------------------
 * The canonical network `_occ` (NeRFOriginal, model.py:227-296) is the 24-tensor network above evaluated at
 * explicit sample positions pts[N*S,3] = x + dx (model.py:148-150); its backward can also return d_pts
 * (the positional encoding sits inside the autograd graph there).
 * The deformation network `_time` / `_time_out` (model.py:113-136; 18 tensors: _time.i.weight/.bias i=0..7,
 * _time_out.weight/.bias) maps (x, t) -> dx[N*S,3].  `time_embedding`: PE(t), 1 + 2 time_L floats (21 for L=10; one
 * time per call, run_dnerf.py:53). */
int swnerf_tc_mlp_fwd_points(const float* rays, int ray_stride, int view_col, const float* pts, int64_t n_rays,
                             int n_samples, const void* packed, int enc, float* raw, void* workspace, int training,
                             void* stream);
int swnerf_tc_mlp_bwd_points(const float* d_raw, int64_t n_rays, int n_samples, const void* packed,
                             const void* packed_t, int enc, const float* const* params, void* workspace,
                             float* const* grads, float grad_scale, const float* pts, float* d_pts, void* stream);
int swnerf_tc_pack_weights_time(const float* const* params, const float* time_embedding_host, int enc, void* packed,
                                void* stream);
int swnerf_tc_pack_weights_time_t(const float* const* params, int enc, const void* packed, void* packed_t, void* stream);
int swnerf_tc_time_fwd(const float* rays, int ray_stride, int view_col, const float* z_vals, int64_t n_rays,
                       int n_samples, const void* packed_time, int enc, float* dx, void* workspace, int training,
                       void* stream);
int swnerf_tc_time_bwd(const float* d_dx, int64_t n_rays, int n_samples, const void* packed_time,
                       const void* packed_time_t, int enc, const float* const* params, const float* time_embedding_dev,
                       void* workspace, float* const* grads, float grad_scale, void* stream);

/* ---- next rows, one step either side of the path ------------------------------------------
 * f1: ray assembly (ray.py:10-38 get_rays, nerf/run.py:137-158): pixel p = j*W + i ->
 *     rays[k] = [o(3), d(3), near, far, (frame_time), (unit viewdir(3))].  pixels == NULL: all H*W pixels in
 *     row-major order.  c2w_host12: HOST pointer to the 3x4 camera-to-world matrix, row-major.  ndc != 0 applies
 *     ndc_rays(H, W, ndc_focal, ndc_near, o, d) (ray.py:75-92) to origin and direction after the unit viewdir was taken
 *     from the camera-space direction, as render() does for the LLFF configs (nerf/run.py:137-147; ndc_near = 1);
 *     ndc_focal is K[0][0] in double precision (the reference forms -1/(W/(2 focal)) in Python doubles). */
int swnerf_make_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w_host12,
                     const int64_t* pixels, int64_t n_rays, float nearv, float farv, float frame_time, int has_time,
                     int with_viewdirs, int ndc, float ndc_near, double ndc_focal, float* rays, int ray_stride,
                     void* stream);
/* f1: the per-step training batch of nerf/run.py:652-681 in one kernel: n_rand DISTINCT pixels drawn from the crop
 *     [crop_y0, crop_y0+crop_h) x [crop_x0, crop_x0+crop_w) of the H x W image (the whole image, or the centre crop of the
 *     first precrop_iters iterations), their ray rows (as swnerf_make_rays) and their target colours gathered from
 *     image[H*W, 3] (device; optional together with `target`).  The draw is pixel k = perm_seed(k), a keyed bijection of
 *     the crop (Feistel network, cycle-walked): without replacement by construction like np.random.choice(...,
 *     replace=False), stateless, no host round trip.  pixels_out (optional) receives the flat ids y*W + x. */
int swnerf_pick_batch(int H, int W, float fx, float fy, float cx, float cy, const float* c2w_host12, const float* image,
                      int crop_y0, int crop_x0, int crop_h, int crop_w, uint64_t seed, int64_t n_rand, float nearv,
                      float farv, float frame_time, int has_time, int with_viewdirs, int ndc, float ndc_near,
                      double ndc_focal, float* rays, int ray_stride, float* target, int64_t* pixels_out, void* stream);
/* f3: torch.optim.Adam (nerf/run.py:254; no weight decay, no amsgrad) on ONE flat buffer; `step` counts from 1. */
int swnerf_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, int64_t step, void* stream);
/* f3: loss = scale * (sum (a-t)^2 [+ sum (b-t)^2]) with scale = 1/(global element count) (img2mse, utils.py:12;
 * nerf/run.py:689-697) and its gradients da, db (optional).  loss: one device float. */
int swnerf_mse2(const float* a, const float* b, const float* target, int64_t n, float scale, float* da, float* db,
                float* loss, void* stream);

/* Selects the forward kernel (process-wide): -1 (default) and 1 run on CTA pairs (cta_group::2, two tile slots per
 * CTA, tensor-map weight loads); 0 runs the first-generation kernel, one CTA per 128-sample tile.  Both produce
 * bit-identical outputs (tests/test_gpu_render.py::test_cta_pair_forward_variant_matches_default). */
int swnerf_tc_set_fwd_variant(int variant);

/* Selects the backward of swnerf_tc_mlp_bwd (process-wide): -1 / 0 (default) = the two-kernel backward (data gradients,
 * then weight gradients over the dy images in HBM); 1 = the layer-pipelined kernel for launches of more than two tiles
 * per SM: one persistent kernel whose CTAs each own one layer's data- or weight-gradient role, the dy images travelling
 * between them through an L2-resident ring, so only the saved activations are read from HBM (3x less DRAM traffic;
 * not yet faster: profiles/r2_lw_backward.md).  Both produce the same gradients up to the accumulation order
 * (tests/test_gpu_render.py::test_layer_pipelined_backward_matches_two_kernel_backward). */
int swnerf_tc_set_bwd_variant(int variant);

/* Per-kernel device timing of the last swnerf_tc_mlp_bwd on this thread (bench.py's roofline): when
 * profiling is on, CUDA events are recorded on the launching stream around the data-gradient and the
 * weight-gradient kernels; swnerf_tc_last_bwd_ms synchronises on them. */
int swnerf_tc_set_profiling(int on);
int swnerf_tc_last_bwd_ms(float* data_ms, float* weight_ms);

/* Hardware self-test of the tcgen05 building blocks (tests only): one 128-row tile on one CTA.
 *  mode 0: D[128,N] = A[128,K] . B[N,K]^T  (K-major operands, K in {64,128,192,256})
 *  mode 1: D[128,N] = A[K,128]^T . B[K,N]   (MN-major operands, K = 128 samples)
 * A, B, D fp32 row-major; scratch >= 256 KiB. */
int swnerf_tc_selftest(int mode, const float* A, const float* B, float* D, int N, int K, void* scratch,
                       void* stream);

/* Tensor-pipe issue-rate probe (tools only): one CTA per SM issues 4*iters back-to-back 128 x N x 16 MMAs;
 * cycles_per_mma[sm_count] receives SM clock cycles per MMA.  variant 0: A from shared memory, 1: A from TMEM. */
int swnerf_tc_probe(int variant, int N, int iters, float* cycles_per_mma, void* stream);

/* CTA-pair (cta_group::2) self-test and rate probe (tests / tools only): D[256,N] = A[256,K] . B[N,K]^T on a cluster
 * of two CTAs, M = 256 MMAs issued by the leader.  iters > 1 repeats the K loop; n_pairs clusters run the same problem;
 * cycles_per_mma[n_pairs] (optional) receives SM cycles per MMA.  D is written by pair 0 (may be NULL).
 * iters == 0 selects the MN-major form (the weight-gradient shape): D[256,N] = A[K,256]^T . B[K,N], N in {128,256}. */
int swnerf_tc_selftest_pair(const float* A, const float* B, float* D, int N, int K, int iters, int n_pairs,
                            float* cycles_per_mma, void* scratch, void* stream);

/* Number of kernels the library has launched in this process since the last reset (bench.py's
 * gpu_launches). */
int64_t swnerf_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* SWNERF_B200_H_ */
