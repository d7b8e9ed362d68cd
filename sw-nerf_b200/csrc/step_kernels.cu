// Kernels for the callers on either side of the render path (SURVEY.md 8f "next" rows):
//  f1  ray assembly: pixel -> [o, d, near, far, (time), unit viewdir] rows of the flat ray batch
//      (ray.py:10-38 get_rays + nerf/run.py:137-158 viewdir normalisation and concatenation)
//  f3  Adam on one flat parameter / gradient buffer (torch.optim.Adam(lr, betas=(0.9, 0.999)) as created by
//      create_nerf, nerf/run.py:254), and the two-loss MSE with its gradient (nerf/run.py:689-697)
#include "common.cuh"
#include "../../include/swnerf_b200.h"

namespace swnerf {

struct Cam {
  float fx, fy, cx, cy;
  float c2w[12];      // rows of the 3x4 camera-to-world matrix
  int ndc;            // ray.py:75-92 applied to (o, d) after the viewdirs were taken (nerf/run.py:137-147)
  float ndc_near;     // the near plane of ndc_rays (nerf/run.py:147 passes 1.)
  float ndc_c0, ndc_c1;   // -1 / (W / (2 focal)), -1 / (H / (2 focal)): Python-side scalars, rounded to fp32 once
};

// one ray row [o(3), d(3), near, far, (time), (unit viewdir(3))] for pixel p
__device__ __forceinline__ void write_ray(const Cam& cam, int W, int64_t p, float nearv, float farv, float frame_time,
                                          int has_time, int with_viewdirs, float* __restrict__ o) {
  float i = (float)(p % W), j = (float)(p / W);
  float dx = __fdiv_rn(__fsub_rn(i, cam.cx), cam.fx);
  float dy = -__fdiv_rn(__fsub_rn(j, cam.cy), cam.fy);
  float dz = -1.f;
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)     // sum(dirs[..., None, :] * c2w[:3, :3], -1): ((a + b) + c), no fma
    d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam.c2w[4 * r]), __fmul_rn(dy, cam.c2w[4 * r + 1])), __fmul_rn(dz, cam.c2w[4 * r + 2]));
  float org[3] = {cam.c2w[3], cam.c2w[7], cam.c2w[11]};
  int c = 8;
  if (has_time) o[c++] = frame_time;
  if (with_viewdirs) {          // from the camera-space direction, BEFORE the NDC warp (nerf/run.py:137-147)
    float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    o[c] = __fdiv_rn(d[0], nrm); o[c + 1] = __fdiv_rn(d[1], nrm); o[c + 2] = __fdiv_rn(d[2], nrm);
  }
  if (cam.ndc) {                // ray.py:75-92, operation by operation
    const float n = cam.ndc_near;
    const float tt = __fdiv_rn(-__fadd_rn(n, org[2]), d[2]);                    // t = -(near + o_z) / d_z
#pragma unroll
    for (int r = 0; r < 3; ++r) org[r] = __fadd_rn(org[r], __fmul_rn(tt, d[r]));     // shift the origin to the near plane
    const float o0 = __fdiv_rn(__fmul_rn(cam.ndc_c0, org[0]), org[2]);
    const float o1 = __fdiv_rn(__fmul_rn(cam.ndc_c1, org[1]), org[2]);
    const float o2 = __fadd_rn(1.f, __fdiv_rn(2.f * n, org[2]));
    const float d0 = __fmul_rn(cam.ndc_c0, __fsub_rn(__fdiv_rn(d[0], d[2]), __fdiv_rn(org[0], org[2])));
    const float d1 = __fmul_rn(cam.ndc_c1, __fsub_rn(__fdiv_rn(d[1], d[2]), __fdiv_rn(org[1], org[2])));
    const float d2 = __fdiv_rn(-2.f * n, org[2]);
    org[0] = o0; org[1] = o1; org[2] = o2; d[0] = d0; d[1] = d1; d[2] = d2;
  }
  o[0] = org[0]; o[1] = org[1]; o[2] = org[2];
  o[3] = d[0]; o[4] = d[1]; o[5] = d[2];
  o[6] = nearv; o[7] = farv;
}

__global__ void make_rays_kernel(Cam cam, int W, const int64_t* __restrict__ pix, int64_t n, float nearv, float farv,
                                 float frame_time, int has_time, int with_viewdirs, float* __restrict__ out, int stride) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  write_ray(cam, W, pix ? pix[t] : t, nearv, farv, frame_time, has_time, with_viewdirs, out + t * stride);
}

// f1, the per-step batch (nerf/run.py:652-681): N_rand DISTINCT pixels of one training image - inside the centre crop
// during the first precrop_iters iterations (:660-668) - their rays and their target colours, in one kernel.
// The reference draws np.random.choice(n, N_rand, replace=False) on the host and gathers with three fancy-index ops.
// Here pixel k of the batch is perm_seed(k), where perm is a keyed bijection of [0, n): a 4-round Feistel network on
// the smallest even number of bits covering n, cycle-walked back into range.  Distinctness is by construction, the
// draw needs no state, no host round trip and no O(n) shuffle.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {          // lowbias32 finaliser
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint64_t feistel_perm(uint64_t k, uint64_t n, int half_bits, uint64_t seed) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint64_t x = k;
  do {
    uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      const uint32_t f = mix32(r ^ (uint32_t)(seed >> (16 * (round & 1))) ^ (0x9e3779b9u * (round + 1)) ^ (uint32_t)(seed >> 32)) & mask;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    x = ((uint64_t)l << half_bits) | r;
  } while (x >= n);                     // the domain is < 4 n: the walk ends after < 4 steps on average
  return x;
}

__global__ void pick_batch_kernel(Cam cam, int W, const float* __restrict__ image, int crop_y0, int crop_x0, int crop_h,
                                  int crop_w, uint64_t seed, int half_bits, int64_t n_rand, float nearv, float farv,
                                  float frame_time, int has_time, int with_viewdirs, float* __restrict__ rays, int stride,
                                  float* __restrict__ target, int64_t* __restrict__ pix_out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rand) return;
  const uint64_t q = feistel_perm((uint64_t)t, (uint64_t)crop_h * crop_w, half_bits, seed);
  const int64_t y = crop_y0 + (int64_t)(q / crop_w), x = crop_x0 + (int64_t)(q % crop_w);
  const int64_t p = y * W + x;
  write_ray(cam, W, p, nearv, farv, frame_time, has_time, with_viewdirs, rays + t * stride);
  if (target) {
    target[t * 3] = __ldg(image + p * 3); target[t * 3 + 1] = __ldg(image + p * 3 + 1); target[t * 3 + 2] = __ldg(image + p * 3 + 2);
  }
  if (pix_out) pix_out[t] = p;
}

__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                                 float bc2_sqrt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  float mi = m[i] = fmaf(b1, m[i], (1.f - b1) * gi);          // exp_avg.lerp_(grad, 1 - beta1)
  float vi = v[i] = fmaf(b2, v[i], (1.f - b2) * gi * gi);
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// loss = mean((a - t)^2) [+ mean((b - t)^2)], da = 2 (a - t) * scale, db likewise; scale = 1 / count_global
__global__ void mse2_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ t,
                            int64_t n, float scale, float* __restrict__ da, float* __restrict__ db, float* __restrict__ loss) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float ti = t[i];
    float ea = a[i] - ti;
    acc += ea * ea;
    if (da) da[i] = 2.f * ea * scale;
    if (b) {
      float eb = b[i] - ti;
      acc += eb * eb;
      if (db) db[i] = 2.f * eb * scale;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc * scale);
}

}  // namespace swnerf

using namespace swnerf;

extern "C" {

static Cam make_cam(int H, int W, float fx, float fy, float cx, float cy, const float* c2w_host12, int ndc, float ndc_near,
                    double ndc_focal) {
  Cam cam;
  cam.fx = fx; cam.fy = fy; cam.cx = cx; cam.cy = cy;
  for (int i = 0; i < 12; ++i) cam.c2w[i] = c2w_host12[i];
  cam.ndc = ndc; cam.ndc_near = ndc_near;
  // ray.py:81-82: -1./(W/(2.*focal)) is a Python double; multiplying a float32 tensor by it rounds it to fp32 once
  // (ndc_focal is the caller's K[0][0] at full double precision; fx above is its fp32 rounding)
  cam.ndc_c0 = (float)(-1.0 / ((double)W / (2.0 * ndc_focal)));
  cam.ndc_c1 = (float)(-1.0 / ((double)H / (2.0 * ndc_focal)));
  return cam;
}

int swnerf_make_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w_host12,
                     const int64_t* pixels, int64_t n_rays, float nearv, float farv, float frame_time, int has_time,
                     int with_viewdirs, int ndc, float ndc_near, double ndc_focal, float* rays, int ray_stride, void* stream) {
  SW_REQUIRE(c2w_host12 && rays, "make_rays: null pointer");
  SW_REQUIRE(!ndc || ndc_focal != 0.0, "make_rays: ndc needs the focal length");
  SW_REQUIRE(H > 0 && W > 0 && fx != 0.f && fy != 0.f, "make_rays: bad intrinsics");
  SW_REQUIRE(ray_stride >= 8 + (has_time ? 1 : 0) + (with_viewdirs ? 3 : 0), "make_rays: ray_stride too small");
  SW_REQUIRE(pixels || n_rays == (int64_t)H * W, "make_rays: without a pixel list n_rays must be H*W");
  if (n_rays == 0) return SWNERF_OK;
  const Cam cam = make_cam(H, W, fx, fy, cx, cy, c2w_host12, ndc, ndc_near, ndc_focal);
  make_rays_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      cam, W, pixels, n_rays, nearv, farv, frame_time, has_time, with_viewdirs, rays, ray_stride);
  return check_launch("make_rays");
}

int swnerf_pick_batch(int H, int W, float fx, float fy, float cx, float cy, const float* c2w_host12, const float* image,
                      int crop_y0, int crop_x0, int crop_h, int crop_w, uint64_t seed, int64_t n_rand, float nearv,
                      float farv, float frame_time, int has_time, int with_viewdirs, int ndc, float ndc_near,
                      double ndc_focal, float* rays, int ray_stride, float* target, int64_t* pixels_out, void* stream) {
  SW_REQUIRE(c2w_host12 && rays, "pick_batch: null pointer");
  SW_REQUIRE(!ndc || ndc_focal != 0.0, "pick_batch: ndc needs the focal length");
  SW_REQUIRE(!target || image, "pick_batch: target colours need the image");
  SW_REQUIRE(H > 0 && W > 0 && fx != 0.f && fy != 0.f, "pick_batch: bad intrinsics");
  SW_REQUIRE(crop_h > 0 && crop_w > 0 && crop_y0 >= 0 && crop_x0 >= 0 && crop_y0 + crop_h <= H && crop_x0 + crop_w <= W,
             "pick_batch: the crop must lie inside the image");
  SW_REQUIRE(n_rand >= 0 && n_rand <= (int64_t)crop_h * crop_w, "pick_batch: cannot draw more distinct pixels than the crop holds");
  SW_REQUIRE(ray_stride >= 8 + (has_time ? 1 : 0) + (with_viewdirs ? 3 : 0), "pick_batch: ray_stride too small");
  if (n_rand == 0) return SWNERF_OK;
  const Cam cam = make_cam(H, W, fx, fy, cx, cy, c2w_host12, ndc, ndc_near, ndc_focal);
  const uint64_t n = (uint64_t)crop_h * crop_w;
  int bits = 2;
  while (bits < 62 && (1ull << bits) < n) bits += 2;
  pick_batch_kernel<<<(unsigned)((n_rand + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      cam, W, image, crop_y0, crop_x0, crop_h, crop_w, seed, bits / 2, n_rand, nearv, farv, frame_time, has_time,
      with_viewdirs, rays, ray_stride, target, pixels_out);
  return check_launch("pick_batch");
}

int swnerf_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, int64_t step, void* stream) {
  SW_REQUIRE(params && grads && exp_avg && exp_avg_sq, "adam_flat: null pointer");
  SW_REQUIRE(step >= 1, "adam_flat: step counts from 1");
  if (n == 0) return SWNERF_OK;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_flat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2));
  return check_launch("adam_flat");
}

int swnerf_mse2(const float* a, const float* b, const float* target, int64_t n, float scale, float* da, float* db,
                float* loss, void* stream) {
  SW_REQUIRE(a && target && loss, "mse2: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (n == 0) return SWNERF_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  mse2_kernel<<<(unsigned)blocks, 256, 0, s>>>(a, b, target, n, scale, da, db, loss);
  return check_launch("mse2");
}

}  // extern "C"
