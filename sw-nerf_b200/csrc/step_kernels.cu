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
};

__global__ void make_rays_kernel(Cam cam, int W, const int64_t* __restrict__ pix, int64_t n, float nearv, float farv,
                                 float frame_time, int has_time, int with_viewdirs, float* __restrict__ out, int stride) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int64_t p = pix ? pix[t] : t;
  float i = (float)(p % W), j = (float)(p / W);
  float dx = __fdiv_rn(__fsub_rn(i, cam.cx), cam.fx);
  float dy = -__fdiv_rn(__fsub_rn(j, cam.cy), cam.fy);
  float dz = -1.f;
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)     // sum(dirs[..., None, :] * c2w[:3, :3], -1): ((a + b) + c), no fma
    d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam.c2w[4 * r]), __fmul_rn(dy, cam.c2w[4 * r + 1])), __fmul_rn(dz, cam.c2w[4 * r + 2]));
  float* o = out + t * stride;
  o[0] = cam.c2w[3]; o[1] = cam.c2w[7]; o[2] = cam.c2w[11];
  o[3] = d[0]; o[4] = d[1]; o[5] = d[2];
  o[6] = nearv; o[7] = farv;
  int c = 8;
  if (has_time) o[c++] = frame_time;
  if (with_viewdirs) {
    float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    o[c] = __fdiv_rn(d[0], nrm); o[c + 1] = __fdiv_rn(d[1], nrm); o[c + 2] = __fdiv_rn(d[2], nrm);
  }
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

int swnerf_make_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w_host12,
                     const int64_t* pixels, int64_t n_rays, float nearv, float farv, float frame_time, int has_time,
                     int with_viewdirs, float* rays, int ray_stride, void* stream) {
  SW_REQUIRE(c2w_host12 && rays, "make_rays: null pointer");
  SW_REQUIRE(H > 0 && W > 0 && fx != 0.f && fy != 0.f, "make_rays: bad intrinsics");
  SW_REQUIRE(ray_stride >= 8 + (has_time ? 1 : 0) + (with_viewdirs ? 3 : 0), "make_rays: ray_stride too small");
  SW_REQUIRE(pixels || n_rays == (int64_t)H * W, "make_rays: without a pixel list n_rays must be H*W");
  if (n_rays == 0) return SWNERF_OK;
  Cam cam;
  cam.fx = fx; cam.fy = fy; cam.cx = cx; cam.cy = cy;
  for (int i = 0; i < 12; ++i) cam.c2w[i] = c2w_host12[i];
  make_rays_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      cam, W, pixels, n_rays, nearv, farv, frame_time, has_time, with_viewdirs, rays, ray_stride);
  return check_launch("make_rays");
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
