// HBM-bound per-ray kernels: stratified sampling, positional encoding, alpha compositing
// (forward / backward), hierarchical resampling (pdf -> cdf -> searchsorted -> lerp -> sort)
// and the torchsearchsorted-compatible batched binary search.
//
// Reference behaviour restated (never copied): nerf/run.py:361-385 (sampling, points),
// embedder.py:12-59, ray.py:96-153 (sample_pdf), ray.py:155-198 (raw2outputs),
// nerf/run.py:400,416 (sort, z_std), d_nerf/torchsearchsorted/src/cuda/searchsorted_cuda_kernel.cu.
//
// Mapping: one warp per ray, lanes stride the samples (coalesced float4 / float loads),
// transmittance and cdf by warp shuffles scans with a running carry, so any S works; the resampling of the
// reference configs' 64 + 128 shape runs on eight lanes per ray (resample64q_kernel).
#include "common.cuh"
#include "../../include/swnerf_b200.h"

namespace swnerf {

constexpr int kWarpsPerBlock = 8;

// ---------------------------------------------------------------------------------------------
// a2  stratified z-values                                    nerf/run.py:361-383
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float linspace01(int i, int S) {
  // torch.linspace(0, 1, S): symmetric evaluation, step = 1/(S-1)
  if (S == 1) return 0.f;
  float step = 1.0f / (float)(S - 1);
  return (i < S / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(S - 1 - i)));
}

__device__ __forceinline__ float z_at(float nearv, float farv, int i, int S, int lindisp) {
  float t = linspace01(i, S);
  if (!lindisp) return __fadd_rn(__fmul_rn(nearv, __fsub_rn(1.f, t)), __fmul_rn(farv, t));
  float a = __fmul_rn(__fdiv_rn(1.f, nearv), __fsub_rn(1.f, t));
  float b = __fmul_rn(__fdiv_rn(1.f, farv), t);
  return __fdiv_rn(1.f, __fadd_rn(a, b));
}

__global__ void stratified_kernel(const float* __restrict__ rays, int ray_stride, int near_col,
                                  const float* __restrict__ t_rand, float* __restrict__ z_out,
                                  int64_t N, int S, int lindisp, int perturb) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * S) return;
  int64_t r = idx / S;
  int i = (int)(idx - r * S);
  float nearv = rays[r * ray_stride + near_col], farv = rays[r * ray_stride + near_col + 1];
  float z = z_at(nearv, farv, i, S, lindisp);
  if (perturb) {
    float zl = (i > 0) ? z_at(nearv, farv, i - 1, S, lindisp) : z;
    float zu = (i < S - 1) ? z_at(nearv, farv, i + 1, S, lindisp) : z;
    float lower = (i > 0) ? __fmul_rn(0.5f, __fadd_rn(z, zl)) : z;
    float upper = (i < S - 1) ? __fmul_rn(0.5f, __fadd_rn(zu, z)) : z;
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[idx]));
  }
  z_out[idx] = z;
}

// ---------------------------------------------------------------------------------------------
// a4  positional encoding, standalone (Embedder drop-in)       embedder.py:33-42
// ---------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t M, int d, int L) {
  int od = d * (1 + 2 * L);
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * od) return;
  int64_t r = idx / od;
  int c = (int)(idx - r * od);
  int blk = c / d, j = c - blk * d;
  float v = x[r * d + j];
  if (blk > 0) {
    int k = (blk - 1) >> 1;
    float a = v * exp2f((float)k);            // exact power-of-two scaling
    v = ((blk - 1) & 1) ? cosf(a) : sinf(a);
  }
  y[idx] = v;
}

__global__ void embed_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                 float* __restrict__ dx, int64_t M, int d, int L) {
  int od = d * (1 + 2 * L);
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * d) return;
  int64_t r = idx / d;
  int j = (int)(idx - r * d);
  float v = x[idx];
  const float* g = dy + r * od;
  float acc = g[j];
  for (int k = 0; k < L; ++k) {
    float f = exp2f((float)k);
    float s, c;
    sincosf(v * f, &s, &c);
    acc += f * (c * g[d * (1 + 2 * k) + j] - s * g[d * (2 + 2 * k) + j]);
  }
  dx[idx] = acc;
}

// a3 + a4 fused for the fp32 layered path: rays + z -> [N*S, in_pts + in_views] embedded rows
// (nerf/run.py:385 points, :76-83 embed + expand viewdirs + cat).
__global__ void encode_points_kernel(const float* __restrict__ rays, int ray_stride, int view_col,
                                     const float* __restrict__ z, float* __restrict__ out,
                                     int64_t N, int S, int L_pos, int L_dir, int out_stride) {
  int in_pts = 3 * (1 + 2 * (L_pos < 0 ? 0 : L_pos));
  int in_views = (view_col >= 0) ? 3 * (1 + 2 * (L_dir < 0 ? 0 : L_dir)) : 0;
  int od = in_pts + in_views;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * S * od) return;
  int64_t row = idx / od;
  int c = (int)(idx - row * od);
  int64_t r = row / S;
  const float* ray = rays + r * ray_stride;
  float v;
  int L, cc;
  if (c < in_pts) {
    cc = c; L = L_pos;
    int j = cc % 3;
    v = __fadd_rn(ray[j], __fmul_rn(ray[3 + j], z[row]));   // o + d*z, no fma (matches eager mul,add)
  } else {
    cc = c - in_pts; L = L_dir;
    v = ray[view_col + cc % 3];
  }
  int blk = cc / 3;
  if (blk > 0 && L > 0) {
    int k = (blk - 1) >> 1;
    float a = v * exp2f((float)k);
    v = ((blk - 1) & 1) ? cosf(a) : sinf(a);
  }
  out[row * out_stride + c] = v;
}

// ---------------------------------------------------------------------------------------------
// a8  raw2outputs forward                                     ray.py:168-196
// ---------------------------------------------------------------------------------------------
struct RayGeom {
  float norm;
};

__device__ __forceinline__ float ray_norm(const float* __restrict__ rays, int64_t r, int ray_stride, int d_col) {
  const float* d = rays + r * ray_stride + d_col;
  float x = d[0], y = d[1], z = d[2];
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// per-sample quantities shared by forward and backward
__device__ __forceinline__ void sample_alpha(float sigma, float noise, float zi, float zn, bool last, float norm,
                                             float& dist, float& e, float& alpha, float& u, bool& on) {
  dist = last ? 1e10f : __fsub_rn(zn, zi);
  dist = __fmul_rn(dist, norm);
  float s = sigma + noise;
  on = s > 0.f;
  s = on ? s : 0.f;
  e = __expf(-__fmul_rn(s, dist));        // exp(-relu(sigma)*dists); MUFU.EX2 path: <= 2 ulp + 1 ulp of the argument
                                          // scaling, far inside the 1e-5 check-mode tolerance and 5x fewer instructions
  alpha = __fsub_rn(1.f, e);
  u = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
}

__device__ __forceinline__ float sigmoidf_(float x) { return __frcp_rn(1.f + __expf(-x)); }

// All loads of a ray are issued before any arithmetic (NC = compile-time number of 32-sample chunks, registers
// hold the whole ray: 6 x (float4 + z + noise) for S = 192), so every warp keeps ~5 KB in flight instead of
// 0.6 KB and the kernel is HBM- rather than latency-bound.  z[i+1] comes from the neighbouring lane.
// raw_ch = channels per sample in `raw` (4: one float4 per sample; > 4: the reference's output_ch = 5 networks,
// nerf/run.py:231 - ray.py:175-186 reads channels 0..3 only, so do we, with scalar loads).
template <int NC>
__device__ __forceinline__ void load_ray(const float* __restrict__ rawr, int raw_ch, const float* __restrict__ zr,
                                         const float* __restrict__ nr, int S, int lane, float4 (&q)[NC], float (&zi)[NC],
                                         float (&nz)[NC]) {
  const float4* raw4 = reinterpret_cast<const float4*>(rawr);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int i = c * 32 + lane;
    bool valid = i < S;
    if (raw_ch == 4) {
      q[c] = valid ? __ldg(raw4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float* p = rawr + (size_t)i * raw_ch;
      q[c] = valid ? make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    zi[c] = valid ? __ldg(zr + i) : 0.f;
    nz[c] = (valid && nr) ? __ldg(nr + i) : 0.f;
  }
}

template <int NC>
__device__ __forceinline__ float z_next(const float (&zi)[NC], int c, int lane) {
  float up = __shfl_down_sync(0xffffffffu, zi[c], 1);                 // z[i+1] of lanes 0..30
  float first_of_next = (c + 1 < NC) ? __shfl_sync(0xffffffffu, zi[c + 1 < NC ? c + 1 : c], 0) : 0.f;
  return lane == 31 ? first_of_next : up;
}

template <int NC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(const float* __restrict__ raw, int raw_ch, const float* __restrict__ z, const float* __restrict__ rays,
                     int ray_stride, int d_col, const float* __restrict__ noise, int white_bkgd, int64_t N, int S,
                     float* __restrict__ rgb_map, float* __restrict__ disp_map, float* __restrict__ acc_map,
                     float* __restrict__ weights, float* __restrict__ depth_map) {
  int lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= N) return;
  const float* rawr = raw + r * S * raw_ch;
  const float* zr = z + r * S;
  const float* nr = noise ? noise + r * S : nullptr;
  float4 q[NC]; float zi[NC], nz[NC];
  load_ray<NC>(rawr, raw_ch, zr, nr, S, lane, q, zi, nz);
  float norm = ray_norm(rays, r, ray_stride, d_col);
  float carry = 1.f;
  float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int i = c * 32 + lane;
    bool valid = i < S;
    float zn = z_next<NC>(zi, c, lane);
    float dist, e, alpha, u; bool on;
    sample_alpha(q[c].w, nz[c], zi[c], zn, i == S - 1, norm, dist, e, alpha, u, on);
    if (!valid) { alpha = 0.f; u = 1.f; }
    float incl = warp_scan_mul(u, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    float T = carry * excl;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    float w = alpha * T;
    if (valid) {
      weights[r * S + i] = w;
      sr += w * sigmoidf_(q[c].x);
      sg += w * sigmoidf_(q[c].y);
      sb += w * sigmoidf_(q[c].z);
      sd += w * zi[c];
      sa += w;
    }
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
  if (lane == 0) {
    if (white_bkgd) { float bg = 1.f - sa; sr += bg; sg += bg; sb += bg; }
    rgb_map[r * 3 + 0] = sr; rgb_map[r * 3 + 1] = sg; rgb_map[r * 3 + 2] = sb;
    depth_map[r] = sd;
    acc_map[r] = sa;
    // torch.max(1e-10, NaN) propagates the 0/0 NaN of an empty ray (ray.py:192); fmaxf would not.
    float qd = sd / sa;
    disp_map[r] = (qd != qd) ? qd : 1.f / fmaxf(1e-10f, qd);
  }
}

// ---------------------------------------------------------------------------------------------
// a8  raw2outputs backward (autograd of ray.py:168-196)
//   G_i = dL/dw_i = g_rgb.c_i + g_depth z_i + g_acc' + g_w[i]
//   dL/dalpha_i = G_i T_i - R_i / u_i,  R_i = sum_{k>i} G_k w_k   (T_k carries the factor u_i)
//   dL/dsigma_i = dL/dalpha_i * dist_i * e_i * [sigma_i + noise_i > 0],   e_i = exp(-relu(.) dist_i)
//   dL/draw_rgb = w_i g_rgb c (1 - c)
// The ray is loaded once into registers; sweep A walks it forward and keeps T_i and G_i w_i per sample, sweep B
// walks the chunks BACKWARD and forms R_i as a true suffix sum (reverse warp scan + carry), so no cancellation.
// e_i/u_i <= 1 is formed first so that a tiny u_i never amplifies the rounding of R_i.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {   // inclusive suffix sum
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

template <int NC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(const float* __restrict__ raw, int raw_ch, const float* __restrict__ z, const float* __restrict__ rays,
                     int ray_stride, int d_col, const float* __restrict__ noise, int white_bkgd, int64_t N, int S,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_disp, const float* __restrict__ g_acc,
                     const float* __restrict__ g_w, const float* __restrict__ g_depth,
                     const float* __restrict__ acc_map, const float* __restrict__ depth_map,
                     float* __restrict__ d_raw) {
  int lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= N) return;
  const float* rawr = raw + r * S * raw_ch;
  float* outr = d_raw + r * S * raw_ch;
  const float* zr = z + r * S;
  const float* nr = noise ? noise + r * S : nullptr;
  const float* gw = g_w ? g_w + r * S : nullptr;
  float4 q[NC]; float zi[NC], nz[NC], gwv[NC];
  load_ray<NC>(rawr, raw_ch, zr, nr, S, lane, q, zi, nz);
#pragma unroll
  for (int c = 0; c < NC; ++c) gwv[c] = (gw && c * 32 + lane < S) ? __ldg(gw + c * 32 + lane) : 0.f;
  float norm = ray_norm(rays, r, ray_stride, d_col);
  float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f;
  if (g_rgb) { gr = g_rgb[r * 3]; gg = g_rgb[r * 3 + 1]; gb = g_rgb[r * 3 + 2]; }
  if (g_depth) gd = g_depth[r];
  if (g_acc) ga = g_acc[r];
  if (g_disp) {   // disp = 1 / max(1e-10, depth/acc)
    float acc = acc_map[r], dep = depth_map[r];
    float qd = dep / acc;
    if (qd > 1e-10f) {
      float gq = -g_disp[r] / (qd * qd);
      gd += gq / acc;
      ga += -gq * dep / (acc * acc);
    }
  }
  if (white_bkgd) ga -= (gr + gg + gb);

  // sweep A (forward): per sample T, and the quantities sweep B needs
  float Tv[NC], ev[NC], uv[NC], distv[NC], Gv[NC], wv[NC];
  bool onv[NC];
  float carry = 1.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int i = c * 32 + lane;
    bool valid = i < S;
    float zn = z_next<NC>(zi, c, lane);
    float alpha;
    sample_alpha(q[c].w, nz[c], zi[c], zn, i == S - 1, norm, distv[c], ev[c], alpha, uv[c], onv[c]);
    if (!valid) { alpha = 0.f; uv[c] = 1.f; }
    float incl = warp_scan_mul(uv[c], lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    Tv[c] = carry * excl;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    wv[c] = alpha * Tv[c];
    float cr = sigmoidf_(q[c].x), cg = sigmoidf_(q[c].y), cb = sigmoidf_(q[c].z);
    Gv[c] = gr * cr + gg * cg + gb * cb + gd * zi[c] + ga + gwv[c];
    q[c].x = cr; q[c].y = cg; q[c].z = cb;           // keep the colours for sweep B
  }
  // sweep B (backward): suffix sums of G_k w_k
  float rcarry = 0.f;
#pragma unroll
  for (int c = NC - 1; c >= 0; --c) {
    int i = c * 32 + lane;
    bool valid = i < S;
    float Gw = valid ? Gv[c] * wv[c] : 0.f;
    float suf = warp_rscan_add(Gw, lane);                       // sum_{k>=i} within the chunk
    float nxt = __shfl_down_sync(0xffffffffu, suf, 1);
    float R = ((lane == 31) ? 0.f : nxt) + rcarry;               // sum_{k>i} over the whole ray
    rcarry += __shfl_sync(0xffffffffu, suf, 0);
    float dalpha_e = Gv[c] * Tv[c] * ev[c] - R * (ev[c] / uv[c]);   // dL/dalpha * e
    float dsig = onv[c] ? dalpha_e * distv[c] : 0.f;
    if (valid) {
      float4 o;
      o.x = wv[c] * gr * q[c].x * (1.f - q[c].x);
      o.y = wv[c] * gg * q[c].y * (1.f - q[c].y);
      o.z = wv[c] * gb * q[c].z * (1.f - q[c].z);
      o.w = dsig;
      if (raw_ch == 4) {
        reinterpret_cast<float4*>(outr)[i] = o;
      } else {                                   // channels the reference never reads get a zero gradient
        float* p = outr + (size_t)i * raw_ch;
        p[0] = o.x; p[1] = o.y; p[2] = o.z; p[3] = o.w;
        for (int k = 4; k < raw_ch; ++k) p[k] = 0.f;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// a9/a10/a11  hierarchical resampling                         ray.py:96-153, nerf/run.py:396-400,416
// One warp per ray.  smem per warp: cdf[M] | bins[M] | sort buffer[P] (P = pow2 >= S + Ni).
// ---------------------------------------------------------------------------------------------
// IEEE division that keeps a zero numerator off __fdiv_rn's slow path (a ~100-instruction subroutine the whole warp
// executes when ANY lane has a zero / denormal operand: it was half of the resample kernel's instructions)
__device__ __forceinline__ float div_rn_z(float a, float b) {
  const bool z = (a == 0.f);
  const float q = __fdiv_rn(z ? 1.f : a, b);
  return z ? a : q;
}

__device__ __forceinline__ int upper_bound_smem(const float* a, int n, float v) {
  // number of elements <= v  == torch.searchsorted(right=True)
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound_smem(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// build cdf[0..M-1] in smem from weights w[0..M-2]  (ray.py:111-114); all lanes participate
__device__ void warp_build_cdf(const float* __restrict__ w, int M, float* cdf, int lane) {
  float part = 0.f;
  for (int i = lane; i < M - 1; i += 32) part += __fadd_rn(__ldg(w + i), 1e-5f);
  // torch.sum order differs (pairwise); agreement is to rounding, hence "bit-exact given identical CDFs"
  float tot = warp_sum(part);
  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < M - 1; base += 32) {
    int i = base + lane;
    float p = (i < M - 1) ? div_rn_z(__fadd_rn(__ldg(w + i), 1e-5f), tot) : 0.f;
    float inc = warp_scan_add(p, lane);
    if (i < M - 1) cdf[i + 1] = carry + inc;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
}

// The same cdf in the order the reference's CPU ops use, for the check mode and the "given identical weights" parity
// test (tests/test_gpu_kernels.py): ray.py:112 torch.sum over the last dim = ATen's vectorised inner sum (SumKernel.cpp:
// `lanes`-wide vector partial sums, four interleaved accumulators when there are >= 4 vectors, scalar tail first, then
// the lanes of the partial sum in order; lanes = 16 on AVX512 hosts, 8 on AVX2), IEEE division, and ray.py:113
// torch.cumsum = sequential accumulation in double rounded to float at every step.  One lane does the work.
__device__ void warp_build_cdf_ref(const float* __restrict__ w, int M, float* cdf, int lane, int lanes) {
  if (lane == 0) {
    const int n = M - 1;
    const int nvec = n / lanes, nilp = nvec / 4;
    float tot = 0.f;
    for (int k = nvec * lanes; k < n; ++k) tot = __fadd_rn(tot, __fadd_rn(w[k], 1e-5f));       // scalar tail
    for (int l = 0; l < lanes; ++l) {
      float p[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < nilp; ++i)
        for (int k = 0; k < 4; ++k) p[k] = __fadd_rn(p[k], __fadd_rn(w[(4 * i + k) * lanes + l], 1e-5f));
      for (int i = nilp * 4; i < nvec; ++i) p[0] = __fadd_rn(p[0], __fadd_rn(w[i * lanes + l], 1e-5f));
      p[0] = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[2]), p[3]);
      tot = __fadd_rn(tot, p[0]);
    }
    double acc = 0.0;
    cdf[0] = 0.f;
    for (int i = 0; i < n; ++i) {
      acc += (double)div_rn_z(__fadd_rn(w[i], 1e-5f), tot);
      cdf[i + 1] = (float)acc;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int M, float u, int* ind_out) {
  int ind = upper_bound_smem(cdf, M, u);                       // ray.py:136
  int below = max(0, ind - 1), above = min(M - 1, ind);        // ray.py:137-138
  float cb = cdf[below], ca = cdf[above];
  float bb = bins[below], ba = bins[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < 1e-5f) denom = 1.f;                              // ray.py:149
  float t = div_rn_z(__fsub_rn(u, cb), denom);
  *ind_out = ind;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));       // ray.py:151
}

__device__ __forceinline__ float det_u(int j, int Ns) { return linspace01(j, Ns); }

// Generic sample_pdf(bins[N,M], weights[N,M-1] | cdf[N,M]) -> samples[N,Ns] (+ inds int64)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, const float* __restrict__ cdf_in,
                  const float* __restrict__ u_in, int det, int64_t N, int M, int Ns,
                  float* __restrict__ samples, int64_t* __restrict__ inds) {
  extern __shared__ float smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= N) return;
  float* cdf = smem + (size_t)warp * 2 * M;
  float* bs = cdf + M;
  for (int i = lane; i < M; i += 32) bs[i] = __ldg(bins + r * M + i);
  if (cdf_in) {
    for (int i = lane; i < M; i += 32) cdf[i] = __ldg(cdf_in + r * M + i);
    __syncwarp();
  } else {
    warp_build_cdf(weights + r * (M - 1), M, cdf, lane);
  }
  __syncwarp();
  for (int j = lane; j < Ns; j += 32) {
    float u = u_in ? __ldg(u_in + r * Ns + j) : det_u(j, Ns);
    int ind;
    float s = invert_cdf(cdf, bs, M, u, &ind);
    samples[r * Ns + j] = s;
    if (inds) inds[r * Ns + j] = (int64_t)ind;
  }
  (void)det;
}

__device__ void warp_bitonic_sort(float* a, int P, int lane) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < P / 2; t += 32) {
        int i = 2 * t - (t & (j - 1));        // index with bit j clear
        int l = i + j;
        bool up = ((i & k) == 0);
        float x = a[i], y = a[l];
        if ((x > y) == up) { a[i] = y; a[l] = x; }
      }
      __syncwarp();
    }
  }
}

// Fused: z_mid bins + sample_pdf on weights[:,1:-1] + sort(cat(z_vals, z_samples)) + z_std.
// The inverse cdf is monotone, so instead of sorting S+Ni values (nerf/run.py:400) the kernel sorts only the Ni
// uniforms (nothing to sort in the deterministic mode), draws the samples in ascending order and MERGES them with
// the already sorted z_vals by rank (two binary searches per element).  Rounding can invert neighbouring
// samples across a bin boundary by one ulp: two odd-even passes repair that, and a warp vote falls back to a
// full bitonic sort of the samples if anything is still out of order, so z_fine is exactly sort(cat(.)).
// smem per warp: cdf[M] | bins[M] | z[S] | samples[NiP] | merged[S+Ni]   (M = S-1, NiP = pow2 >= Ni)
__device__ __forceinline__ void smem_cswap(float* a, int i, int l) {
  float x = a[i], y = a[l];
  if (x > y) { a[i] = y; a[l] = x; }
}

// One ray on one warp, any S / Ni (the exact algorithm; also the fallback of the specialised kernel below).
// row_smem: cdf[M] | bins[M] | z[S] | samples[NiP] | merged[S+Ni]
struct ResampleCheck {      // test / check-mode extras (all optional; the production launches pass an all-zero struct)
  const float* cdf_in;      // [N, S-1]: use this cdf instead of building it
  int64_t* inds_out;        // [N, Ni]: torch.searchsorted(cdf, u, right=True) per sample, in ascending-u order
  float* cdf_out;           // [N, S-1]: the cdf the samples were drawn from
  int ref_lanes;            // 0: warp-scan cdf; 8 / 16: the reference's CPU summation order (warp_build_cdf_ref)
};

__device__ void resample_row_generic(const float* __restrict__ z_vals, const float* __restrict__ weights,
                                     const float* __restrict__ u_in, int64_t r, int S, int Ni, int NiP,
                                     float* __restrict__ z_samples, float* __restrict__ z_fine, float* __restrict__ z_std,
                                     float* row_smem, int lane, const ResampleCheck ck = ResampleCheck{nullptr, nullptr, nullptr, 0}) {
  const int M = S - 1;                               // bins = z_mid (S-1), weights[1:-1] (S-2)
  float* cdf = row_smem;
  float* bs = cdf + M;
  float* zs = bs + M;
  float* sm = zs + S;
  float* outb = sm + NiP;
  const float* zr = z_vals + r * S;
  for (int i = lane; i < S; i += 32) {
    float zi = __ldg(zr + i);
    zs[i] = zi;
    if (i < M) bs[i] = __fmul_rn(0.5f, __fadd_rn(__ldg(zr + i + 1), zi));    // run.py:396
  }
  const float inf = __int_as_float(0x7f800000);
  if (u_in) {
    for (int j = lane; j < NiP; j += 32) sm[j] = (j < Ni) ? __ldg(u_in + r * Ni + j) : inf;
  } else {
    for (int j = Ni + lane; j < NiP; j += 32) sm[j] = inf;
  }
  if (ck.cdf_in) {
    for (int i = lane; i < M; i += 32) cdf[i] = __ldg(ck.cdf_in + r * M + i);
  } else if (ck.ref_lanes) {
    warp_build_cdf_ref(weights + r * S + 1, M, cdf, lane, ck.ref_lanes);
  } else {
    warp_build_cdf(weights + r * S + 1, M, cdf, lane);
  }
  __syncwarp();
  if (ck.cdf_out)
    for (int i = lane; i < M; i += 32) ck.cdf_out[r * M + i] = cdf[i];
  if (u_in) warp_bitonic_sort(sm, NiP, lane);        // ascending uniforms -> ascending samples
  float s1 = 0.f;
  for (int j = lane; j < Ni; j += 32) {
    float u = u_in ? sm[j] : det_u(j, Ni);
    int ind;
    float sv = invert_cdf(cdf, bs, M, u, &ind);
    sm[j] = sv;
    s1 += sv;
    if (ck.inds_out) ck.inds_out[r * Ni + j] = (int64_t)ind;
  }
  __syncwarp();
  // repair one-ulp inversions between neighbours, then verify
  for (int t = lane; 2 * t + 1 < Ni; t += 32) smem_cswap(sm, 2 * t, 2 * t + 1);
  __syncwarp();
  for (int t = lane; 2 * t + 2 < Ni; t += 32) smem_cswap(sm, 2 * t + 1, 2 * t + 2);
  __syncwarp();
  bool ok = true;
  for (int j = lane; j + 1 < Ni; j += 32) ok = ok && (sm[j] <= sm[j + 1]);
  if (!__all_sync(0xffffffffu, ok)) warp_bitonic_sort(sm, NiP, lane);
  // population std (torch.std unbiased=False): two-pass for accuracy
  float mean = warp_sum(s1) / (float)Ni;
  float s2 = 0.f;
  for (int j = lane; j < Ni; j += 32) { float d = sm[j] - mean; s2 += d * d; }
  s2 = warp_sum(s2);
  if (lane == 0 && z_std) z_std[r] = sqrtf(s2 / (float)Ni);
  // merge by rank: ties put the z_vals element first
  for (int j = lane; j < Ni; j += 32) {
    float sv = sm[j];
    outb[j + upper_bound_smem(zs, S, sv)] = sv;      // + number of z <= s
    if (z_samples) z_samples[r * Ni + j] = sv;
  }
  for (int i = lane; i < S; i += 32) {
    float zv = zs[i];
    outb[i + lower_bound_smem(sm, Ni, zv)] = zv;      // + number of samples < z
  }
  __syncwarp();
  for (int i = lane; i < S + Ni; i += 32) z_fine[r * (S + Ni) + i] = outb[i];
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
resample_kernel(const float* __restrict__ z_vals, const float* __restrict__ weights, const float* __restrict__ u_in,
                int64_t N, int S, int Ni, int NiP, float* __restrict__ z_samples, float* __restrict__ z_fine,
                float* __restrict__ z_std, const ResampleCheck ck) {
  extern __shared__ float smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= N) return;
  resample_row_generic(z_vals, weights, u_in, r, S, Ni, NiP, z_samples, z_fine, z_std,
                       smem + (size_t)warp * (2 * (S - 1) + S + NiP + S + Ni), lane, ck);
}

// Specialisation for the reference configs' shape (64 coarse samples, 128 importance samples): everything a lane owns
// stays in registers in a BLOCKED layout (z / weights 2 per lane, samples 4 per lane, merged output 6 per lane), the
// inverse cdf walks forward from one binary search per lane (the uniforms are ascending), a sample's rank among the
// z_vals follows from its bin (it lies between two z_mids, so only z[below+1] has to be compared), and the z_vals
// drop into the holes the scattered samples leave (one prefix count).  ~4x fewer instructions than the generic
// kernel.  Every step is verified with a warp vote (samples ascending, merged row ascending); a ray that fails -
// rounding across a bin edge, degenerate spacing - is redone by the exact generic routine.
__device__ __forceinline__ void cswap2(float& a, float& b, int& ia, int& ib) {
  if (a > b) { float t = a; a = b; b = t; int ti = ia; ia = ib; ib = ti; }
}

template <bool RANDOM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
resample64_kernel(const float* __restrict__ z_vals, const float* __restrict__ weights, const float* __restrict__ u_in,
                  int64_t N, float* __restrict__ z_samples, float* __restrict__ z_fine, float* __restrict__ z_std) {
  constexpr int S = 64, M = 63, Ni = 128;
  constexpr int ROW = 2 * M + S + Ni + S + Ni;         // the generic routine's footprint (510 floats)
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= N) return;
  float* row = smem + (size_t)warp * ROW;
  float* cdf = row;            // [64] (63 used)
  float* bs = row + 64;        // [64] (63 used)
  float* zs = row + 128;       // [64]
  float* outb = row + 192;     // [192]
  const unsigned full = 0xffffffffu;

  const float2 z2 = __ldg(reinterpret_cast<const float2*>(z_vals + r * S) + lane);
  const float2 w2 = __ldg(reinterpret_cast<const float2*>(weights + r * S) + lane);
  const float znext = __shfl_down_sync(full, z2.x, 1);
  zs[2 * lane] = z2.x; zs[2 * lane + 1] = z2.y;
  bs[2 * lane] = __fmul_rn(0.5f, __fadd_rn(z2.y, z2.x));                         // run.py:396
  if (lane < 31) bs[2 * lane + 1] = __fmul_rn(0.5f, __fadd_rn(znext, z2.y));
  // pdf over weights[1..62] (ray.py:111-114): lane l owns p[2l-1] = w[2l] and p[2l] = w[2l+1]
  float pa = (lane >= 1) ? __fadd_rn(w2.x, 1e-5f) : 0.f;
  float pb = (lane <= 30) ? __fadd_rn(w2.y, 1e-5f) : 0.f;
  const float tot = warp_sum(pa + pb);
  pa = div_rn_z(pa, tot); pb = div_rn_z(pb, tot);
  const float inc = warp_scan_add(pa + pb, lane);
  float ex = __shfl_up_sync(full, inc, 1);
  if (lane == 0) ex = 0.f;
  cdf[2 * lane] = (lane == 0) ? 0.f : fminf(ex + pa, inc);                       // non-decreasing by construction
  if (lane <= 30) cdf[2 * lane + 1] = inc;
  __syncwarp();

  float u[4];
  if (RANDOM) {
    const float4 u4 = __ldg(reinterpret_cast<const float4*>(u_in + r * Ni) + lane);
    u[0] = u4.x; u[1] = u4.y; u[2] = u4.z; u[3] = u4.w;
    // bitonic sort of the 128 uniforms, element e = 4 lane + i: strides < 4 inside the lane, the rest by shuffle
#pragma unroll
    for (int k = 2; k <= Ni; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        if (j >= 4) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = 4 * lane + i;
            const float o = __shfl_xor_sync(full, u[i], j >> 2);
            const bool up = (e & k) == 0, low = (e & j) == 0;
            u[i] = (low == up) ? fminf(u[i], o) : fmaxf(u[i], o);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if ((i & j) == 0) {
              const int e = 4 * lane + i;
              const bool up = (e & k) == 0;
              const float x = u[i], y = u[i | j];
              if ((x > y) == up) { u[i] = y; u[i | j] = x; }
            }
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = det_u(4 * lane + i, Ni);
  }

  // inverse cdf: one binary search, then walk forward (ray.py:136-151)
  float sv[4];
  int bl[4];
  int ind = upper_bound_smem(cdf, M, u[0]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i > 0) while (ind < M && cdf[ind] <= u[i]) ++ind;
    const int below = max(0, ind - 1), above = min(M - 1, ind);
    const float cb = cdf[below], ca = cdf[above], bb = bs[below], ba = bs[above];
    float denom = __fsub_rn(ca, cb);
    if (denom < 1e-5f) denom = 1.f;
    const float t = div_rn_z(__fsub_rn(u[i], cb), denom);
    sv[i] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
    bl[i] = below;
  }
  // one-ulp inversions across a bin edge: two odd-even passes (as the generic routine), then verify
  cswap2(sv[0], sv[1], bl[0], bl[1]);
  cswap2(sv[2], sv[3], bl[2], bl[3]);
  cswap2(sv[1], sv[2], bl[1], bl[2]);
  {
    const float n0 = __shfl_down_sync(full, sv[0], 1), p3 = __shfl_up_sync(full, sv[3], 1);
    const int nb0 = __shfl_down_sync(full, bl[0], 1), pb3 = __shfl_up_sync(full, bl[3], 1);
    const bool sw_hi = lane < 31 && sv[3] > n0, sw_lo = lane > 0 && p3 > sv[0];
    if (sw_hi) { sv[3] = n0; bl[3] = nb0; }
    if (sw_lo) { sv[0] = p3; bl[0] = pb3; }
  }
  bool ok = sv[0] <= sv[1] && sv[1] <= sv[2] && sv[2] <= sv[3];
  {
    const float n0 = __shfl_down_sync(full, sv[0], 1);
    ok = ok && (lane == 31 || sv[3] <= n0);
  }
  // rank among the z_vals: the sample lies in [z_mid[below], z_mid[below+1]], so only z[below+1] can tie or precede it
  int pos[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rank = bl[i] + 1 + (sv[i] >= zs[bl[i] + 1] ? 1 : 0);
    pos[i] = 4 * lane + i + rank;
    ok = ok && sv[i] >= zs[bl[i]] && (bl[i] + 2 >= S || sv[i] < zs[bl[i] + 2]);
  }
  const int SENT = 0x7fc00123;
#pragma unroll
  for (int j = 0; j < 6; ++j) outb[lane + 32 * j] = __int_as_float(SENT);
  __syncwarp();
  if (__all_sync(full, ok)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) outb[pos[i]] = sv[i];
    __syncwarp();
    // the z_vals fill the holes in order: hole ordinal = position - samples before it
    float v[6];
    int ns = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) { v[j] = outb[6 * lane + j]; ns += (__float_as_int(v[j]) != SENT); }
    int before = ns;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(full, before, o); if (lane >= o) before += t; }
    before -= ns;                                   // samples in lower lanes
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (__float_as_int(v[j]) == SENT) v[j] = zs[6 * lane + j - before]; else ++before;
    }
    bool ok2 = v[0] <= v[1] && v[1] <= v[2] && v[2] <= v[3] && v[3] <= v[4] && v[4] <= v[5];
    const float nv = __shfl_down_sync(full, v[0], 1);
    ok2 = ok2 && (lane == 31 || v[5] <= nv);
    if (__all_sync(full, ok2)) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 6; ++j) outb[6 * lane + j] = v[j];
      // population std of the samples (run.py:416), two-pass
      const float mean = warp_sum((sv[0] + sv[1]) + (sv[2] + sv[3])) / (float)Ni;
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float d = sv[i] - mean; s2 += d * d; }
      s2 = warp_sum(s2);
      if (lane == 0 && z_std) z_std[r] = sqrtf(s2 / (float)Ni);
      if (z_samples) reinterpret_cast<float4*>(z_samples + r * Ni)[lane] = make_float4(sv[0], sv[1], sv[2], sv[3]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 6; ++j) z_fine[r * (S + Ni) + lane + 32 * j] = outb[lane + 32 * j];
      return;
    }
  }
  __syncwarp();
  resample_row_generic(z_vals, weights, RANDOM ? u_in : nullptr, r, S, Ni, Ni, z_samples, z_fine, z_std, row, lane);
}

// Second specialisation of the 64+128 shape: EIGHT lanes per ray (four rays per warp), so that every shuffle scan,
// vote and reduction serves four rays and has three steps instead of five.  The kernel is bound by shared-memory
// wavefronts (LSU data pipe ~88 % busy) as much as by instruction issue, so the layout is chosen for conflict-free access:
//   * a lane owns z / weights 8g..8g+7 while the cdf is built, but SAMPLES g, g+8, g+16, ... : at any instruction the
//     eight lanes of a ray look up neighbouring uniforms, i.e. the same or adjacent cdf entries (broadcast or distinct
//     banks), and the four rays of a warp sit 8 banks apart (row stride = 8 mod 32);
//   * the inverse cdf is one BRANCHLESS 6-probe search per sample: the three upper levels compare registers, the three
//     lower ones read one small table per level (cdf[8m+3], cdf[4m+1], cdf[2m]) so that two rays probing neighbouring
//     entries are a few banks apart instead of a multiple of the row skew;
//   * what a sample needs after the search sits in ONE 16-byte record per bin {cdf_b, z_mid_b, 1/denom, z_mid_a -
//     z_mid_b}, stored at k ^ (k >> 3) so that building the records is conflict-free too; 1/denom is the MUFU
//     reciprocal (1 ulp) - t in [0,1] scales a bin width (~0.06), so a sample moves by < 2e-8, far below one ulp of z;
//     the pdf is normalised by the correctly rounded reciprocal of the sum (the sum's own rounding already differs from
//     torch.sum by its order);
//   * the uniforms of the random mode are sorted in registers (16 per lane, all-ascending bitonic network) and
//     transposed to the interleaved ownership through shared memory; deterministic uniforms are computed;
//   * samples scatter to their ranks in a sentinel-filled row, each lane reads 24 consecutive slots back, fills the
//     holes with the z_vals in order (read from a 9-strided copy, conflict-free) and writes them back, so that the row
//     leaves for HBM in 128-byte runs per ray (two merges without the sentinel row - a rank loop, a shared-memory
//     histogram of the ranks - were measured and were not faster: profiles/r1h_ncu_resample64q.md).
// Exactness is by construction, guarded by per-ray checks made while the records are built: z_mid[k] < z[k+1]
// strictly, z_mid_b + (z_mid_a - z_mid_b) <= z_mid_a, cdf non-decreasing, every u >= 0.  With those, t clamped to 1 and
// ascending uniforms, the samples are ascending and a sample's rank among the z_vals is below + 1 + (s >= z[below+1]).
// A ray that fails a check (equal or one-ulp-apart z_vals, NaNs) is redone by resample_row_generic.
__device__ __forceinline__ float lds_f32(unsigned a) {
  float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v;
}
__device__ __forceinline__ float4 lds_f32x4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float rcp_approx(float x) {         // MUFU.RCP, 1 ulp; callers keep x in [1e-5, 1]
  float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}

__device__ __forceinline__ void sort2(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo; b = hi;
}

// rays the eight-lane kernel handed to the generic routine since the last reset (diagnostic: swnerf_resample_fallbacks)
__device__ unsigned long long g_resample_fallbacks = 0ull;

// floats per ray: cdf level tables[64] | z[64] | z 9-strided[72] | rec[64 x 4] (the 136-float transpose buffer of the random mode
// and, later, the merged row outb[192] alias rec).  456 = 8 mod 32: the four rays of a warp start 8 banks apart.
constexpr int kQRow = 456;
constexpr int kQWarps = 4;                 // warps per block (16 rays): 29 KB of shared memory; 5-6 blocks per SM (registers)

// CHECK adds the test entry's cdf-in / inds-out / cdf-out (swnerf_resample_check); the production instances
// (CHECK = false) carry none of it.
// NI = 128 (configs/lego.txt and every D-NeRF / MultiRes config) or 64 (the LLFF configs and hotdog / materials:
// nerf/configs/fern.txt): NPL = NI / 8 samples per lane, a merged row of 64 + NI floats = NF float4 per lane.
template <bool RANDOM, bool CHECK = false, int NI = 128>
__global__ void __launch_bounds__(kQWarps * 32, 5)
resample64q_kernel(const float* __restrict__ z_vals, const float* __restrict__ weights, const float* __restrict__ u_in,
                   int64_t N, float* __restrict__ z_samples, float* __restrict__ z_fine, float* __restrict__ z_std,
                   const ResampleCheck ck = ResampleCheck{nullptr, nullptr, nullptr, 0}) {
  constexpr int S = 64, Ni = NI;
  constexpr int NPL = NI / 8;                // samples per lane
  constexpr int NF = (64 + NI) / 32;         // float4 of the merged row per lane (6 or 4)
  static_assert(NI == 128 || NI == 64, "eight-lane resample kernel: 64 + 128 or 64 + 64 samples");
  extern __shared__ __align__(16) float smem[];
  const unsigned full = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 3, g = lane & 7;
  const int64_t r0 = ((int64_t)blockIdx.x * kQWarps + warp) * 4;
  if (r0 >= N) return;                                         // whole warp out of range
  const int64_t r = r0 + sub;
  const bool valid = r < N;
  const int64_t rr = valid ? r : N - 1;                        // out-of-range groups shadow the last ray, store nothing
  float* wbase = smem + (size_t)warp * 4 * kQRow;
  float* row = wbase + sub * kQRow;
  float* cdf = row;                                            // [64]: lv4[8] | lv2[16] | lv1[32] (see below)
  float* zs = row + 64;                                        // [64]
  float* zsw = row + 128;                                      // z[k] at k + (k >> 3)
  float* recf = row + 200;                                     // record k at (k ^ (k >> 3)) * 4
  float* outb = recf;                                          // [192], after the last read of the records

  float u[NPL];
  if (RANDOM) {
    const float4* up = reinterpret_cast<const float4*>(u_in + rr * Ni) + (NPL / 4) * g;
#pragma unroll
    for (int j = 0; j < NPL / 4; ++j) {
      const float4 t = __ldg(up + j);
      u[4 * j] = t.x; u[4 * j + 1] = t.y; u[4 * j + 2] = t.z; u[4 * j + 3] = t.w;
    }
  }
  float ze[10], q[8];
  {
    const float4* zp = reinterpret_cast<const float4*>(z_vals + rr * S) + 2 * g;
    const float4* wp = reinterpret_cast<const float4*>(weights + rr * S) + 2 * g;
    const float4 a = __ldg(zp), b = __ldg(zp + 1), c = __ldg(wp), d = __ldg(wp + 1);
    ze[0] = a.x; ze[1] = a.y; ze[2] = a.z; ze[3] = a.w; ze[4] = b.x; ze[5] = b.y; ze[6] = b.z; ze[7] = b.w;
    q[0] = c.x; q[1] = c.y; q[2] = c.z; q[3] = c.w; q[4] = d.x; q[5] = d.y; q[6] = d.z; q[7] = d.w;
  }
  if (RANDOM) {
    // bitonic sort of the ray's NI uniforms, element e = NPL g + i, in the all-ascending form: each merge of size k
    // starts with the mirror step (e against e ^ (k-1)) and continues with the strides k/4 .. 1, and the lower
    // index always keeps the minimum, so comparators inside a lane have no run-time direction
#pragma unroll
    for (int k = 2; k <= Ni; k <<= 1) {
      if (k <= NPL) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          if ((i & (k >> 1)) == 0) sort2(u[i], u[i ^ (k - 1)]);
        }
      } else {
        const bool keep_min = (g & (k / (2 * NPL))) == 0;       // bit k/2 of e clear
        float o[NPL];
#pragma unroll
        for (int i = 0; i < NPL; ++i) o[i] = __shfl_xor_sync(full, u[NPL - 1 - i], k / NPL - 1);
#pragma unroll
        for (int i = 0; i < NPL; ++i) u[i] = keep_min ? fminf(u[i], o[i]) : fmaxf(u[i], o[i]);
      }
#pragma unroll
      for (int j = k >> 2; j > 0; j >>= 1) {
        if (j >= NPL) {
          const bool keep_min = (g & (j / NPL)) == 0;
#pragma unroll
          for (int i = 0; i < NPL; ++i) {
            const float o = __shfl_xor_sync(full, u[i], j / NPL);
            u[i] = keep_min ? fminf(u[i], o) : fmaxf(u[i], o);
          }
        } else {
#pragma unroll
          for (int i = 0; i < NPL; ++i) {
            if ((i & j) == 0) sort2(u[i], u[i | j]);
          }
        }
      }
    }
    // blocked (NPL g + i) -> interleaved (g + 8 i) ownership through shared memory, element e at e + e / NPL
#pragma unroll
    for (int i = 0; i < NPL; ++i) recf[(NPL + 1) * g + i] = u[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NPL; ++i) u[i] = recf[g + 8 * i + (NPL == 16 ? (i >> 1) : i)];
    __syncwarp();
  } else {
    // torch.linspace(0, 1, NI), symmetric evaluation (linspace01), sample j = g + 8 i
    const float step = 1.0f / (float)(Ni - 1), fg = (float)g;
#pragma unroll
    for (int i = 0; i < NPL / 2; ++i) u[i] = __fmul_rn(step, fg + (float)(8 * i));
#pragma unroll
    for (int i = NPL / 2; i < NPL; ++i) u[i] = __fsub_rn(1.0f, __fmul_rn(step, (float)(Ni - 1 - 8 * i) - fg));
  }

  bool ok = true;
  ze[8] = __shfl_down_sync(full, ze[0], 1);
  ze[9] = __shfl_down_sync(full, ze[1], 1);
  if (g == 7) { ze[8] = ze[7]; ze[9] = ze[7]; }
  // a lane's two float4 go out in swapped order in the upper half of the group (lanes g and g+4 share bank groups)
  const bool hi = g >= 4;
  {
    const float4 a = make_float4(ze[0], ze[1], ze[2], ze[3]), b = make_float4(ze[4], ze[5], ze[6], ze[7]);
    reinterpret_cast<float4*>(zs)[2 * g + (hi ? 1 : 0)] = hi ? b : a;
    reinterpret_cast<float4*>(zs)[2 * g + (hi ? 0 : 1)] = hi ? a : b;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) zsw[9 * g + i] = ze[i];
  // pdf over weights[1..62] (ray.py:111-112); with q[0] = q[63] = 0, cdf[k] = inclusive prefix of q at k (ray.py:113-114)
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __fadd_rn(q[i], 1e-5f);
  if (g == 0) q[0] = 0.f;
  if (g == 7) q[7] = 0.f;
  float tot = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
  tot += __shfl_xor_sync(full, tot, 1);
  tot += __shfl_xor_sync(full, tot, 2);
  tot += __shfl_xor_sync(full, tot, 4);
  const float rtot = __frcp_rn(tot);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __fmul_rn(q[i], rtot);
#pragma unroll
  for (int i = 1; i < 8; ++i) q[i] = __fadd_rn(q[i - 1], q[i]);          // local inclusive prefix, non-decreasing
  float inc = q[7];
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const float t = __shfl_up_sync(full, inc, o, 8);
    if (g >= o) inc += t;
  }
  float ex = __shfl_up_sync(full, inc, 1, 8);
  if (g == 0) ex = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __fadd_rn(ex, q[i]);
  if (CHECK) {
    if (ck.cdf_in) {                                           // given cdf[0..62]; entry 63 repeats entry 62 (w[63] adds 0)
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = __ldg(ck.cdf_in + rr * 63 + min(8 * g + i, 62));
    }
    if (ck.cdf_out && valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (8 * g + i < 63) ck.cdf_out[r * 63 + 8 * g + i] = q[i];
    }
  }
  // the cdf is stored in search order, one small table per probe level (the three upper levels live in registers):
  // lv4[m] = cdf[8m+3], lv2[m] = cdf[4m+1], lv1[m] = cdf[2m].  Rays that probe neighbouring entries of a level are then
  // a few banks apart, never a multiple of the 8-bank row skew (the linear array made two rays of a warp collide
  // whenever their indices differed by 8)
  cdf[g] = q[3];
  reinterpret_cast<float2*>(cdf + 8)[g] = make_float2(q[1], q[5]);
  reinterpret_cast<float4*>(cdf + 24)[g] = make_float4(q[0], q[2], q[4], q[6]);
  const float c7 = __shfl_sync(full, q[7], 0, 8), c15 = __shfl_sync(full, q[7], 1, 8),
              c23 = __shfl_sync(full, q[7], 2, 8), c31 = __shfl_sync(full, q[7], 3, 8),
              c39 = __shfl_sync(full, q[7], 4, 8), c47 = __shfl_sync(full, q[7], 5, 8),
              c55 = __shfl_sync(full, q[7], 6, 8);
  {
    float cnext = __shfl_down_sync(full, q[0], 1);             // cdf[8g + 8]
    if (g == 7) cnext = q[7];
    ok = ok && (q[7] <= cnext);                                // non-decreasing across lanes (inside: by construction)
    float m[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = __fmul_rn(0.5f, __fadd_rn(ze[i + 1], ze[i]));   // z_mid, run.py:396
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = 8 * g + i;                                 // bin k: below = k, above = k + 1 (ray.py:137-151)
      float denom = __fsub_rn(i < 7 ? q[i + 1] : cnext, q[i]);
      if (denom < 1e-5f) denom = 1.f;                          // ray.py:149
      float rden = rcp_approx(denom);
      float dbin = __fsub_rn(m[i + 1], m[i]);
      const bool last = (g == 7) && (i >= 6);                  // k = 62: below = above (u >= cdf[62]); k = 63 unused
      if (last) { rden = 1.f; dbin = 0.f; }
      ok = ok && ((g == 7 && i == 7) || m[i] < ze[i + 1]);
      ok = ok && (last || __fadd_rn(m[i], dbin) <= m[i + 1]);
      reinterpret_cast<float4*>(recf)[k ^ g] = make_float4(q[i], m[i], rden, dbin);
    }
  }
  __syncwarp();

  // inverse cdf (ray.py:136-151) and the sample's rank among the z_vals
  const unsigned cdf_sa = (unsigned)__cvta_generic_to_shared(cdf);
  const unsigned zs_sa = (unsigned)__cvta_generic_to_shared(zs);
  const unsigned rec_sa = (unsigned)__cvta_generic_to_shared(recf);
  float sv[NPL];
  int pos[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const float uu = u[i];
    // p = number of cdf entries <= u (searchsorted right=True): three probes on registers, three on the level tables
    const bool h1 = c31 <= uu;
    const bool h2 = (h1 ? c47 : c15) <= uu;
    const bool h3 = (h1 ? (h2 ? c55 : c39) : (h2 ? c23 : c7)) <= uu;
    unsigned pc = (h1 ? 32u : 0u) + (h2 ? 16u : 0u) + (h3 ? 8u : 0u);
    pc += (lds_f32(cdf_sa + (pc >> 1)) <= uu) ? 4u : 0u;       // lv4[pc / 8]      = cdf[pc + 3]
    pc += (lds_f32(cdf_sa + 32 + pc) <= uu) ? 2u : 0u;         // lv2[pc / 4]      = cdf[pc + 1]
    pc += (lds_f32(cdf_sa + 96 + 2 * pc) <= uu) ? 1u : 0u;     // lv1[pc / 2]      = cdf[pc]
    if (CHECK) {                                               // ray.py:136 over the 63 real entries
      if (ck.inds_out && valid) ck.inds_out[r * Ni + g + 8 * i] = (int64_t)min(pc, 63u);
    }
    const unsigned pb = 4 * pc;
    ok = ok && (pb != 0);                                      // u >= cdf[0] = 0
    const unsigned x = max(pb, 4u) - 4u;                       // 4 x below (ray.py:137)
    const float z1 = lds_f32(zs_sa + x + 4);                   // z[below + 1]
    const float4 rb = lds_f32x4(rec_sa + ((x << 2) ^ ((x >> 1) & 0x70u)));
    const float t = fminf(__fmul_rn(__fsub_rn(uu, rb.x), rb.z), 1.f);
    const float s = __fadd_rn(rb.y, __fmul_rn(t, rb.w));       // ray.py:151
    sv[i] = s;
    pos[i] = g + 8 * i + (int)(x >> 2) + 1 + (s >= z1 ? 1 : 0);
  }
  const unsigned okmask = __ballot_sync(full, ok);
  __syncwarp();                                                // every read of the records precedes the writes below
  const bool st_ok = valid && ((okmask >> (8 * sub)) & 0xffu) == 0xffu;

  // merged row: samples scatter to their ranks, the z_vals fill the holes in order
  const int SENT = 0x7fc00123;
  {
    const float sn = __int_as_float(SENT);
#pragma unroll
    for (int j = 0; j < NF; ++j) reinterpret_cast<float4*>(outb)[g + 8 * j] = make_float4(sn, sn, sn, sn);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NPL; ++i) outb[pos[i]] = sv[i];          // pos in [0, 64 + NI) whatever the data
  __syncwarp();
  // each lane takes 4 NF consecutive slots (NF float4 at NF g .. NF g + NF - 1).  Lanes whose first float4 falls into
  // the same 16-byte bank group (NF = 6: g and g + 4; NF = 4: g, g + 2, g + 4, g + 6) walk their float4 in rotated
  // order, step j reads float4 (j + rot) mod NF: conflict-free
  constexpr int NROT = NF == 6 ? 2 : 4;
  const int rot = NF == 6 ? (hi ? 1 : 0) : (g >> 1);
  float v[4 * NF];
  {
    float4 t[NF];
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      int sl = j + rot; sl = sl >= NF ? sl - NF : sl;
      t[j] = reinterpret_cast<const float4*>(outb)[NF * g + sl];
    }
#pragma unroll
    for (int j = 0; j < NF; ++j) {                             // float4 j of the lane was read at step (j - rot) mod NF
      float4 a = t[j];
#pragma unroll
      for (int rr_ = 1; rr_ < NROT; ++rr_) {
        const float4 b = t[(j - rr_ + NF) % NF];
        if (rot == rr_) a = b;
      }
      v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
    }
  }
  int ns = 0;
#pragma unroll
  for (int j = 0; j < 4 * NF; ++j) ns += (__float_as_int(v[j]) != SENT) ? 1 : 0;
  int before = ns;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const int t = __shfl_up_sync(full, before, o, 8);
    if (g >= o) before += t;
  }
  int zi = 4 * NF * g - (before - ns);                         // next z_val to place; stays inside the row on bad rays
#pragma unroll
  for (int j = 0; j < 4 * NF; ++j) {
    if (__float_as_int(v[j]) == SENT) { v[j] = zsw[zi + (zi >> 3)]; ++zi; }
  }
  // population std of the samples (run.py:416), two-pass
  float s1 = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) s1 += sv[i];
  s1 += __shfl_xor_sync(full, s1, 1);
  s1 += __shfl_xor_sync(full, s1, 2);
  s1 += __shfl_xor_sync(full, s1, 4);
  const float mean = s1 / (float)Ni;
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) { const float d = sv[i] - mean; s2 += d * d; }
  s2 += __shfl_xor_sync(full, s2, 1);
  s2 += __shfl_xor_sync(full, s2, 2);
  s2 += __shfl_xor_sync(full, s2, 4);
  // the row goes back to shared memory (same rotated order) so that the global stores are 128-byte runs per ray
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    int sl = j + rot; sl = sl >= NF ? sl - NF : sl;            // step j writes float4 (j + rot) mod NF of the lane
    float4 a = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
#pragma unroll
    for (int rr_ = 1; rr_ < NROT; ++rr_) {
      const int q4 = 4 * ((j + rr_) % NF);
      if (rot == rr_) a = make_float4(v[q4], v[q4 + 1], v[q4 + 2], v[q4 + 3]);
    }
    reinterpret_cast<float4*>(outb)[NF * g + sl] = a;
  }
  __syncwarp();
  if (st_ok) {
    if (g == 0 && z_std) z_std[r] = sqrtf(s2 / (float)Ni);
    if (z_samples) {
      float* zo = z_samples + r * Ni + g;                      // 32-B runs per ray and instruction
#pragma unroll
      for (int i = 0; i < NPL; ++i) zo[8 * i] = sv[i];
    }
    float4* fo = reinterpret_cast<float4*>(z_fine + r * (S + Ni));
#pragma unroll
    for (int j = 0; j < NF; ++j) fo[g + 8 * j] = reinterpret_cast<const float4*>(outb)[g + 8 * j];
  }
  if (okmask == full) return;
  // rays that failed a check: the exact generic routine, one ray at a time on the whole warp
  __syncwarp();
#pragma unroll 1
  for (int sb = 0; sb < 4; ++sb) {
    if (((okmask >> (8 * sb)) & 0xffu) == 0xffu || r0 + sb >= N) continue;
    if (lane == 0) atomicAdd(&g_resample_fallbacks, 1ull);
    resample_row_generic(z_vals, weights, RANDOM ? u_in : nullptr, r0 + sb, S, Ni, Ni, z_samples, z_fine, z_std,
                         wbase, lane, ck);
    __syncwarp();
  }
}


// ---------------------------------------------------------------------------------------------
// a13  torchsearchsorted-compatible batched search              searchsorted_cuda_kernel.cu:83-107
// ---------------------------------------------------------------------------------------------
__global__ void searchsorted_kernel(const float* __restrict__ a, const float* __restrict__ v, int64_t* __restrict__ out,
                                    int64_t nrow_res, int64_t nrow_a, int64_t nrow_v, int64_t ncol_a, int64_t ncol_v,
                                    int side_left) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrow_res * ncol_v) return;
  int64_t row = idx / ncol_v, col = idx - row * ncol_v;
  const float* ar = a + ((nrow_a == 1) ? 0 : row) * ncol_a;
  float val = v[((nrow_v == 1) ? 0 : row) * ncol_v + col];
  int64_t lo = 0, hi = ncol_a;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    float am = __ldg(ar + mid);
    bool go_right = side_left ? (am < val) : (am <= val);
    if (go_right) lo = mid + 1; else hi = mid;
  }
  out[idx] = lo;
}

static int g_resample_variant = 1;      // 64+128 shape: 1 = resample64q_kernel (default), 0 = resample64_kernel
static inline int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

}  // namespace swnerf

using namespace swnerf;

extern "C" {

int swnerf_stratified_z(const float* rays, int ray_stride, int near_col, const float* t_rand, float* z_vals,
                        int64_t n_rays, int n_samples, int lindisp, int perturb, void* stream) {
  SW_REQUIRE(rays && z_vals, "stratified_z: null pointer");
  SW_REQUIRE(n_samples >= 1 && n_rays >= 0, "stratified_z: bad sizes");
  SW_REQUIRE(!perturb || t_rand, "stratified_z: perturb requires t_rand");
  if (n_rays == 0) return SWNERF_OK;
  int64_t tot = n_rays * n_samples;
  stratified_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rays, ray_stride, near_col, t_rand, z_vals, n_rays, n_samples, lindisp, perturb);
  return check_launch("stratified_z");
}

int swnerf_embed_fwd(const float* x, float* y, int64_t rows, int dims, int L, void* stream) {
  SW_REQUIRE(x && y, "embed_fwd: null pointer");
  SW_REQUIRE(dims >= 1 && L >= 0 && L <= 32, "embed_fwd: bad dims/L");
  if (rows == 0) return SWNERF_OK;
  int64_t tot = rows * dims * (1 + 2 * L);
  embed_fwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, rows, dims, L);
  return check_launch("embed_fwd");
}

int swnerf_embed_bwd(const float* x, const float* dy, float* dx, int64_t rows, int dims, int L, void* stream) {
  SW_REQUIRE(x && dy && dx, "embed_bwd: null pointer");
  SW_REQUIRE(dims >= 1 && L >= 0 && L <= 32, "embed_bwd: bad dims/L");
  if (rows == 0) return SWNERF_OK;
  int64_t tot = rows * dims;
  embed_bwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, rows, dims, L);
  return check_launch("embed_bwd");
}

int swnerf_encode_points(const float* rays, int ray_stride, int view_col, const float* z_vals, float* out,
                         int64_t n_rays, int n_samples, int L_pos, int L_dir, int out_stride, void* stream) {
  SW_REQUIRE(rays && z_vals && out, "encode_points: null pointer");
  SW_REQUIRE(L_pos <= 32 && L_dir <= 32, "encode_points: L too large");
  if (n_rays == 0) return SWNERF_OK;
  int od = 3 * (1 + 2 * (L_pos < 0 ? 0 : L_pos)) + (view_col >= 0 ? 3 * (1 + 2 * (L_dir < 0 ? 0 : L_dir)) : 0);
  SW_REQUIRE(out_stride >= od, "encode_points: out_stride < embedded width");
  int64_t tot = n_rays * n_samples * od;
  encode_points_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rays, ray_stride, view_col, z_vals, out, n_rays, n_samples, L_pos, L_dir, out_stride);
  return check_launch("encode_points");
}

int swnerf_composite_fwd(const float* raw, int raw_ch, const float* z_vals, const float* rays, int ray_stride, int d_col,
                         const float* noise, int white_bkgd, int64_t n_rays, int n_samples, float* rgb_map,
                         float* disp_map, float* acc_map, float* weights, float* depth_map, void* stream) {
  SW_REQUIRE(raw && z_vals && rays && rgb_map && disp_map && acc_map && weights && depth_map,
             "composite_fwd: null pointer");
  SW_REQUIRE(raw_ch >= 4, "composite_fwd: raw needs at least 4 channels per sample (rgb, sigma), got %d", raw_ch);
  SW_REQUIRE(raw_ch != 4 || aligned16(raw), "composite_fwd: raw must be 16-byte aligned");
  SW_REQUIRE(n_samples >= 1, "composite_fwd: n_samples < 1");
  if (n_rays == 0) return SWNERF_OK;
  SW_REQUIRE(n_samples <= 1024, "composite_fwd: n_samples > 1024");
  unsigned blocks = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const int nc = (n_samples + 31) / 32;
#define SW_FWD(NC)                                                                                              \
  composite_fwd_kernel<NC><<<blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(                            \
      raw, raw_ch, z_vals, rays, ray_stride, d_col, noise, white_bkgd, n_rays, n_samples, rgb_map, disp_map, acc_map, \
      weights, depth_map)
  if (nc <= 2) SW_FWD(2); else if (nc <= 4) SW_FWD(4); else if (nc <= 6) SW_FWD(6); else if (nc <= 8) SW_FWD(8);
  else if (nc <= 16) SW_FWD(16); else SW_FWD(32);
#undef SW_FWD
  return check_launch("composite_fwd");
}

int swnerf_composite_bwd(const float* raw, int raw_ch, const float* z_vals, const float* rays, int ray_stride, int d_col,
                         const float* noise, int white_bkgd, int64_t n_rays, int n_samples, const float* g_rgb,
                         const float* g_disp, const float* g_acc, const float* g_weights, const float* g_depth,
                         const float* acc_map, const float* depth_map, float* d_raw, void* stream) {
  SW_REQUIRE(raw && z_vals && rays && d_raw, "composite_bwd: null pointer");
  SW_REQUIRE(raw_ch >= 4, "composite_bwd: raw needs at least 4 channels per sample (rgb, sigma), got %d", raw_ch);
  SW_REQUIRE(raw_ch != 4 || (aligned16(raw) && aligned16(d_raw)), "composite_bwd: raw/d_raw must be 16-byte aligned");
  SW_REQUIRE(!g_disp || (acc_map && depth_map), "composite_bwd: g_disp needs saved acc/depth maps");
  SW_REQUIRE(n_samples >= 1 && n_samples <= 1024, "composite_bwd: n_samples must be in 1..1024 (as the forward)");
  if (n_rays == 0) return SWNERF_OK;
  unsigned blocks = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const int nc = (n_samples + 31) / 32;
#define SW_BWD(NC)                                                                                              \
  composite_bwd_kernel<NC><<<blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(                            \
      raw, raw_ch, z_vals, rays, ray_stride, d_col, noise, white_bkgd, n_rays, n_samples, g_rgb, g_disp, g_acc, g_weights, \
      g_depth, acc_map, depth_map, d_raw)
  if (nc <= 2) SW_BWD(2); else if (nc <= 4) SW_BWD(4); else if (nc <= 6) SW_BWD(6); else if (nc <= 8) SW_BWD(8);
  else if (nc <= 16) SW_BWD(16); else SW_BWD(32);
#undef SW_BWD
  return check_launch("composite_bwd");
}

int swnerf_sample_pdf(const float* bins, const float* weights, const float* cdf, const float* u, int det,
                      int64_t n_rays, int n_bins, int n_samples, float* samples, int64_t* inds, void* stream) {
  SW_REQUIRE(bins && samples, "sample_pdf: null pointer");
  SW_REQUIRE((weights != nullptr) != (cdf != nullptr), "sample_pdf: give exactly one of weights / cdf");
  SW_REQUIRE(det || u, "sample_pdf: random mode needs caller-supplied u");
  SW_REQUIRE(n_bins >= 2 && n_bins <= 2048 && n_samples >= 1, "sample_pdf: bad sizes");
  if (n_rays == 0) return SWNERF_OK;
  size_t smem = (size_t)kWarpsPerBlock * 2 * n_bins * sizeof(float);
  unsigned blocks = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sample_pdf_kernel<<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      bins, weights, cdf, det ? nullptr : u, det, n_rays, n_bins, n_samples, samples, inds);
  return check_launch("sample_pdf");
}

int swnerf_set_resample_variant(int variant) {
  SW_REQUIRE(variant == 0 || variant == 1, "set_resample_variant: 0 (one warp per ray) or 1 (eight lanes per ray)");
  g_resample_variant = variant;
  return SWNERF_OK;
}

int swnerf_resample_fallbacks(unsigned long long* count, int reset, void* stream) {
  SW_REQUIRE(count, "resample_fallbacks: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(count, g_resample_fallbacks, sizeof(unsigned long long));
  if (e == cudaSuccess && reset) {
    const unsigned long long zero = 0ull;
    e = cudaMemcpyToSymbol(g_resample_fallbacks, &zero, sizeof(zero));
  }
  if (e != cudaSuccess) return set_err(SWNERF_ERR_CUDA, "resample_fallbacks: %s", cudaGetErrorString(e));
  return SWNERF_OK;
}

int swnerf_resample(const float* z_vals, const float* weights, const float* u, int det, int64_t n_rays,
                    int n_samples, int n_importance, float* z_samples, float* z_fine, float* z_std, void* stream) {
  SW_REQUIRE(z_vals && weights && z_fine, "resample: null pointer");
  SW_REQUIRE(det || u, "resample: random mode needs caller-supplied u");
  SW_REQUIRE(n_samples >= 3 && n_importance >= 1, "resample: need n_samples >= 3 and n_importance >= 1");
  int P = next_pow2(n_importance);
  SW_REQUIRE(P <= 1024 && n_samples <= 1024, "resample: n_samples / n_importance > 1024");
  if (n_rays == 0) return SWNERF_OK;
  size_t smem = (size_t)kWarpsPerBlock * (2 * (n_samples - 1) + n_samples + P + n_samples + n_importance) * sizeof(float);
  unsigned blocks = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const bool al = aligned16(z_vals) && aligned16(weights) && (det || aligned16(u)) && (!z_samples || aligned16(z_samples));
  if (n_samples == 64 && (n_importance == 128 || n_importance == 64) && al && aligned16(z_fine) && g_resample_variant != 0) {
    // eight lanes per ray, four rays per warp: the two shapes of the reference's configs (64 + 128, 64 + 64)
    const size_t qsmem = (size_t)kQWarps * 4 * kQRow * sizeof(float);
    if (once_per_device(ONCE_RESAMPLE64Q)) {
      cudaFuncSetAttribute(resample64q_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem);
      cudaFuncSetAttribute(resample64q_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem);
      cudaFuncSetAttribute(resample64q_kernel<false, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem);
      cudaFuncSetAttribute(resample64q_kernel<true, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem);
    }
    const unsigned qblocks = (unsigned)((n_rays + 4 * kQWarps - 1) / (4 * kQWarps));
    cudaStream_t qs = (cudaStream_t)stream;
    if (n_importance == 128) {
      if (det) resample64q_kernel<false><<<qblocks, kQWarps * 32, qsmem, qs>>>(z_vals, weights, nullptr, n_rays, z_samples, z_fine, z_std);
      else resample64q_kernel<true><<<qblocks, kQWarps * 32, qsmem, qs>>>(z_vals, weights, u, n_rays, z_samples, z_fine, z_std);
    } else {
      if (det) resample64q_kernel<false, false, 64><<<qblocks, kQWarps * 32, qsmem, qs>>>(z_vals, weights, nullptr, n_rays, z_samples, z_fine, z_std);
      else resample64q_kernel<true, false, 64><<<qblocks, kQWarps * 32, qsmem, qs>>>(z_vals, weights, u, n_rays, z_samples, z_fine, z_std);
    }
    return check_launch("resample");
  }
  if (n_samples == 64 && n_importance == 128 && al) {
    if (det) resample64_kernel<false><<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
                 z_vals, weights, nullptr, n_rays, z_samples, z_fine, z_std);
    else resample64_kernel<true><<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
             z_vals, weights, u, n_rays, z_samples, z_fine, z_std);
    return check_launch("resample");
  }
  resample_kernel<<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      z_vals, weights, det ? nullptr : u, n_rays, n_samples, n_importance, P, z_samples, z_fine, z_std,
      ResampleCheck{nullptr, nullptr, nullptr, 0});
  return check_launch("resample");
}

int swnerf_resample_check(const float* z_vals, const float* weights, const float* cdf_in, const float* u, int det,
                          int64_t n_rays, int n_samples, int n_importance, int variant, int ref_lanes, float* z_samples,
                          float* z_fine, float* z_std, int64_t* inds_out, float* cdf_out, void* stream) {
  SW_REQUIRE(z_vals && weights && z_fine, "resample_check: null pointer");
  SW_REQUIRE(det || u, "resample_check: random mode needs caller-supplied u");
  SW_REQUIRE(n_samples >= 3 && n_importance >= 1, "resample_check: need n_samples >= 3 and n_importance >= 1");
  SW_REQUIRE(variant == 0 || variant == 1, "resample_check: variant 0 (production kernel of the shape) or 1 (reference order)");
  SW_REQUIRE(variant == 0 || ref_lanes == 8 || ref_lanes == 16, "resample_check: ref_lanes must be 8 (AVX2 host) or 16 (AVX512)");
  int P = next_pow2(n_importance);
  SW_REQUIRE(P <= 1024 && n_samples <= 1024, "resample_check: n_samples / n_importance > 1024");
  if (n_rays == 0) return SWNERF_OK;
  const ResampleCheck ck{cdf_in, inds_out, cdf_out, variant == 1 ? ref_lanes : 0};
  cudaStream_t s = (cudaStream_t)stream;
  if (variant == 0 && n_samples == 64 && n_importance == 128) {
    SW_REQUIRE(aligned16(z_vals) && aligned16(weights) && (det || aligned16(u)) && (!z_samples || aligned16(z_samples)) &&
               aligned16(z_fine), "resample_check: the 64+128 kernel needs 16-byte aligned buffers");
    const size_t qsmem = (size_t)kQWarps * 4 * kQRow * sizeof(float);
    if (once_per_device(ONCE_RESAMPLE64Q_CHECK)) {
      cudaFuncSetAttribute(resample64q_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem);
      cudaFuncSetAttribute(resample64q_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem);
    }
    const unsigned qblocks = (unsigned)((n_rays + 4 * kQWarps - 1) / (4 * kQWarps));
    if (det) resample64q_kernel<false, true><<<qblocks, kQWarps * 32, qsmem, s>>>(z_vals, weights, nullptr, n_rays,
                                                                                 z_samples, z_fine, z_std, ck);
    else resample64q_kernel<true, true><<<qblocks, kQWarps * 32, qsmem, s>>>(z_vals, weights, u, n_rays, z_samples,
                                                                            z_fine, z_std, ck);
    return check_launch("resample_check");
  }
  size_t smem = (size_t)kWarpsPerBlock * (2 * (n_samples - 1) + n_samples + P + n_samples + n_importance) * sizeof(float);
  unsigned blocks = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  resample_kernel<<<blocks, kWarpsPerBlock * 32, smem, s>>>(z_vals, weights, det ? nullptr : u, n_rays, n_samples,
                                                           n_importance, P, z_samples, z_fine, z_std, ck);
  return check_launch("resample_check");
}

int swnerf_searchsorted(const float* a, const float* v, int64_t* out, int64_t nrow_a, int64_t nrow_v,
                        int64_t ncol_a, int64_t ncol_v, int side_left, void* stream) {
  SW_REQUIRE(a && v && out, "searchsorted: null pointer");
  SW_REQUIRE(nrow_a == nrow_v || nrow_a == 1 || nrow_v == 1,
             "searchsorted: `a` and `v` must have the same number of rows or one of them must have only one");
  int64_t rows = nrow_a > nrow_v ? nrow_a : nrow_v;
  int64_t tot = rows * ncol_v;
  if (tot == 0) return SWNERF_OK;
  searchsorted_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      a, v, out, rows, nrow_a, nrow_v, ncol_a, ncol_v, side_left);
  return check_launch("searchsorted");
}

}  // extern "C"
