"""CPU oracle for the SW-NeRF per-ray volumetric rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (sw-nerf_b200/) imports
this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may use it, and only as the checker.

It is a plain fp32 PyTorch restatement of the reference algorithm, run on the CPU by the tests (device-agnostic, like
the reference: bench.py's `gpu_eager_baseline` leg times it unchanged on the B200 - the reference's own deployment) (the
reference itself is fp32 eager PyTorch; its arithmetic lives in torch, which is
un-pinned by the reference - requirements.txt:9).  Each function cites the
reference file:line it restates.  Gradients come from autograd over these
functions, exactly as in the reference (nerf/run.py:699 loss.backward()).

Parity pin: tests/test_oracle_golden.py checks every function here against
golden vectors produced by importing and running the UNMODIFIED reference in the
build container (oracle/make_golden.py -> tests/golden/*.npz).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

F32 = torch.float32


# ----------------------------------------------------------------------------
# a4  positional encoding                       reference: embedder.py:12-59
# ----------------------------------------------------------------------------
def embed(x: torch.Tensor, L: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)].

    embedder.py:21-42: include_input first, then per frequency sin then cos,
    each applied to all input dims; bands 2**linspace(0, L-1, L) are exact
    powers of two.  L < 0 selects the identity (embedder.py:45-46, i == -1).
    """
    if L < 0:
        return x
    outs = [x]
    for k in range(L):
        f = float(2.0 ** k)
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, -1)


def embed_dim(L: int, d: int) -> int:
    return d if L < 0 else d * (1 + 2 * L)


# ----------------------------------------------------------------------------
# a6 / a6d  8x256 skip MLP with viewdir branch   reference: model.py:39-62, 273-296
# ----------------------------------------------------------------------------
def mlp_param_names(D: int = 8, use_viewdirs: bool = True, prefix: str = "") -> List[str]:
    names = []
    for i in range(D):
        names += [f"{prefix}pts_linears.{i}.weight", f"{prefix}pts_linears.{i}.bias"]
    names += [f"{prefix}views_linears.0.weight", f"{prefix}views_linears.0.bias"]
    if use_viewdirs:
        names += [f"{prefix}feature_linear.weight", f"{prefix}feature_linear.bias",
                  f"{prefix}alpha_linear.weight", f"{prefix}alpha_linear.bias",
                  f"{prefix}rgb_linear.weight", f"{prefix}rgb_linear.bias"]
    else:
        names += [f"{prefix}output_linear.weight", f"{prefix}output_linear.bias"]
    return names


def mlp_param_shapes(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4,
                     skips=(4,), use_viewdirs=True, prefix="") -> Dict[str, Tuple[int, ...]]:
    """Shapes of the state_dict of vallina_NeRF / NeRFOriginal (model.py:22-37, 243-269)."""
    sh = {}
    for i in range(D):
        if i == 0:
            k = input_ch
        elif (i - 1) in skips:
            k = W + input_ch
        else:
            k = W
        sh[f"{prefix}pts_linears.{i}.weight"] = (W, k)
        sh[f"{prefix}pts_linears.{i}.bias"] = (W,)
    sh[f"{prefix}views_linears.0.weight"] = (W // 2, input_ch_views + W)
    sh[f"{prefix}views_linears.0.bias"] = (W // 2,)
    if use_viewdirs:
        sh[f"{prefix}feature_linear.weight"] = (W, W)
        sh[f"{prefix}feature_linear.bias"] = (W,)
        sh[f"{prefix}alpha_linear.weight"] = (1, W)
        sh[f"{prefix}alpha_linear.bias"] = (1,)
        sh[f"{prefix}rgb_linear.weight"] = (3, W // 2)
        sh[f"{prefix}rgb_linear.bias"] = (3,)
    else:
        sh[f"{prefix}output_linear.weight"] = (output_ch, W)
        sh[f"{prefix}output_linear.bias"] = (output_ch,)
    return sh


def _lin(p, name, x):
    return x @ p[name + ".weight"].t() + p[name + ".bias"]


def mlp_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, input_ch: int, input_ch_views: int,
                D: int = 8, skips: Sequence[int] = (4,), use_viewdirs: bool = True,
                prefix: str = "") -> torch.Tensor:
    """model.py:39-62.  x = [pts_embedded | views_embedded] -> [rgb(3), sigma(1)] raw."""
    pts = x[..., :input_ch]
    views = x[..., input_ch:input_ch + input_ch_views]
    h = pts
    for i in range(D):
        h = torch.relu(_lin(p, f"{prefix}pts_linears.{i}", h))      # model.py:43-44
        if i in skips:
            h = torch.cat([pts, h], -1)                              # model.py:45-46 (input first)
    if use_viewdirs:
        alpha = _lin(p, f"{prefix}alpha_linear", h)                  # model.py:49
        feature = _lin(p, f"{prefix}feature_linear", h)              # model.py:50
        h = torch.cat([feature, views], -1)                          # model.py:51
        h = torch.relu(_lin(p, f"{prefix}views_linears.0", h))       # model.py:53-55
        rgb = _lin(p, f"{prefix}rgb_linear", h)                      # model.py:57
        return torch.cat([rgb, alpha], -1)                           # model.py:58
    return _lin(p, f"{prefix}output_linear", h)                      # model.py:60


# ----------------------------------------------------------------------------
# a7  DirectTemporalNeRF                           reference: model.py:93-151
# ----------------------------------------------------------------------------
def dnerf_param_shapes(D=8, W=256, input_ch=63, input_ch_views=27, input_ch_time=21,
                       skips=(4,), use_viewdirs=True) -> Dict[str, Tuple[int, ...]]:
    sh = mlp_param_shapes(D, W, input_ch, input_ch_views, 4, skips, use_viewdirs, prefix="_occ.")
    for i in range(D):                                                # model.py:113-126
        if i == 0:
            k = input_ch + input_ch_time
        elif (i - 1) in skips:
            k = W + input_ch
        else:
            k = W
        sh[f"_time.{i}.weight"] = (W, k)
        sh[f"_time.{i}.bias"] = (W,)
    sh["_time_out.weight"] = (3, W)
    sh["_time_out.bias"] = (3,)
    return sh


def dnerf_forward(p, x, emb_t, input_ch, input_ch_views, L_pos, cur_time: float,
                  D=8, skips=(4,), use_viewdirs=True, zero_canonical=True):
    """model.py:138-151.  Returns (raw[M,4], dx[M,3])."""
    pts = x[..., :input_ch]
    views = x[..., input_ch:input_ch + input_ch_views]
    if cur_time == 0.0 and zero_canonical:                           # model.py:144-145
        dx = torch.zeros_like(pts[..., :3])
    else:
        h = torch.cat([pts, emb_t], -1)                              # model.py:129
        for i in range(D):
            h = torch.relu(_lin(p, f"_time.{i}", h))                 # model.py:131-132
            if i in skips:
                h = torch.cat([pts, h], -1)                          # model.py:133-134
        dx = _lin(p, "_time_out", h)                                 # model.py:136
        pts = embed(pts[..., :3] + dx, L_pos)                        # model.py:148-149 (PE inside the graph)
    out = mlp_forward(p, torch.cat([pts, views], -1), input_ch, input_ch_views, D, skips,
                      use_viewdirs, prefix="_occ.")                  # model.py:150
    return out, dx


# ----------------------------------------------------------------------------
# a8  raw2outputs                                 reference: ray.py:155-198
# ----------------------------------------------------------------------------
def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0.0, white_bkgd=False, noise=None):
    """Alpha compositing.  `noise` (optional, [N,S]) is added to sigma before the relu
    (ray.py:176-186); the caller supplies it so both sides consume identical randoms."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]                        # ray.py:170
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)   # ray.py:171
    dists = dists * torch.linalg.norm(rays_d, dim=-1, keepdim=True)   # ray.py:173
    rgb = torch.sigmoid(raw[..., :3])                                  # ray.py:175
    sigma = raw[..., 3]
    if noise is not None and raw_noise_std > 0.0:
        sigma = sigma + noise
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * dists)               # ray.py:168,186
    t = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1), -1)[..., :-1]
    weights = alpha * t                                                # ray.py:188
    rgb_map = (weights[..., None] * rgb).sum(-2)                       # ray.py:189
    depth_map = (weights * z_vals).sum(-1)                             # ray.py:191
    acc_map = weights.sum(-1)                                          # ray.py:193
    disp_map = 1.0 / torch.maximum(torch.full_like(depth_map, 1e-10), depth_map / acc_map)  # ray.py:192
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])                 # ray.py:195-196
    return rgb_map, disp_map, acc_map, weights, depth_map


# ----------------------------------------------------------------------------
# a9 / a13  sample_pdf and batched searchsorted    reference: ray.py:96-153
# ----------------------------------------------------------------------------
def pdf_to_cdf(weights: torch.Tensor) -> torch.Tensor:
    w = weights + 1e-5                                                 # ray.py:111
    pdf = w / w.sum(-1, keepdim=True)                                  # ray.py:112
    cdf = torch.cumsum(pdf, -1)                                        # ray.py:113
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)        # ray.py:114


def torch_cpu_sum_order(w: np.ndarray, lanes: int = 8) -> np.float32:
    """The order in which torch.sum reduces ONE contiguous fp32 row on the reference's CPU path (ray.py:112), restated
    from ATen's vectorised inner sum (aten/src/ATen/native/cpu/SumKernel.cpp: `lanes`-wide vector partial sums, four
    interleaved accumulators when the row holds >= 4 vectors, scalar tail first, then the lanes of the partial sum in
    order).  ATen registers this kernel for AVX2 (8 lanes) on AVX512 hosts too.  Valid below 16 * 4 * lanes elements
    (no cascade level).  The CUDA check mode follows this order (ray_kernels.cu: warp_build_cdf_ref); the test
    test_cpu_sum_order_is_the_one_the_check_mode_follows pins it to torch.sum on the host at hand."""
    f = np.float32
    n = len(w)
    nvec, nilp = n // lanes, (n // lanes) // 4
    tot = f(0)
    for k in range(nvec * lanes, n):
        tot = f(tot + w[k])
    for l in range(lanes):
        p = [f(0)] * 4
        for i in range(nilp):
            for k in range(4):
                p[k] = f(p[k] + w[(4 * i + k) * lanes + l])
        for i in range(nilp * 4, nvec):
            p[0] = f(p[0] + w[i * lanes + l])
        tot = f(tot + f(f(f(p[0] + p[1]) + p[2]) + p[3]))
    return tot


def searchsorted_rows(a: np.ndarray, v: np.ndarray, side: str = "left") -> np.ndarray:
    """Row-wise np.searchsorted with row broadcasting, int64 result.
    d_nerf/torchsearchsorted/src/torchsearchsorted/utils.py:4-14 and
    src/cpu/searchsorted_cpu_wrapper.cpp:4-80 (binary_search returns idx, +1 stored)."""
    rows = max(a.shape[0], v.shape[0])
    out = np.empty((rows, v.shape[1]), dtype=np.int64)
    for r in range(rows):
        ar = a[0] if a.shape[0] == 1 else a[r]
        vr = v[0] if v.shape[0] == 1 else v[r]
        out[r] = np.searchsorted(ar, vr, side=side)
    return out


def sample_from_cdf(bins, cdf, u):
    """Inverse-CDF sampling given an explicit cdf (ray.py:134-151).  Returns (samples, inds)."""
    u = u.contiguous()
    inds = torch.searchsorted(cdf.contiguous(), u, right=True)         # ray.py:136
    below = torch.clamp(inds - 1, min=0)                               # ray.py:137
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)                   # ray.py:138
    cdf_b = torch.gather(cdf, -1, below)
    cdf_a = torch.gather(cdf, -1, above)                               # ray.py:145
    bin_b = torch.gather(bins, -1, below)
    bin_a = torch.gather(bins, -1, above)                              # ray.py:146
    denom = cdf_a - cdf_b                                              # ray.py:148
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)   # ray.py:149
    t = (u - cdf_b) / denom                                            # ray.py:150
    return bin_b + t * (bin_a - bin_b), inds                           # ray.py:151


def sample_pdf(bins, weights, N_samples, det=False, u=None):
    """ray.py:96-153.  `u` overrides the random draws (the pytest hook, ray.py:123-132)."""
    cdf = pdf_to_cdf(weights)
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, N_samples, dtype=F32, device=cdf.device)   # ray.py:118
            u = u.expand(list(cdf.shape[:-1]) + [N_samples])
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [N_samples], dtype=F32, device=cdf.device)   # ray.py:121
    samples, _ = sample_from_cdf(bins, cdf, u)
    return samples


# ----------------------------------------------------------------------------
# a2 / a3  stratified sampling and point generation  reference: nerf/run.py:361-385
# ----------------------------------------------------------------------------
def stratified_z(near, far, N_samples, lindisp=False, perturb=0.0, t_rand=None):
    t = torch.linspace(0.0, 1.0, N_samples, dtype=F32, device=near.device)   # run.py:361
    if not lindisp:
        z = near * (1.0 - t) + far * t                                 # run.py:363
    else:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)             # run.py:365
    z = z.expand([near.shape[0], N_samples])
    if perturb > 0.0:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])                        # run.py:371
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(z.shape, dtype=F32, device=z.device)   # run.py:375
        z = lower + (upper - lower) * t_rand                           # run.py:383
    return z


def points(rays_o, rays_d, z):
    return rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]   # run.py:385


# ----------------------------------------------------------------------------
# a5  run_network                                 reference: nerf/run.py:73-87
# ----------------------------------------------------------------------------
def run_network(p, pts, viewdirs, L_pos, L_dir, **mlp_kw):
    flat = pts.reshape(-1, pts.shape[-1])
    emb = embed(flat, L_pos)                                            # run.py:76-77
    ch = emb.shape[-1]
    chv = 0
    if viewdirs is not None:
        dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)       # run.py:80-81
        ed = embed(dirs, L_dir)
        chv = ed.shape[-1]
        emb = torch.cat([emb, ed], -1)                                  # run.py:83
    out = mlp_forward(p, emb, ch, chv, use_viewdirs=viewdirs is not None, **mlp_kw)
    return out.reshape(list(pts.shape[:-1]) + [out.shape[-1]])


# ----------------------------------------------------------------------------
# a1  render_rays (vanilla)                        reference: nerf/run.py:316-422
# ----------------------------------------------------------------------------
def render_rays(ray_batch, p_coarse, p_fine, N_samples, N_importance, L_pos=10, L_dir=4,
                lindisp=False, perturb=0.0, white_bkgd=False, raw_noise_std=0.0,
                t_rand=None, u=None, noise0=None, noise1=None, retraw=False, **mlp_kw):
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]              # run.py:355
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None  # run.py:357
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]                   # run.py:358-359
    z_vals = stratified_z(near, far, N_samples, lindisp, perturb, t_rand)
    raw = run_network(p_coarse, points(rays_o, rays_d, z_vals), viewdirs, L_pos, L_dir, **mlp_kw)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(
        raw, z_vals, rays_d, raw_noise_std, white_bkgd, noise0)        # run.py:390
    ret = {}
    if N_importance > 0:
        rgb0, disp0, acc0 = rgb_map, disp_map, acc_map
        z_mid = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])              # run.py:396
        z_samples = sample_pdf(z_mid, weights[..., 1:-1], N_importance,
                               det=(perturb == 0.0), u=u).detach()      # run.py:397-398
        z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)  # run.py:400
        p_run = p_coarse if p_fine is None else p_fine                  # run.py:403
        raw = run_network(p_run, points(rays_o, rays_d, z_vals), viewdirs, L_pos, L_dir, **mlp_kw)
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(
            raw, z_vals, rays_d, raw_noise_std, white_bkgd, noise1)    # run.py:407
        ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0,
                   z_std=torch.std(z_samples, dim=-1, unbiased=False))  # run.py:412-416
    ret.update(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map)
    ret["z_vals"] = z_vals
    ret["weights"] = weights
    ret["depth_map"] = depth_map
    if retraw:
        ret["raw"] = raw
    return ret


# ----------------------------------------------------------------------------
# a1d / a5d  render_rays + run_network (D-NeRF)    reference: d_nerf/run_dnerf.py:46-83, 354-480
# ----------------------------------------------------------------------------
def run_network_dnerf(p, pts, viewdirs, frame_time: float, L_pos, L_time, L_dir,
                      zero_canonical=True, **kw):
    flat = pts.reshape(-1, 3)
    emb = embed(flat, L_pos)                                            # run_dnerf.py:57-58
    t = torch.full((flat.shape[0], 1), float(frame_time), dtype=F32, device=flat.device)    # run_dnerf.py:62-64
    emb_t = embed(t, L_time)                                            # run_dnerf.py:65
    ch, chv = emb.shape[-1], 0
    if viewdirs is not None:
        dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
        ed = embed(dirs, L_dir)                                         # run_dnerf.py:72-75
        chv = ed.shape[-1]
        emb = torch.cat([emb, ed], -1)
    out, dx = dnerf_forward(p, emb, emb_t, ch, chv, L_pos, float(frame_time),
                            use_viewdirs=viewdirs is not None, zero_canonical=zero_canonical, **kw)
    return out.reshape(list(pts.shape[:-1]) + [4]), dx.reshape(list(pts.shape[:-1]) + [3])


def render_rays_dnerf(ray_batch, p, N_samples, N_importance, L_pos=10, L_time=10, L_dir=4,
                      lindisp=False, perturb=0.0, white_bkgd=False, raw_noise_std=0.0,
                      t_rand=None, u=None, z_vals=None, p_fine=None, use_two_models_for_fine=False,
                      zero_canonical=True, retraw=False):
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 9 else None   # run_dnerf.py:402
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    frame_time = float(ray_batch[0, 8])                                 # run_dnerf.py:403-404 (+ :53 single time)
    q = lambda pp, pts: run_network_dnerf(pp, pts, viewdirs, frame_time, L_pos, L_time, L_dir,
                                          zero_canonical=zero_canonical)
    ret = {}
    z_samples = None
    if z_vals is None:                                                  # run_dnerf.py:408
        z_vals = stratified_z(near, far, N_samples, lindisp, perturb, t_rand)
        pts = points(rays_o, rays_d, z_vals)
        if N_importance > 0:
            if use_two_models_for_fine:                                 # run_dnerf.py:441-443
                raw, dx0 = q(p, pts)
                rgb0, disp0, acc0, weights, _ = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd)
                ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0, position_delta_0=dx0)
            else:
                with torch.no_grad():                                   # run_dnerf.py:446-448
                    raw, _ = q(p, pts)
                    _, _, _, weights, _ = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd)
            z_mid = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
            z_samples = sample_pdf(z_mid, weights[..., 1:-1], N_importance,
                                   det=(perturb == 0.0), u=u).detach()  # run_dnerf.py:451-452
            z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
    pts = points(rays_o, rays_d, z_vals)                                # run_dnerf.py:455
    p_run = p if p_fine is None else p_fine
    raw, dx = q(p_run, pts)
    rgb_map, disp_map, acc_map, weights, _ = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd)
    ret.update(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map, z_vals=z_vals, position_delta=dx)
    if retraw:
        ret["raw"] = raw
    if z_samples is not None:
        ret["z_std"] = torch.std(z_samples, dim=-1, unbiased=False)
    return ret


# ----------------------------------------------------------------------------
# f4  T-NeRF                 reference: model.py:152-210, t_nerf/run_tnerf.py:45-86, 396-500
# ----------------------------------------------------------------------------
def tnerf_param_shapes(depth=8, in_feat=63, dir_feat=27, time_feat=21, net_dim=128,
                       skip_layer=4) -> Dict[str, Tuple[int, ...]]:
    """state_dict of model.TNeRF (model.py:153-186): layers.{i}.0, density.0, feature.0, layer_9.0, color.0."""
    sh: Dict[str, Tuple[int, ...]] = {}
    for i in range(depth):
        k = (in_feat + time_feat) if i == 0 else net_dim
        if i % (skip_layer + 1) == 0 and i > 0:                         # model.py:163
            k = net_dim + in_feat + time_feat
        sh[f"layers.{i}.0.weight"] = (net_dim, k)
        sh[f"layers.{i}.0.bias"] = (net_dim,)
    sh["density.0.weight"], sh["density.0.bias"] = (1, net_dim), (1,)
    sh["feature.0.weight"], sh["feature.0.bias"] = (net_dim, net_dim), (net_dim,)
    sh["layer_9.0.weight"], sh["layer_9.0.bias"] = (net_dim // 2, net_dim + dir_feat), (net_dim // 2,)
    sh["color.0.weight"], sh["color.0.bias"] = (3, net_dim // 2), (3,)
    return sh


def tnerf_forward(p, inp, vdir, dyn_t, depth=8, in_feat=63, skip_layer=4):
    """model.py:188-210.  inp [M, >= in_feat], vdir [M, dir_feat], dyn_t [M, time_feat] -> [M, 4] (rgb, sigma)."""
    x0 = torch.cat([inp[:, :in_feat], dyn_t], -1)                       # model.py:189-190
    x = x0
    for i in range(depth):
        x = torch.nn.functional.elu(_lin(p, f"layers.{i}.0", x))        # model.py:197
        if i % skip_layer == 0 and i > 0:                               # model.py:198-199
            x = torch.cat([x0, x], -1)
    sigma = _lin(p, "density.0", x)                                     # model.py:201
    x = _lin(p, "feature.0", x)                                         # model.py:202
    x = torch.nn.functional.elu(_lin(p, "layer_9.0", torch.cat([x, vdir], -1)))   # model.py:203-204
    rgb = torch.relu(_lin(p, "color.0", x))                             # model.py:205
    return torch.cat([rgb, sigma], -1)                                  # model.py:209


def run_network_tnerf(p, pts, viewdirs, frame_time: float, L_pos=10, L_time=10, L_dir=4, **kw):
    """run_tnerf.py:45-86."""
    flat = pts.reshape(-1, 3)
    emb = embed(flat, L_pos)                                            # :56-57
    t = torch.full((flat.shape[0], 1), float(frame_time), dtype=F32, device=flat.device)    # :61-63
    emb_t = embed(t, L_time)                                            # :64
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    ed = embed(dirs, L_dir)                                             # :70-73
    out = tnerf_forward(p, torch.cat([emb, ed], -1), ed, emb_t, in_feat=emb.shape[-1], **kw)
    return out.reshape(list(pts.shape[:-1]) + [4])


def render_rays_tnerf(ray_batch, p, N_samples, L_pos=10, L_time=10, L_dir=4, lindisp=False, perturb=0.0,
                      white_bkgd=False, raw_noise_std=0.0, t_rand=None, z_vals=None, retraw=False, depth=8):
    """run_tnerf.py:396-500 (single network, single pass)."""
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:]                                        # :440 (use_viewdirs configs)
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    frame_time = float(ray_batch[0, 8])                                 # :441-442 (+ :53 single time)
    if z_vals is None:
        z_vals = stratified_z(near, far, N_samples, lindisp, perturb, t_rand)   # :447-470
    pts = points(rays_o, rays_d, z_vals)
    raw = run_network_tnerf(p, pts, viewdirs, frame_time, L_pos, L_time, L_dir, depth=depth)
    rgb_map, disp_map, acc_map, weights, _ = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd)
    ret = dict(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map, z_vals=z_vals)
    if retraw:
        ret["raw"] = raw
    return ret


# ----------------------------------------------------------------------------
# Deterministic parameters / inputs shared by the golden generator and the tests
# ----------------------------------------------------------------------------
def make_params(shapes: Dict[str, Tuple[int, ...]], seed: int, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Weights from numpy's legacy MT19937 RandomState (bit-stable across numpy versions and
    machines), U(-b, b) with b = gain/sqrt(fan_in) - the range of nn.Linear's default init
    (kaiming_uniform a=sqrt(5)).  The golden files store only `seed`, not the 4.8 MB of weights."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, sh in shapes.items():
        fan_in = sh[1] if len(sh) == 2 else None
        if fan_in is None:
            wname = name.replace(".bias", ".weight")
            fan_in = shapes[wname][1]
        b = gain / math.sqrt(fan_in)
        out[name] = torch.from_numpy(rs.uniform(-b, b, size=sh).astype(np.float32))
    # The default-init net renders an empty scene (sigma ~ 0 +- 0.05): every map is degenerate.
    # Scale the two heads so that sigma ~ 0.5 +- 3 and rgb logits ~ +-2: a non-trivial synthetic scene.
    for name in out:
        if name.endswith("alpha_linear.weight"):
            out[name] = out[name] * 24.0
        elif name.endswith("alpha_linear.bias"):
            out[name] = out[name] * 0.0 + 0.5
        elif name.endswith("rgb_linear.weight"):
            out[name] = out[name] * 6.0
        elif name.endswith("_time_out.weight"):
            out[name] = out[name] * 2.0
        elif name.endswith("density.0.weight"):                 # TNeRF heads, same purpose
            out[name] = out[name] * 24.0
        elif name.endswith("density.0.bias"):
            out[name] = out[name] * 0.0 + 0.5
        elif name.endswith("color.0.weight"):
            out[name] = out[name] * 6.0
    return out


def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    """dataloader/load_blender.py:9-35 (trans_t, rot_phi, rot_theta, the axis flip)."""
    t = np.eye(4, dtype=np.float64); t[2, 3] = radius
    ph = phi_deg / 180.0 * np.pi
    rp = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1.0]])
    th = theta_deg / 180.0 * np.pi
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1.0]])
    c2w = rt @ rp @ t
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    return (flip @ c2w).astype(np.float32)


def blender_rays(n_rays: int, seed: int, H: int = 800, W: int = 800, near: float = 2.0, far: float = 6.0,
                 frame_time: Optional[float] = None) -> np.ndarray:
    """Synthetic Blender-shaped pinhole rays: focal from camera_angle_x=0.6911112
    (dataloader/load_blender.py:133-134), c2w = pose_spherical(theta, -30, 4), directions as in
    get_rays_np (ray.py:42-72), near=2 far=6 (nerf/run.py:466-467), unit viewdirs
    (nerf/run.py:137-158).  Returns the flat [N, 11] (or [N, 12] with frame_time) ray batch."""
    rs = np.random.RandomState(seed)
    focal = 0.5 * W / np.tan(0.5 * 0.6911112)
    theta = rs.uniform(-180.0, 180.0)
    c2w = pose_spherical(theta, -30.0, 4.0)
    pix = rs.randint(0, H * W, size=n_rays)
    i = (pix % W).astype(np.float32)
    j = (pix // W).astype(np.float32)
    dirs = np.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -np.ones_like(i)], -1).astype(np.float32)
    rays_d = np.sum(dirs[:, None, :] * c2w[:3, :3], -1).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(np.float32)
    vd = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)
    cols = [rays_o, rays_d, np.full((n_rays, 1), near, np.float32), np.full((n_rays, 1), far, np.float32)]
    if frame_time is not None:
        cols.append(np.full((n_rays, 1), frame_time, np.float32))
    cols.append(vd.astype(np.float32))
    return np.concatenate(cols, -1).astype(np.float32)
