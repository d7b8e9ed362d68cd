"""Drop-in for the reference's ray.py: sample_pdf and raw2outputs on the warp-per-ray kernels,
plus the ray helpers get_rays / get_rays_np / ndc_rays (host-side, one step before the hot path).

Reference: ray.py:10-92 (helpers), :96-153 (sample_pdf), :155-198 (raw2outputs).
"""
import numpy as np
import torch

from . import ops

img2mse = lambda x, y: torch.mean((x - y) ** 2)
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)


def get_rays(H, W, focal_or_K, c2w):
    """ray.py:10-38."""
    dev = c2w.device if isinstance(c2w, torch.Tensor) else None
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev),
                          indexing='ij')
    i, j = i.t(), j.t()
    if isinstance(focal_or_K, float):
        focal = focal_or_K
        dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
    else:
        K = focal_or_K
        dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    c2w = torch.as_tensor(c2w, dtype=torch.float32, device=dev)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def _camera(H, W, focal_or_K, c2w):
    import ctypes
    if isinstance(focal_or_K, (float, int)):
        fx = fy = float(focal_or_K); cx, cy = W * 0.5, H * 0.5
    else:
        K = focal_or_K
        fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    m = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()[:3, :4].contiguous()
    return fx, fy, cx, cy, (ctypes.c_float * 12)(*m.reshape(-1).tolist())


def make_ray_batch(H, W, focal_or_K, c2w, near, far, pixels=None, frame_time=None, use_viewdirs=True, device=None,
                   ndc=False, ndc_near=1.):
    """Flat ray batch straight from the camera: what get_rays (ray.py:10-38) plus the
    viewdir normalisation / NDC warp / near-far / concatenation of render() (nerf/run.py:137-158) build with a dozen
    eager ops, as ONE kernel.  `pixels`: int64 tensor of flat pixel ids j*W + i (None = the whole frame).
    ndc=True applies ndc_rays(H, W, K[0][0], ndc_near, o, d) (ray.py:75-92) like render() does for the LLFF configs."""
    if device is None:
        device = pixels.device if pixels is not None else (c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda
                                                           else torch.device("cuda"))
    fx, fy, cx, cy, c2w12 = _camera(H, W, focal_or_K, c2w)
    n = H * W if pixels is None else pixels.numel()
    stride = 8 + (1 if frame_time is not None else 0) + (3 if use_viewdirs else 0)
    rays = torch.empty((n, stride), dtype=torch.float32, device=device)
    from ._lib import call, ptr, stream
    call("swnerf_make_rays", H, W, fx, fy, cx, cy, c2w12,
         None if pixels is None else ptr(pixels.reshape(-1), torch.int64, "pixels"), n, float(near), float(far),
         0.0 if frame_time is None else float(frame_time), int(frame_time is not None), int(use_viewdirs),
         int(bool(ndc)), float(ndc_near), float(fx), rays.data_ptr(), stride, stream())
    return rays


def pick_batch(H, W, focal_or_K, c2w, image, N_rand, seed, near, far, precrop_frac=None, frame_time=None,
               use_viewdirs=True, ndc=False, ndc_near=1., return_pixels=False):
    """The per-step training batch of nerf/run.py:652-681 in ONE kernel: N_rand distinct random
    pixels of `image` [H, W, 3] (a CUDA tensor) - inside the centre crop of half-size H//2 * precrop_frac when
    `precrop_frac` is given (run.py:660-668) - as (ray_batch [N_rand, 8|9|11|12], target_s [N_rand, 3]).
    `seed`: any integer, e.g. the global step; equal seeds give equal batches (the reference uses numpy's global RNG)."""
    if not (isinstance(image, torch.Tensor) and image.is_cuda):
        raise RuntimeError("swnerf_b200: pick_batch needs the target image on the GPU - this library has no CPU path")
    if precrop_frac is not None:
        dH, dW = int(H // 2 * precrop_frac), int(W // 2 * precrop_frac)
        y0, x0, ch, cw = H // 2 - dH, W // 2 - dW, 2 * dH, 2 * dW
    else:
        y0, x0, ch, cw = 0, 0, H, W
    fx, fy, cx, cy, c2w12 = _camera(H, W, focal_or_K, c2w)
    img = image.reshape(H * W, -1)[:, :3].float().contiguous()
    stride = 8 + (1 if frame_time is not None else 0) + (3 if use_viewdirs else 0)
    rays = torch.empty((N_rand, stride), dtype=torch.float32, device=image.device)
    target = torch.empty((N_rand, 3), dtype=torch.float32, device=image.device)
    pix = torch.empty((N_rand,), dtype=torch.int64, device=image.device) if return_pixels else None
    from ._lib import call, stream
    call("swnerf_pick_batch", H, W, fx, fy, cx, cy, c2w12, img.data_ptr(), y0, x0, ch, cw,
         int(seed) & 0xFFFFFFFFFFFFFFFF, N_rand, float(near), float(far),
         0.0 if frame_time is None else float(frame_time), int(frame_time is not None), int(use_viewdirs),
         int(bool(ndc)), float(ndc_near), float(fx), rays.data_ptr(), stride, target.data_ptr(),
         None if pix is None else pix.data_ptr(), stream())
    return (rays, target, pix) if return_pixels else (rays, target)


def get_rays_np(H, W, focal_or_K, c2w):
    """ray.py:42-72."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    if isinstance(focal_or_K, float):
        focal = focal_or_K
        dirs = np.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -np.ones_like(i)], -1)
    else:
        K = focal_or_K
        dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """ray.py:75-92."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def pytest_uniform(shape, device):
    """The reference's deterministic-random hook: np.random.seed(0); np.random.rand(*shape)
    (nerf/run.py:377-381, ray.py:124-132, 180-184)."""
    np.random.seed(0)
    return torch.Tensor(np.random.rand(*shape)).to(device)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """ray.py:96-153.  bins [N, M], weights [N, M-1] -> samples [N, N_samples]."""
    lead = bins.shape[:-1]
    b2 = bins.reshape(-1, bins.shape[-1]).contiguous()
    w2 = weights.detach().reshape(-1, weights.shape[-1]).contiguous()
    u = None
    if pytest and not det:
        u = pytest_uniform([b2.shape[0], N_samples], bins.device)
    out = ops.sample_pdf(b2.detach(), w2, N_samples, det=det, u=u)
    return out.reshape(list(lead) + [N_samples])


def raw_noise(shape, raw_noise_std, device, pytest=False):
    """ray.py:176-184 (note the pytest branch is uniform, the normal path Gaussian)."""
    if not raw_noise_std > 0.:
        return None
    if pytest:
        return pytest_uniform(list(shape), device) * raw_noise_std
    return torch.randn(shape, device=device) * raw_noise_std


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """ray.py:155-198 -> (rgb_map, disp_map, acc_map, weights, depth_map)."""
    noise = raw_noise(tuple(raw.shape[:-1]), raw_noise_std, raw.device, pytest)
    rd = rays_d if rays_d.is_contiguous() else rays_d.contiguous()
    z = z_vals if z_vals.is_contiguous() else z_vals.contiguous()
    return ops.composite(raw, z, rd, 0, noise, white_bkgd)
