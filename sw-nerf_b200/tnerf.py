"""T-NeRF render path: batchify, run_network, batchify_rays, render_rays,
render and create_nerf of t_nerf/run_tnerf.py (:25-86, :89-176, :242-346, :396-500) on the library's kernels.

T-NeRF is one time-conditioned network (model.TNeRF: ELU, width 128) evaluated at the N_samples stratified
points of each ray - no hierarchical pass (`N_importance` is forced to 0 by create_nerf, run_tnerf.py:319).
The network runs on the fp32 GEMM kernels with the ELU epilogue; sampling, encoding and compositing are the
kernels the vanilla path uses.  As in dnerf.py the "all rays share one time" check is ONE host scalar per
render_rays call instead of two torch.unique syncs per query (run_tnerf.py:53-54).
"""
import os

import torch

from . import ops
from .embedder import get_embedder
from .model import TNeRF
from .ray import get_rays, ndc_rays, raw_noise, pytest_uniform
from .dnerf import _host_time

DEBUG = False


def batchify(fn, chunk):
    """run_tnerf.py:25-42.  TNeRF.forward returns [1, m, 4] (model.py:205-208) and the reference concatenates the
    slabs along dim 0, which only works when every slab has the same length (it raises on a ragged last slab);
    here 3-D slabs are joined along the sample axis - the same flat order run_network's final reshape sees."""
    if chunk is None:
        return fn

    def ret(inputs_pos, viewdirs, dyn_t):
        out_list = []
        for i in range(0, inputs_pos.shape[0], chunk):
            out_list += [fn(inputs_pos[i:i + chunk], viewdirs[i:i + chunk], dyn_t[i:i + chunk])]
        return torch.cat(out_list, 1 if out_list[0].dim() == 3 else 0)
    return ret


def run_network(inputs, viewdirs, frame_time, fn, embed_fn, embeddirs_fn, embedtime_fn, netchunk=1024 * 64,
                embd_time_discr=True):
    """run_tnerf.py:45-86 (reference signature and host checks)."""
    assert len(torch.unique(frame_time)) == 1, "Only accepts all points from same time"
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if embd_time_discr:
        B, N, _ = inputs.shape
        input_frame_time = frame_time[:, None].expand([B, N, 1])
        input_frame_time_flat = torch.reshape(input_frame_time, [-1, 1])
        embedded_times = embedtime_fn(input_frame_time_flat)
    else:
        raise NotImplementedError
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded_dirs = embeddirs_fn(input_dirs_flat)
        embedded = torch.cat([embedded, embedded_dirs], -1)
    else:
        embedded_dirs = None
    outputs_flat = batchify(fn, netchunk)(embedded, embedded_dirs, embedded_times)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


class TNerfNetworkQuery:
    """network_query_fn of run_tnerf.py:281-286 as an object, plus the ray entry render_rays prefers: points and
    both encodings come from one kernel (no [N,S,3] points tensor, no expand of viewdirs) and the time encoding
    is one broadcast row."""

    def __init__(self, embed_fn, embeddirs_fn, embedtime_fn, netchunk=1024 * 64, embd_time_discr=True,
                 precision=None):
        self.embed_fn, self.embeddirs_fn, self.embedtime_fn = embed_fn, embeddirs_fn, embedtime_fn
        self.netchunk, self.embd_time_discr = netchunk, embd_time_discr
        # "tc": layers on the tcgen05 GEMM (fp16 operands, fp32 accumulate - the precision of the fused vanilla path);
        # "fp32": the fp32 SIMT GEMM (the <= 1e-5 check mode)
        self.precision = precision or os.environ.get("SWNERF_PRECISION", "tc")
        if self.precision not in ("tc", "fp32"):
            raise ValueError("precision must be 'tc' or 'fp32'")

    def _arm(self, network_fn):
        if hasattr(network_fn, "tc_gemm"):
            network_fn.tc_gemm = self.precision == "tc"

    def __call__(self, inputs, viewdirs, ts, network_fn):
        self._arm(network_fn)
        return run_network(inputs, viewdirs, ts, network_fn, embed_fn=self.embed_fn,
                           embeddirs_fn=self.embeddirs_fn, embedtime_fn=self.embedtime_fn,
                           netchunk=self.netchunk, embd_time_discr=self.embd_time_discr)

    def query_rays(self, ray_batch, z_vals, network_fn, view_col, cur_time: float):
        N, S = z_vals.shape
        dev = z_vals.device
        self._arm(network_fn)
        if view_col < 0 or getattr(self.embed_fn, "L", None) is None or getattr(self.embeddirs_fn, "L", None) is None:
            rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
            pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
            viewdirs = ray_batch[:, view_col:view_col + 3] if view_col >= 0 else None
            return self(pts, viewdirs, ray_batch[:, 8:9], network_fn)
        emb = ops.encode_points(ray_batch, z_vals, self.embed_fn.L, self.embeddirs_fn.L, view_col)
        n_pts = network_fn.in_feat
        t1 = torch.full((1, 1), float(cur_time), dtype=torch.float32, device=dev)
        emb_t = self.embedtime_fn(t1).expand(N * S, -1)          # one row, broadcast with stride 0
        out = network_fn(emb, emb[:, n_pts:], emb_t)
        return out.reshape(N, S, out.shape[-1])


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """run_tnerf.py:89-102."""
    all_ret = {}
    ft = getattr(rays_flat, "_swnerf_frame_time", None)
    for i in range(0, rays_flat.shape[0], chunk):
        rb = rays_flat[i:i + chunk]
        if ft is not None:
            rb._swnerf_frame_time = ft
        ret = render_rays(rb, **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: torch.cat(all_ret[k], 0) for k in all_ret}


def render(H, W, focal, chunk=1024 * 32, rays=None, frame_time=None, c2w=None, ndc=True, near=0., far=1.,
           use_viewdirs=False, c2w_staticcam=None, **kwargs):
    """run_tnerf.py:105-176."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, focal, c2w)
    else:
        rays_o, rays_d = rays
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, focal, c2w_staticcam)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    sh = rays_d.shape
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, focal, 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near, far = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    ft_host = float(frame_time) if not isinstance(frame_time, torch.Tensor) or frame_time.numel() == 1 else None
    frame_time = float(frame_time) * torch.ones_like(rays_d[..., :1]) if ft_host is not None else \
        frame_time * torch.ones_like(rays_d[..., :1])
    rays = torch.cat([rays_o, rays_d, near, far, frame_time], -1)
    if use_viewdirs:
        rays = torch.cat([rays, viewdirs], -1)
    if ft_host is not None:
        rays._swnerf_frame_time = ft_host
    all_ret = batchify_rays(rays, chunk, **kwargs)
    for k in all_ret:
        k_sh = list(sh[:-1]) + list(all_ret[k].shape[1:])
        all_ret[k] = torch.reshape(all_ret[k], k_sh)
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False,
                pytest=False, z_vals=None, use_two_models_for_fine=False):
    """run_tnerf.py:396-500.  ray_batch [N, 9 | 12]: o, d, near, far, frame_time[, unit viewdir].  One network, one
    pass; N_importance > 0 only prints the reference's warning (:470-471)."""
    ft = getattr(ray_batch, "_swnerf_frame_time", None)
    if not ray_batch.is_contiguous():
        ray_batch = ray_batch.contiguous()
    ray_batch = ray_batch.float()
    if ft is not None:
        ray_batch._swnerf_frame_time = ft
    N_rays, C = ray_batch.shape
    dev = ray_batch.device
    view_col = C - 3 if C > 9 else -1                                             # run_tnerf.py:440
    cur_time = _host_time(ray_batch)
    if z_vals is None:                                                            # run_tnerf.py:447
        t_rand = None
        if perturb > 0.:
            if pytest:                                                            # run_tnerf.py:466-469 (as shipped:
                t_rand = pytest_uniform([N_rays, N_samples], dev) * raw_noise_std  # the draw is scaled by the noise std)
            else:
                t_rand = torch.rand((N_rays, N_samples), device=dev)
        z_vals = ops.stratified_z(ray_batch, N_samples, lindisp, perturb, t_rand, near_col=6)
        if N_importance > 0:
            print("Warning: N_importance is set but only a single model is used.")
    else:
        z_vals = z_vals.contiguous().float()
    if hasattr(network_query_fn, "query_rays"):
        raw = network_query_fn.query_rays(ray_batch, z_vals, network_fn, view_col, cur_time)
    else:
        rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        viewdirs = ray_batch[:, view_col:view_col + 3] if view_col >= 0 else None
        raw = network_query_fn(pts, viewdirs, ray_batch[:, 8:9], network_fn)
    noise = raw_noise(tuple(z_vals.shape), raw_noise_std, dev, pytest)
    rgb_map, disp_map, acc_map, weights, _ = ops.composite(raw.contiguous(), z_vals, ray_batch, 3, noise, white_bkgd)
    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map, 'z_vals': z_vals}
    if retraw:
        ret['raw'] = raw
    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


def create_nerf(args, device=None):
    """run_tnerf.py:242-346: embedders, ONE TNeRF (width 128, skip 4), query closure, Adam, checkpoint reload.
    Returns (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer)."""
    device = torch.device(device if device is not None else getattr(args, "device", "cuda"))
    embed_fn, input_ch = get_embedder(args.multires, 3, args.i_embed)
    embedtime_fn, input_ch_time = get_embedder(args.multires, 1, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, 3, args.i_embed)
    model = TNeRF(depth=args.netdepth, in_feat=input_ch, dir_feat=input_ch_views, time_feat=input_ch_time,
                  net_dim=128, skip_layer=4).to(device)
    grad_vars = list(model.parameters())
    network_query_fn = TNerfNetworkQuery(embed_fn, embeddirs_fn, embedtime_fn, netchunk=args.netchunk,
                                         embd_time_discr=getattr(args, "nerf_type", "tnerf") != "temporal",
                                         precision=getattr(args, "swnerf_precision", None))
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    if getattr(args, "do_half_precision", False):
        raise NotImplementedError("apex amp (run_tnerf.py:291-294) is not part of this path")
    start = 0
    basedir, expname = args.basedir, args.expname
    if getattr(args, "ft_path", None) is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        d = os.path.join(basedir, expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if 'tar' in f] if os.path.isdir(d) else []
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        ckpt_path = ckpts[-1]
        print('Reloading from', ckpt_path)
        ckpt = torch.load(ckpt_path, map_location=device)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
    render_kwargs_train = {
        'network_query_fn': network_query_fn,
        'perturb': args.perturb,
        'N_importance': 0,                                                        # run_tnerf.py:319
        'network_fn': model,
        'N_samples': args.N_samples,
        'use_viewdirs': args.use_viewdirs,
        'white_bkgd': args.white_bkgd,
        'raw_noise_std': args.raw_noise_std,
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    render_kwargs_test = {k: render_kwargs_train[k] for k in render_kwargs_train}
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer
