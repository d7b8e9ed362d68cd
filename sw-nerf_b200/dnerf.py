"""D-NeRF render path: batchify, run_network, batchify_rays, render_rays, create_nerf, render with
`frame_time` (d_nerf/run_dnerf.py:24-99, 102-172, 238-351, 354-480), and MultiRes D-NeRF's
create_nerf(args, channels, layer) (multires_dnerf/multires_dnerf.py:242-354).

Differences from the reference that do not change results: the "all rays share one time" check and
the `t == 0` canonical test use ONE host scalar per render_rays call (the reference issues two
torch.unique syncs per network query, run_dnerf.py:53-54, plus model.py:142-144); with
N_importance <= 0 the reference queries the network twice with identical inputs (run_dnerf.py:437-439
then :455-458) - the redundant first query is skipped.
"""
import os

import torch

from . import ops, tc
from .embedder import get_embedder
from .model import NeRF
from .ray import get_rays, ndc_rays, raw_noise, pytest_uniform

DEBUG = False


def batchify(fn, chunk):
    """run_dnerf.py:24-43."""
    if chunk is None:
        return fn

    def ret(inputs_pos, inputs_time):
        out_list, dx_list = [], []
        for i in range(0, inputs_pos.shape[0], chunk):
            out, dx = fn(inputs_pos[i:i + chunk], [inputs_time[0][i:i + chunk], inputs_time[1][i:i + chunk]])
            out_list += [out]
            dx_list += [dx]
        return torch.cat(out_list, 0), torch.cat(dx_list, 0)
    return ret


def run_network(inputs, viewdirs, frame_time, fn, embed_fn, embeddirs_fn, embedtime_fn, netchunk=1024 * 64,
                embd_time_discr=True):
    """run_dnerf.py:46-83 (reference signature; this generic entry keeps the reference's host checks)."""
    assert len(torch.unique(frame_time)) == 1, "Only accepts all points from same time"
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if embd_time_discr:
        B, N, _ = inputs.shape
        input_frame_time = frame_time[:, None].expand([B, N, 1])
        input_frame_time_flat = torch.reshape(input_frame_time, [-1, 1])
        embedded_time = embedtime_fn(input_frame_time_flat)
        embedded_times = [embedded_time, embedded_time]
    else:
        assert NotImplementedError
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded_dirs = embeddirs_fn(input_dirs_flat)
        embedded = torch.cat([embedded, embedded_dirs], -1)
    outputs_flat, position_delta_flat = batchify(fn, netchunk)(embedded, embedded_times)
    outputs = torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])
    position_delta = torch.reshape(position_delta_flat, list(inputs.shape[:-1]) + [position_delta_flat.shape[-1]])
    return outputs, position_delta


class DNerfNetworkQuery:
    """network_query_fn of run_dnerf.py:279-284 as an object, plus the fused ray entry."""

    def __init__(self, embed_fn, embeddirs_fn, embedtime_fn, netchunk=1024 * 64, embd_time_discr=True,
                 precision=None):
        self.embed_fn, self.embeddirs_fn, self.embedtime_fn = embed_fn, embeddirs_fn, embedtime_fn
        self.netchunk, self.embd_time_discr = netchunk, embd_time_discr
        self.precision = precision or os.environ.get("SWNERF_PRECISION", "tc")
        if self.precision not in ("tc", "fp32"):
            raise ValueError("precision must be 'tc' or 'fp32'")
        self.allow_fused = True          # False: 'tc' runs layer by layer on the tcgen05 GEMM even where a fused kernel exists (tests)

    def uses_tc(self, network_fn, has_views):
        """Fused tcgen05 kernels for 8x256 DirectTemporalNeRF networks: the D-NeRF configs (PE 10 / 10 / 4) and every
        level of the MultiRes pyramid ((pos, time, view) = (20, 8, 20), (10, 4, 10), identity; multires_dnerf.py:665)."""
        return (self.precision == "tc" and self.allow_fused and tc.available() and tc.bwd_available() and has_views
                and hasattr(network_fn, "_time") and tc.dnerf_tc_eligible(network_fn)
                and self._enc(network_fn) is not None)

    def _enc(self, network_fn):
        return tc.enc_for(self.embed_fn, self.embeddirs_fn, self.embedtime_fn, network_fn)

    def _arm(self, network_fn):
        """Shapes without a fused kernel (MultiRes encoding widths, ...) run layer by layer: on the tcgen05 GEMM in 'tc'
        precision, on the fp32 SIMT GEMM in 'fp32'."""
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
        if self.uses_tc(network_fn, view_col >= 0):
            return tc.dnerf_query(network_fn, ray_batch, z_vals, view_col, float(cur_time), enc=self._enc(network_fn))
        self._arm(network_fn)
        L_pos = self.embed_fn.L
        L_dir = self.embeddirs_fn.L if view_col >= 0 else -1
        emb = ops.encode_points(ray_batch, z_vals, L_pos, L_dir, view_col)
        t1 = torch.full((1, 1), float(cur_time), dtype=torch.float32, device=dev)
        emb_t = self.embedtime_fn(t1).expand(N * S, -1)          # one row, broadcast with stride 0
        if hasattr(network_fn, "_time"):
            out, dx = network_fn(emb, [emb_t, emb_t], cur_time=float(cur_time))
        else:
            out, dx = network_fn(emb, [emb_t, emb_t])
        return out.reshape(N, S, out.shape[-1]), dx.reshape(N, S, 3)


def _host_time(ray_batch):
    ft = getattr(ray_batch, "_swnerf_frame_time", None)
    if ft is None:
        col = ray_batch[:, 8]
        ft = float(col[0])
        assert bool((col == ft).all()), "Only accepts all points from same time"   # run_dnerf.py:53
    return ft


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """run_dnerf.py:86-99."""
    all_ret = {}
    ft = getattr(rays_flat, "_swnerf_frame_time", None)
    zv = kwargs.pop("z_vals", None)
    for i in range(0, rays_flat.shape[0], chunk):
        rb = rays_flat[i:i + chunk]
        if ft is not None:
            rb._swnerf_frame_time = ft
        kw = dict(kwargs)
        if zv is not None:
            kw["z_vals"] = zv[i:i + chunk]
        ret = render_rays(rb, **kw)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: torch.cat(all_ret[k], 0) for k in all_ret}


def render(H, W, focal, chunk=1024 * 32, rays=None, frame_time=None, c2w=None, ndc=True, near=0., far=1.,
           use_viewdirs=False, c2w_staticcam=None, **kwargs):
    """run_dnerf.py:102-172."""
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


def _query(network_query_fn, ray_batch, z_vals, network, view_col, cur_time):
    if hasattr(network_query_fn, "query_rays"):
        return network_query_fn.query_rays(ray_batch, z_vals, network, view_col, cur_time)
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    viewdirs = ray_batch[:, view_col:view_col + 3] if view_col >= 0 else None
    return network_query_fn(pts, viewdirs, ray_batch[:, 8:9], network)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False,
                pytest=False, z_vals=None, use_two_models_for_fine=False):
    """run_dnerf.py:354-480.  ray_batch [N, 9 | 12]: o, d, near, far, frame_time[, unit viewdir]."""
    ft = getattr(ray_batch, "_swnerf_frame_time", None)
    if not ray_batch.is_contiguous():
        ray_batch = ray_batch.contiguous()
    ray_batch = ray_batch.float()
    if ft is not None:
        ray_batch._swnerf_frame_time = ft
    N_rays, C = ray_batch.shape
    dev = ray_batch.device
    view_col = C - 3 if C > 9 else -1                                             # run_dnerf.py:402
    cur_time = _host_time(ray_batch)
    z_std = None
    rgb_map_0 = disp_map_0 = acc_map_0 = position_delta_0 = None

    if z_vals is None:                                                            # run_dnerf.py:408
        t_rand = None
        if perturb > 0.:
            t_rand = pytest_uniform([N_rays, N_samples], dev) if pytest else \
                torch.rand((N_rays, N_samples), device=dev)
        z_vals = ops.stratified_z(ray_batch, N_samples, lindisp, perturb, t_rand, near_col=6)
        if N_importance > 0:
            if use_two_models_for_fine:                                           # run_dnerf.py:441-443
                raw, position_delta_0 = _query(network_query_fn, ray_batch, z_vals, network_fn, view_col, cur_time)
                noise = raw_noise((N_rays, N_samples), raw_noise_std, dev, pytest)
                rgb_map_0, disp_map_0, acc_map_0, weights, _ = ops.composite(raw, z_vals, ray_batch, 3, noise,
                                                                             white_bkgd)
            else:
                with torch.no_grad():                                             # run_dnerf.py:446-448
                    raw, _ = _query(network_query_fn, ray_batch, z_vals, network_fn, view_col, cur_time)
                    noise = raw_noise((N_rays, N_samples), raw_noise_std, dev, pytest)
                    _, _, _, weights, _ = ops.composite(raw, z_vals, ray_batch, 3, noise, white_bkgd)
            det = (perturb == 0.)
            u = pytest_uniform([N_rays, N_importance], dev) if (pytest and not det) else None
            _, z_vals, z_std = ops.resample(z_vals, weights.detach(), N_importance, det=det, u=u, want_samples=False,
                                            exact=getattr(network_query_fn, "precision", None) == "fp32")
    else:
        z_vals = z_vals.contiguous().float()

    run_fn = network_fn if network_fine is None else network_fine
    raw, position_delta = _query(network_query_fn, ray_batch, z_vals, run_fn, view_col, cur_time)
    noise = raw_noise(tuple(z_vals.shape), raw_noise_std, dev, pytest)
    rgb_map, disp_map, acc_map, weights, _ = ops.composite(raw, z_vals, ray_batch, 3, noise, white_bkgd)

    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map, 'z_vals': z_vals,
           'position_delta': position_delta}
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        if rgb_map_0 is not None:
            ret['rgb0'] = rgb_map_0
        if disp_map_0 is not None:
            ret['disp0'] = disp_map_0
        if acc_map_0 is not None:
            ret['acc0'] = acc_map_0
        if position_delta_0 is not None:
            ret['position_delta_0'] = position_delta_0
        if z_std is not None:
            ret['z_std'] = z_std
    return ret


def _finish_create(args, model, model_fine, grad_vars, network_query_fn, device, model_key="network_fn_state_dict",
                   fine_key="network_fine_state_dict", opt_key="optimizer_state_dict"):
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    if getattr(args, "do_half_precision", False):
        raise NotImplementedError("apex amp (do_half_precision) is not supported; every reference config has it off")
    start = 0
    basedir, expname = args.basedir, args.expname
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        ckpts = [os.path.join(basedir, expname, f) for f in sorted(os.listdir(os.path.join(basedir, expname)))
                 if 'tar' in f]
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        ckpt_path = ckpts[-1]
        print('Reloading from', ckpt_path)
        ckpt = torch.load(ckpt_path, map_location=device)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt[opt_key])
        model.load_state_dict(ckpt[model_key])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt[fine_key])
    render_kwargs_train = {
        'network_query_fn': network_query_fn,
        'perturb': args.perturb,
        'N_importance': args.N_importance,
        'network_fine': model_fine,
        'N_samples': args.N_samples,
        'network_fn': model,
        'use_viewdirs': args.use_viewdirs,
        'white_bkgd': args.white_bkgd,
        'raw_noise_std': args.raw_noise_std,
        'use_two_models_for_fine': args.use_two_models_for_fine,
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    render_kwargs_test = {k: render_kwargs_train[k] for k in render_kwargs_train}
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


def _build_models(args, embed_fn, input_ch, embedtime_fn, input_ch_time, embeddirs_fn, input_ch_views, device):
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    mk = lambda D, W: NeRF.get_by_name(args.nerf_type, D=D, W=W, input_ch=input_ch, output_ch=output_ch,
                                       skips=skips, input_ch_views=input_ch_views, input_ch_time=input_ch_time,
                                       use_viewdirs=args.use_viewdirs, embed_fn=embed_fn,
                                       zero_canonical=not args.not_zero_canonical).to(device)
    model = mk(args.netdepth, args.netwidth)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.use_two_models_for_fine:
        model_fine = mk(args.netdepth_fine, args.netwidth_fine)
        grad_vars += list(model_fine.parameters())
    q = DNerfNetworkQuery(embed_fn, embeddirs_fn, embedtime_fn, args.netchunk,
                          embd_time_discr=args.nerf_type != "temporal",
                          precision=getattr(args, "swnerf_precision", None))
    return model, model_fine, grad_vars, q


def create_nerf(args, device=None):
    """run_dnerf.py:238-351."""
    if device is None:
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    embed_fn, input_ch = get_embedder(args.multires, 3, args.i_embed)
    embedtime_fn, input_ch_time = get_embedder(args.multires, 1, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, 3, args.i_embed)
    model, model_fine, grad_vars, q = _build_models(args, embed_fn, input_ch, embedtime_fn, input_ch_time,
                                                    embeddirs_fn, input_ch_views, device)
    return _finish_create(args, model, model_fine, grad_vars, q, device)


def create_nerf_multires(args, channels=None, layer=None, device=None):
    """multires_dnerf.py:242-354: per-level PE sizes `channels = (pos, time, dir)`, -1 = identity
    (get_embedder(L, dims, i=L), :256-262), per-level checkpoint keys network_fn_{layer} (:320-325)."""
    if device is None:
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    embed_fn, input_ch = get_embedder(channels[0], 3, channels[0])
    embedtime_fn, input_ch_time = get_embedder(channels[1], 1, channels[1])
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(channels[2], 3, channels[2])
    model, model_fine, grad_vars, q = _build_models(args, embed_fn, input_ch, embedtime_fn, input_ch_time,
                                                    embeddirs_fn, input_ch_views, device)
    return _finish_create(args, model, model_fine, grad_vars, q, device, model_key=f"network_fn_{layer}",
                          fine_key=f"network_fine_{layer}", opt_key=f"optimizer_{layer}")
