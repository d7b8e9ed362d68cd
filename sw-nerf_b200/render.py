"""Vanilla NeRF render path: batchify, run_network, batchify_rays, render_rays, create_nerf, render.

Mirrors nerf/run.py:63-102, 105-170, 222-311, 316-422 (same names, arguments, return dicts and error
behaviour) so that nerf/run.py can import these instead of defining its own.  render_rays hands the
flat ray batch straight to the fused kernels (stratified z -> points+PE+MLP -> compositing ->
resample -> points+PE+MLP -> compositing) without materialising points or embeddings in HBM.
"""
import os

import torch

from . import ops, tc
from .embedder import get_embedder
from .model import vallina_NeRF as NeRF
from .ray import get_rays, ndc_rays, raw_noise, pytest_uniform, make_ray_batch
from .parallel import render_path  # noqa: F401  (nerf/run.py:172-219, ray-sharded; defined next to the other DP code)

DEBUG = False


def batchify(fn, chunk):
    """nerf/run.py:63-70."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """nerf/run.py:73-87: embed points (+ expanded viewdirs), apply `fn` in netchunk slabs."""
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded_dirs = embeddirs_fn(input_dirs_flat)
        embedded = torch.cat([embedded, embedded_dirs], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


class NetworkQuery:
    """The `network_query_fn` closure of create_nerf (nerf/run.py:248-251) as an object.

    Calling it keeps the reference signature `(inputs, viewdirs, network_fn)`.  `query_rays` is the
    fused entry render_rays uses: it takes the ray batch and z-values, so points / encodings never
    touch HBM.  precision: 'tc' = fused tcgen05 kernels (fp16 operands, fp32 accumulate) when the
    network has the shape they are built for; 'fp32' = the fp32-accumulate check path."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk=1024 * 64, precision=None):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk
        self.precision = precision or os.environ.get("SWNERF_PRECISION", "tc")
        if self.precision not in ("tc", "fp32"):
            raise ValueError("precision must be 'tc' or 'fp32'")
        self.allow_fused = True          # False: 'tc' runs layer by layer on the tcgen05 GEMM even where a fused kernel exists (tests)

    def __call__(self, inputs, viewdirs, network_fn):
        """Reference signature (nerf/run.py:248).  Also accepts the 2-D point lists of the mesh tools
        (nerf/load_model.py:56-74, nerf/extract_mesh.py:176: positions [M,3], viewdirs [M,3])."""
        if viewdirs is not None and inputs.is_cuda and self.uses_tc(network_fn, True):
            # fused kernel in explicit-points mode: one "ray" per row of viewdirs, S points each
            squeeze = inputs.dim() == 2
            pts3 = inputs[:, None] if squeeze else inputs
            n, s_ = pts3.shape[0], pts3.shape[1]
            vd = viewdirs.reshape(n, 3).float().contiguous()
            params = network_fn.param_list()
            training = torch.is_grad_enabled() and (inputs.requires_grad or any(p.requires_grad for p in params))
            out = tc.TcOccPointsFn.apply(network_fn, vd, pts3.reshape(-1, 3).float(), s_, 0, 0.0, training,
                                         self._enc(network_fn), *params)
            return out.reshape(n, 4) if squeeze else out
        self._arm(network_fn)
        if inputs.dim() == 2:                                  # load_model.py:57-58
            inputs = inputs[:, None]
            if viewdirs is not None and viewdirs.dim() == 3:
                viewdirs = viewdirs[:, 0]
            out = run_network(inputs, viewdirs, network_fn, embed_fn=self.embed_fn,
                              embeddirs_fn=self.embeddirs_fn, netchunk=self.netchunk)
            return out
        return run_network(inputs, viewdirs, network_fn, embed_fn=self.embed_fn,
                           embeddirs_fn=self.embeddirs_fn, netchunk=self.netchunk)

    def _arm(self, network_fn):
        """Shapes without a fused kernel run layer by layer: on the tcgen05 GEMM in 'tc' precision, fp32 SIMT in 'fp32'."""
        if hasattr(network_fn, "tc_gemm"):
            network_fn.tc_gemm = self.precision == "tc"

    def uses_tc(self, network_fn, has_views):
        if torch.is_grad_enabled() and not tc.bwd_available() and \
                any(p.requires_grad for p in network_fn.parameters()):
            return False
        return (self.precision == "tc" and self.allow_fused and tc.available() and has_views
                and getattr(network_fn, "tc_eligible", lambda: False)() and self._enc(network_fn) is not None)

    def _enc(self, network_fn):
        """Encoding code of the fused kernels for this query's embedders (None: not served, or widths do not match)."""
        return tc.enc_for(self.embed_fn, self.embeddirs_fn, None, network_fn)

    def query_rays(self, ray_batch, z_vals, network_fn, view_col):
        """raw[N, S, out] for points o + d*z of every ray (nerf/run.py:385-389 fused)."""
        N, S = z_vals.shape
        if self.uses_tc(network_fn, view_col >= 0):
            return tc.mlp_query(network_fn, ray_batch, z_vals, view_col, enc=self._enc(network_fn))
        self._arm(network_fn)
        L_pos = getattr(self.embed_fn, "L", None)
        L_dir = getattr(self.embeddirs_fn, "L", -1) if view_col >= 0 else -1
        if L_pos is None:
            raise TypeError("network_query_fn.embed_fn is not a swnerf_b200 embedder")
        emb = ops.encode_points(ray_batch, z_vals, L_pos, L_dir, view_col)
        in_pts = network_fn.input_ch
        x_views = emb[:, in_pts:] if (view_col >= 0 and network_fn.use_viewdirs) else None
        out = ops.mlp_fp32(network_fn.spec, emb[:, :in_pts], None, x_views, network_fn.param_list())
        return out.reshape(N, S, out.shape[-1])


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """nerf/run.py:90-102."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: torch.cat(all_ret[k], 0) for k in all_ret}


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """nerf/run.py:105-170."""
    if (c2w is not None and c2w_staticcam is None and torch.cuda.is_available()
            and not torch.is_tensor(near) and not torch.is_tensor(far)):
        # full-frame fast path: the whole ray assembly below (NDC warp included) is one kernel (same values,
        # tests/test_gpu_next_rows.py)
        rays = make_ray_batch(H, W, K, c2w, near, far, use_viewdirs=bool(use_viewdirs), ndc=bool(ndc))
        all_ret = batchify_rays(rays, chunk, **kwargs)
        for k in all_ret:
            all_ret[k] = torch.reshape(all_ret[k], [H, W] + list(all_ret[k].shape[1:]))
        k_extract = ['rgb_map', 'disp_map', 'acc_map']
        return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    sh = rays_d.shape
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near, far = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    rays = torch.cat([rays_o, rays_d, near, far], -1)
    if use_viewdirs:
        rays = torch.cat([rays, viewdirs], -1)
    all_ret = batchify_rays(rays, chunk, **kwargs)
    for k in all_ret:
        k_sh = list(sh[:-1]) + list(all_ret[k].shape[1:])
        all_ret[k] = torch.reshape(all_ret[k], k_sh)
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


def _query(network_query_fn, ray_batch, z_vals, network, view_col):
    if hasattr(network_query_fn, "query_rays"):
        return network_query_fn.query_rays(ray_batch, z_vals, network, view_col)
    # foreign closure with the reference signature: materialise the points (nerf/run.py:385)
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    viewdirs = ray_batch[:, view_col:view_col + 3] if view_col >= 0 else None
    return network_query_fn(pts, viewdirs, network)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False,
                pytest=False):
    """nerf/run.py:316-422.  ray_batch [N, 8 | 11]: o, d, near, far[, unit viewdir]."""
    if not ray_batch.is_contiguous():
        ray_batch = ray_batch.contiguous()
    ray_batch = ray_batch.float()
    N_rays, C = ray_batch.shape
    dev = ray_batch.device
    view_col = C - 3 if C > 8 else -1                                             # run.py:357

    t_rand = None
    if perturb > 0.:
        t_rand = pytest_uniform([N_rays, N_samples], dev) if pytest else \
            torch.rand((N_rays, N_samples), device=dev)                           # run.py:375-381
    z_vals = ops.stratified_z(ray_batch, N_samples, lindisp, perturb, t_rand, near_col=6)

    raw = _query(network_query_fn, ray_batch, z_vals, network_fn, view_col)
    noise = raw_noise((N_rays, N_samples), raw_noise_std, dev, pytest)
    rgb_map, disp_map, acc_map, weights, depth_map = ops.composite(raw, z_vals, ray_batch, 3, noise, white_bkgd)

    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0 = rgb_map, disp_map, acc_map
        det = (perturb == 0.)
        u = pytest_uniform([N_rays, N_importance], dev) if (pytest and not det) else None
        exact = getattr(network_query_fn, "precision", None) == "fp32"      # check mode: the reference's summation order
        _, z_vals, z_std = ops.resample(z_vals, weights.detach(), N_importance, det=det, u=u, want_samples=False,
                                        exact=exact)
        run_fn = network_fn if network_fine is None else network_fine
        raw = _query(network_query_fn, ray_batch, z_vals, run_fn, view_col)
        noise = raw_noise((N_rays, N_samples + N_importance), raw_noise_std, dev, pytest)
        rgb_map, disp_map, acc_map, weights, depth_map = ops.composite(raw, z_vals, ray_batch, 3, noise, white_bkgd)

    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map}
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        ret['rgb0'] = rgb_map_0
        ret['disp0'] = disp_map_0
        ret['acc0'] = acc_map_0
        ret['z_std'] = z_std                                                      # run.py:416
    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


def create_nerf(args, device=None):
    """nerf/run.py:222-311 -> (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer)."""
    if device is None:
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    embed_fn, input_ch = get_embedder(args.multires, input_dims=3, i=args.i_embed)
    input_ch_views = 0
    embeddirs_fn = None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, input_dims=3, i=args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    model = NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, output_ch=output_ch, skips=skips,
                 input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(device)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, input_ch=input_ch, output_ch=output_ch,
                          skips=skips, input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(device)
        grad_vars += list(model_fine.parameters())

    network_query_fn = NetworkQuery(embed_fn, embeddirs_fn, args.netchunk,
                                    precision=getattr(args, "swnerf_precision", None))
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

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
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])

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
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        print('Not ndc!')
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    render_kwargs_test = {k: render_kwargs_train[k] for k in render_kwargs_train}
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer
