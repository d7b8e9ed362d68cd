"""Host-side operators over the C ABI: thin wrappers and torch.autograd.Functions.

PyTorch is plumbing here (device memory, streams, autograd graph); every computation below is a
hand-written sm_100a kernel reached through include/swnerf_b200.h.  Nothing falls back to eager
PyTorch or to the CPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream

F32 = torch.float32


def _mat(t: torch.Tensor, name="matrix") -> Tuple[int, int]:
    """(device pointer, row stride) of a 2-D fp32 CUDA matrix with unit column stride.
    Row stride may exceed the width (a column slice of a wider buffer) or be 0 (broadcast row)."""
    if t.dim() != 2 or t.dtype != F32 or not t.is_cuda:
        raise TypeError("%s must be a 2-D fp32 CUDA tensor" % name)
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError("%s must have unit column stride" % name)
    return t.data_ptr(), t.stride(0)


def _rays(ray_batch: torch.Tensor) -> Tuple[int, int]:
    if ray_batch.dim() != 2:
        raise ValueError("ray batch must be [N, C]")
    return ptr(ray_batch, F32, "ray_batch"), ray_batch.shape[1]


# ------------------------------------------------------------------------------------------------
# a2  stratified sampling                                                    nerf/run.py:361-383
# ------------------------------------------------------------------------------------------------
def stratified_z(ray_batch, n_samples: int, lindisp=False, perturb=0.0, t_rand=None, near_col=6):
    p, stride = _rays(ray_batch)
    N = ray_batch.shape[0]
    z = torch.empty((N, n_samples), dtype=F32, device=ray_batch.device)
    do_perturb = perturb > 0.0
    if do_perturb and t_rand is None:
        t_rand = torch.rand((N, n_samples), dtype=F32, device=ray_batch.device)   # run.py:375
    call("swnerf_stratified_z", p, stride, near_col, ptr(t_rand, F32, "t_rand") if do_perturb else None,
         z.data_ptr(), N, n_samples, int(bool(lindisp)), int(do_perturb), stream())
    return z


# ------------------------------------------------------------------------------------------------
# a4  Embedder                                                               embedder.py:33-42
# ------------------------------------------------------------------------------------------------
class EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, L):
        shape = x.shape
        d = shape[-1]
        x2 = x.reshape(-1, d)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y = torch.empty((x2.shape[0], d * (1 + 2 * L)), dtype=F32, device=x.device)
        call("swnerf_embed_fwd", ptr(x2, F32, "x"), y.data_ptr(), x2.shape[0], d, L, stream())
        ctx.save_for_backward(x2)
        ctx.L, ctx.shape = L, shape
        return y.reshape(list(shape[:-1]) + [y.shape[1]])

    @staticmethod
    def backward(ctx, dy):
        (x2,) = ctx.saved_tensors
        d = x2.shape[1]
        dy2 = dy.reshape(x2.shape[0], -1).contiguous()
        dx = torch.empty_like(x2)
        call("swnerf_embed_bwd", x2.data_ptr(), ptr(dy2, F32, "dy"), dx.data_ptr(), x2.shape[0], d, ctx.L, stream())
        return dx.reshape(ctx.shape), None


def embed(x, L: int):
    if L < 0:
        return x
    return EmbedFn.apply(x, L)


def encode_points(ray_batch, z_vals, L_pos: int, L_dir: int, view_col: int):
    """points + PE(points) + PE(viewdir) -> [N*S, in_pts + in_views] (nerf/run.py:385, 76-83)."""
    p, stride = _rays(ray_batch)
    N, S = z_vals.shape
    in_pts = 3 * (1 + 2 * max(L_pos, 0))
    in_views = 3 * (1 + 2 * max(L_dir, 0)) if view_col >= 0 else 0
    out = torch.empty((N * S, in_pts + in_views), dtype=F32, device=ray_batch.device)
    call("swnerf_encode_points", p, stride, view_col, ptr(z_vals, F32, "z_vals"), out.data_ptr(), N, S, L_pos,
         L_dir, out.shape[1], stream())
    return out


# ------------------------------------------------------------------------------------------------
# a8  raw2outputs                                                            ray.py:155-198
# ------------------------------------------------------------------------------------------------
class CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z_vals, rays, d_col, noise, white_bkgd):
        N, S = z_vals.shape
        dev = raw.device
        if raw.dim() != 3 or tuple(raw.shape[:2]) != (N, S) or raw.shape[-1] < 4:
            raise ValueError("raw must be [N, S, C >= 4] matching z_vals [N, S] (ray.py:175-186 reads channels 0..3); "
                             "got %s for z_vals %s" % (tuple(raw.shape), tuple(z_vals.shape)))
        raw = raw if raw.is_contiguous() else raw.contiguous()
        rgb = torch.empty((N, 3), dtype=F32, device=dev)
        disp = torch.empty((N,), dtype=F32, device=dev)
        acc = torch.empty((N,), dtype=F32, device=dev)
        depth = torch.empty((N,), dtype=F32, device=dev)
        weights = torch.empty((N, S), dtype=F32, device=dev)
        rp, stride = _rays(rays)
        call("swnerf_composite_fwd", ptr(raw, F32, "raw"), raw.shape[-1], ptr(z_vals, F32, "z_vals"), rp, stride, d_col,
             ptr(noise, F32, "noise", allow_none=True), int(bool(white_bkgd)), N, S, rgb.data_ptr(),
             disp.data_ptr(), acc.data_ptr(), weights.data_ptr(), depth.data_ptr(), stream())
        ctx.save_for_backward(raw, z_vals, rays, noise, acc, depth)
        ctx.d_col, ctx.white = d_col, bool(white_bkgd)
        ctx.mark_non_differentiable()
        # outputs the loss does not use (disp, acc, weights, depth in training) reach backward as None instead of
        # zero-filled tensors: four fill kernels and four gradient streams less per call
        ctx.set_materialize_grads(False)
        return rgb, disp, acc, weights, depth

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_w, g_depth):
        raw, z_vals, rays, noise, acc, depth = ctx.saved_tensors
        N, S = z_vals.shape

        def c(g):
            return None if g is None else ptr(g.contiguous(), F32, "grad")
        keep = [None if g is None else g.contiguous() for g in (g_rgb, g_disp, g_acc, g_w, g_depth)]
        if all(g is None for g in keep):
            return None, None, None, None, None, None
        d_raw = torch.empty_like(raw)
        rp, stride = _rays(rays)
        call("swnerf_composite_bwd", raw.data_ptr(), raw.shape[-1], z_vals.data_ptr(), rp, stride, ctx.d_col,
             None if noise is None else noise.data_ptr(), int(ctx.white), N, S,
             *[None if g is None else g.data_ptr() for g in keep],
             acc.data_ptr(), depth.data_ptr(), d_raw.data_ptr(), stream())
        return d_raw, None, None, None, None, None


def composite(raw, z_vals, rays, d_col=3, noise=None, white_bkgd=False):
    """rays: any [N, C] tensor whose columns d_col:d_col+3 hold rays_d."""
    return CompositeFn.apply(raw, z_vals, rays, d_col, noise, white_bkgd)


# ------------------------------------------------------------------------------------------------
# a9 / a10 / a11 / a13
# ------------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_samples: int, det=False, u=None, cdf=None, return_inds=False):
    N, M = bins.shape
    dev = bins.device
    if not det and u is None:
        u = torch.rand((N, n_samples), dtype=F32, device=dev)                    # ray.py:121
    samples = torch.empty((N, n_samples), dtype=F32, device=dev)
    inds = torch.empty((N, n_samples), dtype=torch.int64, device=dev) if return_inds else None
    call("swnerf_sample_pdf", ptr(bins, F32, "bins"), ptr(weights, F32, "weights", allow_none=True),
         ptr(cdf, F32, "cdf", allow_none=True), None if det else ptr(u, F32, "u"), int(bool(det)), N, M,
         n_samples, samples.data_ptr(), None if inds is None else inds.data_ptr(), stream())
    return (samples, inds) if return_inds else samples


# SIMD lanes of the ATen CPU sum kernel behind the reference's torch.sum (ray.py:112): ATen registers its sum kernel
# for AVX2 only (8 fp32 lanes), AVX512 hosts included (pinned against torch.sum on the host by the CPU tests)
REF_SUM_LANES = 8


def resample(z_vals, weights, n_importance: int, det=False, u=None, want_samples=True, exact=False):
    """Returns (z_samples, z_fine, z_std): nerf/run.py:396-400 and :416.  z_samples comes back in ascending
    order (the reference uses it only through std() and the sort, both order-free); render_rays passes
    want_samples=False - it only needs z_std, which the kernel computes itself - and gets None for it.
    exact=True (the precision='fp32' check mode): the reference-order routine (swnerf_resample_check variant 1),
    whose z_fine is bit-identical to the reference's for identical weights."""
    N, S = z_vals.shape
    dev = z_vals.device
    if not det and u is None:
        u = torch.rand((N, n_importance), dtype=F32, device=dev)
    z_samples = torch.empty((N, n_importance), dtype=F32, device=dev) if want_samples else None
    z_fine = torch.empty((N, S + n_importance), dtype=F32, device=dev)
    z_std = torch.empty((N,), dtype=F32, device=dev)
    if exact and S <= 512:
        if det:     # the reference's u on its CPU path: torch.linspace's vectorised kernel (ray.py:118), not the closed form
            u, det = _ref_linspace(n_importance, dev).expand(N, n_importance).contiguous(), False
        call("swnerf_resample_check", ptr(z_vals, F32, "z_vals"), ptr(weights, F32, "weights"), None,
             None if det else ptr(u, F32, "u"), int(bool(det)), N, S, n_importance, 1, REF_SUM_LANES,
             z_samples.data_ptr() if want_samples else None, z_fine.data_ptr(), z_std.data_ptr(), None, None, stream())
        return z_samples, z_fine, z_std
    call("swnerf_resample", ptr(z_vals, F32, "z_vals"), ptr(weights, F32, "weights"),
         None if det else ptr(u, F32, "u"), int(bool(det)), N, S, n_importance,
         z_samples.data_ptr() if want_samples else None, z_fine.data_ptr(), z_std.data_ptr(), stream())
    return z_samples, z_fine, z_std


_LINSPACE = {}


def _ref_linspace(n, dev):
    """torch.linspace(0, 1, n) as the reference's CPU path produces it (ray.py:118): ATen's vectorised kernel evaluates
    base + k * step per SIMD block, one ulp off the closed form the device kernels (and torch's CUDA linspace) use."""
    key = (n, str(dev))
    if key not in _LINSPACE:
        _LINSPACE[key] = torch.linspace(0.0, 1.0, n, dtype=F32).to(dev)
    return _LINSPACE[key]


def resample_check(z_vals, weights, n_importance: int, det=False, u=None, cdf=None, variant=0, ref_lanes=REF_SUM_LANES):
    """Test entry (swnerf_resample_check): returns dict(z_samples, z_fine, z_std, inds, cdf).  `cdf` [N, S-1]
    replaces the cdf built from the weights; `inds` are torch.searchsorted(cdf, sort(u), right=True)."""
    N, S = z_vals.shape
    dev = z_vals.device
    if det and variant == 1:
        u, det = _ref_linspace(n_importance, dev).expand(N, n_importance).contiguous(), False
    out = dict(z_samples=torch.empty((N, n_importance), dtype=F32, device=dev),
               z_fine=torch.empty((N, S + n_importance), dtype=F32, device=dev),
               z_std=torch.empty((N,), dtype=F32, device=dev),
               inds=torch.empty((N, n_importance), dtype=torch.int64, device=dev),
               cdf=torch.empty((N, S - 1), dtype=F32, device=dev))
    call("swnerf_resample_check", ptr(z_vals, F32, "z_vals"), ptr(weights, F32, "weights"),
         ptr(cdf, F32, "cdf", allow_none=True), None if det else ptr(u, F32, "u"), int(bool(det)), N, S, n_importance,
         int(variant), int(ref_lanes), out["z_samples"].data_ptr(), out["z_fine"].data_ptr(), out["z_std"].data_ptr(),
         out["inds"].data_ptr(), out["cdf"].data_ptr(), stream())
    return out


def searchsorted(a, v, out=None, side="left"):
    """torchsearchsorted.searchsorted (searchsorted.py:20-52), same asserts and result."""
    assert len(a.shape) == 2, "input `a` must be 2-D."
    assert len(v.shape) == 2, "input `v` mus(t be 2-D."
    assert (a.shape[0] == v.shape[0] or a.shape[0] == 1 or v.shape[0] == 1), (
        "`a` and `v` must have the same number of rows or one of them must have only one ")
    assert a.device == v.device, "`a` and `v` must be on the same device"
    result_shape = (max(a.shape[0], v.shape[0]), v.shape[1])
    if out is not None:
        assert out.device == a.device, "`out` must be on the same device as `a`"
        assert out.dtype == torch.long, "out.dtype must be torch.long"
        assert out.shape == result_shape, "If the output tensor is provided, its shape must be correct."
    else:
        out = torch.empty(result_shape, device=v.device, dtype=torch.long)
    call("swnerf_searchsorted", ptr(a, F32, "a"), ptr(v, F32, "v"), ptr(out, torch.int64, "out"), a.shape[0],
         v.shape[0], a.shape[1], v.shape[1], 1 if side == "left" else 0, stream())
    return out


# ------------------------------------------------------------------------------------------------
# a6 / a7  fp32 layered MLP (check path + general shapes)                    model.py:39-62, 128-136
# ------------------------------------------------------------------------------------------------
_ACT = {None: 0, False: 0, True: 1, "relu": 1, "elu": 2}


def _gemm(op, A, B, C, M, N, K, bias=None, accumulate=False, relu=False, mask=None, mask_act="relu", tc=False,
          a_scale=1.0, a_scale_dev=None):
    """A, B, C, mask: (ptr, ld) pairs.  relu: False / True / "relu" / "elu" (the epilogue activation);
    mask_act: which activation produced `mask` (its derivative scales the result).  tc: run ops 0 / 1 on the
    tensor cores (fp16 operands, fp32 accumulation) when the shape allows, else on the fp32 SIMT kernel."""
    flags = _ACT[relu] | ((1 if mask_act == "elu" else 0) << 4)
    if tc and op in (0, 1) and 16 <= N <= 256 and 1 <= K <= 256:
        call("swnerf_hgemm_tc", op, A[0], A[1], B[0], B[1], C[0], C[1], M, N, K, bias, int(accumulate), flags,
             None if mask is None else mask[0], 0 if mask is None else mask[1], float(a_scale), a_scale_dev, stream())
        return
    if tc and op == 2 and accumulate and 32 <= M <= 256 and 1 <= N <= 256:
        # C[M = out channels, N = in channels] += A[K = samples, M]^T . B[K, N]
        call("swnerf_hgemm_tc_wgrad", A[0], A[1], B[0], B[1], C[0], C[1], K, M, N, float(a_scale), a_scale_dev, stream())
        return
    call("swnerf_sgemm", op, A[0], A[1], B[0], B[1], C[0], C[1], M, N, K, bias, int(accumulate), flags,
         None if mask is None else mask[0], 0 if mask is None else mask[1], stream())


def _colsum(x, rows, cols, out, accumulate=True):
    call("swnerf_colsum", x[0], x[1], rows, cols, out, int(accumulate), stream())


def _off(m, cols):
    """(ptr, ld) of the column block starting at column `cols` of matrix m=(ptr, ld)."""
    return (m[0] + 4 * cols, m[1])


class MLPSpec:
    """Shape of one reference MLP (model.py:22-37, 113-126).

    head = 'viewdirs' (feature/alpha/views/rgb, model.py:48-58), 'output' (output_linear, :60)
    or 'linear' (the D-NeRF _time_out layer, model.py:126,136).
    Parameter order: trunk (w, b) x D, then for 'viewdirs': views_linears.0, feature_linear,
    alpha_linear, rgb_linear; for 'output' / 'linear': the output layer."""

    def __init__(self, D=8, W=256, in_pts=63, in_extra=0, in_views=27, skips=(4,), head="viewdirs", out_ch=4,
                 act="relu", skip_extra=False, rgb_relu=False, tc=False):
        """act: trunk / view-branch activation ("relu", or "elu" for TNeRF, model.py:163-171,182);
        skip_extra: the skip concatenation re-injects [pts | extra] instead of pts alone (TNeRF, model.py:189-199);
        rgb_relu: the colour head ends in a ReLU (TNeRF, model.py:183-186)."""
        self.D, self.W, self.in_pts, self.in_extra, self.in_views = D, W, in_pts, in_extra, in_views
        self.skips, self.head, self.out_ch = tuple(skips), head, out_ch
        self.act, self.skip_extra, self.rgb_relu = act, bool(skip_extra), bool(rgb_relu)
        self.tc = bool(tc)              # forward layers on the tcgen05 GEMM (fp16 operands) instead of fp32 SIMT
        if act not in ("relu", "elu"):
            raise ValueError("act must be 'relu' or 'elu'")

    @property
    def skip_in(self):
        return self.in_pts + (self.in_extra if self.skip_extra else 0)

    def trunk_in(self, i):
        if i == 0:
            return self.in_pts + self.in_extra
        return self.W + self.skip_in if (i - 1) in self.skips else self.W

    @property
    def out_dim(self):
        return 4 if self.head == "viewdirs" else self.out_ch

    @property
    def n_params(self):
        return 2 * self.D + (8 if self.head == "viewdirs" else 2)


class MLPFp32Fn(torch.autograd.Function):
    """Whole-network forward/backward, one GEMM per layer: on the fp32 SIMT kernel (spec.tc False: the check mode, true
    fp32 like the reference's nn.Linear), or with the forward and data-gradient GEMMs on the tcgen05 kernel (spec.tc
    True: fp16 operands, fp32 accumulation; weight gradients stay fp32)."""

    @staticmethod
    def forward(ctx, spec: MLPSpec, need_grad, x_pts, x_extra, x_views, *params):
        s = spec
        dev = x_pts.device
        M = x_pts.shape[0]
        W = s.W
        P = [(p.data_ptr(), p.stride(0) if p.dim() == 2 else 0) for p in params]
        for p in params:
            ptr(p, F32, "parameter")
        xp = _mat(x_pts, "x_pts")
        xe = _mat(x_extra, "x_extra") if s.in_extra else None
        xv = _mat(x_views, "x_views") if s.head == "viewdirs" and s.in_views else None
        hs: List[torch.Tensor] = []
        h_prev = None
        act = s.act
        tc = s.tc
        n_buf = s.D if need_grad else 2
        bufs = [torch.empty((M, W), dtype=F32, device=dev) for _ in range(min(n_buf, s.D))]
        for i in range(s.D):
            h = bufs[i % len(bufs)]
            hm = (h.data_ptr(), W)
            wi, bi = P[2 * i], P[2 * i + 1][0]
            if i == 0:
                if s.in_extra:
                    _gemm(0, xp, wi, hm, M, W, s.in_pts, tc=tc)
                    _gemm(0, xe, _off(wi, s.in_pts), hm, M, W, s.in_extra, bias=bi, accumulate=True, relu=act, tc=tc)
                else:
                    _gemm(0, xp, wi, hm, M, W, s.in_pts, bias=bi, relu=act, tc=tc)
            elif (i - 1) in s.skips:
                _gemm(0, xp, wi, hm, M, W, s.in_pts, tc=tc)
                if s.skip_extra and s.in_extra:
                    _gemm(0, xe, _off(wi, s.in_pts), hm, M, W, s.in_extra, accumulate=True, tc=tc)
                _gemm(0, h_prev, _off(wi, s.skip_in), hm, M, W, W, bias=bi, accumulate=True, relu=act, tc=tc)
            else:
                _gemm(0, h_prev, wi, hm, M, W, W, bias=bi, relu=act, tc=tc)
            h_prev = hm
            hs.append(h)
        out = torch.empty((M, s.out_dim), dtype=F32, device=dev)
        om = (out.data_ptr(), s.out_dim)
        feat = hv = None
        k = 2 * s.D
        if s.head == "viewdirs":
            wv, bv, wf, bf, wa, ba, wr, br = P[k], P[k + 1][0], P[k + 2], P[k + 3][0], P[k + 4], P[k + 5][0], \
                P[k + 6], P[k + 7][0]
            feat = torch.empty((M, W), dtype=F32, device=dev)
            hv = torch.empty((M, W // 2), dtype=F32, device=dev)
            fm, hvm = (feat.data_ptr(), W), (hv.data_ptr(), W // 2)
            _gemm(0, h_prev, wa, _off(om, 3), M, 1, W, bias=ba, tc=tc)                       # model.py:49
            _gemm(0, h_prev, wf, fm, M, W, W, bias=bf, tc=tc)                                # model.py:50
            if xv is not None:
                _gemm(0, fm, wv, hvm, M, W // 2, W, tc=tc)                                   # model.py:51-55
                _gemm(0, xv, _off(wv, W), hvm, M, W // 2, s.in_views, bias=bv, accumulate=True, relu=act, tc=tc)
            else:
                _gemm(0, fm, wv, hvm, M, W // 2, W, bias=bv, relu=act, tc=tc)
            _gemm(0, hvm, wr, om, M, 3, W // 2, bias=br, relu=s.rgb_relu, tc=tc)             # model.py:57 / :183-186
        else:
            wo, bo = P[k], P[k + 1][0]
            _gemm(0, h_prev, wo, om, M, s.out_dim, W, bias=bo, tc=tc)                        # model.py:60 / :136
        if need_grad:
            ctx.spec = s
            ctx.x = (x_pts, x_extra, x_views)
            ctx.saved = (hs, feat, hv)
            ctx.out = out if s.rgb_relu else None
            ctx.params = params
            ctx.pts_grad = x_pts.requires_grad
        return out

    @staticmethod
    def backward(ctx, d_out):
        s: MLPSpec = ctx.spec
        x_pts, x_extra, x_views = ctx.x
        hs, feat, hv = ctx.saved
        params = ctx.params
        dev = d_out.device
        M, W = x_pts.shape[0], s.W
        d_out = d_out.contiguous()
        P = [(p.data_ptr(), p.stride(0) if p.dim() == 2 else 0) for p in params]
        G = [torch.zeros_like(p) for p in params]
        GP = [(g.data_ptr(), g.stride(0) if g.dim() == 2 else 0) for g in G]
        xp = _mat(x_pts)
        xe = _mat(x_extra) if s.in_extra else None
        xv = _mat(x_views) if s.head == "viewdirs" and s.in_views else None
        act = s.act
        if s.head == "viewdirs" and s.rgb_relu:                 # d_rgb through the colour head's ReLU (model.py:185)
            d_in = d_out
            d_out = d_in.clone()
            call("swnerf_act_bwd", d_in.data_ptr(), s.out_dim, ctx.out.data_ptr(), s.out_dim, M, 3, 0,
                 d_out.data_ptr(), s.out_dim, stream())
        dm = (d_out.data_ptr(), s.out_dim)
        # data-gradient GEMMs on the tensor cores: one power-of-two scale, chosen on the device from max|d_out|, lifts
        # the gradients into fp16's normal range (the weight-gradient GEMMs stay fp32)
        tc, sc, scale_t = s.tc, None, None
        if tc:
            scale_t = torch.empty(1, dtype=F32, device=dev)
            call("swnerf_pow2_scale", d_out.data_ptr(), M * s.out_dim, 32.0, scale_t.data_ptr(), stream())
            sc = scale_t.data_ptr()
        ga = torch.empty((M, W), dtype=F32, device=dev)
        gb = torch.empty((M, W), dtype=F32, device=dev)
        g = (ga.data_ptr(), W)          # grad wrt pre-activation of the current trunk layer
        g_next = (gb.data_ptr(), W)
        hlast = (hs[-1].data_ptr(), W)
        k = 2 * s.D
        if s.head == "viewdirs":
            wv, wf, wa, wr = P[k], P[k + 2], P[k + 4], P[k + 6]
            fm, hvm = (feat.data_ptr(), W), (hv.data_ptr(), W // 2)
            d_hv = torch.empty((M, W // 2), dtype=F32, device=dev)
            dhm = (d_hv.data_ptr(), W // 2)
            d_alpha = _off(dm, 3)
            _gemm(2, dm, hvm, GP[k + 6], 3, W // 2, M, accumulate=True, tc=tc, a_scale_dev=sc)                # dW_rgb
            _colsum(dm, M, 3, G[k + 7].data_ptr())
            _gemm(1, dm, wr, dhm, M, W // 2, 3, mask=hvm, mask_act=act, tc=tc, a_scale_dev=sc)               # d_hv (masked)
            _gemm(2, dhm, fm, GP[k], W // 2, W, M, accumulate=True, tc=tc, a_scale_dev=sc)                    # dW_v[:, :W]
            if xv is not None:
                _gemm(2, dhm, xv, _off(GP[k], W), W // 2, s.in_views, M, accumulate=True, tc=tc, a_scale_dev=sc)
            _colsum(dhm, M, W // 2, G[k + 1].data_ptr())
            d_feat = torch.empty((M, W), dtype=F32, device=dev)
            dfm = (d_feat.data_ptr(), W)
            _gemm(1, dhm, wv, dfm, M, W, W // 2, tc=tc, a_scale_dev=sc)                                       # d_feature
            _gemm(2, dfm, hlast, GP[k + 2], W, W, M, accumulate=True, tc=tc, a_scale_dev=sc)                  # dW_f
            _colsum(dfm, M, W, G[k + 3].data_ptr())
            _gemm(2, d_alpha, hlast, GP[k + 4], 1, W, M, accumulate=True, tc=tc, a_scale_dev=sc)              # dW_a
            _colsum(d_alpha, M, 1, G[k + 5].data_ptr())
            _gemm(1, dfm, wf, g, M, W, W, tc=tc, a_scale_dev=sc)
            _gemm(1, d_alpha, wa, g, M, W, 1, accumulate=True, mask=hlast, mask_act=act, tc=tc, a_scale_dev=sc)
        else:
            wo = P[k]
            _gemm(2, dm, hlast, GP[k], s.out_dim, W, M, accumulate=True, tc=tc, a_scale_dev=sc)
            _colsum(dm, M, s.out_dim, G[k + 1].data_ptr())
            _gemm(1, dm, wo, g, M, W, s.out_dim, mask=hlast, mask_act=act, tc=tc, a_scale_dev=sc)
        d_pts = torch.zeros((M, s.in_pts), dtype=F32, device=dev) if ctx.pts_grad else None
        dpm = (d_pts.data_ptr(), s.in_pts) if d_pts is not None else None
        for i in range(s.D - 1, -1, -1):
            wi = P[2 * i]
            _colsum(g, M, W, G[2 * i + 1].data_ptr())
            hp = (hs[i - 1].data_ptr(), W) if i > 0 else None
            if i == 0:
                _gemm(2, g, xp, GP[0], W, s.in_pts, M, accumulate=True, tc=tc, a_scale_dev=sc)
                if s.in_extra:
                    _gemm(2, g, xe, _off(GP[0], s.in_pts), W, s.in_extra, M, accumulate=True, tc=tc, a_scale_dev=sc)
                if dpm is not None:
                    _gemm(1, g, wi, dpm, M, s.in_pts, W, accumulate=True, tc=tc, a_scale_dev=sc)
            elif (i - 1) in s.skips:
                _gemm(2, g, xp, GP[2 * i], W, s.in_pts, M, accumulate=True, tc=tc, a_scale_dev=sc)
                if s.skip_extra and s.in_extra:
                    _gemm(2, g, xe, _off(GP[2 * i], s.in_pts), W, s.in_extra, M, accumulate=True, tc=tc, a_scale_dev=sc)
                _gemm(2, g, hp, _off(GP[2 * i], s.skip_in), W, W, M, accumulate=True, tc=tc, a_scale_dev=sc)
                if dpm is not None:
                    _gemm(1, g, wi, dpm, M, s.in_pts, W, accumulate=True, tc=tc, a_scale_dev=sc)
                _gemm(1, g, _off(wi, s.skip_in), g_next, M, W, W, mask=hp, mask_act=act, tc=tc, a_scale_dev=sc)
                g, g_next = g_next, g
            else:
                _gemm(2, g, hp, GP[2 * i], W, W, M, accumulate=True, tc=tc, a_scale_dev=sc)
                _gemm(1, g, wi, g_next, M, W, W, mask=hp, mask_act=act, tc=tc, a_scale_dev=sc)
                g, g_next = g_next, g
        return (None, None, d_pts, None, None) + tuple(G)


def mlp_fp32(spec: MLPSpec, x_pts, x_extra, x_views, params: Sequence[torch.Tensor]):
    need_grad = torch.is_grad_enabled() and (x_pts.requires_grad or any(p.requires_grad for p in params))
    return MLPFp32Fn.apply(spec, need_grad, x_pts, x_extra, x_views, *params)
