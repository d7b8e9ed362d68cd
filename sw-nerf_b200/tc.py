"""Host side of the fused tcgen05 path (points + PE + 8x256 MLP, forward and backward).

The kernels live in csrc/mlp_tc*.cu behind swnerf_tc_* (include/swnerf_b200.h).  The module keeps a
packed fp16 image of each network's weights, re-packed whenever a parameter changed (the optimizer
bumps the tensors' version counters); the fp32 nn.Parameters stay the master copy.
"""
import weakref

import torch

from . import _lib
from ._lib import call, ptr, ptr_array, stream

F32 = torch.float32


# Encoding codes of the C ABI (SWNERF_TC_ENC in include/swnerf_b200.h): frequency counts of the position / view / time
# embedders, identity (get_embedder(..., i=-1)) = 0.  The kernels instantiate position / view encoders for L in
# {0, 4, 10, 20} (every embedder the reference's runners build: configs 10/4, the MultiRes pyramid 20/8/20, 10/4/10, identity).
ENC_SUPPORTED_L = (0, 4, 10, 20)


def enc_code(pos_L, view_L, time_L=10) -> int:
    return max(int(pos_L), 0) | (max(int(view_L), 0) << 8) | (max(int(time_L), 0) << 16)


ENC_DEFAULT = enc_code(10, 4, 10)


def enc_for(embed_fn, embeddirs_fn, embedtime_fn=None, network=None):
    """The code for a set of embedders, or None when the fused kernels do not serve it (or `network`'s input widths do
    not match the embedders: then the caller falls back to the layer-wise path)."""
    def freq(f):
        L = getattr(f, "L", None)
        return None if L is None else max(int(L), 0)
    Lp, Lv = freq(embed_fn), freq(embeddirs_fn)
    Lt = 10 if embedtime_fn is None else freq(embedtime_fn)
    if Lp not in ENC_SUPPORTED_L or Lv not in ENC_SUPPORTED_L or Lt is None or Lt > 20:
        return None
    if network is not None:
        if getattr(network, "input_ch", None) != 3 * (1 + 2 * Lp) or getattr(network, "input_ch_views", None) != 3 * (1 + 2 * Lv):
            return None
        if embedtime_fn is not None and getattr(network, "input_ch_time", None) != 1 + 2 * Lt:
            return None
    return enc_code(Lp, Lv, Lt)


def available() -> bool:
    return int(_lib.lib().swnerf_tc_packed_bytes()) > 0


def bwd_available() -> bool:
    return int(_lib.lib().swnerf_tc_packed_t_bytes()) > 0


class _Token:
    """Held by the autograd ctx of a training forward until its backward has run (or the graph is dropped)."""
    __slots__ = ("__weakref__",)


class _Packed:
    """Per-network derived state: packed fp16 weight images and the training workspace.  A forward that will be
    differentiated takes a lease on both; while a lease is alive a re-pack (parameters changed, another frame time)
    or another training forward gets FRESH buffers instead of overwriting the ones the pending backward will read."""
    __slots__ = ("versions", "fwd", "bwd", "bwd_versions", "leases", "ws", "ws_leases")

    def __init__(self):
        self.versions = None
        self.fwd = None
        self.bwd = None
        self.bwd_versions = None
        self.leases = []
        self.ws = None
        self.ws_leases = []

    @staticmethod
    def _alive(leases):
        leases[:] = [r for r in leases if r() is not None]
        return bool(leases)

    def lease(self):
        tok = _Token()
        self.leases.append(weakref.ref(tok))
        return tok

    def release_if_leased(self):
        """Before overwriting the packed images: drop them (the pending backward keeps its own references)."""
        if self._alive(self.leases):
            self.fwd = self.bwd = None
            self.bwd_versions = None
            self.leases = []

    def workspace(self, nbytes, dev):
        """A workspace of at least nbytes that no pending backward reads: the cached one when it is free (the steady
        state of a training loop: one allocation for the whole run), otherwise a fresh one."""
        if self.ws is None or self.ws.numel() < nbytes or self.ws.device != dev or self._alive(self.ws_leases):
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self.ws_leases = []
        tok = _Token()
        self.ws_leases.append(weakref.ref(tok))
        return self.ws, tok


# bumped whenever parameters change behind autograd's back: parallel.FlatAdam.step(), and invalidate() below
GENERATION = 0


def invalidate(network=None):
    """Mark the packed fp16 images stale.  Needed only after edits the tensors' version counters do not see:
    `p.data.copy_ / mul_` (EMA updates), `p.data = ...` (load_weights_from_keras).  optimizer.step(), load_state_dict
    and every in-place op on the parameter itself are detected without it."""
    global GENERATION
    GENERATION += 1


def _versions(params):
    return (GENERATION,) + tuple((p.data_ptr(), p._version) for p in params)


def direct_grads(params, on=True):
    """Opt in (parallel.FlatGrads does) to the backward kernels accumulating straight into the dense fp32 `.grad` of
    these parameters instead of returning gradients to autograd: no zero-fill + AccumulateGrad per tensor, and the
    flat all-reduce buffer is written in place.  Autograd hooks on these parameters do NOT fire in this mode and
    torch.autograd.grad() sees None for them, which is why it is never inferred."""
    for p in params:
        p._swnerf_direct_grad = bool(on)


def _zero_grads(params):
    """Zero-initialised gradient tensors for a backward that returns them to autograd: views into ONE buffer (one fill
    kernel instead of one per parameter tensor; AccumulateGrad takes each view over as the parameter's .grad)."""
    total = sum(p.numel() for p in params)
    flat = torch.zeros(total, dtype=F32, device=params[0].device)
    out, off = [], 0
    for p in params:
        n = p.numel()
        out.append(flat[off:off + n].view(p.shape))
        off += n
    return out


def _grad_targets(params):
    direct = all(getattr(p, "_swnerf_direct_grad", False) and p.grad is not None and p.grad.dtype == F32
                 and p.grad.is_contiguous() and p.grad.device == p.device for p in params)
    return direct, ([p.grad for p in params] if direct else _zero_grads(params))


def _take_ws(ctx):
    ws = ctx.ws
    if ws is None:
        raise RuntimeError("swnerf_b200: backward through this network query ran twice (retain_graph=True?) - the "
                           "saved activations were released after the first pass; call the forward again")
    ctx.ws = ctx.ws_lease = ctx.lease = None
    return ws


def packed_weights(network, need_bwd=False, enc=ENC_DEFAULT):
    params = network.param_list()
    st = getattr(network, "_swnerf_packed", None)
    if st is None:
        st = _Packed()
        object.__setattr__(network, "_swnerf_packed", st)
    v = (enc,) + _versions(params)
    dev = params[0].device
    if st.versions != v or st.fwd is None:
        for p in params:
            ptr(p, F32, "parameter")
        st.release_if_leased()
        if st.fwd is None or st.fwd.device != dev:
            st.fwd = torch.empty(int(_lib.lib().swnerf_tc_packed_bytes()), dtype=torch.uint8, device=dev)
        call("swnerf_tc_pack_weights", ptr_array([p.detach() for p in params]), enc, st.fwd.data_ptr(), stream())
        st.versions = v
    if need_bwd and (st.bwd_versions != v or st.bwd is None):
        if st.bwd is None or st.bwd.device != dev:
            st.bwd = torch.empty(int(_lib.lib().swnerf_tc_packed_t_bytes()), dtype=torch.uint8, device=dev)
        call("swnerf_tc_pack_weights_t", ptr_array([p.detach() for p in params]), enc, st.fwd.data_ptr(),
             st.bwd.data_ptr(), stream())
        st.bwd_versions = v
    return st


class TcMlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, network, ray_batch, z_vals, view_col, grad_scale, training, enc, *params):
        N, S = z_vals.shape
        dev = z_vals.device
        st = packed_weights(network, need_bwd=training, enc=enc)
        raw = torch.empty((N, S, 4), dtype=F32, device=dev)
        ws = None
        if training:
            ws, ctx.ws_lease = st.workspace(int(_lib.lib().swnerf_tc_workspace_bytes(N * S, 1, enc)), dev)
        call("swnerf_tc_mlp_fwd", ptr(ray_batch, F32, "ray_batch"), ray_batch.shape[1], view_col,
             ptr(z_vals, F32, "z_vals"), N, S, st.fwd.data_ptr(), enc, raw.data_ptr(),
             None if ws is None else ws.data_ptr(), int(training), stream())
        if training:
            ctx.network, ctx.ws, ctx.shape, ctx.params, ctx.grad_scale = network, ws, (N, S), params, grad_scale
            ctx.enc = enc
            ctx.packed, ctx.lease = (st.fwd, st.bwd), st.lease()
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        N, S = ctx.shape
        params = ctx.params
        d_raw = d_raw.contiguous()
        # The kernels ACCUMULATE into the buffers they are given.  For parameters that opted in (direct_grads():
        # parallel.FlatGrads' views into the flat all-reduce buffer) the gradients go straight into .grad - the same
        # result loss.backward() leaves there, without 24 zero-fills and 24 AccumulateGrad adds per network.
        direct, grads = _grad_targets(params)
        if direct and torch.is_grad_enabled():                   # create_graph=True: autograd needs real outputs
            direct, grads = False, _zero_grads(params)
        fwd, bwd = ctx.packed
        ws = _take_ws(ctx)
        call("swnerf_tc_mlp_bwd", ptr(d_raw, F32, "d_raw"), N, S, fwd.data_ptr(), bwd.data_ptr(), ctx.enc,
             ptr_array([p.detach() for p in params]), ws.data_ptr(), ptr_array(grads),
             float(ctx.grad_scale), stream())
        if direct:
            return (None,) * (7 + len(params))
        return (None,) * 7 + tuple(grads)


def mlp_query(network, ray_batch, z_vals, view_col, grad_scale=None, enc=ENC_DEFAULT):
    if grad_scale is None:
        grad_scale = getattr(network, "grad_scale", 0.0)     # 0 = automatic (device-side max|d_raw|)
    params = network.param_list()
    training = torch.is_grad_enabled() and any(p.requires_grad for p in params)   # (grad mode is off inside forward)
    return TcMlpFn.apply(network, ray_batch, z_vals, view_col, grad_scale, training, enc, *params)


# ------------------------------------------------------------------------------------------------
# D-NeRF on the fused kernels: deformation network (x, t) -> dx and canonical network at x + dx
# ------------------------------------------------------------------------------------------------
def time_embedding(t: float, L: int = 10):
    """PE(t) as 1 + 2L fp32 values (embedder.py:33-42 applied to the single scalar time of the call)."""
    import numpy as np
    t32 = np.float32(t)
    vals = [t32]
    for k in range(L):
        a = np.float32(t32 * np.float32(2.0 ** k))
        vals += [np.sin(a, dtype=np.float32), np.cos(a, dtype=np.float32)]
    return [float(v) for v in vals]


def _tpe_device(model, t, L, dev):
    """PE(t) on the device, cached per (time, L, device): frame times repeat (one per training image), and a cached tensor
    keeps the call free of host-to-device copies, so a training step can be captured into a CUDA graph."""
    cache = getattr(model, "_swnerf_tpe", None)
    if cache is None:
        cache = {}
        object.__setattr__(model, "_swnerf_tpe", cache)
    key = (float(t), int(L), str(dev))
    v = cache.get(key)
    if v is None:
        if len(cache) > 4096:
            cache.clear()
        v = cache[key] = torch.tensor(time_embedding(t, L), dtype=F32, device=dev)
    return v


def dnerf_tc_eligible(model) -> bool:
    """DirectTemporalNeRF in the shape every reference D-NeRF / MultiRes config uses (8x256, skips [4]); the encoding
    widths are checked against the embedders by enc_for()."""
    occ = getattr(model, "_occ", None)
    return (occ is not None and occ.tc_eligible() and model.D == 8 and model.W == 256 and list(model.skips) == [4]
            and getattr(model.embed_fn, "L", None) is not None
            and model.input_ch == 3 * (1 + 2 * max(int(model.embed_fn.L), 0)))


def packed_time_weights(model, t: float, need_bwd=False, enc=ENC_DEFAULT):
    import ctypes
    params = model.time_param_list()
    st = getattr(model, "_swnerf_packed_time", None)
    if st is None:
        st = _Packed()
        object.__setattr__(model, "_swnerf_packed_time", st)
    v = (float(t), enc) + _versions(params)
    dev = params[0].device
    if st.versions != v or st.fwd is None:
        for p in params:
            ptr(p, F32, "parameter")
        st.release_if_leased()          # e.g. the tv-loss render at a neighbouring time before the backward
        if st.fwd is None or st.fwd.device != dev:
            st.fwd = torch.empty(int(_lib.lib().swnerf_tc_packed_bytes()), dtype=torch.uint8, device=dev)
        tpe = time_embedding(t, (enc >> 16) & 255)
        call("swnerf_tc_pack_weights_time", ptr_array([p.detach() for p in params]), (ctypes.c_float * len(tpe))(*tpe),
             enc, st.fwd.data_ptr(), stream())
        st.versions = v
    if need_bwd and (st.bwd_versions != v or st.bwd is None):
        if st.bwd is None or st.bwd.device != dev:
            st.bwd = torch.empty(int(_lib.lib().swnerf_tc_packed_t_bytes()), dtype=torch.uint8, device=dev)
        call("swnerf_tc_pack_weights_time_t", ptr_array([p.detach() for p in params]), enc, st.fwd.data_ptr(),
             st.bwd.data_ptr(), stream())
        st.bwd_versions = v
    return st


class TcTimeFn(torch.autograd.Function):
    """dx[N,S,3] = deformation network at (o + d z, t)  (model.py:128-136) on the fused forward kernel."""

    @staticmethod
    def forward(ctx, model, ray_batch, z_vals, view_col, t, grad_scale, training, enc, *params):
        N, S = z_vals.shape
        dev = z_vals.device
        st = packed_time_weights(model, t, need_bwd=training, enc=enc)
        dx = torch.empty((N, S, 3), dtype=F32, device=dev)
        ws = None
        if training:
            ws, ctx.ws_lease = st.workspace(int(_lib.lib().swnerf_tc_workspace_bytes(N * S, 1, enc)), dev)
        call("swnerf_tc_time_fwd", ptr(ray_batch, F32, "ray_batch"), ray_batch.shape[1], view_col,
             ptr(z_vals, F32, "z_vals"), N, S, st.fwd.data_ptr(), enc, dx.data_ptr(),
             None if ws is None else ws.data_ptr(), int(training), stream())
        if training:
            ctx.ws, ctx.shape, ctx.params, ctx.grad_scale = ws, (N, S), params, grad_scale
            ctx.packed, ctx.lease = (st.fwd, st.bwd), st.lease()
            ctx.tpe = _tpe_device(model, t, (enc >> 16) & 255, dev)
            ctx.enc = enc
        return dx

    @staticmethod
    def backward(ctx, d_dx):
        N, S = ctx.shape
        params = ctx.params
        d_dx = d_dx.contiguous()
        direct, grads = _grad_targets(params)
        fwd, bwd = ctx.packed
        ws = _take_ws(ctx)
        call("swnerf_tc_time_bwd", ptr(d_dx, F32, "d_dx"), N, S, fwd.data_ptr(), bwd.data_ptr(), ctx.enc,
             ptr_array([p.detach() for p in params]), ctx.tpe.data_ptr(), ws.data_ptr(), ptr_array(grads),
             float(ctx.grad_scale), stream())
        return (None,) * 8 + ((None,) * len(params) if direct else tuple(grads))


class TcOccPointsFn(torch.autograd.Function):
    """raw[N,S,4] = canonical network at explicit positions pts[N*S,3] (model.py:148-150), with d/d pts."""

    @staticmethod
    def forward(ctx, network, ray_batch, pts, n_samples, view_col, grad_scale, training, enc, *params):
        P = pts.shape[0]
        N = P // n_samples
        dev = pts.device
        pts = pts.contiguous()
        st = packed_weights(network, need_bwd=training, enc=enc)
        raw = torch.empty((N, n_samples, 4), dtype=F32, device=dev)
        ws = None
        if training:
            ws, ctx.ws_lease = st.workspace(int(_lib.lib().swnerf_tc_workspace_bytes(P, 1, enc)), dev)
        call("swnerf_tc_mlp_fwd_points", ptr(ray_batch, F32, "ray_batch"), ray_batch.shape[1], view_col,
             ptr(pts, F32, "pts"), N, n_samples, st.fwd.data_ptr(), enc, raw.data_ptr(),
             None if ws is None else ws.data_ptr(), int(training), stream())
        if training:
            ctx.ws, ctx.shape, ctx.params, ctx.grad_scale = ws, (N, n_samples), params, grad_scale
            ctx.enc = enc
            ctx.packed, ctx.lease = (st.fwd, st.bwd), st.lease()
            ctx.pts = pts
            ctx.pts_grad = ctx.needs_input_grad[2]
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        N, S = ctx.shape
        params = ctx.params
        d_raw = d_raw.contiguous()
        direct, grads = _grad_targets(params)
        fwd, bwd = ctx.packed
        d_pts = torch.empty_like(ctx.pts) if ctx.pts_grad else None
        ws = _take_ws(ctx)
        call("swnerf_tc_mlp_bwd_points", ptr(d_raw, F32, "d_raw"), N, S, fwd.data_ptr(), bwd.data_ptr(), ctx.enc,
             ptr_array([p.detach() for p in params]), ws.data_ptr(), ptr_array(grads), float(ctx.grad_scale),
             ctx.pts.data_ptr() if ctx.pts_grad else None, None if d_pts is None else d_pts.data_ptr(), stream())
        return (None, None, d_pts, None, None, None, None, None) + ((None,) * len(params) if direct else tuple(grads))


def dnerf_query(model, ray_batch, z_vals, view_col, cur_time: float, grad_scale=0.0, enc=ENC_DEFAULT):
    """(raw[N,S,4], dx[N,S,3]) of DirectTemporalNeRF.forward (model.py:138-151) for the rays' sample points."""
    N, S = z_vals.shape
    occ = model._occ
    occ_params = occ.param_list()
    grad_on = torch.is_grad_enabled()
    if cur_time == 0. and model.zero_canonical:                                  # model.py:144-145
        raw = mlp_query(occ, ray_batch, z_vals, view_col, grad_scale, enc)
        return raw, torch.zeros((N, S, 3), dtype=F32, device=z_vals.device)
    tparams = model.time_param_list()
    t_train = grad_on and any(p.requires_grad for p in tparams)
    dx = TcTimeFn.apply(model, ray_batch, z_vals, view_col, float(cur_time), grad_scale, t_train, enc, *tparams)
    base = ray_batch[:, None, 0:3] + ray_batch[:, None, 3:6] * z_vals[..., None]   # run_dnerf.py:455, mul then add
    pts = (base + dx).reshape(-1, 3)                                              # model.py:148
    o_train = grad_on and (pts.requires_grad or any(p.requires_grad for p in occ_params))
    raw = TcOccPointsFn.apply(occ, ray_batch, pts, S, view_col, grad_scale, o_train, enc, *occ_params)
    return raw, dx
