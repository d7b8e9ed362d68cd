"""Host side of the fused tcgen05 path (points + PE + 8x256 MLP, forward and backward).

The kernels live in csrc/mlp_tc*.cu behind swnerf_tc_* (include/swnerf_b200.h).  The module keeps a
packed fp16 image of each network's weights, re-packed whenever a parameter changed (the optimizer
bumps the tensors' version counters); the fp32 nn.Parameters stay the master copy.
"""
import torch

from . import _lib
from ._lib import call, ptr, ptr_array, stream

F32 = torch.float32


def available() -> bool:
    return int(_lib.lib().swnerf_tc_packed_bytes()) > 0


def bwd_available() -> bool:
    return int(_lib.lib().swnerf_tc_packed_t_bytes()) > 0


class _Packed:
    __slots__ = ("versions", "fwd", "bwd", "bwd_versions")

    def __init__(self):
        self.versions = None
        self.fwd = None
        self.bwd = None
        self.bwd_versions = None


# bumped by optimizers that update the parameters behind autograd's back (parallel.FlatAdam)
GENERATION = 0


def _versions(params):
    return (GENERATION,) + tuple((p.data_ptr(), p._version) for p in params)


def packed_weights(network, need_bwd=False):
    params = network.param_list()
    st = getattr(network, "_swnerf_packed", None)
    if st is None:
        st = _Packed()
        object.__setattr__(network, "_swnerf_packed", st)
    v = _versions(params)
    dev = params[0].device
    if st.versions != v or st.fwd is None:
        for p in params:
            ptr(p, F32, "parameter")
        if st.fwd is None or st.fwd.device != dev:
            st.fwd = torch.empty(int(_lib.lib().swnerf_tc_packed_bytes()), dtype=torch.uint8, device=dev)
        call("swnerf_tc_pack_weights", ptr_array([p.detach() for p in params]), st.fwd.data_ptr(), stream())
        st.versions = v
    if need_bwd and (st.bwd_versions != v or st.bwd is None):
        if st.bwd is None or st.bwd.device != dev:
            st.bwd = torch.empty(int(_lib.lib().swnerf_tc_packed_t_bytes()), dtype=torch.uint8, device=dev)
        call("swnerf_tc_pack_weights_t", ptr_array([p.detach() for p in params]), st.fwd.data_ptr(),
             st.bwd.data_ptr(), stream())
        st.bwd_versions = v
    return st


class TcMlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, network, ray_batch, z_vals, view_col, grad_scale, training, *params):
        N, S = z_vals.shape
        dev = z_vals.device
        st = packed_weights(network, need_bwd=training)
        raw = torch.empty((N, S, 4), dtype=F32, device=dev)
        ws = None
        if training:
            nbytes = int(_lib.lib().swnerf_tc_workspace_bytes(N * S, 1))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        call("swnerf_tc_mlp_fwd", ptr(ray_batch, F32, "ray_batch"), ray_batch.shape[1], view_col,
             ptr(z_vals, F32, "z_vals"), N, S, st.fwd.data_ptr(), raw.data_ptr(),
             None if ws is None else ws.data_ptr(), int(training), stream())
        if training:
            ctx.network, ctx.ws, ctx.shape, ctx.params, ctx.grad_scale = network, ws, (N, S), params, grad_scale
            ctx.packed = (st.fwd, st.bwd)
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        N, S = ctx.shape
        params = ctx.params
        d_raw = d_raw.contiguous()
        # The kernels ACCUMULATE into the buffers they are given.  When every parameter already owns a
        # dense .grad (optimizer.zero_grad(set_to_none=False), or parallel.FlatGrads' views into the flat
        # all-reduce buffer) the gradients go straight there - the same result loss.backward() leaves in
        # .grad, without 24 zero-fills and 24 AccumulateGrad adds per network.
        direct = all(p.grad is not None and p.grad.dtype == F32 and p.grad.is_contiguous()
                     and p.grad.device == p.device for p in params) and not torch.is_grad_enabled()
        grads = [p.grad for p in params] if direct else [torch.zeros_like(p) for p in params]
        fwd, bwd = ctx.packed
        call("swnerf_tc_mlp_bwd", ptr(d_raw, F32, "d_raw"), N, S, fwd.data_ptr(), bwd.data_ptr(),
             ptr_array([p.detach() for p in params]), ctx.ws.data_ptr(), ptr_array(grads),
             float(ctx.grad_scale), stream())
        ctx.ws = None
        if direct:
            return (None,) * (6 + len(params))
        return (None, None, None, None, None, None) + tuple(grads)


def mlp_query(network, ray_batch, z_vals, view_col, grad_scale=None):
    if grad_scale is None:
        grad_scale = getattr(network, "grad_scale", 0.0)     # 0 = automatic (device-side max|d_raw|)
    params = network.param_list()
    training = torch.is_grad_enabled() and any(p.requires_grad for p in params)   # (grad mode is off inside forward)
    return TcMlpFn.apply(network, ray_batch, z_vals, view_col, grad_scale, training, *params)
