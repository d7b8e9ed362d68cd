"""Drop-in for the reference's model.py: vallina_NeRF, NeRFOriginal, DirectTemporalNeRF, NeRF.get_by_name.

Same constructor signatures, same state_dict keys/shapes and the same default initialisation as the
reference (model.py:10-62, 93-151, 214-296), so reference checkpoints load and vice versa.  The
modules only HOLD fp32 master parameters (nn.Linear is used as a container); forward runs on the
library's kernels: the fp32 SIMT GEMM chain here, or - through render.NetworkQuery - the fused
tcgen05 kernel that takes rays instead of embedded points.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .ops import MLPSpec


class _NerfBase(nn.Module):
    # layer-at-a-time path on the tcgen05 GEMM (fp16 operands) instead of the fp32 SIMT GEMM: set by the network query
    # objects from their `precision` when the shape has no fused kernel
    tc_gemm = False

    def _build(self, D, W, input_ch, input_ch_views, output_ch, skips, use_viewdirs, output_color_ch=3):
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips, self.use_viewdirs = skips, use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, output_color_ch)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self.output_ch = output_ch

    @property
    def spec(self) -> MLPSpec:
        head = "viewdirs" if self.use_viewdirs else "output"
        return MLPSpec(self.D, self.W, self.input_ch, 0, self.input_ch_views, tuple(self.skips), head,
                       self.output_ch, tc=self.tc_gemm)

    def param_list(self):
        """Parameters in the order of include/swnerf_b200.h (trunk, views, feature, alpha, rgb)."""
        ps = []
        for l in self.pts_linears:
            ps += [l.weight, l.bias]
        if self.use_viewdirs:
            ps += [self.views_linears[0].weight, self.views_linears[0].bias,
                   self.feature_linear.weight, self.feature_linear.bias,
                   self.alpha_linear.weight, self.alpha_linear.bias,
                   self.rgb_linear.weight, self.rgb_linear.bias]
        else:
            ps += [self.output_linear.weight, self.output_linear.bias]
        return ps

    def tc_eligible(self) -> bool:
        """Shape the fused tcgen05 kernels are instantiated for (8x256, skips [4], view branch; configs/lego.txt).  The
        input widths must be those of encodings the kernels serve (tc.enc_for checks them against the embedders)."""
        widths = (3, 27, 63, 123)                              # 3 (1 + 2 L), L in {identity, 4, 10, 20}
        return (self.use_viewdirs and self.D == 8 and self.W == 256 and list(self.skips) == [4]
                and self.input_ch in widths and self.input_ch_views in widths and self.rgb_linear.out_features == 3)

    def _forward_embedded(self, x):
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        x_pts = x2[:, :self.input_ch]
        x_views = x2[:, self.input_ch:self.input_ch + self.input_ch_views] if self.use_viewdirs else None
        out = ops.mlp_fp32(self.spec, x_pts, None, x_views, self.param_list())
        return out.reshape(list(x.shape[:-1]) + [out.shape[-1]])

    def load_weights_from_keras(self, weights):                                     # model.py:64-91
        assert self.use_viewdirs, "Not implemented if use_viewdirs=False"
        t = lambda a: torch.from_numpy(np.transpose(a))
        for i in range(self.D):
            self.pts_linears[i].weight.data = t(weights[2 * i])
            self.pts_linears[i].bias.data = t(weights[2 * i + 1])
        k = 2 * self.D
        self.feature_linear.weight.data, self.feature_linear.bias.data = t(weights[k]), t(weights[k + 1])
        self.views_linears[0].weight.data, self.views_linears[0].bias.data = t(weights[k + 2]), t(weights[k + 3])
        self.rgb_linear.weight.data, self.rgb_linear.bias.data = t(weights[k + 4]), t(weights[k + 5])
        self.alpha_linear.weight.data, self.alpha_linear.bias.data = t(weights[k + 6]), t(weights[k + 7])


class vallina_NeRF(_NerfBase):
    """model.py:10-62."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self._build(D, W, input_ch, input_ch_views, output_ch, skips, use_viewdirs)

    def forward(self, x):
        return self._forward_embedded(x)


class NeRFOriginal(_NerfBase):
    """model.py:227-296: the same network, kaiming-normal weights, forward(x, ts) -> (out, zeros)."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, input_ch_time=1, output_ch=4, skips=[4],
                 use_viewdirs=False, memory=[], embed_fn=None, output_color_ch=3, zero_canonical=True):
        super().__init__()
        if any(i in memory for i in range(D - 1)):
            raise NotImplementedError
        self._build(D, W, input_ch, input_ch_views, output_ch, skips, use_viewdirs, output_color_ch)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in')                 # model.py:270-272

    def forward(self, x, ts):
        out = self._forward_embedded(x)
        return out, torch.zeros_like(x[..., :3])


class DirectTemporalNeRF(nn.Module):
    """model.py:93-151: deformation network (x, t) -> dx, then the canonical NeRFOriginal at x + dx."""
    tc_gemm = False           # as _NerfBase.tc_gemm; forwarded to the canonical network

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, input_ch_time=1, output_ch=4, skips=[4],
                 use_viewdirs=False, memory=[], embed_fn=None, zero_canonical=True):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views, self.input_ch_time = input_ch, input_ch_views, input_ch_time
        self.skips, self.use_viewdirs, self.memory = skips, use_viewdirs, memory
        self.embed_fn, self.zero_canonical = embed_fn, zero_canonical
        self._occ = NeRFOriginal(D=D, W=W, input_ch=input_ch, input_ch_views=input_ch_views,
                                 input_ch_time=input_ch_time, output_ch=output_ch, skips=skips,
                                 use_viewdirs=use_viewdirs, memory=memory, embed_fn=embed_fn, output_color_ch=3)
        self._time, self._time_out = self.create_time_net()

    def create_time_net(self):                                                     # model.py:112-126
        layers = [nn.Linear(self.input_ch + self.input_ch_time, self.W)]
        for i in range(self.D - 1):
            if i in self.memory:
                raise NotImplementedError
            in_channels = self.W + (self.input_ch if i in self.skips else 0)
            layers += [nn.Linear(in_channels, self.W)]
        return nn.ModuleList(layers), nn.Linear(self.W, 3)

    @property
    def time_spec(self) -> MLPSpec:
        return MLPSpec(self.D, self.W, self.input_ch, self.input_ch_time, 0, tuple(self.skips), "linear", 3,
                       tc=self.tc_gemm)

    def time_param_list(self):
        ps = []
        for l in self._time:
            ps += [l.weight, l.bias]
        return ps + [self._time_out.weight, self._time_out.bias]

    def query_time(self, new_pts, t, net=None, net_final=None):                    # model.py:128-136
        return ops.mlp_fp32(self.time_spec, new_pts, t, None, self.time_param_list())

    def forward(self, x, ts, cur_time=None):
        """x: embedded [pts | views]; ts[0]: embedded time [M, input_ch_time] (all rows equal).
        `cur_time` lets the caller pass the scalar it already holds on the host, which removes the
        reference's two device->host syncs per query (model.py:142-144, run_dnerf.py:53-54)."""
        x = x.reshape(-1, x.shape[-1])
        input_pts = x[:, :self.input_ch]
        input_views = x[:, self.input_ch:self.input_ch + self.input_ch_views]
        t = ts[0]
        if cur_time is None:
            assert len(torch.unique(t[:, :1])) == 1, "Only accepts all points from same time"
            cur_time = float(t[0, 0])
        if cur_time == 0. and self.zero_canonical:
            dx = torch.zeros_like(input_pts[:, :3])
        else:
            dx = self.query_time(input_pts, t)
            input_pts = self.embed_fn(input_pts[:, :3] + dx)                       # model.py:148-149
        occ = self._occ
        occ.tc_gemm = self.tc_gemm
        out = ops.mlp_fp32(occ.spec, input_pts, None, input_views if occ.use_viewdirs else None, occ.param_list())
        return out, dx


class TNeRF(nn.Module):
    """model.py:152-210: time-conditioned NeRF of t_nerf/run_tnerf.py - `depth` ELU layers of width `net_dim` on
    [emb_pts | emb_t], the input re-injected after layer 4, density / feature heads, a view branch (ELU) and a
    colour head that ends in a ReLU.  Same sub-module names as the reference, so state_dicts interchange.  The
    forward runs on the library's fp32 GEMM kernels with the ELU epilogue (ops.mlp_fp32)."""

    def __init__(self, depth, in_feat, dir_feat, time_feat, net_dim=128, skip_layer=4):
        super().__init__()
        self.depth, self.skip_layer, self.in_feat = depth, skip_layer, in_feat
        self.dir_feat, self.time_feat, self.net_dim = dir_feat, time_feat, net_dim
        # the reference WIDENS layer i when i % (skip_layer+1) == 0 (model.py:163) but CONCATENATES after layer i when
        # i % skip_layer == 0 (model.py:198): the two agree only for one re-injection after layer 4 (6 <= depth <= 8)
        widened = [i for i in range(1, depth) if i % (skip_layer + 1) == 0]
        concat_after = [i for i in range(1, depth) if i % skip_layer == 0]
        if [i + 1 for i in concat_after] != widened:
            raise ValueError("TNeRF: depth=%d, skip_layer=%d is inconsistent in the reference itself "
                             "(model.py:163 vs :198)" % (depth, skip_layer))
        self._skips = tuple(concat_after)
        units = [in_feat + time_feat] + [net_dim] * (depth + 1)
        self.layers = nn.ModuleList([])
        self.bnorm_layers = nn.ModuleList([])
        for i in range(depth):
            fan_in = units[i] + (in_feat + time_feat if i in widened else 0)
            self.layers.append(nn.Sequential(nn.Linear(fan_in, units[i + 1]), nn.ELU()))
        self.density = nn.Sequential(nn.Linear(net_dim, 1))
        self.feature = nn.Sequential(nn.Linear(net_dim, net_dim))
        self.layer_9 = nn.Sequential(nn.Linear(net_dim + dir_feat, net_dim // 2), nn.ELU())
        self.color = nn.Sequential(nn.Linear(net_dim // 2, 3), nn.ReLU())
        self.tc_gemm = False      # layers on the tcgen05 GEMM (fp16 operands); set by tnerf.TNerfNetworkQuery

    @property
    def spec(self) -> MLPSpec:
        return MLPSpec(self.depth, self.net_dim, self.in_feat, self.time_feat, self.dir_feat, self._skips,
                       "viewdirs", 4, act="elu", skip_extra=True, rgb_relu=True, tc=self.tc_gemm)

    def param_list(self):
        """trunk, view branch (layer_9), feature, density, colour - the order ops.MLPSpec documents."""
        ps = []
        for l in self.layers:
            ps += [l[0].weight, l[0].bias]
        return ps + [self.layer_9[0].weight, self.layer_9[0].bias, self.feature[0].weight, self.feature[0].bias,
                     self.density[0].weight, self.density[0].bias, self.color[0].weight, self.color[0].bias]

    def forward(self, inp, vdir, dyn_t):
        """inp [M, >= in_feat] (embedded points first), vdir [M, dir_feat], dyn_t [M, time_feat] -> [1, M, 4]
        (rgb after the ReLU, sigma), the reference's odd leading 1 included (model.py:205-208)."""
        M = inp.shape[0]
        x_pts = inp[:, :self.in_feat]
        out = ops.mlp_fp32(self.spec, x_pts, dyn_t, vdir if self.dir_feat else None, self.param_list())
        return out.reshape(-1, M, 4)


class NeRF:
    @staticmethod
    def get_by_name(type, *args, **kwargs):                                        # model.py:214-225
        print("NeRF type selected: %s" % type)
        if type == "original":
            model = NeRFOriginal(*args, **kwargs)
        elif type == "direct_temporal":
            model = DirectTemporalNeRF(*args, **kwargs)
        else:
            raise ValueError("Type %s not recognized." % type)
        return model
