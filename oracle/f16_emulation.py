"""CPU emulation of the fused tcgen05 path's ARITHMETIC (fp16 operands, fp32 accumulation) for one 8x256 network.

TEST INFRASTRUCTURE ONLY (same rule as nerf_oracle.py: tests/, smoke() and bench.py's baseline legs may import it).

Why it exists: the north_star allows <= 1e-3 "with bf16/TF32 MMA" against the fp32 reference.  With 11-bit operand
mantissas that bound holds for the maps but NOT for the parameter gradients of the early layers - whatever kernel
computes them.  This file restates model.py:39-62 (and its autograd) with exactly the roundings the kernels apply:

  forward   x = fp16([PE(o + d z) | PE(viewdir)]);  h_l = fp16(relu(fp16(W_l) . h_{l-1} + b_l)) with fp32 accumulation;
            skip layer on [PE | h_4] (model.py:45-46); feature_linear folded into views_linears in fp32, THEN rounded
            (mlp_tc.cu: fold_head_kernel); sigma and the view branch from the same folded head; rgb_linear in fp32 on
            the UNROUNDED view-branch activation (mlp_tc.cu head epilogue); the backward sees fp16(h9).
  backward  d_raw scaled by 2^floor(log2(32 / max|d_raw|)); every dy_l rounded to fp16 before it is used as an operand
            (data gradient: dy_l . fp16(W_l); weight gradient: dy_l^T . h_{l-1}; bias gradient: column sums of the fp16
            dy_l), fp32 accumulation, un-scaled at the end; the folded head is un-folded in fp32 (unfold_head_kernel).

tests/test_parity_floors.py prints the per-tensor distance of this emulation from the fp32 oracle (the FLOOR of the
operand format); tests/test_gpu_render.py holds the CUDA kernels to this emulation at a tight tolerance and to the fp32
oracle at the floor.  The activations' accumulation ORDER differs from the tensor cores' (irrelevant at 1e-6).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import nerf_oracle as O

F32 = torch.float32


def h(x: torch.Tensor) -> torch.Tensor:
    """round to fp16 (saturating like cvt.rn.satfinite) and back"""
    return x.clamp(-65504.0, 65504.0).half().float()


def encode(rays: torch.Tensor, z: torch.Tensor):
    """[P,63] and [P,27] encodings of the sample points / view directions (nerf/run.py:385, :76-83), fp32."""
    o, d, vd = rays[:, 0:3], rays[:, 3:6], rays[:, -3:]
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).reshape(-1, 3)
    dirs = vd[:, None, :].expand(z.shape[0], z.shape[1], 3).reshape(-1, 3)
    return O.embed(pts, 10), O.embed(dirs, 4)


class Net:
    """One 8x256 network with the kernels' derived weights."""

    def __init__(self, p: Dict[str, torch.Tensor], prefix: str = "", round_fwd: bool = True, round_bwd: bool = True):
        """round_fwd / round_bwd = False keep that pass in fp32: attributes the gradient error to the pass that causes
        it (tests/test_parity_floors.py: it is the FORWARD's rounding - ReLU units whose pre-activation lies within the
        fp16 rounding of zero switch - not the backward's operand format)."""
        self.hf = h if round_fwd else (lambda t: t)
        self.hb = h if round_bwd else (lambda t: t)
        g = lambda n: p[prefix + n].detach().float()
        self.W = [g(f"pts_linears.{i}.weight") for i in range(8)]
        self.b = [g(f"pts_linears.{i}.bias") for i in range(8)]
        self.Wv, self.bv = g("views_linears.0.weight"), g("views_linears.0.bias")
        self.Wf, self.bf = g("feature_linear.weight"), g("feature_linear.bias")
        self.Wa, self.ba = g("alpha_linear.weight"), g("alpha_linear.bias")
        self.Wr, self.br = g("rgb_linear.weight"), g("rgb_linear.bias")
        # folded head (fp32, then rounded): W_fv = W_v[:, :256] W_f ; b_fv = W_v[:, :256] b_f + b_v
        self.Wfv = self.Wv[:, :256] @ self.Wf
        self.bfv = self.Wv[:, :256] @ self.bf + self.bv

    # ------------------------------------------------------------------ forward
    def forward(self, pe: torch.Tensor, ve: torch.Tensor):
        """pe [P,63], ve [P,27] fp32 -> raw [P,4]; keeps what the backward needs."""
        hf = self.hf
        x = hf(pe)
        v = hf(ve)
        acts, masks = [], []
        cur = x
        for l in range(8):
            W = hf(self.W[l])
            pre = cur @ W.t() + self.b[l]
            masks.append(pre >= 0)                 # the kernel's sign-bit mask counts +0 as active
            a = hf(torch.relu(pre))
            acts.append(a)
            cur = torch.cat([x, a], -1) if l == 4 else a
        h7 = acts[7]
        pre9 = h7 @ hf(self.Wfv).t() + v @ hf(self.Wv[:, 256:]).t() + self.bfv
        h9 = torch.relu(pre9)                      # fp32 in the epilogue registers
        sigma = h7 @ hf(self.Wa).t() + self.ba
        rgb = h9 @ self.Wr.t() + self.br           # fp32 weights, fp32 activation
        self.saved = dict(x=x, v=v, acts=acts, masks=masks, h9=hf(h9), mask9=pre9 > 0)
        return torch.cat([rgb, sigma], -1)

    # ------------------------------------------------------------------ backward
    def backward(self, d_raw: torch.Tensor) -> Dict[str, torch.Tensor]:
        s = self.saved
        h = self.hb                                                # noqa: F841 (shadows the module-level rounding)
        hw = self.hf                                               # the weight images are the forward's
        mx = float(d_raw.abs().max())
        scale = 2.0 ** max(-40.0, min(40.0, math.floor(math.log2(32.0 / mx)))) if mx > 0 else 1.0
        dr = d_raw * scale
        d_rgb, d_sig = dr[:, :3], dr[:, 3:4]
        dy9 = h((d_rgb @ self.Wr) * s["mask9"])                   # fp32 product, gated, rounded
        d_rgb16, d_sig16 = h(d_rgb), h(d_sig)
        G = {}
        G["rgb_linear.weight"] = d_rgb16.t() @ s["h9"]
        G["rgb_linear.bias"] = d_rgb16.sum(0)
        G["alpha_linear.weight"] = d_sig16.t() @ s["acts"][7]
        G["alpha_linear.bias"] = d_sig16.sum(0)
        gWfv = dy9.t() @ s["acts"][7]                             # d W_fv
        gbfv = dy9.sum(0)
        gWv_views = dy9.t() @ s["v"]
        # un-fold (fp32): dW_f = W_v[:, :256]^T G ; dW_v[:, :256] = G W_f^T + gb b_f^T ; db_f = W_v[:, :256]^T gb
        G["feature_linear.weight"] = self.Wv[:, :256].t() @ gWfv
        G["feature_linear.bias"] = self.Wv[:, :256].t() @ gbfv
        G["views_linears.0.weight"] = torch.cat([gWfv @ self.Wf.t() + gbfv[:, None] * self.bf[None, :], gWv_views], -1)
        G["views_linears.0.bias"] = gbfv
        dh = dy9 @ hw(self.Wfv) + d_sig16 @ hw(self.Wa)
        for l in range(7, -1, -1):
            dy = h(dh * s["masks"][l])
            xin = s["x"] if l == 0 else (torch.cat([s["x"], s["acts"][4]], -1) if l == 5 else s["acts"][l - 1])
            G[f"pts_linears.{l}.weight"] = dy.t() @ xin
            G[f"pts_linears.{l}.bias"] = dy.sum(0)
            if l > 0:
                W = hw(self.W[l])
                dh = dy @ (W[:, 63:] if l == 5 else W)
        return {k: v / scale for k, v in G.items()}


def render_fine_given_z(rays: torch.Tensor, z: torch.Tensor, p: Dict[str, torch.Tensor], cot_fn, white_bkgd=True,
                        round_fwd=True, round_bwd=True):
    """One network at given sample positions, composited by the fp32 oracle (ray.py:155-198); `cot_fn(maps) -> loss`.
    Returns (maps dict, raw, grads dict) with the gradients of the emulated arithmetic."""
    net = Net(p, round_fwd=round_fwd, round_bwd=round_bwd)
    pe, ve = encode(rays, z)
    raw = net.forward(pe, ve).reshape(z.shape[0], z.shape[1], 4).detach().requires_grad_()
    rgb, disp, acc, w, depth = O.raw2outputs(raw, z, rays[:, 3:6], 0.0, white_bkgd)
    maps = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, weights=w, depth_map=depth)
    loss = cot_fn(maps)
    loss.backward()
    grads = net.backward(raw.grad.reshape(-1, 4))
    return {k: v.detach() for k, v in maps.items()}, raw.detach(), grads
