"""Committed FLOOR measurements behind every tolerance in tests/test_gpu_render.py that is looser than the
north_star's wording (<= 1e-3 with 11-bit-mantissa MMA operands, <= 1e-5 in the fp32 check mode).  CPU only: these
are properties of the reference algorithm and of the operand format, not of any kernel.

1. Hierarchical resampling is ill-conditioned in the coarse weights: moving every coarse weight by ONE fp32 ulp - less
   than what a different GEMM summation order does - moves z_fine by ~5e-4 and the fine maps by ~2e-4; the reference
   evaluated in fp64 differs from itself in fp32 by ~2e-3.  So "maps <= 1e-5 end to end through the fine pass" is not a
   property the reference has against itself; it holds (and is tested) for the coarse maps and for the fine pass GIVEN
   the same sample positions.
2. Parameter gradients with fp16 (or TF32 / bf16) operands: the emulation of the operand roundings
   (oracle/f16_emulation.py) is 4e-3 .. 3e-2 away from the fp32 gradients per tensor whatever computes it; maps stay
   within 1e-3.  The CUDA path is held to the emulation (tight) and to the oracle at this floor.
"""
import numpy as np
import torch

from oracle import nerf_oracle as O
from oracle import f16_emulation as E


def _ulp_perturb(p, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in p.items():
        s = torch.randint(-1, 2, v.shape, generator=g).numpy().astype(np.int32)
        out[k] = torch.from_numpy((v.numpy().view(np.int32) + s).view(np.float32).copy())
    return out


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_resampling_conditioning_floor(capsys):
    N = 128
    rays = torch.from_numpy(O.blender_rays(N, 1))
    shapes = O.mlp_param_shapes()
    pc, pf = O.make_params(shapes, 23), O.make_params(shapes, 43)
    with torch.no_grad():
        ref = O.render_rays(rays, pc, pf, 64, 128, white_bkgd=True)
        per = O.render_rays(rays, _ulp_perturb(pc, 1), pf, 64, 128, white_bkgd=True)
        r64 = O.render_rays(rays.double(), {k: v.double() for k, v in pc.items()},
                            {k: v.double() for k, v in pf.items()}, 64, 128, white_bkgd=True)
    floor = {k: _rel(per[k], ref[k]) for k in ("rgb0", "rgb_map", "acc_map")}
    floor64 = {k: _rel(r64[k].float(), ref[k]) for k in ("rgb0", "rgb_map", "acc_map")}
    dz = float((per["z_vals"] - ref["z_vals"]).abs().max())
    with capsys.disabled():
        print("\n[floor] one-ulp coarse-weight perturbation: ", {k: "%.2e" % v for k, v in floor.items()}, "max |dz_fine| %.2e" % dz)
        print("[floor] fp64 reference vs fp32 reference:     ", {k: "%.2e" % v for k, v in floor64.items()})
    assert floor["rgb0"] < 1e-5 and floor64["rgb0"] < 1e-5          # the coarse pass IS a 1e-5-class computation
    assert floor["rgb_map"] > 1e-5 and floor64["rgb_map"] > 1e-4      # the fine pass is not, for the reference itself
    assert floor64["rgb_map"] < 5e-3 and floor64["acc_map"] < 5e-3    # ... which is the bound the fp32-mode tests use


def grad_floor_table(N=96, seed_rays=46, seed_net=55, round_fwd=True, round_bwd=True):
    """(per-tensor rel-L2 of the fp16-operand emulation against the fp32 oracle, flat rel-L2, map errors)."""
    rays = torch.from_numpy(O.blender_rays(N, seed_rays))
    shapes = O.mlp_param_shapes()
    pc, pf = O.make_params(shapes, 21), O.make_params(shapes, seed_net)
    pfr = {k: v.clone().requires_grad_() for k, v in pf.items()}
    ref = O.render_rays(rays, pc, pfr, 64, 128, white_bkgd=True, retraw=True)
    cot = torch.from_numpy(np.random.RandomState(1).normal(size=(N, 3)).astype(np.float32))
    (ref["rgb_map"] * cot).sum().backward()
    z = ref["z_vals"].detach()
    maps, raw, grads = E.render_fine_given_z(rays, z, pf, lambda m: (m["rgb_map"] * cot).sum(),
                                             round_fwd=round_fwd, round_bwd=round_bwd)
    table = {}
    for n in O.mlp_param_names():
        table[n] = float((grads[n].reshape(-1) - pfr[n].grad.reshape(-1)).norm() / pfr[n].grad.norm())
    ge = torch.cat([grads[n].reshape(-1) for n in O.mlp_param_names()])
    gr = torch.cat([pfr[n].grad.reshape(-1) for n in O.mlp_param_names()])
    flat = float((ge - gr).norm() / gr.norm())
    maperr = {k: _rel(maps[k], ref[k].detach()) for k in ("rgb_map", "acc_map", "depth_map")}
    rawerr = float((raw - ref["raw"].detach()).norm() / ref["raw"].detach().norm())
    return table, flat, maperr, rawerr


def test_fp16_operand_gradient_floor(capsys):
    table, flat, maperr, rawerr = grad_floor_table()
    with capsys.disabled():
        print("\n[floor] fp16-operand emulation vs fp32 oracle, same sample positions: flat gradient rel-L2 %.2e, "
              "raw rel-L2 %.2e, maps %s" % (flat, rawerr, {k: "%.1e" % v for k, v in maperr.items()}))
        for n, v in table.items():
            if n.endswith("weight"):
                print("[floor]   %-28s %.2e" % (n, v))
    assert all(v < 1e-3 for v in maperr.values())                     # maps: the north_star's 1e-3 holds
    assert 1e-3 < flat < 1e-2                                          # gradients: 1e-3 is below the operand format's floor
    assert table["pts_linears.0.weight"] > 5e-3                        # ... and the early layers are the worst
    assert max(table.values()) < 6e-2


def test_gradient_floor_comes_from_the_forward_rounding(capsys):
    """Attribution: with an fp32 forward and fp16 operands in the backward only, the gradients are 10x closer to the
    oracle; with an fp16-operand forward and an fp32 backward they are as far as with both.  The floor is the FORWARD's:
    ReLU units whose pre-activation lies within the fp16 rounding of zero switch, which no backward can undo - so a
    wider backward format (hi/lo operand splits, fp32 dy) would not buy the north_star's 1e-3."""
    _, flat_b, _, _ = grad_floor_table(round_fwd=False, round_bwd=True)
    _, flat_f, _, _ = grad_floor_table(round_fwd=True, round_bwd=False)
    _, flat, _, _ = grad_floor_table()
    with capsys.disabled():
        print("\n[floor] flat gradient rel-L2: fp16 backward only %.2e | fp16 forward only %.2e | both %.2e" % (flat_b, flat_f, flat))
    assert flat_b < 5e-4 and flat_f > 0.8 * flat
