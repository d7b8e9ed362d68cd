"""GPU parity of the T-NeRF path (model.py:152-210, t_nerf/run_tnerf.py) against the
golden vectors of the unmodified reference (tests/golden/render_rays_tnerf.npz, oracle/make_golden.py) and the
oracle.  T-NeRF has no resampling step, so the fp32 path is compared at fp32-rounding tolerances end to end
(north_star: <= 1e-5 class in the fp32-accumulate mode; written below per quantity)."""
import os
from argparse import Namespace

import numpy as np
import pytest
import torch

import swnerf_b200 as S
from swnerf_b200 import ops, tnerf, _lib
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def relmax(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _args(tmp):
    os.makedirs(os.path.join(str(tmp), "e"), exist_ok=True)
    return Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=0, N_samples=64,
                     netdepth=8, netwidth=256, netchunk=65536, lrate=5e-4, ft_path=None, basedir=str(tmp),
                     expname="e", no_reload=True, perturb=1.0, white_bkgd=True, raw_noise_std=0.0,
                     dataset_type="blender", no_ndc=False, lindisp=False, nerf_type="tnerf",
                     do_half_precision=False, swnerf_precision="fp32")


def _model(seed):
    m = S.TNeRF(depth=8, in_feat=63, dir_feat=27, time_feat=21).to(DEV)
    m.load_state_dict({k: v.to(DEV) for k, v in O.make_params(O.tnerf_param_shapes(), seed).items()})
    return m


def test_tnerf_module_forward_golden(golden):
    g = golden("render_rays_tnerf")
    m = _model(int(g["seed"]))
    x, t = T(g["mlp/x"]), T(g["mlp/t"])
    before = _lib.launch_count()
    out = m(x, x[:, 63:], t)
    assert _lib.launch_count() > before                              # the library's kernels ran
    assert tuple(out.shape) == (1, x.shape[0], 4)                    # model.py:205-208
    # sigma carries a x24 gain in the synthetic scene; fp32 GEMM with another summation order than MKL
    assert relmax(out, torch.from_numpy(g["mlp/out"])) < 2e-5
    assert float(out[..., :3].detach().min()) >= 0.0


def test_tnerf_mlp_backward_vs_oracle():
    """ELU trunk, [pts | t] re-injection, ReLU colour head: gradients of every parameter and of the embedded
    points against autograd on the oracle's restatement."""
    rs = np.random.RandomState(3)
    M = 301
    x = rs.uniform(-1, 1, size=(M, 90)).astype(np.float32)
    tt = rs.uniform(-1, 1, size=(M, 21)).astype(np.float32)
    gout = rs.normal(size=(M, 4)).astype(np.float32)
    p = {k: v.requires_grad_() for k, v in O.make_params(O.tnerf_param_shapes(), 77).items()}
    xo = torch.from_numpy(x).requires_grad_()
    out_o = O.tnerf_forward(p, xo, xo[:, 63:].detach(), torch.from_numpy(tt))
    (out_o * torch.from_numpy(gout)).sum().backward()
    m = _model(77)
    xg = T(x[:, :63]).requires_grad_()
    out = ops.mlp_fp32(m.spec, xg, T(tt), T(x[:, 63:]), m.param_list())
    assert relmax(out, out_o) < 2e-5
    (out * T(gout)).sum().backward()
    for n, prm in m.named_parameters():
        ref = p[n].grad
        err = float((prm.grad.cpu() - ref).norm() / ref.norm().clamp_min(1e-12))
        assert err < 5e-4, (n, err)            # fp32, units within rounding of the ELU / ReLU kinks excepted
    errx = float((xg.grad.cpu() - xo.grad[:, :63]).norm() / xo.grad[:, :63].norm())
    assert errx < 5e-4, errx


@pytest.mark.parametrize("tag", ["det", "pert"])
def test_render_rays_tnerf_golden(golden, tag, tmp_path):
    g = golden("render_rays_tnerf")
    kw_train, kw_test, start, grad_vars, opt = tnerf.create_nerf(_args(tmp_path), device=torch.device(DEV))
    model = kw_train["network_fn"]
    assert isinstance(model, S.TNeRF) and kw_train["N_importance"] == 0 and start == 0
    assert kw_test["perturb"] is False and kw_test["raw_noise_std"] == 0.
    model.load_state_dict({k: v.to(DEV) for k, v in O.make_params(O.tnerf_param_shapes(), int(g["seed"])).items()})
    kw = dict(kw_test if tag == "det" else kw_train)
    kw.pop("use_viewdirs"); kw.pop("ndc")
    if tag == "pert":
        kw["raw_noise_std"] = 0.5              # as in oracle/make_golden.py (the pytest draw is scaled by it)
    rays, target = T(g[f"{tag}/rays"]), T(g[f"{tag}/target"])
    ret = tnerf.render_rays(rays, retraw=True, pytest=True, **kw)
    assert set(ret) == {"rgb_map", "disp_map", "acc_map", "z_vals", "raw"}       # run_tnerf.py:487-489
    assert relmax(ret["z_vals"], torch.from_numpy(g[f"{tag}/z_vals"])) < 1e-6
    assert relmax(ret["raw"], torch.from_numpy(g[f"{tag}/raw"])) < 5e-5           # x24 sigma gain, L=10 encoding
    for k in ["rgb_map", "acc_map", "disp_map"]:
        assert relmax(ret[k], torch.from_numpy(g[f"{tag}/{k}"])) < 2e-5, k
    loss = torch.mean((ret["rgb_map"] - target) ** 2)
    assert abs(loss.item() - float(g[f"{tag}/loss"])) < 1e-6
    loss.backward()
    num = den = 0.0
    for n, p in model.named_parameters():
        sub = p.grad.reshape(-1)[::251].cpu().double()
        ref = torch.from_numpy(g[f"{tag}/gsub/{n}"]).double()
        num += float((sub - ref).pow(2).sum()); den += float(ref.pow(2).sum())
        gn = float(g[f"{tag}/gnorm/{n}"])
        assert abs(float(p.grad.double().norm()) - gn) <= 1e-3 * max(gn, 1e-12), n
    assert (num / max(den, 1e-30)) ** 0.5 < 1e-3


def test_tnerf_reference_signature_query_equals_ray_entry(tmp_path):
    """network_query_fn(pts, viewdirs, frame_time, net) - what run_tnerf.py:474 calls - against the ray entry
    render_rays uses (encoding from rays, broadcast time row), plus chunked evaluation (netchunk)."""
    kw, _, _, _, _ = tnerf.create_nerf(_args(tmp_path), device=torch.device(DEV))
    model, q = kw["network_fn"], kw["network_query_fn"]
    model.load_state_dict({k: v.to(DEV) for k, v in O.make_params(O.tnerf_param_shapes(), 9).items()})
    rays = T(O.blender_rays(50, seed=4, frame_time=0.25))
    z = ops.stratified_z(rays, 64, near_col=6)
    with torch.no_grad():
        a = q.query_rays(rays, z, model, 9, 0.25)
        pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
        b = q(pts, rays[:, 9:12], rays[:, 8:9], model)
        q.netchunk = 1000                                           # ragged chunks
        c = q(pts, rays[:, 9:12], rays[:, 8:9], model)
    assert tuple(a.shape) == tuple(b.shape) == (50, 64, 4)
    assert relmax(a, b) < 2e-5 and relmax(c, b) < 1e-6
    with pytest.raises(AssertionError):                              # run_tnerf.py:53
        bad = rays[:, 8:9].clone(); bad[0] = 0.5
        q(pts, rays[:, 9:12], bad, model)


def test_tnerf_render_frame_and_checkpoint_roundtrip(tmp_path):
    args = _args(tmp_path)
    kw, kw_test, _, grad_vars, opt = tnerf.create_nerf(args, device=torch.device(DEV))
    model = kw["network_fn"]
    assert len(grad_vars) == len(list(model.parameters())) == 24
    H = W = 24
    focal = 0.5 * W / np.tanh(0.5 * 0.6911112)
    c2w = T(O.pose_spherical(30.0, -30.0, 4.0)[:3, :4].astype(np.float32))
    with torch.no_grad():
        rgb, disp, acc, extras = tnerf.render(H, W, focal, chunk=200, c2w=c2w, frame_time=0.4, near=2., far=6.,
                                              **kw_test)
    assert tuple(rgb.shape) == (H, W, 3) and tuple(disp.shape) == (H, W) and tuple(extras["z_vals"].shape) == (H, W, 64)
    # one optimiser step, save in the reference's checkpoint layout (run_tnerf.py:748-757), reload
    rays = T(O.blender_rays(64, seed=5, frame_time=0.4))
    ret = tnerf.render_rays(rays, **{k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")})
    ret["rgb_map"].square().mean().backward()
    opt.step()
    path = os.path.join(str(tmp_path), "e", "000001.tar")
    torch.save({"global_step": 1, "network_fn_state_dict": model.state_dict(),
                "optimizer_state_dict": opt.state_dict()}, path)
    args.no_reload = False
    kw2, _, start, _, _ = tnerf.create_nerf(args, device=torch.device(DEV))
    assert start == 1
    for (n, a), (_, b) in zip(model.state_dict().items(), kw2["network_fn"].state_dict().items()):
        assert torch.equal(a, b), n


def test_hgemm_tc_vs_fp32_gemm():
    """The layer-at-a-time tcgen05 GEMM against the fp32 SIMT GEMM: both operand orders, ragged M, N and K that are not
    multiples of the tile, strided / unaligned / broadcast operands, bias + accumulate + ELU / ReLU, mask derivative.
    Tolerance: fp16 operands (2^-11 relative each) over K <= 256 terms, fp32 accumulation."""
    rs = np.random.RandomState(11)
    for op, M, N, K, act, acc, mask_act in [(0, 1000, 128, 84, "elu", False, None), (0, 333, 256, 256, "relu", True, None),
                                             (0, 129, 64, 155, "elu", False, None), (0, 5, 16, 3, False, False, None),
                                             (1, 700, 63, 128, False, True, None), (1, 300, 256, 256, False, False, "relu"),
                                             (1, 257, 128, 64, False, False, "elu"), (0, 4096, 128, 128, "elu", False, None)]:
        wide = torch.from_numpy(rs.uniform(-1, 1, size=(M, K + 7)).astype(np.float32)).to(DEV)
        A = wide[:, 3:3 + K]                                         # unaligned base, stride K + 7
        Wm = torch.from_numpy(rs.uniform(-1, 1, size=((N, K) if op == 0 else (K, N))).astype(np.float32)).to(DEV) / np.sqrt(K)
        bias = torch.from_numpy(rs.uniform(-1, 1, size=(N,)).astype(np.float32)).to(DEV) if op == 0 else None
        C0 = torch.from_numpy(rs.uniform(-1, 1, size=(M, N)).astype(np.float32)).to(DEV)
        mask = torch.from_numpy(rs.uniform(-1, 1, size=(M, N)).astype(np.float32)).to(DEV) if mask_act else None
        outs = []
        for tcflag in (False, True):
            C = C0.clone()
            ops._gemm(op, (A.data_ptr(), A.stride(0)), (Wm.data_ptr(), Wm.stride(0)), (C.data_ptr(), N), M, N, K,
                      bias=None if bias is None else bias.data_ptr(), accumulate=acc, relu=act,
                      mask=None if mask is None else (mask.data_ptr(), N), mask_act=mask_act or "relu", tc=tcflag,
                      a_scale=4.0 if op == 1 else 1.0)
            outs.append(C)
        torch.cuda.synchronize()
        err = float((outs[1] - outs[0]).abs().max())
        assert err < 2e-3, (op, M, N, K, err)
    # weight gradients: G[n_out, k_in] += dY^T X over ragged sample counts, gradient-sized dY lifted by a power of two
    for n_out, k_in, M in [(128, 84, 1000), (256, 256, 700), (64, 155, 333), (128, 128, 20000), (256, 63, 129)]:
        wide = torch.from_numpy(rs.uniform(-1, 1, size=(M, k_in + 5)).astype(np.float32)).to(DEV)
        X = wide[:, 1:1 + k_in]
        dY = torch.from_numpy(rs.normal(size=(M, n_out)).astype(np.float32)).to(DEV) * 1e-6
        scale = torch.empty(1, device=DEV)
        _lib.call("swnerf_pow2_scale", dY.data_ptr(), dY.numel(), 32.0, scale.data_ptr(), _lib.stream())
        assert 16.0 <= float(scale.item() * dY.abs().max().item()) <= 32.0
        Gs = [torch.zeros(n_out, k_in, device=DEV) for _ in range(2)]
        for G, tcflag in zip(Gs, (False, True)):
            ops._gemm(2, (dY.data_ptr(), n_out), (X.data_ptr(), X.stride(0)), (G.data_ptr(), k_in), n_out, k_in, M,
                      accumulate=True, tc=tcflag, a_scale_dev=scale.data_ptr())
        err = float((Gs[0] - Gs[1]).abs().max() / Gs[0].abs().max())
        assert err < 2e-3, (n_out, k_in, M, err)
    # a broadcast row as A (the time encoding of one frame, stride 0)
    row = torch.from_numpy(rs.uniform(-1, 1, size=(1, 21)).astype(np.float32)).to(DEV)
    A = row.expand(200, -1)
    Wm = torch.from_numpy(rs.uniform(-1, 1, size=(128, 21)).astype(np.float32)).to(DEV)
    Cs = [torch.zeros(200, 128, device=DEV) for _ in range(2)]
    for C, tcflag in zip(Cs, (False, True)):
        ops._gemm(0, (A.data_ptr(), 0), (Wm.data_ptr(), 21), (C.data_ptr(), 128), 200, 128, 21, tc=tcflag)
    assert float((Cs[0] - Cs[1]).abs().max()) < 2e-3


@pytest.mark.parametrize("tag", ["det", "pert"])
def test_render_rays_tnerf_tc_golden(golden, tag, tmp_path):
    """precision='tc' (the default): the network's layers on the tcgen05 GEMM.  north_star: <= 1e-3 on the maps with
    fp16 / TF32-class operands."""
    g = golden("render_rays_tnerf")
    args = _args(tmp_path); args.swnerf_precision = "tc"
    kw_train, kw_test, _, _, _ = tnerf.create_nerf(args, device=torch.device(DEV))
    model = kw_train["network_fn"]
    model.load_state_dict({k: v.to(DEV) for k, v in O.make_params(O.tnerf_param_shapes(), int(g["seed"])).items()})
    kw = dict(kw_test if tag == "det" else kw_train)
    kw.pop("use_viewdirs"); kw.pop("ndc")
    if tag == "pert":
        kw["raw_noise_std"] = 0.5
    rays, target = T(g[f"{tag}/rays"]), T(g[f"{tag}/target"])
    before = _lib.launch_count()
    ret = tnerf.render_rays(rays, retraw=True, pytest=True, **kw)
    assert model.tc_gemm and _lib.launch_count() > before
    for k in ["rgb_map", "acc_map"]:
        assert float((ret[k].cpu() - torch.from_numpy(g[f"{tag}/{k}"])).abs().max()) < 1e-3, k
    loss = torch.mean((ret["rgb_map"] - target) ** 2)
    assert abs(loss.item() - float(g[f"{tag}/loss"])) < 1e-4
    loss.backward()
    num = den = 0.0
    for n, p in model.named_parameters():
        sub = p.grad.reshape(-1)[::251].cpu().double()
        ref = torch.from_numpy(g[f"{tag}/gsub/{n}"]).double()
        num += float((sub - ref).pow(2).sum()); den += float(ref.pow(2).sum())
    assert (num / max(den, 1e-30)) ** 0.5 < 1e-2       # the tolerance of the fused vanilla path's gradients
