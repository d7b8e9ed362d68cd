"""Pins oracle/nerf_oracle.py against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only.  Tolerances: both sides are fp32 torch on CPU running the
same op sequence, so agreement is to a few ulp; indices are bit-exact."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O

T = torch.from_numpy


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol, equal_nan=True)


def test_embed(golden):
    g = golden("embed")
    for L, xk, tag in [(10, "x3", "L10_d3"), (4, "x3", "L4_d3"), (20, "x3", "L20_d3"), (10, "x1", "L10_d1"),
                       (8, "x1", "L8_d1"), (4, "x1", "L4_d1")]:
        y = O.embed(T(g[xk]), L).numpy()
        assert y.shape[1] == O.embed_dim(L, g[xk].shape[1])
        np.testing.assert_array_equal(y, g[f"y_{tag}"])
    np.testing.assert_array_equal(O.embed(T(g["x3"]), -1).numpy(), g["y_identity"])


def test_mlp(golden):
    g = golden("mlp")
    p = O.make_params(O.mlp_param_shapes(), int(g["seed"]))
    y = O.mlp_forward(p, T(g["x"]), 63, 27).numpy()
    close(y, g["y_vanilla"])
    close(y, g["y_original"])
    pn = O.make_params(O.mlp_param_shapes(input_ch_views=0, output_ch=4, use_viewdirs=False), int(g["seed_nv"]))
    close(O.mlp_forward(pn, T(g["x"][:, :63].copy()), 63, 0, use_viewdirs=False).numpy(), g["y_noview"])


def test_dnerf_model(golden):
    g = golden("mlp")
    p = O.make_params(O.dnerf_param_shapes(), int(g["seed_dnerf"]))
    pts, vd = T(g["d_pts"]), T(g["d_vd"])
    x = torch.cat([O.embed(pts, 10), O.embed(vd, 4)], -1)
    for tval, tag in [(0.37, "t037"), (0.0, "t0")]:
        et = O.embed(torch.full((pts.shape[0], 1), tval), 10)
        out, dx = O.dnerf_forward(p, x, et, 63, 27, 10, tval)
        close(out.numpy(), g[f"d_out_{tag}"], rtol=2e-5, atol=2e-6)
        close(dx.numpy(), g[f"d_dx_{tag}"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("tag", ["S64", "S192", "S5"])
def test_raw2outputs(golden, tag):
    g = golden("raw2outputs")
    raw, z, rd = T(g[f"raw_{tag}"]), T(g[f"z_{tag}"]), T(g[f"rd_{tag}"])
    for wb in (0, 1):
        outs = O.raw2outputs(raw, z, rd, 0.0, bool(wb))
        for name, o in zip(["rgb", "disp", "acc", "weights", "depth"], outs):
            close(o.numpy(), g[f"{name}_{tag}_wb{wb}"])
    outs = O.raw2outputs(raw, z, rd, 1.0, True, noise=T(g[f"noise_{tag}"]))
    for name, o in zip(["rgb", "disp", "acc", "weights", "depth"], outs):
        close(o.numpy(), g[f"{name}_{tag}_noise"])
    # the reference's 0/0 -> NaN disparity for a fully transparent ray is preserved
    assert np.isnan(g[f"disp_{tag}_wb0"][0]) and np.isnan(O.raw2outputs(raw, z, rd)[1].numpy()[0])


def test_sample_pdf(golden):
    g = golden("sample_pdf")
    bins, w = T(g["bins"]), T(g["weights"])
    np.testing.assert_array_equal(O.pdf_to_cdf(w).numpy(), g["cdf"])
    close(O.sample_pdf(bins, w, 128, det=True).numpy(), g["samples_det128"])
    close(O.sample_pdf(bins, w, 64, det=True).numpy(), g["samples_det64"])
    close(O.sample_pdf(bins, w, 128, u=T(g["u_rand128"])).numpy(), g["samples_rand128"])
    u = torch.linspace(0., 1., 128).expand(bins.shape[0], 128)
    _, inds = O.sample_from_cdf(bins, T(g["cdf"]), u)
    np.testing.assert_array_equal(inds.numpy(), g["inds_det128"])
    _, inds = O.sample_from_cdf(bins, T(g["cdf"]), T(g["u_rand128"]))
    np.testing.assert_array_equal(inds.numpy(), g["inds_rand128"])


@pytest.mark.parametrize("mode", ["det", "rand"])
def test_resample(golden, mode):
    """nerf/run.py:396-400, :416 as the reference ran it: the oracle reproduces cdf, indices, samples and the merged
    z_fine bit for bit on this host (same torch ops)."""
    g = golden("resample")
    z, w = T(g["z_vals"]), T(g["weights"])
    z_mid = 0.5 * (z[:, 1:] + z[:, :-1])
    cdf = O.pdf_to_cdf(w[:, 1:-1])
    np.testing.assert_array_equal(cdf.numpy(), g["cdf"])
    det = mode == "det"
    u = T(g["u_rand"]) if not det else torch.linspace(0.0, 1.0, 128).expand(z.shape[0], 128).contiguous()
    zs, inds = O.sample_from_cdf(z_mid, cdf, u)
    np.testing.assert_array_equal(inds.numpy(), g[f"{mode}/inds"])
    np.testing.assert_array_equal(zs.numpy(), g[f"{mode}/z_samples"])
    zf = torch.sort(torch.cat([z, zs], -1), -1)[0]
    np.testing.assert_array_equal(zf.numpy(), g[f"{mode}/z_fine"])
    close(torch.std(zs, dim=-1, unbiased=False).numpy(), g[f"{mode}/z_std"])


def test_cpu_sum_order_is_the_one_the_check_mode_follows(golden):
    """ray.py:112's torch.sum on this host == the 8-lane order restated in O.torch_cpu_sum_order (and implemented by
    the CUDA check mode, ops.REF_SUM_LANES): on the golden weights and on random rows of several lengths."""
    g = golden("resample")
    w = (g["weights"][:, 1:-1] + np.float32(1e-5)).astype(np.float32)
    ref = T(w).sum(-1).numpy()
    mine = np.array([O.torch_cpu_sum_order(r) for r in w], dtype=np.float32)
    np.testing.assert_array_equal(mine, ref)
    rs = np.random.RandomState(3)
    for n in (3, 14, 30, 62, 126, 190, 254, 510):
        x = (rs.rand(40, n).astype(np.float32) ** 3).astype(np.float32)
        np.testing.assert_array_equal(np.array([O.torch_cpu_sum_order(r) for r in x], dtype=np.float32), T(x).sum(-1).numpy())
    # torch.cumsum on CPU: sequential accumulation in double, rounded to float at every step
    p = T(w / ref[:, None])
    acc = np.zeros(p.shape[0], dtype=np.float64)
    seq = []
    for i in range(p.shape[1]):
        acc = acc + p[:, i].numpy().astype(np.float64)
        seq.append(acc.astype(np.float32))
    np.testing.assert_array_equal(np.stack(seq, -1), torch.cumsum(p, -1).numpy())


def test_searchsorted(golden):
    g = golden("searchsorted")
    for k in range(int(g["n"])):
        side = "left" if int(g[f"side{k}"]) else "right"
        np.testing.assert_array_equal(O.searchsorted_rows(g[f"a{k}"], g[f"v{k}"], side), g[f"out{k}"])


def _grad_check(g, tag, named, rtol):
    for k, grad in named.items():
        gd = grad.detach().double().reshape(-1)
        ref_norm = float(g[f"{tag}/gnorm/{k}"])
        assert abs(gd.norm().item() - ref_norm) <= rtol * max(ref_norm, 1e-12), k
        sub = g[f"{tag}/gsub/{k}"]
        np.testing.assert_allclose(gd[::251].float().numpy(), sub, rtol=0, atol=rtol * max(np.abs(sub).max(), 1e-12) * 10)


@pytest.mark.parametrize("tag", ["det", "pert", "noise", "lindisp"])
def test_render_rays(golden, tag):
    g = golden("render_rays")
    rays, target = T(g["rays"]), T(g["target"])
    shapes = O.mlp_param_shapes()
    pc = {k: v.requires_grad_() for k, v in O.make_params(shapes, int(g["seed_coarse"])).items()}
    pf = {k: v.requires_grad_() for k, v in O.make_params(shapes, int(g["seed_fine"])).items()}
    N = rays.shape[0]
    perturb = 0.0 if tag in ("det", "lindisp") else 1.0
    kw = {}
    if perturb > 0:   # the pytest hooks re-seed numpy with 0 before every draw (run.py:377-381, ray.py:124-132,180-184)
        np.random.seed(0); kw["t_rand"] = torch.Tensor(np.random.rand(N, 64))
        np.random.seed(0); kw["u"] = torch.Tensor(np.random.rand(N, 128))
    std = 1.0 if tag == "noise" else 0.0
    if std > 0:
        np.random.seed(0); kw["noise0"] = torch.Tensor(np.random.rand(N, 64) * std)
        np.random.seed(0); kw["noise1"] = torch.Tensor(np.random.rand(N, 192) * std)
    ret = O.render_rays(rays, pc, pf, 64, 128, perturb=perturb, white_bkgd=True, raw_noise_std=std,
                        lindisp=(tag == "lindisp"), retraw=True, **kw)
    for k in ["rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "z_std"]:
        close(ret[k].detach().numpy(), g[f"{tag}/{k}"], rtol=2e-4, atol=2e-5)
    # raw logits: the reference evaluates the MLP in netchunk slabs, the oracle in one GEMM (different
    # MKL blocking -> different fp32 summation order), and the sigma head carries a x24 gain
    close(ret["raw"].detach().numpy(), g[f"{tag}/raw"], rtol=2e-4, atol=1e-3)
    loss = torch.mean((ret["rgb_map"] - target) ** 2) + torch.mean((ret["rgb0"] - target) ** 2)
    assert abs(loss.item() - float(g[f"{tag}/loss"])) < 1e-5
    loss.backward()
    named = {"coarse." + k: v.grad for k, v in pc.items()}
    named.update({"fine." + k: v.grad for k, v in pf.items()})
    _grad_check(g, tag, named, rtol=1e-3)


@pytest.mark.parametrize("tag", ["t037", "t0"])
def test_render_rays_dnerf(golden, tag):
    g = golden("render_rays_dnerf")
    rays, target = T(g[f"{tag}/rays"]), T(g[f"{tag}/target"])
    p = {k: v.requires_grad_() for k, v in O.make_params(O.dnerf_param_shapes(), int(g["seed"])).items()}
    N = rays.shape[0]
    np.random.seed(0); t_rand = torch.Tensor(np.random.rand(N, 64))
    np.random.seed(0); u = torch.Tensor(np.random.rand(N, 128))
    ret = O.render_rays_dnerf(rays, p, 64, 128, perturb=1.0, white_bkgd=True, t_rand=t_rand, u=u, retraw=True)
    for k in ["rgb_map", "disp_map", "acc_map", "z_vals", "position_delta", "raw", "z_std"]:
        close(ret[k].detach().numpy(), g[f"{tag}/{k}"], rtol=2e-4, atol=2e-5)
    loss = torch.mean((ret["rgb_map"] - target) ** 2)
    if tag == "t037":
        rays2 = rays.clone(); rays2[:, 8] = 0.37 + 0.01
        ret2 = O.render_rays_dnerf(rays2, p, 64, 128, perturb=1.0, white_bkgd=True, z_vals=ret["z_vals"].detach())
        close(ret2["position_delta"].detach().numpy(), g[f"{tag}/position_delta_next"], rtol=2e-4, atol=2e-5)
        loss = loss + 0.1 * torch.sum((ret["position_delta"] - ret2["position_delta"]) ** 2)
    assert abs(loss.item() - float(g[f"{tag}/loss"])) < 1e-5 * max(1.0, abs(float(g[f"{tag}/loss"])))
    loss.backward()
    named = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    _grad_check(g, tag, named, rtol=1e-3)


# ---------------------------------------------------------------- f4 T-NeRF (model.py:152-210, run_tnerf.py:396-500)
def test_tnerf_mlp(golden):
    g = golden("render_rays_tnerf")
    p = O.make_params(O.tnerf_param_shapes(), int(g["seed"]))
    x, t = T(g["mlp/x"]), T(g["mlp/t"])
    out = O.tnerf_forward(p, x, x[:, 63:], t)
    assert g["mlp/out"].shape == (1, x.shape[0], 4)                 # the reference's [-1, M, 4] reshape, model.py:205-208
    close(out.numpy(), g["mlp/out"][0], rtol=1e-5, atol=1e-6)
    assert float(out[:, :3].min()) >= 0.0                           # colour head ends in a ReLU


@pytest.mark.parametrize("tag", ["det", "pert"])
def test_render_rays_tnerf(golden, tag):
    g = golden("render_rays_tnerf")
    rays, target = T(g[f"{tag}/rays"]), T(g[f"{tag}/target"])
    p = {k: v.requires_grad_() for k, v in O.make_params(O.tnerf_param_shapes(), int(g["seed"])).items()}
    N = rays.shape[0]
    kw = {}
    std = 0.0
    if tag == "pert":     # run_tnerf.py:466-469 as shipped: the pytest draw of t_rand is scaled by raw_noise_std
        std = 0.5
        np.random.seed(0); kw["t_rand"] = torch.Tensor(np.random.rand(N, 64) * std)
    ret = O.render_rays_tnerf(rays, p, 64, perturb=1.0 if tag == "pert" else 0.0, white_bkgd=True, retraw=True, **kw)
    if std > 0:           # the same hook draws the density noise (run_tnerf.py:372-376)
        np.random.seed(0); noise = torch.Tensor(np.random.rand(N, 64) * std)
        rgb_map, disp_map, acc_map, _, _ = O.raw2outputs(ret["raw"], ret["z_vals"], rays[:, 3:6], std, True, noise=noise)
        ret.update(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map)
    for k in ["rgb_map", "disp_map", "acc_map", "z_vals", "raw"]:
        close(ret[k].detach().numpy(), g[f"{tag}/{k}"], rtol=2e-4, atol=2e-5)
    loss = torch.mean((ret["rgb_map"] - target) ** 2)
    assert abs(loss.item() - float(g[f"{tag}/loss"])) < 1e-5
    loss.backward()
    _grad_check(g, tag, {k: v.grad for k, v in p.items()}, rtol=1e-3)
