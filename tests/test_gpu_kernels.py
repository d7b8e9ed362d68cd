"""GPU parity tests of the per-ray kernels, through the C ABI, against the CPU oracle and the golden
vectors of the unmodified reference.  Tolerances are stated per test: bit-exact for indices and
sorted order; fp32 rounding (different summation order on the GPU) elsewhere."""
import itertools

import numpy as np
import pytest
import torch

import swnerf_b200 as S
from swnerf_b200 import ops
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(a, b, rtol=2e-5, atol=2e-6):
    np.testing.assert_allclose(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a,
                               b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b,
                               rtol=rtol, atol=atol, equal_nan=True)


# ---------------------------------------------------------------- a13 searchsorted (bit-exact)
def test_searchsorted_golden(golden):
    g = golden("searchsorted")
    for k in range(int(g["n"])):
        side = "left" if int(g[f"side{k}"]) else "right"
        out = S.searchsorted(T(g[f"a{k}"]), T(g[f"v{k}"]), side=side)
        assert out.dtype == torch.long
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"out{k}"])


@pytest.mark.parametrize("Ba,Bv,A,V,side", list(itertools.product([1, 100, 200], [1, 100, 200], [1, 50, 500],
                                                                  [1, 12, 120], ["left", "right"])))
def test_searchsorted_reference_grid(Ba, Bv, A, V, side):
    """The reference extension's own test grid (test/test_searchsorted.py:27-44), 5 repeats per cell."""
    if Ba > 1 and Bv > 1 and Ba != Bv:
        return
    for _ in range(5):
        a = torch.sort(torch.rand(Ba, A, device=DEV), dim=1)[0]
        v = torch.rand(Bv, V, device=DEV)
        out = S.searchsorted(a, v, side=side).cpu().numpy()
        np.testing.assert_array_equal(out, O.searchsorted_rows(a.cpu().numpy(), v.cpu().numpy(), side))
    out = torch.empty((max(Ba, Bv), V), dtype=torch.long, device=DEV)
    assert S.searchsorted(a, v, out, side=side) is out


def test_searchsorted_large():
    a = torch.sort(torch.rand(50000, 300, device=DEV), dim=1)[0]      # examples/test.py sizes
    v = torch.rand(50000, 100, device=DEV)
    out = S.searchsorted(a, v, side="left")
    ref = torch.searchsorted(a.cpu(), v.cpu(), right=False)
    assert torch.equal(out.cpu(), ref)


# ---------------------------------------------------------------- a2 stratified sampling
@pytest.mark.parametrize("lindisp,perturb", [(False, 0.0), (False, 1.0), (True, 0.0), (True, 1.0)])
def test_stratified(lindisp, perturb):
    rays = O.blender_rays(257, 3)
    rays[:, 6] = np.random.RandomState(0).uniform(0.5, 2.5, 257)
    t_rand = torch.rand(257, 64)
    z_ref = O.stratified_z(torch.from_numpy(rays[:, 6:7]), torch.from_numpy(rays[:, 7:8]), 64, lindisp, perturb, t_rand)
    z = ops.stratified_z(T(rays), 64, lindisp, perturb, t_rand.to(DEV))
    close(z, z_ref, rtol=1e-6, atol=1e-6)
    assert float((z.cpu() - z_ref).abs().max()) <= 1e-6            # 2 ulp at z ~ 6


# ---------------------------------------------------------------- a4 positional encoding
@pytest.mark.parametrize("L,d", [(10, 3), (4, 3), (20, 3), (10, 1), (8, 1), (0, 3)])
def test_embed_fwd_bwd(L, d):
    x = (torch.rand(1001, d) * 12 - 6)
    xg = x.clone().requires_grad_()
    y_ref = O.embed(xg, L)
    w = torch.randn_like(y_ref)
    (y_ref * w).sum().backward()
    xc = x.to(DEV).requires_grad_()
    fn, od = S.get_embedder(L, d, 0)
    y = fn(xc)
    assert y.shape[1] == od
    close(y, y_ref, rtol=0, atol=2e-6 if L <= 10 else 1e-5)
    (y * w.to(DEV)).sum().backward()
    scale = float(xg.grad.abs().max())
    close(xc.grad, xg.grad, rtol=0, atol=scale * 2e-6)


def test_embed_golden(golden):
    g = golden("embed")
    for L, xk, tag in [(10, "x3", "L10_d3"), (4, "x3", "L4_d3"), (20, "x3", "L20_d3"), (10, "x1", "L10_d1")]:
        close(ops.embed(T(g[xk]), L), g[f"y_{tag}"], rtol=0, atol=2e-6 if L <= 10 else 1e-5)


# ---------------------------------------------------------------- a8 raw2outputs
@pytest.mark.parametrize("tag", ["S64", "S192", "S5"])
def test_raw2outputs_golden(golden, tag):
    g = golden("raw2outputs")
    raw, z, rd = T(g[f"raw_{tag}"]), T(g[f"z_{tag}"]), T(g[f"rd_{tag}"])
    for wb in (0, 1):
        outs = S.raw2outputs(raw, z, rd, 0, bool(wb))
        for name, o in zip(["rgb", "disp", "acc", "weights", "depth"], outs):
            close(o, g[f"{name}_{tag}_wb{wb}"], rtol=2e-5, atol=2e-6)
    outs = ops.composite(raw, z, rd, 0, T(g[f"noise_{tag}"]), True)
    for name, o in zip(["rgb", "disp", "acc", "weights", "depth"], outs):
        close(o, g[f"{name}_{tag}_noise"], rtol=2e-5, atol=2e-6)
    assert torch.isnan(S.raw2outputs(raw, z, rd)[1][0])      # empty ray: reference's 0/0 NaN kept


@pytest.mark.parametrize("S_,wb,noisy", [(64, True, False), (192, True, True), (192, False, False), (37, True, False),
                                        (300, False, True)])
def test_raw2outputs_backward(S_, wb, noisy):
    rs = np.random.RandomState(S_)
    N = 203
    raw = (rs.normal(size=(N, S_, 4)) * 2).astype(np.float32)
    raw[0, :, 3] = -1.0
    raw[1, 3:, 3] = 40.0
    z = np.sort(rs.uniform(2, 6, size=(N, S_)).astype(np.float32), -1)
    rd = rs.normal(size=(N, 3)).astype(np.float32)
    noise = (rs.rand(N, S_)).astype(np.float32) if noisy else None
    cot = [rs.normal(size=s).astype(np.float32) for s in [(N, 3), (N,), (N,), (N, S_), (N,)]]
    cot[1][0] = 0.0          # NaN disparity of the empty ray carries no cotangent

    def run(dev, fn):
        r = torch.from_numpy(raw).to(dev).requires_grad_()
        outs = fn(r, torch.from_numpy(z).to(dev), torch.from_numpy(rd).to(dev),
                  None if noise is None else torch.from_numpy(noise).to(dev))
        loss = sum((o * torch.from_numpy(c).to(dev)).nan_to_num().sum() for o, c in zip(outs, cot))
        loss.backward()
        return r.grad
    g_ref = run("cpu", lambda r, zz, d, n: O.raw2outputs(r, zz, d, 1.0 if noisy else 0.0, wb, n))
    g_gpu = run(DEV, lambda r, zz, d, n: ops.composite(r, zz, d, 0, n, wb))
    scale = float(g_ref[2:].abs().max())
    close(g_gpu[2:], g_ref[2:], rtol=1e-4, atol=scale * 5e-7)


def test_raw2outputs_full_size_properties():
    """BASELINE config sizes (4096 rays x 192): acc = sum(weights), weights in [0,1], linear in g."""
    N, S_ = 4096, 192
    raw = torch.randn(N, S_, 4, device=DEV) * 3
    z = torch.sort(torch.rand(N, S_, device=DEV) * 4 + 2, -1)[0]
    rd = torch.randn(N, 3, device=DEV)
    rgb, disp, acc, w, depth = ops.composite(raw, z, rd, 0, None, True)
    assert float((w.sum(-1) - acc).abs().max()) < 1e-5
    assert float(w.min()) >= 0 and float(acc.max()) <= 1 + 1e-5
    assert float((rgb - ((w[..., None] * torch.sigmoid(raw[..., :3])).sum(1) + 1 - acc[:, None])).abs().max()) < 1e-5
    assert float(((w * z).sum(-1) - depth).abs().max()) < 1e-4


# ---------------------------------------------------------------- a9 sample_pdf / a10 sort / a11 z_std
def test_sample_pdf_indices_bit_exact_given_cdf(golden):
    g = golden("sample_pdf")
    bins, cdf = T(g["bins"]), T(g["cdf"])
    s, inds = ops.sample_pdf(bins, None, 128, det=True, cdf=cdf, return_inds=True)
    np.testing.assert_array_equal(inds.cpu().numpy(), g["inds_det128"])
    close(s, g["samples_det128"], rtol=1e-5, atol=2e-5)     # (u - cdf_b)/denom amplifies 1 ulp when denom ~ 1e-5
    s, inds = ops.sample_pdf(bins, None, 128, det=False, u=T(g["u_rand128"]), cdf=cdf, return_inds=True)
    np.testing.assert_array_equal(inds.cpu().numpy(), g["inds_rand128"])
    close(s, g["samples_rand128"], rtol=1e-5, atol=2e-5)


def test_sample_pdf_from_weights(golden):
    g = golden("sample_pdf")
    bins, w = T(g["bins"]), T(g["weights"])
    # the cdf is rebuilt on the GPU (warp scan, different rounding): values agree to fp32 rounding of the
    # cdf amplified by the bin width / pdf; rows with flat cdf stretches excluded from the tight check
    close(S.sample_pdf(bins, w, 128, det=True), g["samples_det128"], rtol=0, atol=2e-4)
    close(S.sample_pdf(bins, w, 64, det=True), g["samples_det64"], rtol=0, atol=2e-4)
    close(ops.sample_pdf(bins, w, 128, det=False, u=T(g["u_rand128"])), g["samples_rand128"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("N,S_,Ni,det", [(513, 64, 128, True), (513, 64, 128, False), (77, 64, 64, False),
                                        (513, 64, 64, True), (4099, 64, 64, False),          # the LLFF configs' 64 + 64
                                        (33, 16, 7, False), (4096, 64, 128, False)])
@pytest.mark.parametrize("variant", [1, 0])
def test_resample(N, S_, Ni, det, variant):
    """variant: 1 = eight lanes per ray (the default; serves 64 + 128 and 64 + 64), 0 = one warp per ray (64 + 128)
    / the generic kernel; other shapes always run the generic kernel."""
    if variant == 0 and (S_, Ni) != (64, 128):
        pytest.skip("variant only concerns the 64+128 shape")
    rs = np.random.RandomState(N + Ni)
    z = np.sort(rs.uniform(2, 6, size=(N, S_)).astype(np.float32), -1)
    w = (rs.uniform(0, 1, size=(N, S_)).astype(np.float32)) ** 6
    w[0] = 0
    # degenerate rays (the verified fast paths must hand them to the exact routine): all z equal, z one ulp apart,
    # one dominant bin, and a pair of equal neighbours
    u = rs.rand(N, Ni).astype(np.float32)
    if N >= 77:
        z[1] = 3.0
        z[2] = np.float32(3.0) + np.arange(S_, dtype=np.float32) * np.spacing(np.float32(3.0))
        w[3] = 0; w[3, S_ // 2] = 1.0
        z[4, S_ // 2 + 1] = z[4, S_ // 2]
        u[5, 0] = 0.0
    from swnerf_b200 import _lib
    _lib.call("swnerf_set_resample_variant", variant)
    try:
        zs, zf, zstd = ops.resample(T(z), T(w), Ni, det=det, u=None if det else T(u))
        torch.cuda.synchronize()
    finally:
        _lib.call("swnerf_set_resample_variant", 1)
    zt, wt = torch.from_numpy(z), torch.from_numpy(w)
    z_mid = 0.5 * (zt[:, 1:] + zt[:, :-1])
    zs_ref = O.sample_pdf(z_mid, wt[:, 1:-1], Ni, det=det, u=None if det else torch.from_numpy(u))
    # the cdf is rebuilt on the GPU (warp scan instead of a sequential cumsum): a 1-ulp cdf difference can
    # move a sample across a near-empty bin ("bit-exact given identical CDFs"), so allow <= 0.1% outliers
    bad = ((zs.cpu() - torch.sort(zs_ref, -1)[0]).abs() > 3e-4).float().mean().item()   # returned ascending
    assert bad < 1e-3, bad
    # exact properties: z_fine is sorted and is exactly the multiset {z_vals} U {z_samples}
    zf_c, zs_c = zf.cpu(), zs.cpu()
    assert bool((zf_c[:, 1:] >= zf_c[:, :-1]).all())
    assert torch.equal(zf_c, torch.sort(torch.cat([zt, zs_c], -1), -1)[0])
    close(zstd, torch.std(zs_c, dim=-1, unbiased=False), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("mode", ["det", "rand"])
def test_resample64q_indices_bit_exact_given_cdf(golden, mode):
    """The north_star's index bar on the PRODUCTION kernel of the 64+128 shape (resample64q_kernel, the one
    render_rays launches): fed the reference's cdf, its branchless 6-probe search returns exactly
    torch.searchsorted(cdf, u, right=True) (ray.py:136) for every sample of every ray - degenerate rows included."""
    from swnerf_b200 import _lib
    g = golden("resample")
    z, w, cdf = T(g["z_vals"]), T(g["weights"]), T(g["cdf"])
    det = mode == "det"
    u = None if det else T(g["u_rand"])
    _lib.resample_fallbacks(reset=True)
    out = ops.resample_check(z, w, 128, det=det, u=u, cdf=cdf, variant=0)
    inds_ref = g[f"{mode}/inds"]
    if not det:      # the kernel draws the samples in ascending-u order
        order = np.argsort(g["u_rand"], axis=-1, kind="stable")
        inds_ref = np.take_along_axis(inds_ref, order, -1)
    np.testing.assert_array_equal(out["inds"].cpu().numpy(), inds_ref)
    np.testing.assert_array_equal(out["cdf"].cpu().numpy(), g["cdf"])
    # samples from the reference's cdf: equal up to the kernel's reciprocal (1 / denom, MUFU: <= 2 ulp of t, a 0.06-wide
    # bin) - except where denom ~ 1e-5 amplifies it (the flat-cdf rows), as in test_sample_pdf_indices_bit_exact_given_cdf
    zs_ref = np.sort(g[f"{mode}/z_samples"], -1)
    close(out["z_samples"], zs_ref, rtol=1e-6, atol=2e-5)
    assert float((out["z_samples"].cpu() - torch.from_numpy(zs_ref)).abs().median()) < 5e-7
    zf = out["z_fine"].cpu()
    assert torch.equal(zf, torch.sort(torch.cat([torch.from_numpy(g["z_vals"]), out["z_samples"].cpu()], -1), -1)[0])
    assert _lib.resample_fallbacks() <= 8          # the degenerate rows may take the exact routine; the rest must not


@pytest.mark.parametrize("mode", ["det", "rand"])
def test_resample_reference_order_is_bit_identical(golden, mode):
    """Check mode (what precision='fp32' runs): given the reference's coarse weights, the whole stage - pdf, cdf,
    searchsorted, the inverse-cdf arithmetic, the merge with z_vals - reproduces the reference's tensors BIT FOR BIT
    (ray.py:96-153 + nerf/run.py:396-400 run unmodified by oracle/make_golden.py)."""
    g = golden("resample")
    z, w = T(g["z_vals"]), T(g["weights"])
    det = mode == "det"
    out = ops.resample_check(z, w, 128, det=det, u=None if det else T(g["u_rand"]), variant=1)
    np.testing.assert_array_equal(out["cdf"].cpu().numpy(), g["cdf"])
    inds_ref = g[f"{mode}/inds"]
    if not det:
        inds_ref = np.take_along_axis(inds_ref, np.argsort(g["u_rand"], axis=-1, kind="stable"), -1)
    np.testing.assert_array_equal(out["inds"].cpu().numpy(), inds_ref)
    np.testing.assert_array_equal(out["z_samples"].cpu().numpy(), np.sort(g[f"{mode}/z_samples"], -1))
    np.testing.assert_array_equal(out["z_fine"].cpu().numpy(), g[f"{mode}/z_fine"])
    close(out["z_std"], g[f"{mode}/z_std"], rtol=1e-5, atol=1e-7)
    # and through the public op render_rays uses in the check mode
    _, zf, _ = ops.resample(z, w, 128, det=det, u=None if det else T(g["u_rand"]), want_samples=False, exact=True)
    np.testing.assert_array_equal(zf.cpu().numpy(), g[f"{mode}/z_fine"])


def test_resample_reference_order_other_shapes():
    """Variant 1 on shapes the 64+128 kernels do not cover, against the live oracle (its torch.sum / cumsum order)."""
    for N, S_, Ni in ((19, 16, 7), (9, 128, 64), (5, 33, 130)):
        rs = np.random.RandomState(N)
        z = np.sort(rs.uniform(2, 6, size=(N, S_)).astype(np.float32), -1)
        w = (rs.uniform(0, 1, size=(N, S_)).astype(np.float32)) ** 5
        zt, wt = torch.from_numpy(z), torch.from_numpy(w)
        zs_ref = O.sample_pdf(0.5 * (zt[:, 1:] + zt[:, :-1]), wt[:, 1:-1], Ni, det=True)
        zf_ref = torch.sort(torch.cat([zt, zs_ref], -1), -1)[0]
        out = ops.resample_check(T(z), T(w), Ni, det=True, variant=1)
        np.testing.assert_array_equal(out["cdf"].cpu().numpy(), O.pdf_to_cdf(wt[:, 1:-1]).numpy())
        np.testing.assert_array_equal(out["z_fine"].cpu().numpy(), zf_ref.numpy())


def test_resample_variants_agree_at_render_size():
    """The two kernels of the 64+128 shape on a 32,768-ray render chunk (+3 so the last warp is ragged): identical
    merged rows up to the samples' own rounding (the pdf is normalised by a reciprocal in one and a division in the
    other), both exactly sorted and exactly the multiset union."""
    from swnerf_b200 import _lib
    N = 32768 + 3
    _lib.resample_fallbacks(reset=True)
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.sort(torch.rand(N, 64, device="cuda", generator=g) * 4 + 2, -1)[0]
    w = torch.rand(N, 64, device="cuda", generator=g) ** 4
    u = torch.rand(N, 128, device="cuda", generator=g)
    outs = {}
    for det in (True, False):
        for variant in (0, 1):
            _lib.call("swnerf_set_resample_variant", variant)
            try:
                outs[(det, variant)] = ops.resample(z, w, 128, det=det, u=None if det else u)
            finally:
                _lib.call("swnerf_set_resample_variant", 1)
        (zs0, zf0, sd0), (zs1, zf1, sd1) = outs[(det, 0)], outs[(det, 1)]
        for zs_, zf_ in ((zs0, zf0), (zs1, zf1)):
            assert bool((zf_[:, 1:] >= zf_[:, :-1]).all())
            assert torch.equal(zf_, torch.sort(torch.cat([z, zs_], -1), -1)[0])
        assert float(((zs0 - zs1).abs() > 1e-4).float().mean()) < 1e-3
        # z_std follows the samples: a sample that changes bin moves by a bin width
        assert float(((sd0 - sd1).abs() > 1e-3 * sd0.abs() + 1e-5).float().mean()) < 2e-2
    assert _lib.resample_fallbacks() < 0.01 * N          # well-formed rays stay on the fast path
