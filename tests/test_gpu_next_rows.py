"""GPU: the rows SURVEY.md 8f marks "next" - ray assembly (f1), loss + Adam on flat buffers (f3),
sharded frame gather helper (f2) - each against the oracle / PyTorch at the same parity bar."""
import numpy as np
import pytest
import torch

import swnerf_b200 as S
from swnerf_b200 import parallel, ray as sray
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference_batch(H, W, K, c2w, near, far):
    """get_rays_np (ray.py:42-72) + render()'s assembly (nerf/run.py:137-158) in numpy/torch on the CPU."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = torch.sum(torch.from_numpy(dirs)[..., None, :] * torch.from_numpy(c2w[:3, :3]), -1).reshape(-1, 3)
    rays_o = torch.from_numpy(c2w[:3, -1]).expand(rays_d.shape)
    vd = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    return torch.cat([rays_o, rays_d, near * torch.ones_like(rays_d[:, :1]), far * torch.ones_like(rays_d[:, :1]), vd], -1)


def test_make_ray_batch_matches_get_rays():
    H, W = 37, 53
    focal = 0.5 * W / np.tan(0.5 * 0.6911112)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    c2w = O.pose_spherical(33.0, -30.0, 4.0)[:3, :4]
    ref = _reference_batch(H, W, K, c2w, 2.0, 6.0)
    got = sray.make_ray_batch(H, W, K, torch.from_numpy(c2w), 2.0, 6.0).cpu()
    assert got.shape == (H * W, 11)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-7, atol=2e-7)
    pix = torch.from_numpy(np.random.RandomState(0).randint(0, H * W, 500)).to(DEV)
    sel = sray.make_ray_batch(H, W, K, torch.from_numpy(c2w), 2.0, 6.0, pixels=pix, frame_time=0.25).cpu()
    assert sel.shape == (500, 12)
    np.testing.assert_allclose(sel[:, :8].numpy(), ref[pix.cpu()][:, :8].numpy(), rtol=2e-7, atol=2e-7)
    assert float((sel[:, 8] - 0.25).abs().max()) == 0.0
    np.testing.assert_allclose(sel[:, 9:].numpy(), ref[pix.cpu()][:, 8:].numpy(), rtol=2e-7, atol=2e-7)


def test_render_full_frame_fast_path_equals_generic_path():
    H, W = 24, 20
    focal = 0.5 * W / np.tan(0.5 * 0.6911112)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    c2w = torch.from_numpy(O.pose_spherical(-70.0, -30.0, 4.0)[:3, :4]).to(DEV)
    shapes = O.mlp_param_shapes()
    mc = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mc.load_state_dict(O.make_params(shapes, 21)); mc.to(DEV)
    q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision="fp32")
    kw = dict(network_fn=mc, network_query_fn=q, N_samples=64, N_importance=0, white_bkgd=True)
    with torch.no_grad():
        rgb, disp, acc, _ = S.render(H, W, K, chunk=256, c2w=c2w, ndc=False, near=2., far=6., use_viewdirs=True, **kw)
        rays_o, rays_d = S.get_rays(H, W, K, c2w)
        rgb2, _, acc2, _ = S.render(H, W, K, chunk=256, rays=(rays_o, rays_d), ndc=False, near=2., far=6.,
                                    use_viewdirs=True, **kw)
    assert rgb.shape == (H, W, 3)
    assert float((rgb - rgb2).abs().max()) < 1e-5 and float((acc - acc2).abs().max()) < 1e-5


def test_flat_adam_matches_torch_adam():
    torch.manual_seed(0)
    m1 = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True).to(DEV)
    m2 = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True).to(DEV)
    m2.load_state_dict(m1.state_dict())
    opt1 = torch.optim.Adam(m1.parameters(), lr=5e-4, betas=(0.9, 0.999))
    flat = parallel.FlatParams(list(m2.parameters()))
    assert flat.check_param_views() and flat.check_views()
    opt2 = parallel.FlatAdam(flat, lr=5e-4)
    g = torch.Generator(device=DEV).manual_seed(1)
    for it in range(5):
        opt1.zero_grad(); opt2.zero_grad()
        for p1, p2 in zip(m1.parameters(), m2.parameters()):
            gr = torch.randn(p1.shape, device=DEV, generator=g) * (10.0 ** (it - 3))
            p1.grad = gr.clone()
            p2.grad.copy_(gr)
        opt1.step(); opt2.step()
    for (n, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
        assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-7), n
    assert set(m2.state_dict().keys()) == set(m1.state_dict().keys())


def test_two_loss_mse_matches_img2mse():
    g = torch.Generator(device=DEV).manual_seed(2)
    a = torch.rand(1000, 3, device=DEV, generator=g, requires_grad=True)
    b = torch.rand(1000, 3, device=DEV, generator=g, requires_grad=True)
    t = torch.rand(1000, 3, device=DEV, generator=g)
    ref = torch.mean((a - t) ** 2) + torch.mean((b - t) ** 2)
    ga, gb = torch.autograd.grad(ref, (a, b))
    loss = parallel.two_loss_mse(a, b, t)
    ha, hb = torch.autograd.grad(loss * 3.0, (a, b))
    assert abs(loss.item() - ref.item()) < 1e-6
    assert torch.allclose(ha, 3.0 * ga, rtol=1e-6, atol=1e-9) and torch.allclose(hb, 3.0 * gb, rtol=1e-6, atol=1e-9)


def test_flat_adam_invalidates_packed_weights():
    from swnerf_b200 import tc
    if not tc.available():
        pytest.skip("tcgen05 path not built")
    m = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); m.load_state_dict(O.make_params(O.mlp_param_shapes(), 21)); m.to(DEV)
    flat = parallel.FlatParams(list(m.parameters()))
    opt = parallel.FlatAdam(flat, lr=1e-2)
    q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision="tc")
    rays = torch.from_numpy(O.blender_rays(16, 3)).to(DEV)
    z = S.ops.stratified_z(rays, 64)
    with torch.no_grad():
        a = q.query_rays(rays, z, m, 8).clone()
    flat.flat.fill_(1.0)
    opt.step()
    with torch.no_grad():
        b = q.query_rays(rays, z, m, 8)
    assert float((a - b).abs().max()) > 1e-3        # the fused kernel saw the updated weights


# ---------------------------------------------------------------- round 2: the rest of f1 / f2
def _K(H, W):
    focal = 0.5 * W / np.tan(0.5 * 0.6911112)
    return focal, np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])      # float64, as nerf/run.py:498-502


def test_make_ray_batch_ndc_matches_ndc_rays():
    """LLFF / NDC configs (10 of the 18 vanilla configs): the kernel's NDC warp against get_rays + ndc_rays
    (ray.py:10-38, 75-92) + render()'s assembly (nerf/run.py:137-158) in eager torch, operation for operation."""
    H, W = 31, 45
    focal, K = _K(H, W)
    c2w = torch.from_numpy(O.pose_spherical(12.0, -20.0, 1.5)[:3, :4])
    rays_o, rays_d = S.get_rays(H, W, K, c2w)
    vd = (rays_d / torch.norm(rays_d, dim=-1, keepdim=True)).reshape(-1, 3)
    o2, d2 = S.ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
    ref = torch.cat([o2.reshape(-1, 3), d2.reshape(-1, 3), torch.zeros(H * W, 1), torch.ones(H * W, 1), vd], -1)
    got = sray.make_ray_batch(H, W, K, c2w, 0., 1., ndc=True).cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-6, atol=1e-6)
    got8 = sray.make_ray_batch(H, W, K, c2w, 0., 1., ndc=True, use_viewdirs=False).cpu()
    assert got8.shape == (H * W, 8) and torch.equal(got8, got[:, :8])
    # and through render(): the NDC frame takes the one-kernel path and equals the generic path
    shapes = O.mlp_param_shapes()
    mc = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mc.load_state_dict(O.make_params(shapes, 21)); mc.to(DEV)
    q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision="fp32")
    kw = dict(network_fn=mc, network_query_fn=q, N_samples=64, N_importance=0, white_bkgd=False)
    with torch.no_grad():
        a = S.render(H, W, K, chunk=512, c2w=c2w.to(DEV), ndc=True, near=0., far=1., use_viewdirs=True, **kw)
        ro, rd = S.get_rays(H, W, K, c2w.to(DEV))
        b = S.render(H, W, K, chunk=512, rays=(ro, rd), ndc=True, near=0., far=1., use_viewdirs=True, **kw)
    # (the generic path assembles its rays with eager CUDA ops whose sums round differently by an ulp; the NDC warp
    # and the L=10 encoding amplify that - the exact check of the kernel is the ray-level one above)
    assert float((a[0] - b[0]).abs().max()) < 2e-4 and float((a[2] - b[2]).abs().max()) < 2e-4


@pytest.mark.parametrize("precrop", [None, 0.5])
def test_pick_batch(precrop):
    """nerf/run.py:652-681: N_rand DISTINCT pixels (inside the centre crop while precropping), their rays equal
    get_rays at those pixels, their targets equal the image there; the draw depends on the seed only."""
    H, W, N = 400, 400, 4096
    focal, K = _K(H, W)
    c2w = torch.from_numpy(O.pose_spherical(-33.0, -30.0, 4.0)[:3, :4])
    img = torch.rand(H, W, 3, device=DEV)
    rays, tgt, pix = sray.pick_batch(H, W, K, c2w, img, N, seed=17, near=2., far=6., precrop_frac=precrop, return_pixels=True)
    p = pix.cpu().numpy()
    assert len(np.unique(p)) == N                                   # replace=False
    y, x = p // W, p % W
    if precrop is not None:
        dH, dW = int(H // 2 * precrop), int(W // 2 * precrop)
        assert y.min() >= H // 2 - dH and y.max() <= H // 2 + dH - 1 and x.min() >= W // 2 - dW and x.max() <= W // 2 + dW - 1
        assert y.max() - y.min() > 1.8 * dH and x.max() - x.min() > 1.8 * dW
    else:
        # spread over the whole frame: every 100x100 cell holds about N/16 = 256 picks
        cells = np.bincount((y // 100) * 4 + x // 100, minlength=16)
        assert cells.min() > 180 and cells.max() < 340, cells
    assert torch.equal(tgt, img.reshape(-1, 3)[pix])
    assert torch.equal(rays, sray.make_ray_batch(H, W, K, c2w, 2., 6., pixels=pix))
    rays2, tgt2, pix2 = sray.pick_batch(H, W, K, c2w, img, N, seed=17, near=2., far=6., precrop_frac=precrop, return_pixels=True)
    assert torch.equal(pix, pix2)
    _, _, pix3 = sray.pick_batch(H, W, K, c2w, img, N, seed=18, near=2., far=6., precrop_frac=precrop, return_pixels=True)
    common = len(np.intersect1d(p, pix3.cpu().numpy()))
    expect = N * N / (len(np.unique(p)) and ((2 * int(H // 2 * precrop)) ** 2 if precrop else H * W))
    assert common < 3 * expect + 30                                  # independent draws overlap by ~N^2/n pixels
    # every pixel can be drawn exactly once: N_rand = the whole (small) crop is a permutation of it
    r, t, pall = sray.pick_batch(16, 12, 10.0, c2w, torch.rand(16, 12, 3, device=DEV), 16 * 12, seed=5, near=2., far=6.,
                                 return_pixels=True)
    assert sorted(pall.cpu().tolist()) == list(range(16 * 12))
    with pytest.raises(RuntimeError):
        sray.pick_batch(16, 12, 10.0, c2w, torch.rand(16, 12, 3, device=DEV), 16 * 12 + 1, seed=5, near=2., far=6.)


def test_render_path_frames_equal_render():
    """nerf/run.py:172-219: (rgbs, disps) numpy stacks; every frame equals render() on that pose; frames come back in
    order although the device->host copies are double-buffered behind the next frame's rendering."""
    H, W = 40, 36
    focal, K = _K(H, W)
    shapes = O.mlp_param_shapes()
    mc = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mc.load_state_dict(O.make_params(shapes, 23)); mc.to(DEV)
    mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(O.make_params(shapes, 43)); mf.to(DEV)
    q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision="fp32")
    kw = dict(network_fn=mc, network_query_fn=q, N_samples=64, N_importance=128, network_fine=mf, white_bkgd=True,
              perturb=False, raw_noise_std=0., near=2., far=6., use_viewdirs=True, ndc=False)
    poses = [torch.from_numpy(O.pose_spherical(a, -30.0, 4.0)) for a in (-100.0, -20.0, 60.0, 140.0, 175.0)]
    seen = []
    rgbs, disps = S.render_path(poses, (H, W, focal), K, 500, kw, on_frame=lambda i, r, d: seen.append(i))
    assert rgbs.shape == (5, H, W, 3) and disps.shape == (5, H, W) and seen == [0, 1, 2, 3, 4]
    with torch.no_grad():
        for i, c2w in enumerate(poses):
            rgb, disp, acc, _ = S.render(H, W, K, chunk=500, c2w=c2w[:3, :4].to(DEV), **kw)
            assert float((torch.from_numpy(rgbs[i]) - rgb.cpu()).abs().max()) < 1e-6, i
            m = ~torch.isnan(disp.cpu())
            assert float((torch.from_numpy(disps[i])[m] - disp.cpu()[m]).abs().max()) < 1e-5
    # render_factor (nerf/run.py:185-189)
    r2, d2 = S.render_path(poses[:1], (H, W, focal), K, 500, kw, render_factor=2)
    assert r2.shape == (1, H // 2, W // 2, 3)
