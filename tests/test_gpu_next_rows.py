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
