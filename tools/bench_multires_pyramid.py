"""MultiRes D-NeRF pyramid step (config #5: multires_dnerf.py:665 - four level networks with PE (20,8,20), (10,4,10),
(10,4,10), identity and 1024 / 256 / 64 / 16 rays per step, one loss.backward over all levels, Adam):
`python tools/bench_multires_pyramid.py [rays_scale]`.  'fp32' = fp32 SIMT GEMMs, 'tc' = the fused tcgen05 kernels (every level's encoding widths: SWNERF_TC_ENC)."""
import sys, os, tempfile
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import dnerf

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
LEVELS = [((20, 8, 20), 1024 * scale), ((10, 4, 10), 256 * scale), ((10, 4, 10), 64 * scale), ((-1, -1, -1), 16 * scale)]
dev = torch.device("cuda")
tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "e"), exist_ok=True)
H = W = 800
focal = 0.5 * W / np.tan(0.5 * 0.6911112)
c2w = torch.eye(4, device=dev)[:3, :4].clone(); c2w[2, 3] = 4.0
rays_o, rays_d = S.get_rays(H, W, focal, c2w)


def make_rays(n):
    sel = torch.randperm(H * W, device=dev)[:n]
    o, d = rays_o.reshape(-1, 3)[sel], rays_d.reshape(-1, 3)[sel]
    r = torch.cat([o, d, torch.full((n, 1), 2.0, device=dev), torch.full((n, 1), 6.0, device=dev),
                   torch.full((n, 1), 0.37, device=dev), d / d.norm(dim=-1, keepdim=True)], -1).contiguous()
    r._swnerf_frame_time = 0.37
    return r


res = {}
for prec in ("fp32", "tc"):
    args = Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64,
                     netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=1 << 30, lrate=5e-4,
                     ft_path=None, basedir=tmp, expname="e", no_reload=True, perturb=1.0, white_bkgd=True,
                     raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False, nerf_type="direct_temporal",
                     use_two_models_for_fine=False, not_zero_canonical=False, do_half_precision=False,
                     swnerf_precision=prec)
    torch.manual_seed(0)
    levels, params = [], []
    for li, (ch, n) in enumerate(LEVELS):
        kw, _, _, gv, _ = dnerf.create_nerf_multires(args, ch, li, device=dev)
        kw = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}
        levels.append((kw, make_rays(n), torch.rand(n, 3, device=dev)))
        params += gv
    opt = torch.optim.Adam(params, lr=5e-4)

    def step():
        opt.zero_grad()
        loss = 0.
        for kw, rays, tgt in levels:
            ret = dnerf.render_rays(rays, **kw)
            loss = loss + torch.mean((ret["rgb_map"] - tgt) ** 2)
        loss.backward()                                   # multires_dnerf.py:1005: one backward over all levels
        opt.step()
        return loss
    for _ in range(3):
        l = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        l = step()
    e1.record(); torch.cuda.synchronize()
    res[prec] = e0.elapsed_time(e1) / reps
    n_rays = sum(n for _, n in LEVELS)
    print("MultiRes pyramid, %d rays (%s): %s %.2f ms/step (%.0f rays/s), loss %.5f"
          % (n_rays, "/".join(str(n) for _, n in LEVELS), prec, res[prec], n_rays / res[prec] * 1e3, l.item()))
print("tc / fp32 speed-up: %.2fx" % (res["fp32"] / res["tc"]))
