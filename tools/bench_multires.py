"""MultiRes D-NeRF level step (multires_dnerf.py pyramid, config #5): one level network with the encoding widths of
levels 1-2 (PE 10 / 4 / 10) or level 0 (20 / 8 / 20), N rays x (64 + 128) samples, render + backward.  No fused
kernel exists for these widths: 'fp32' = fp32 SIMT GEMMs, 'tc' = layer-at-a-time tcgen05 GEMM (forward + data
gradients).  `python tools/bench_multires.py [N] [Lp Lt Ld]`"""
import sys, os, tempfile
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import dnerf

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
channels = tuple(int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (10, 4, 10)
dev = torch.device("cuda")
tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "e"), exist_ok=True)
H = W = 800
focal = 0.5 * W / np.tan(0.5 * 0.6911112)
c2w = torch.eye(4, device=dev)[:3, :4].clone(); c2w[2, 3] = 4.0
rays_o, rays_d = S.get_rays(H, W, focal, c2w)
sel = torch.randperm(H * W, device=dev)[:N]
o, d = rays_o.reshape(-1, 3)[sel], rays_d.reshape(-1, 3)[sel]
rays = torch.cat([o, d, torch.full((N, 1), 2.0, device=dev), torch.full((N, 1), 6.0, device=dev),
                  torch.full((N, 1), 0.37, device=dev), d / d.norm(dim=-1, keepdim=True)], -1).contiguous()
rays._swnerf_frame_time = 0.37
tgt = torch.rand(N, 3, device=dev)
res = {}
for prec in ("fp32", "tc"):
    args = Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64,
                     netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=1 << 30, lrate=5e-4,
                     ft_path=None, basedir=tmp, expname="e", no_reload=True, perturb=1.0, white_bkgd=True,
                     raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False, nerf_type="direct_temporal",
                     use_two_models_for_fine=False, not_zero_canonical=False, do_half_precision=False,
                     swnerf_precision=prec)
    torch.manual_seed(0)
    kw, _, _, gv, opt = dnerf.create_nerf_multires(args, channels, 1, device=dev)
    kw = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}

    def step():
        opt.zero_grad()
        torch.manual_seed(1)
        ret = dnerf.render_rays(rays, **kw)
        loss = torch.mean((ret["rgb_map"] - tgt) ** 2)
        loss.backward()
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
    print("MultiRes level PE %s, %d rays: %s %.2f ms/step (%.0f rays/s), loss %.5f"
          % (channels, N, prec, res[prec], N / res[prec] * 1e3, l.item()))
print("tc / fp32 speed-up: %.2fx" % (res["fp32"] / res["tc"]))
