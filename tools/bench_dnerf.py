"""D-NeRF training step (config #4 shape: direct_temporal, N_rand rays, 64+128, single model with no_grad coarse pass,
tv-loss second render): fused tcgen05 path vs the fp32 GEMM path, and their agreement."""
import sys, os, tempfile
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import dnerf
from swnerf_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda")
tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "e"), exist_ok=True)
res = {}
for prec in ("fp32", "tc"):
    args = Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64,
                     netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, lrate=5e-4,
                     ft_path=None, basedir=tmp, expname="e", no_reload=True, perturb=1.0, white_bkgd=True,
                     raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False, nerf_type="direct_temporal",
                     use_two_models_for_fine=False, not_zero_canonical=False, do_half_precision=False,
                     swnerf_precision=prec)
    kw, _, _, gv, opt = dnerf.create_nerf(args, device=dev)
    model = kw["network_fn"]
    model.load_state_dict(synth.scene_params(model, 332)); model.to(dev)
    kw = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}
    rays = torch.from_numpy(synth.blender_rays(N, 31, frame_time=0.37)).to(dev)
    rays2 = rays.clone(); rays2[:, 8] = 0.38
    tgt = torch.rand(N, 3, device=dev)

    def step():
        opt.zero_grad()
        torch.manual_seed(0)
        ret = dnerf.render_rays(rays, **kw)
        loss = torch.mean((ret["rgb_map"] - tgt) ** 2)
        ret2 = dnerf.render_rays(rays2, z_vals=ret["z_vals"].detach(), **kw)          # tv-loss render (run_dnerf.py:690-710)
        loss = loss + 0.1 * torch.sum((ret["position_delta"] - ret2["position_delta"]) ** 2)
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
    g = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    res[prec] = (e0.elapsed_time(e1) / reps, l.item(), g.clone())
    print("%s: %.2f ms/step  (%.0f rays/s)  loss %.6f" % (prec, res[prec][0], N / res[prec][0] * 1e3, l.item()))
a, b = res["tc"][2], res["fp32"][2]
print("tc vs fp32: loss diff %.2e, grad relL2 %.3e" % (abs(res["tc"][1] - res["fp32"][1]), float((a - b).norm() / b.norm())))
