"""Three eager D-NeRF training steps (config #4: N_rand=500, 64+128, coarse no-grad, tv-loss second render) for launch lists."""
import os, sys, tempfile
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import swnerf_b200 as S
from swnerf_b200 import dnerf, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda")
tmp = tempfile.mkdtemp(); os.makedirs(os.path.join(tmp, "e"), exist_ok=True)
args = Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64,
                 netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, lrate=5e-4,
                 ft_path=None, basedir=tmp, expname="e", no_reload=True, perturb=1.0, white_bkgd=True,
                 raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False, nerf_type="direct_temporal",
                 use_two_models_for_fine=False, not_zero_canonical=False, do_half_precision=False, swnerf_precision="tc")
kw, _, _, gv, opt = dnerf.create_nerf(args, device=dev)
model = kw["network_fn"]; model.load_state_dict(synth.scene_params(model, 332)); model.to(dev)
kw = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}
rays = torch.from_numpy(synth.blender_rays(N, 31, frame_time=0.37)).to(dev); rays._swnerf_frame_time = 0.37
rays2 = rays.clone(); rays2[:, 8] = 0.38; rays2._swnerf_frame_time = 0.38
tgt = torch.rand(N, 3, device=dev)
for _ in range(3):
    opt.zero_grad()
    ret = dnerf.render_rays(rays, **kw)
    loss = torch.mean((ret["rgb_map"] - tgt) ** 2)
    ret2 = dnerf.render_rays(rays2, z_vals=ret["z_vals"].detach(), **kw)
    loss = loss + 0.1 * torch.sum((ret["position_delta"] - ret2["position_delta"]) ** 2)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("done")
