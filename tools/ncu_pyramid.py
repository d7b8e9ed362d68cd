"""Three eager MultiRes pyramid steps (config #5: four level networks, 1024 / 256 / 64 / 16 rays, coarse no-grad + fine
pass each, one backward) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys, tempfile
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import dnerf, parallel, synth
dev = torch.device("cuda")
LEVELS = [((20, 8, 20), 1024), ((10, 4, 10), 256), ((10, 4, 10), 64), ((-1, -1, -1), 16)]
tmp = tempfile.mkdtemp(); os.makedirs(os.path.join(tmp, "e"), exist_ok=True)
args = Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64,
                 netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=1 << 30, lrate=5e-4,
                 ft_path=None, basedir=tmp, expname="e", no_reload=True, perturb=1.0, white_bkgd=True,
                 raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False, nerf_type="direct_temporal",
                 use_two_models_for_fine=False, not_zero_canonical=False, do_half_precision=False, swnerf_precision="tc")
levels, params = [], []
for li, (ch, n) in enumerate(LEVELS):
    kw, _, _, gv, _ = dnerf.create_nerf_multires(args, ch, li, device=dev)
    m = kw["network_fn"]; m.load_state_dict(synth.scene_params(m, 700 + li)); m.to(dev)
    kw = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}
    rays = torch.from_numpy(synth.blender_rays(n, 710 + li, frame_time=0.37)).to(dev); rays._swnerf_frame_time = 0.37
    tgt = torch.rand(n, 3, device=dev)
    levels.append((kw, rays, tgt, n)); params += gv
flat = parallel.FlatParams(params); opt = parallel.FlatAdam(flat, lr=5e-4)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    flat.zero_()
    loss = 0.
    for kw, rays, tgt, n in levels:
        ret = dnerf.render_rays(rays, **kw)
        loss = loss + parallel.sharded_mse(ret["rgb_map"], tgt, n)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("done")
