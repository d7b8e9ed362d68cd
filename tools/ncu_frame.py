"""One full 800 x 800 frame through render_path (config #3) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import parallel, synth
dev = "cuda"
mc = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mc.load_state_dict(synth.scene_params(mc, 21)); mc.to(dev)
mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(synth.scene_params(mf, 55)); mf.to(dev)
q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision="tc")
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 800
focal = 0.5 * W / np.tan(0.5 * 0.6911112)
K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
poses = [torch.from_numpy(synth.pose_spherical(40.0, -30.0, 4.0)[:3, :4])]
kw = dict(network_fn=mc, network_query_fn=q, N_samples=64, perturb=0.0, N_importance=128, network_fine=mf,
          white_bkgd=True, raw_noise_std=0.0)
for _ in range(2):
    parallel.render_path(poses, (H, W, focal), K, 1024 * 32, kw, near=2.0, far=6.0)
torch.cuda.synchronize()
print("done")
