"""Three eager training steps at the bench shape (4096 rays, 64 + 128 samples, coarse + fine, forward + backward) for
profiling: `ncu -k regex:'mlp_(fwd4|bwd_data_pair|bwd_weight)' -s 16 -c 8 python tools/ncu_step.py` captures the eight
MLP launches of the third step (forward coarse / fine, then per network data-, pair weight-, edge weight-gradient)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swnerf_b200 as S
from swnerf_b200 import parallel, synth

dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mc = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mc.load_state_dict(synth.scene_params(mc, 21)); mc.to(dev)
mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(synth.scene_params(mf, 55)); mf.to(dev)
q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision="tc")
flat = parallel.FlatParams(list(mc.parameters()) + list(mf.parameters()))
rays = torch.from_numpy(synth.blender_rays(N, 100)).to(dev)
tgt = torch.from_numpy(np.random.RandomState(200).uniform(0, 1, (N, 3)).astype(np.float32)).to(dev)
for _ in range(steps):
    flat.zero_()
    ret = S.render_rays(rays, network_fn=mc, network_query_fn=q, N_samples=64, perturb=1.0, N_importance=128,
                        network_fine=mf, white_bkgd=True, raw_noise_std=0.0)
    parallel.two_loss_mse(ret["rgb_map"], ret["rgb0"], tgt, N).backward()
torch.cuda.synchronize()
print("done")
