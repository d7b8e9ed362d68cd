import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import _lib, synth
dev = 'cuda'
N = int(sys.argv[1]); variant = int(sys.argv[2])
mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(synth.scene_params(mf, 55)); mf.to(dev)
q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision='tc')
rays = torch.from_numpy(synth.blender_rays(N, 7)).to(dev)
z = torch.sort(torch.rand(N, 192, device=dev) * 4 + 2, -1)[0]
cot = torch.randn(N, 192, 4, device=dev)
_lib.call('swnerf_tc_set_bwd_variant', variant)
for it in range(2):
    raw = q.query_rays(rays, z, mf, 8)
    (raw * cot).sum().backward()
torch.cuda.synchronize()
print('done')
