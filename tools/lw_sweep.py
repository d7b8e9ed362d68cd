"""Time the layer-pipelined backward for several role allocations (SWNERF_LW_ROLES is read per launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import _lib, synth
dev = 'cuda'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(synth.scene_params(mf, 55)); mf.to(dev)
q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision='tc')
rays = torch.from_numpy(synth.blender_rays(N, 7)).to(dev)
z = torch.sort(torch.rand(N, 192, device=dev) * 4 + 2, -1)[0]
cot = torch.randn(N, 192, 4, device=dev)
def run(variant, roles=None):
    if roles: os.environ["SWNERF_LW_ROLES"] = roles
    _lib.call('swnerf_tc_set_bwd_variant', variant)
    ts = []
    for _ in range(4):
        raw = q.query_rays(rays, z, mf, 8)
        l = (raw * cot).sum()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); l.backward(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
print('two-kernel: %.3f ms' % run(0))
for roles in sys.argv[2:]:
    assert sum(int(x) for x in roles.split(',')) <= 148 and len(roles.split(',')) == 18, roles
    print('LW %-60s %.3f ms' % (roles, run(1, roles)))
