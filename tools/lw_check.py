import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import _lib, ops, synth, tc
dev = 'cuda'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(synth.scene_params(mf, 55)); mf.to(dev)
q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision='tc')
rays = torch.from_numpy(synth.blender_rays(N, 7)).to(dev)
z = torch.sort(torch.rand(N, 192, device=dev) * 4 + 2, -1)[0]
cot = torch.randn(N, 192, 4, device=dev)
res = {}
for variant in (0, 1):
    _lib.call('swnerf_tc_set_bwd_variant', variant)
    for p in mf.parameters(): p.grad = None
    raw = q.query_rays(rays, z, mf, 8)
    (raw * cot).sum().backward()
    torch.cuda.synchronize()
    res[variant] = torch.cat([p.grad.reshape(-1) for p in mf.param_list()]).clone()
    # timing
    ts = []
    for _ in range(5):
        raw = q.query_rays(rays, z, mf, 8)
        l = (raw * cot).sum()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); l.backward(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print('variant', variant, 'backward ms (incl. autograd glue):', ['%.3f' % t for t in ts], flush=True)
a, b = res[0], res[1]
print('rel l2 LW vs two-kernel: %.3e  max abs %.3e (max |g| %.3e)' % (float((a - b).norm() / a.norm()), float((a - b).abs().max()), float(a.abs().max())))
names = [n for n, _ in mf.named_parameters()]
off = 0
for p_, n in zip(mf.param_list(), ['pts%d.%s' % (i // 2, 'w' if i % 2 == 0 else 'b') for i in range(16)] + ['views.w', 'views.b', 'feat.w', 'feat.b', 'alpha.w', 'alpha.b', 'rgb.w', 'rgb.b']):
    k = p_.numel(); d = (a[off:off + k] - b[off:off + k]).norm() / a[off:off + k].norm().clamp_min(1e-30); off += k
    print('  %-8s %.2e' % (n, float(d)))
