"""Per-tensor gradient error of the fused tcgen05 path vs the CPU oracle at identical sample positions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import ops
from oracle import nerf_oracle as O
DEV = "cuda"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rays = O.blender_rays(N, 46)
shapes = O.mlp_param_shapes()
pc, pf = O.make_params(shapes, 21), O.make_params(shapes, 55)
pfr = {k: v.clone().requires_grad_() for k, v in pf.items()}
ref = O.render_rays(torch.from_numpy(rays), pc, pfr, 64, 128, white_bkgd=True, retraw=True)
z_fine = ref["z_vals"].detach()
cot = torch.from_numpy(np.random.RandomState(1).normal(size=(N, 3)).astype(np.float32))
(ref["rgb_map"] * cot).sum().backward()
for prec in ("fp32", "tc"):
    mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(pf); mf.to(DEV)
    q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision=prec)
    raw = q.query_rays(T(rays), z_fine.to(DEV).contiguous(), mf, 8)
    rgb, disp, acc, w, depth = ops.composite(raw, z_fine.to(DEV).contiguous(), T(rays), 3, None, True)
    (rgb * cot.to(DEV)).sum().backward()
    print("==", prec, "raw relL2 %.2e  rgb relmax %.2e" % (float((raw.cpu() - ref["raw"]).norm() / ref["raw"].norm()),
          float((rgb.cpu() - ref["rgb_map"]).abs().max() / ref["rgb_map"].abs().max())))
    gmax = max(float(pfr[n].grad.abs().max()) for n in pfr)
    tot_n = tot_d = 0.0
    for n, p in mf.named_parameters():
        g, r = p.grad.cpu().double(), pfr[n].grad.double()
        tot_n += float((g - r).pow(2).sum()); tot_d += float(r.pow(2).sum())
        print("  %-26s relL2 %.2e   maxerr/gmax %.2e   |g| %.2e" % (n, float((g - r).norm() / r.norm()), float((g - r).abs().max()) / gmax, float(r.norm())))
    print("  flat relL2 %.3e" % (tot_n / tot_d) ** 0.5)
