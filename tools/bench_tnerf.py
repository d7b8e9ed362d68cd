"""T-NeRF training step (t_nerf/run_tnerf.py shape: one TNeRF, N_rand rays x 64 stratified samples, fwd + bwd + Adam)
on the fp32 GEMM kernels with the ELU epilogue.  `python tools/bench_tnerf.py [N_rand]`"""
import sys, os, tempfile
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import tnerf, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda")
tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "e"), exist_ok=True)
args = Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=0, N_samples=64, netdepth=8,
                 netwidth=256, netchunk=1 << 30, lrate=5e-4, ft_path=None, basedir=tmp, expname="e", no_reload=True,
                 perturb=1.0, white_bkgd=True, raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False,
                 nerf_type="tnerf", do_half_precision=False)
torch.manual_seed(0)
kw, kw_test, _, gv, opt = tnerf.create_nerf(args, device=dev)
kw = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}
# synthetic 800x800 Blender-shaped rays of one camera at one time (load_blender.py:133-134 focal)
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


def timeit(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def fwd_bwd():
    for p in gv:
        if p.grad is not None:
            p.grad.zero_()
    ret = tnerf.render_rays(rays, **kw)
    loss = torch.mean((ret["rgb_map"] - tgt) ** 2)
    loss.backward()
    return loss


# 1. forward + backward captured once in a CUDA graph (static ray / target buffers, as bench.py does for the vanilla
#    step), Adam after the replay.  This comes FIRST: autograd ties a parameter's gradient accumulation to the stream it
#    first ran on, and a capture cannot wait on the legacy default stream.
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        fwd_bwd()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    gl = fwd_bwd()


def graphed():
    g.replay()
    opt.step()
    return gl


for _ in range(3):
    graphed()
ms_g, lg = timeit(graphed)
print("T-NeRF step, %d rays x 64 samples, CUDA graph of fwd+bwd: %.2f ms (%.0f rays/s), loss %.5f"
      % (N, ms_g, N / ms_g * 1e3, lg.item()))


# 2. the eager step: bound by the host (61 library calls + autograd bookkeeping per step)
def step():
    opt.zero_grad()
    ret = tnerf.render_rays(rays, **kw)
    loss = torch.mean((ret["rgb_map"] - tgt) ** 2)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = _lib.launch_count()
ms, l = timeit(step)
flop = 3 * 2 * N * 64 * sum(p.numel() for p in gv if p.dim() == 2)       # fwd + dgrad + wgrad
print("T-NeRF step, %d rays x 64 samples, eager: %.2f ms (%.0f rays/s), %d library launches/step; graph: %.1f TFLOP/s "
      "fp32 SIMT" % (N, ms, N / ms * 1e3, (_lib.launch_count() - n0) // 20, flop / ms_g / 1e9))
