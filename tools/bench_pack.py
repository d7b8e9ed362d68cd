"""Warm-cache timing of the per-step weight re-pack (fold + pack, forward and transposed images) and of a small backward
(un-fold included): `python tools/bench_pack.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import swnerf_b200 as S
from swnerf_b200 import tc, synth
dev = "cuda"
m = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); m.load_state_dict(synth.scene_params(m, 21)); m.to(dev)


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def repack():
    tc.GENERATION += 1
    tc.packed_weights(m, need_bwd=True)


print("re-pack (fold + pack_fwd + pack_bwd), warm L2, graph replay: %.1f us per network" % timed(repack))
q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision="tc")
rays = torch.from_numpy(synth.blender_rays(64, 7)).to(dev)
z = torch.sort(torch.rand(64, 64, device=dev) * 4 + 2, -1)[0]
cot = torch.randn(64, 64, 4, device=dev)


def fb():
    raw = q.query_rays(rays, z, m, 8)
    (raw * cot).sum().backward()


print("forward + backward of 32 tiles (absmax, data, weight, un-fold), graph replay: %.1f us" % timed(fb, 50))
