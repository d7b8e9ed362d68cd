"""Cycle counters of the layer-pipelined backward, per role (needs a library built with
SWNERF_NVCC_EXTRA=-DSWNERF_LW_DEBUG python sw-nerf_b200/build.py --force):  python tools/lw_profile.py [rays]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import _lib, synth
dev = 'cuda'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = _lib.lib()
if not hasattr(L, "swnerf_tc_lw_debug"):
    sys.exit("library was not built with -DSWNERF_LW_DEBUG")
L.swnerf_tc_lw_debug.argtypes = [ctypes.c_void_p]
mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(synth.scene_params(mf, 55)); mf.to(dev)
q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision='tc')
rays = torch.from_numpy(synth.blender_rays(N, 7)).to(dev)
z = torch.sort(torch.rand(N, 192, device=dev) * 4 + 2, -1)[0]
cot = torch.randn(N, 192, 4, device=dev)
_lib.call('swnerf_tc_set_bwd_variant', 1)
dbg = torch.zeros(148 * 12, dtype=torch.int64, device=dev)
for it in range(3):
    raw = q.query_rays(rays, z, mf, 8)
    l = (raw * cot).sum()
    dbg.zero_()
    L.swnerf_tc_lw_debug(dbg.data_ptr() if it == 2 else None)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); l.backward(); e1.record(); torch.cuda.synchronize()
    print('backward %.3f ms' % e0.elapsed_time(e1))
d = dbg.cpu().numpy().reshape(148, 12)
names = ['D(H)'] + ['D(%d)' % l for l in range(7, 0, -1)] + ['W(H)'] + ['W(%d)' % l for l in range(7, 0, -1)] + ['W(0)', 'W(5p)']
print('role   ctas tiles/cta  total_kclk | per tile (clk): total  ready-wait  imgfree-wait  d_full-wait  mma-wait-img  mma-wait-w | prep  epilogue  store-wait-out  store')
for r in range(18):
    rows = d[d[:, 7] == r]
    rows = rows[rows[:, 6] > 0]
    if len(rows) == 0: continue
    t = rows[:, 6].astype(np.float64)
    print('%-6s %4d %8.1f %10.0f | %8.0f %10.0f %12.0f %12.0f %12.0f %10.0f | %6.0f %8.0f %10.0f %8.0f' % (
        names[r], len(rows), t.mean(), rows[:, 0].mean() / 1e3, (rows[:, 0] / t).mean(), (rows[:, 1] / t).mean(),
        (rows[:, 2] / t).mean(), (rows[:, 3] / t).mean(), (rows[:, 4] / t).mean(), (rows[:, 5] / t).mean(),
        (rows[:, 8] / t).mean(), (rows[:, 9] / t).mean(), (rows[:, 10] / t).mean(), (rows[:, 11] / t).mean()))
