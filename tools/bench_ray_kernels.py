"""HBM roofline of the per-ray kernels (compositing fwd/bwd, resample, stratified, ray assembly) at render-chunk
and training sizes.  Algorithmic bytes per ray from BASELINE.md section 4."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import ops, _lib
from swnerf_b200 import synth

dev = "cuda"
peak = 6525.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > L2 (126 MB)


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()                                   # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


call, st = _lib.call, _lib.stream
for N in (4096, 32768, 262144):
    rays = torch.from_numpy(synth.blender_rays(N, 1)).to(dev)
    for Ssamp in (64, 192):
        raw = torch.randn(N, Ssamp, 4, device=dev)
        z = torch.sort(torch.rand(N, Ssamp, device=dev) * 4 + 2, -1)[0]
        rgb = torch.empty(N, 3, device=dev); disp = torch.empty(N, device=dev); acc = torch.empty(N, device=dev)
        dep = torch.empty(N, device=dev); w = torch.empty(N, Ssamp, device=dev); d_raw = torch.empty_like(raw)
        g = torch.randn(N, 3, device=dev)
        fwd = lambda: call("swnerf_composite_fwd", raw.data_ptr(), 4, z.data_ptr(), rays.data_ptr(), 11, 3, None, 1, N, Ssamp,
                           rgb.data_ptr(), disp.data_ptr(), acc.data_ptr(), w.data_ptr(), dep.data_ptr(), st())
        bwd = lambda: call("swnerf_composite_bwd", raw.data_ptr(), 4, z.data_ptr(), rays.data_ptr(), 11, 3, None, 1, N, Ssamp,
                           g.data_ptr(), None, None, None, None, acc.data_ptr(), dep.data_ptr(), d_raw.data_ptr(), st())
        ms = timeit(fwd)
        b = N * (24 * Ssamp + 36)
        print("composite_fwd  N=%6d S=%3d  %8.3f ms  %7.1f GB/s  %.2f of %.0f" % (N, Ssamp, ms, b / ms / 1e6, b / ms / 1e6 / peak, peak))
        ms = timeit(bwd)
        bb = N * (36 * Ssamp + 24)          # reads raw(16S) + z(4S) + g_rgb, writes d_raw(16S)
        print("composite_bwd  N=%6d S=%3d  %8.3f ms  %7.1f GB/s  %.2f" % (N, Ssamp, ms, bb / ms / 1e6, bb / ms / 1e6 / peak))
    z = torch.sort(torch.rand(N, 64, device=dev) * 4 + 2, -1)[0]
    w = torch.rand(N, 64, device=dev)
    zs = torch.empty(N, 128, device=dev); zf = torch.empty(N, 192, device=dev); zstd = torch.empty(N, device=dev)
    u = torch.rand(N, 128, device=dev)
    for variant in (0, 1):        # 0: one warp per ray, 1: eight lanes per ray (default)
        call("swnerf_set_resample_variant", variant)
        ms = timeit(lambda: call("swnerf_resample", z.data_ptr(), w.data_ptr(), None, 1, N, 64, 128, zs.data_ptr(), zf.data_ptr(), zstd.data_ptr(), st()))
        b = N * 4 * (64 + 64 + 128 + 192 + 1)
        print("resample(det)  v%d N=%6d        %8.3f ms  %7.1f GB/s  %.2f" % (variant, N, ms, b / ms / 1e6, b / ms / 1e6 / peak))
        ms = timeit(lambda: call("swnerf_resample", z.data_ptr(), w.data_ptr(), u.data_ptr(), 0, N, 64, 128, zs.data_ptr(), zf.data_ptr(), zstd.data_ptr(), st()))
        b = N * 4 * (64 + 64 + 128 + 128 + 192 + 1)
        print("resample(rand) v%d N=%6d        %8.3f ms  %7.1f GB/s  %.2f" % (variant, N, ms, b / ms / 1e6, b / ms / 1e6 / peak))
    call("swnerf_set_resample_variant", 1)
    # as render_rays calls it: no z_samples output (only z_std is consumed, run.py:416)
    ms = timeit(lambda: call("swnerf_resample", z.data_ptr(), w.data_ptr(), None, 1, N, 64, 128, None, zf.data_ptr(), zstd.data_ptr(), st()))
    b = N * 4 * (64 + 64 + 192 + 1)
    print("resample(det, no z_samples) v1 N=%6d  %8.3f ms  %7.1f GB/s  %.2f" % (N, ms, b / ms / 1e6, b / ms / 1e6 / peak))
    ms = timeit(lambda: call("swnerf_resample", z.data_ptr(), w.data_ptr(), u.data_ptr(), 0, N, 64, 128, None, zf.data_ptr(), zstd.data_ptr(), st()))
    b = N * 4 * (64 + 64 + 128 + 192 + 1)
    print("resample(rand, no z_samples) v1 N=%6d %8.3f ms  %7.1f GB/s  %.2f" % (N, ms, b / ms / 1e6, b / ms / 1e6 / peak))
