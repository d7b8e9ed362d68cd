"""Profiling driver for the resample kernels: `python tools/prof_resample.py [N] [det|rand] [variant]` launches the
kernel three times (run it under ncu) and prints how many rays took the generic fallback."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swnerf_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
det = (sys.argv[2] != "rand") if len(sys.argv) > 2 else True
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = "cuda"
z = torch.sort(torch.rand(N, 64, device=dev) * 4 + 2, -1)[0]
w = torch.rand(N, 64, device=dev)
u = torch.rand(N, 128, device=dev)
zs = torch.empty(N, 128, device=dev); zf = torch.empty(N, 192, device=dev); zstd = torch.empty(N, device=dev)
_lib.call("swnerf_set_resample_variant", variant)
_lib.resample_fallbacks(reset=True)
for _ in range(3):
    _lib.call("swnerf_resample", z.data_ptr(), w.data_ptr(), None if det else u.data_ptr(), int(det), N, 64, 128,
              zs.data_ptr(), zf.data_ptr(), zstd.data_ptr(), _lib.stream())
torch.cuda.synchronize()
print("N=%d det=%s variant=%d: fallback rays per launch = %.1f" % (N, det, variant, _lib.resample_fallbacks() / 3.0))
