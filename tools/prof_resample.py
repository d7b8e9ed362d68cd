import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swnerf_b200 import _lib
N = 65536
dev = "cuda"
z = torch.sort(torch.rand(N, 64, device=dev) * 4 + 2, -1)[0]
w = torch.rand(N, 64, device=dev)
zs = torch.empty(N, 128, device=dev); zf = torch.empty(N, 192, device=dev); zstd = torch.empty(N, device=dev)
for _ in range(3):
    _lib.call("swnerf_resample", z.data_ptr(), w.data_ptr(), None, 1, N, 64, 128, zs.data_ptr(), zf.data_ptr(), zstd.data_ptr(), _lib.stream())
torch.cuda.synchronize()
