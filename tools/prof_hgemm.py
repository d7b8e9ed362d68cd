"""Profiling driver for the layer-wise tcgen05 GEMMs (run under ncu): one 262,144 x 256 x 256 forward layer with bias +
ReLU and the matching weight gradient, three launches each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swnerf_b200 import ops
M, N, K = 262144, 256, 256
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
C = torch.empty(M, N, device="cuda"); b = torch.randn(N, device="cuda"); G = torch.zeros(N, K, device="cuda")
for _ in range(3):
    ops._gemm(0, (A.data_ptr(), K), (W.data_ptr(), K), (C.data_ptr(), N), M, N, K, bias=b.data_ptr(), relu="relu", tc=True)
    ops._gemm(2, (C.data_ptr(), N), (A.data_ptr(), K), (G.data_ptr(), K), N, K, M, accumulate=True, tc=True)
torch.cuda.synchronize()
