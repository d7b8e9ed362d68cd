import sys
sys.path.insert(0, "/root/repo")
import torch
from swnerf_b200 import ops
M,N,K=262144,128,128
A=torch.randn(M,K,device="cuda"); W=torch.randn(N,K,device="cuda")/K**0.5; C=torch.empty(M,N,device="cuda"); b=torch.randn(N,device="cuda")
for _ in range(3):
    ops._gemm(0,(A.data_ptr(),K),(W.data_ptr(),K),(C.data_ptr(),N),M,N,K,bias=b.data_ptr(),relu="elu",tc=True)
torch.cuda.synchronize()
