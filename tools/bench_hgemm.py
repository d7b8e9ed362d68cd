import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
import swnerf_b200 as S
from swnerf_b200 import ops, _lib
dev="cuda"
for (M,N,K) in [(262144,128,128),(262144,256,256),(262144,128,84),(32000,128,128)]:
    A=torch.randn(M,K,device=dev); W=torch.randn(N,K,device=dev)/K**0.5; C=torch.empty(M,N,device=dev); b=torch.randn(N,device=dev)
    for tc in (False, True):
        f=lambda: ops._gemm(0,(A.data_ptr(),K),(W.data_ptr(),K),(C.data_ptr(),N),M,N,K,bias=b.data_ptr(),relu="elu",tc=tc)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        t0=time.time(); e0.record()
        for _ in range(20): f()
        e1.record(); t1=time.time(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/20
        print("M=%d N=%d K=%d tc=%s: %.3f ms/call (host %.3f ms/call)  %.1f TFLOP/s  %.0f GB/s" % (M,N,K,tc,ms,(t1-t0)*1e3/20, 2*M*N*K/ms/1e9, 4*M*(N+K)/ms/1e6))

# weight gradients: G[n_out, k_in] += dY^T X
for (M, n_out, k_in) in [(262144, 128, 128), (262144, 256, 256), (262144, 128, 84)]:
    dY = torch.randn(M, n_out, device=dev); X = torch.randn(M, k_in, device=dev); G = torch.zeros(n_out, k_in, device=dev)
    for tc in (False, True):
        f = lambda: ops._gemm(2, (dY.data_ptr(), n_out), (X.data_ptr(), k_in), (G.data_ptr(), k_in), n_out, k_in, M, accumulate=True, tc=tc)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("wgrad M=%d n_out=%d k_in=%d tc=%s: %.3f ms/call  %.1f TFLOP/s  %.0f GB/s" % (M, n_out, k_in, tc, ms, 2 * M * n_out * k_in / ms / 1e9, 4 * M * (n_out + k_in) / ms / 1e6))
