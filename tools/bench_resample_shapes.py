import sys; sys.path.insert(0,'/root/repo')
import torch, swnerf_b200 as S
from swnerf_b200 import ops
dev='cuda'
for N,Sc,Ni in ((32768,64,128),(32768,64,64),(262144,64,64),(262144,64,128)):
    z=torch.sort(torch.rand(N,Sc,device=dev)*4+2,-1)[0]; w=torch.rand(N,Sc,device=dev)
    for det in (True,False):
        u=None
        f=lambda: ops.resample(z,w,Ni,det=det,u=u,want_samples=False)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/20
        byt=N*4*(Sc+Sc+(Sc+Ni))
        print(N,Sc,Ni,'det' if det else 'rand','%.3f ms %.2f TB/s'%(ms,byt/ms/1e9))
