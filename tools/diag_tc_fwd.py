import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import ops, tc, _lib
from oracle import nerf_oracle as O
DEV="cuda"
N=int(sys.argv[1]); Sm=int(sys.argv[2])
rays=torch.from_numpy(O.blender_rays(N,71)).to(DEV)
m=S.vallina_NeRF(8,256,63,27,5,[4],True); m.load_state_dict(O.make_params(O.mlp_param_shapes(),21)); m.to(DEV)
emb=(S.get_embedder(10,3,0)[0], S.get_embedder(4,3,0)[0])
q32=S.NetworkQuery(*emb, precision="fp32"); qtc=S.NetworkQuery(*emb, precision="tc")
z=ops.stratified_z(rays, Sm)
with torch.no_grad():
    ref=q32.query_rays(rays,z,m,8)
    inf=qtc.query_rays(rays,z,m,8)
trn=qtc.query_rays(rays,z,m,8)
torch.cuda.synchronize()
for name,t in (("infer",inf),("train",trn.detach())):
    err=(t-ref).abs().reshape(-1,4).max(-1)[0]
    tiles=err.reshape(-1,128).max(-1)[0] if err.numel()%128==0 else None
    print(name,"max err %.3e  relL2 %.3e"%(float(err.max()), float((t-ref).norm()/ref.norm())))
    if tiles is not None:
        bad=(tiles>0.05).nonzero().flatten()
        print("  bad tiles:",bad.numel(),"of",tiles.numel(), bad[:40].tolist())
