"""Micro-benchmark of the fused tcgen05 forward (inference mode): TFLOP/s on algorithmic FLOPs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swnerf_b200 as S
from swnerf_b200 import ops, tc, _lib
from swnerf_b200 import synth

dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
Ssamp = int(sys.argv[2]) if len(sys.argv) > 2 else 192
train = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rays = torch.from_numpy(synth.blender_rays(N, 1)).to(dev)
m = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); m.load_state_dict(synth.scene_params(m, 21)); m.to(dev)
q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision="tc")
z = torch.sort(torch.rand(N, Ssamp, device=dev) * 4 + 2, -1)[0]
st = tc.packed_weights(m)
raw = torch.empty(N, Ssamp, 4, device=dev)
ws = None
if train:
    ws = torch.empty(int(_lib.lib().swnerf_tc_workspace_bytes(N * Ssamp, 1, tc.ENC_DEFAULT)), dtype=torch.uint8, device=dev)
def run():
    _lib.call("swnerf_tc_mlp_fwd", rays.data_ptr(), 11, 8, z.data_ptr(), N, Ssamp, st.fwd.data_ptr(), tc.ENC_DEFAULT, raw.data_ptr(),
              None if ws is None else ws.data_ptr(), train, _lib.stream())
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flop = 2 * 593408 * N * Ssamp
print("N=%d S=%d train=%d: %.3f ms  %.1f TFLOP/s (algorithmic)  %.1f Mpts/s" % (N, Ssamp, train, ms, flop / ms / 1e9, N * Ssamp / ms / 1e3))
