"""Tensor-pipe issue-rate probe: cycles per 128 x N x 16 fp16 MMA, A from shared memory vs tensor memory."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swnerf_b200 import _lib

out = torch.zeros(512, device="cuda")
cases = [(0, 256), (1, 256), (0, 128), (1, 128), (1 + 16, 128), (0 + 16, 256), (1 + 16, 256)]
for variant, N in cases:
    for _ in (0,):
        for _ in range(2):
            _lib.call("swnerf_tc_probe", variant, N, 4000, out.data_ptr(), _lib.stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("swnerf_tc_probe", variant, N, 4000, out.data_ptr(), _lib.stream())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        c = out[:148]
        flops = 148 * 16000 * 2 * 128 * N * 16
        print("A from %s  N=%3d: %.1f cycles/MMA (min %.1f max %.1f)  %.3f ms  %.0f TFLOP/s" %
              (("TMEM" if variant & 1 else "smem") + " commits/4=%d" % ((variant >> 4) & 3), N, c.mean().item(), c.min().item(), c.max().item(), ms, flops / ms / 1e9))

# CTA pair (cta_group::2): M = 256 per MMA, each SM reads its A rows and HALF of B
A = torch.randn(256, 256, device="cuda"); B = torch.randn(256, 256, device="cuda")
scratch = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
cyc = torch.zeros(128, device="cuda")
for N, nc in ((256, 0), (128, 0), (128, 1), (128, 2), (256, 1)):
    it = 1000 + (nc << 20)
    for _ in range(2):
        _lib.call("swnerf_tc_selftest_pair", A.data_ptr(), B.data_ptr(), None, N, 256, it, 74, cyc.data_ptr(), scratch.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    print("CTA pair  M=256 N=%3d commits/4=%d: %.1f cycles/MMA (min %.1f max %.1f)" %
          (N, nc, cyc[:74].mean().item(), cyc[:74].min().item(), cyc[:74].max().item()))
