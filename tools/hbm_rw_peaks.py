"""Read-only, write-only and copy bandwidth of this B200's HBM (torch ops over 8 GiB, CUDA events, best of 10): the
training forward and the data-gradient kernel are WRITE streams, the weight-gradient kernels READ streams; the copy figure
in MEASURED_PEAKS.json (read + write bytes) is the roof of neither."""
import json
import torch
dev = "cuda"
n = 2 << 30            # 2 Gi floats = 8 GiB
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
a.fill_(1.0); b.fill_(2.0)


def best(fn, bytes_, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return bytes_ / (min(ts) * 1e-3) / 1e9


res = {"write_only_fill_gbs": best(lambda: a.fill_(3.0), n * 4),
       "write_only_memset_gbs": best(lambda: a.zero_(), n * 4),
       "read_only_sum_gbs": best(lambda: a.sum(), n * 4),
       "copy_read_plus_write_gbs": best(lambda: b.copy_(a), 2 * n * 4)}
print(json.dumps(res))
