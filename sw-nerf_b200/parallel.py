"""Ray-sharded data parallelism over the GPUs of one box (new functionality - the
reference is single-process).

One process per GPU.  The rays of a step are independent through the whole path, so rank r renders
rays [r*N/R, (r+1)*N/R) with no data-path collective; the only exchange is ONE all-reduce (SUM, fp32)
of a flat gradient buffer per step (1,191,688 floats = 4.77 MB for coarse+fine).  Every parameter's
.grad is a view into that buffer, so the backward kernels accumulate straight into it and no pack
copy precedes the collective.  The local loss is sum_sq / (N_global * 3), so the summed gradients
equal the gradients of the reference's mean loss (nerf/run.py:689-697).  Backend: nccl on GPUs,
gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_total: int, rank: int, world: int):
    """Contiguous block of rank `rank`; the first (n_total % world) ranks get one extra ray."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGrads:
    """One flat fp32 gradient buffer; p.grad of every parameter is a view into it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def check_views(self):
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat


def sharded_mse(pred, target, n_global: int):
    """sum of squares over the local rays / (global element count): sums to the reference's
    img2mse over the full batch (utils.py:12) after the gradient all-reduce."""
    return torch.sum((pred - target) ** 2) / float(n_global * pred.shape[-1])


def gather_rows(local: torch.Tensor, n_total: int, group=None):
    """All-gather row blocks of a sharded frame render (rgb[rays/R, 3] etc.) onto every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    mx = max(b - a for a, b in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:b - a] for o, (a, b) in zip(outs, sizes)], 0)
