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

from . import _lib, tc
from ._lib import call, stream


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
        self.offsets = {}
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self.offsets[id(p)] = (off, off + p.numel())
            off += p.numel()
        tc.direct_grads(self.params, True)     # the fused backward accumulates straight into these views

    def span(self, params):
        """(lo, hi) of the contiguous slice of the flat buffer that holds `params` (e.g. one network)."""
        spans = sorted(self.offsets[id(p)] for p in params if id(p) in self.offsets)
        lo, hi = spans[0][0], spans[-1][1]
        if sum(b - a for a, b in spans) != hi - lo:
            raise ValueError("these parameters are not contiguous in the flat buffer")
        return lo, hi

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


# ------------------------------------------------------------------------------------------------
# the step after the path - loss and optimizer on flat buffers
# ------------------------------------------------------------------------------------------------
class FlatParams(FlatGrads):
    """Parameters AND gradients as views into two flat fp32 buffers (state_dict names/shapes unchanged:
    each nn.Parameter keeps its identity, only its storage moves)."""

    def __init__(self, params):
        params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in params)
        dev = params[0].device
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            view = self.flat_p[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            off += p.numel()
        super().__init__(params)

    def check_param_views(self):
        base = self.flat_p.untyped_storage().data_ptr()
        return all(p.data.untyped_storage().data_ptr() == base for p in self.params)


class FlatAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) as created by create_nerf (nerf/run.py:254), as ONE
    kernel over the flat buffers (the reference runs ~48 small tensors through a multi-tensor kernel)."""

    def __init__(self, flat: FlatParams, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        self.flat, self.lr, self.betas, self.eps = flat, lr, betas, eps
        self.exp_avg = torch.zeros_like(flat.flat_p)
        self.exp_avg_sq = torch.zeros_like(flat.flat_p)
        self.step_count = 0
        self.param_groups = [{"lr": lr}]          # the runners decay the rate through param_groups (nerf/run.py:704-708)

    def zero_grad(self, set_to_none=False):
        self.flat.zero_()

    def step(self):
        self.step_count += 1
        call("swnerf_adam_flat", self.flat.flat_p.data_ptr(), self.flat.flat.data_ptr(), self.exp_avg.data_ptr(),
             self.exp_avg_sq.data_ptr(), self.flat.flat_p.numel(), float(self.param_groups[0]["lr"]),
             float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_count, stream())
        tc.GENERATION += 1                        # packed fp16 weight images are stale now

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0]["lr"] = sd["lr"]


class _TwoLossMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, target, n_global):
        a, target = a.contiguous(), target.contiguous()
        b = None if b is None else b.contiguous()
        da = torch.empty_like(a)
        db = None if b is None else torch.empty_like(b)
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        call("swnerf_mse2", _lib.ptr(a), None if b is None else _lib.ptr(b), _lib.ptr(target), a.numel(),
             1.0 / float(n_global * a.shape[-1]), da.data_ptr(), None if db is None else db.data_ptr(),
             loss.data_ptr(), stream())
        ctx.save_for_backward(da, db) if db is not None else ctx.save_for_backward(da)
        ctx.has_b = db is not None
        return loss

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        da = saved[0] * g
        db = saved[1] * g if ctx.has_b else None
        return da, db, None, None


def two_loss_mse(rgb, rgb0, target, n_global=None):
    """img2mse(rgb, target) + img2mse(rgb0, target) (nerf/run.py:689-697) in one kernel; n_global = number of
    rays of the WHOLE step when the batch is sharded over ranks (defaults to the local count)."""
    if n_global is None:
        n_global = rgb.shape[0]
    return _TwoLossMSE.apply(rgb, rgb0, target, n_global)


# ------------------------------------------------------------------------------------------------
# the frame loop of render_path (nerf/run.py:172-219), ray-sharded over the ranks
# ------------------------------------------------------------------------------------------------
def render_frame_sharded(H, W, K, c2w, near, far, chunk, render_rays_fn, group=None, **render_kwargs):
    """One full frame: every rank builds and renders a contiguous block of the H*W rays (ray assembly kernel +
    batchify over `chunk`), then the maps are all-gathered so each rank holds the whole frame.
    Returns (rgb[H,W,3], disp[H,W], acc[H,W]) like render() (nerf/run.py:166-170)."""
    from .ray import make_ray_batch
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n = H * W
    lo, hi = shard_bounds(n, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    pix = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    rays = make_ray_batch(H, W, K, c2w, near, far, pixels=pix, use_viewdirs=True, device=dev)
    outs = {"rgb_map": [], "disp_map": [], "acc_map": []}
    for i in range(0, rays.shape[0], chunk):
        ret = render_rays_fn(rays[i:i + chunk], **render_kwargs)
        for k in outs:
            outs[k].append(ret[k])
    maps = []
    for k in ("rgb_map", "disp_map", "acc_map"):
        local = torch.cat(outs[k], 0) if outs[k] else torch.empty((0,), device=dev)
        if local.dim() == 1:
            local = local[:, None]
        maps.append(gather_rows(local, n, group))
    rgb, disp, acc = maps
    return rgb.reshape(H, W, 3), disp.reshape(H, W), acc.reshape(H, W)
