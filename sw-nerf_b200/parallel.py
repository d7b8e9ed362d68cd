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


def two_loss_mse_grads(rgb, rgb0, target, n_global=None):
    """The same kernel without the autograd wrapper: (loss, d loss / d rgb, d loss / d rgb0).  For callers that run the
    two losses' backward passes separately - torch.autograd.backward([rgb], [da]) then ([rgb0], [db]) - e.g. to
    all-reduce the fine network's gradients while the coarse network's backward is still running (bench.py)."""
    if n_global is None:
        n_global = rgb.shape[0]
    a, t = rgb.detach().contiguous(), target.contiguous()
    b = None if rgb0 is None else rgb0.detach().contiguous()
    da = torch.empty_like(a)
    db = None if b is None else torch.empty_like(b)
    loss = torch.empty((), dtype=torch.float32, device=a.device)
    call("swnerf_mse2", _lib.ptr(a), None if b is None else _lib.ptr(b), _lib.ptr(t), a.numel(),
         1.0 / float(n_global * a.shape[-1]), da.data_ptr(), None if db is None else db.data_ptr(), loss.data_ptr(), stream())
    return loss, da, db


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


class _HostFrames:
    """Double-buffered device->host copies of rendered frames: frame i is copied on a side stream into one of two
    pinned buffers while frame i+1 renders; the host only waits for a buffer when it needs it again (the reference
    synchronises on every frame: rgb.cpu().numpy(), nerf/run.py:199-200)."""

    def __init__(self, shapes, device):
        self.stream = torch.cuda.Stream(device=device)
        self.bufs = [[torch.empty(s, dtype=torch.float32).pin_memory() for s in shapes] for _ in range(2)]
        self.events = [None, None]
        self.pending = [None, None]          # (frame index, callback) waiting in buffer k
        self.i = 0

    def _drain(self, k, sink):
        if self.pending[k] is not None:
            self.events[k].synchronize()
            sink(self.pending[k], [b.numpy().copy() for b in self.bufs[k]])
            self.pending[k] = None

    def push(self, idx, tensors, sink):
        k = self.i & 1
        self._drain(k, sink)
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for b, t in zip(self.bufs[k], tensors):
                b.copy_(t, non_blocking=True)
                t.record_stream(self.stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.events[k], self.pending[k] = ev, idx
        self.i += 1

    def finish(self, sink):
        for k in ((self.i & 1), (self.i & 1) ^ 1):
            self._drain(k, sink)


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0, near=None,
                far=None, group=None, on_frame=None):
    """nerf/run.py:172-219 -> (rgbs [F,H,W,3], disps [F,H,W]) as numpy arrays.

    Same arguments and return value as the reference; near / far come from render_kwargs like there (train() puts them
    in, nerf/run.py:577-579) or from the keywords.  What differs is how: every frame's rays are built by the ray assembly
    kernel and rendered in `chunk`-ray slabs, the rays of a frame are SHARDED over the ranks of `group` and the maps
    all-gathered (every rank returns the full frames; only rank 0 copies them to the host unless on_frame is given),
    and the device->host copies are double-buffered on a side stream instead of a blocking .cpu() per frame.
    `on_frame(i, rgb_np, disp_np)` is called as soon as frame i has landed in host memory (the reference writes its PNG
    there); savedir needs imageio, which the caller's environment must provide."""
    from .render import render_rays
    import numpy as np
    H, W, focal = hwf
    if render_factor != 0:                                   # nerf/run.py:185-189
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    kw = {k: v for k, v in render_kwargs.items() if k not in ("near", "far", "use_viewdirs", "ndc")}
    near = render_kwargs.get("near", near)
    far = render_kwargs.get("far", far)
    if near is None or far is None:
        raise ValueError("render_path: near / far must be in render_kwargs (nerf/run.py:577-579) or passed as keywords")
    if render_kwargs.get("ndc", False):
        raise NotImplementedError("render_path: the sharded frame loop renders non-NDC rays; use render() for LLFF/NDC")
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    writer = None
    if savedir is not None:
        import imageio                                           # noqa: F401  (I/O dependency of the caller, as in the reference)
        import os
        from .ray import to8b
        writer = lambda i, rgb: imageio.imwrite(os.path.join(savedir, "{:03d}.png".format(i)), to8b(rgb))
    to_host = rank == 0 or on_frame is not None
    frames = {}

    def sink(i, arrays):
        frames[i] = arrays
        if on_frame is not None:
            on_frame(i, arrays[0], arrays[1])
        if writer is not None and rank == 0:
            writer(i, arrays[0])
    host = _HostFrames([(H, W, 3), (H, W)], dev) if to_host else None
    n = 0
    with torch.no_grad():
        for i, c2w in enumerate(render_poses):
            c2w = torch.as_tensor(c2w)[:3, :4]
            rgb, disp, acc = render_frame_sharded(H, W, K, c2w, float(near), float(far), chunk, render_rays, group=group, **kw)
            if host is not None:
                host.push(i, [rgb, disp], sink)
            n += 1
    if host is not None:
        host.finish(sink)
        return np.stack([frames[i][0] for i in range(n)], 0), np.stack([frames[i][1] for i in range(n)], 0)
    return None, None
