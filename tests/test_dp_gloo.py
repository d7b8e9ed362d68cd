"""CPU, world_size 2, gloo: host-side logic of the ray-sharded data-parallel path (parallel.py).

The kernels cannot run here, so the per-rank "render + loss" is a small differentiable stand-in; what
is checked is the contract bench.py and a DP trainer rely on: contiguous ray shards, the local loss
normalised by the GLOBAL element count, .grad views into one flat buffer, one all-reduce(SUM) giving the
gradients of the reference's mean loss over the whole batch (nerf/run.py:689-697), and the row gather of
a sharded frame render."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import swnerf_b200  # noqa: F401
from swnerf_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(11, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        rays = torch.randn(n_total, 11, generator=g)
        target = torch.rand(n_total, 3, generator=g)
        model = _model()
        flat = parallel.FlatGrads(list(model.parameters()))
        lo, hi = parallel.shard_bounds(n_total, rank, world)
        flat.zero_()
        pred = model(rays[lo:hi])
        loss = parallel.sharded_mse(pred, target[lo:hi], n_total)
        loss.backward()
        assert flat.check_views()
        flat.all_reduce()
        rows = parallel.gather_rows(pred.detach(), n_total)
        if rank == 0:
            torch.save({"flat": flat.flat.clone(), "rows": rows}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37])
def test_dp_two_ranks_equal_single_process(tmp_path, n_total):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_total, out), nprocs=2, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(1)
    rays = torch.randn(n_total, 11, generator=g)
    target = torch.rand(n_total, 3, generator=g)
    model = _model()
    pred = model(rays)
    torch.mean((pred - target) ** 2).backward()          # img2mse over the full batch
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(got["flat"], ref, rtol=1e-5, atol=1e-7)
    assert torch.allclose(got["rows"], pred.detach(), rtol=1e-6, atol=1e-7)


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 4096, 640000):
        for w in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


# ---- config #5 (MultiRes pyramid): several level networks of different shapes, every level's rays sharded over the
#      ranks - the 16-ray level leaves a rank with 0 rays at world_size 3+ and an uneven split at 2 - ONE backward over all
#      levels (multires_dnerf.py:1005), ONE all-reduce of a flat buffer that holds all models
_LEVELS = [(11, 24, 40), (11, 16, 10), (11, 16, 4), (11, 8, 1)]          # (in, hidden, rays of the level)


def _level_models():
    torch.manual_seed(3)
    return [torch.nn.Sequential(torch.nn.Linear(i, h), torch.nn.ReLU(), torch.nn.Linear(h, 3)) for i, h, _ in _LEVELS]


def _level_data():
    g = torch.Generator().manual_seed(4)
    return [(torch.randn(n, i, generator=g), torch.rand(n, 3, generator=g)) for i, _, n in _LEVELS]


def _pyramid_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        models, data = _level_models(), _level_data()
        flat = parallel.FlatGrads([p for m in models for p in m.parameters()])
        flat.zero_()
        loss = None
        for m, (rays, tgt) in zip(models, data):
            n = rays.shape[0]
            lo, hi = parallel.shard_bounds(n, rank, world)
            if hi == lo:
                continue                                   # this rank holds no ray of the level
            l = parallel.sharded_mse(m(rays[lo:hi]), tgt[lo:hi], n)
            loss = l if loss is None else loss + l
        if loss is not None:
            loss.backward()
        assert flat.check_views()
        flat.all_reduce()
        if rank == 0:
            torch.save(flat.flat.clone(), out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dp_pyramid_levels_one_allreduce(tmp_path, world):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_pyramid_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    models, data = _level_models(), _level_data()
    loss = sum(torch.mean((m(r) - t) ** 2) for m, (r, t) in zip(models, data))     # F.mse_loss per level, summed
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for m in models for p in m.parameters()])
    assert got.numel() == ref.numel()
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-7)
