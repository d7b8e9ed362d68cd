#!/usr/bin/env python
"""Benchmark of the SW-NeRF per-ray rendering hot path on B200 (BASELINE.json metric: train rays/s at 64+128
samples, fwd+bwd, AND 800x800 render ms/frame, at 1-8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision tc|fp32] [--scaling weak|strong]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Step = one vanilla-NeRF lego training step on one batch of synthetic Blender-shaped rays per GPU (configs[1]:
N_rand=4096, 64 coarse + 128 fine samples, coarse+fine 8x256 networks, forward + backward + gradient all-reduce + Adam),
weak scaling (4096 rays per GPU; --scaling strong divides 4096 rays over the ranks).  Prints ONE JSON line, which also
carries, measured in the same run:
  render_ms_per_frame_800x800   config #3: one full frame, rays sharded over the ranks (the strong-scaling half of the metric)
  multires_dp                   config #5: MultiRes D-NeRF pyramid step (4 level networks, one backward, ONE all-reduce)
  dp_check                      N > 1: the all-reduced sharded gradient equals the single-process gradient of the
                                concatenated batch, parameters stay bit-identical across ranks after Adam steps
  gpu_eager_baseline            the reference's algorithm in eager fp32 PyTorch ON THE SAME B200 (its real deployment)
  cpu_baseline                  the same on the host cores (the reference's CPU path), full 4096-ray steps
  roofline                      the fused forward kernel against the measured tensor peak (burst and sustained)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAND = 4096
N_SAMPLES, N_IMPORTANCE = 64, 128
MAC_PER_EVAL = 593408                      # BASELINE.md section 4 (model.py:22-35): algorithmic
MAC_EXEC_PER_EVAL = 30 * 256 * 64 + 5 * 144 * 64 + 384     # what the fused kernel issues: 35 K-chunks on the tensor pipe
                                                           # (feature_linear folded into the view layer) + rgb_linear in fp32
FLOP_PER_EVAL_FWD = 2 * MAC_PER_EVAL


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "src": "MEASURED_PEAKS.json"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "src": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


def measured_traffic():
    """DRAM bytes per launch of the dominant kernels from the committed ncu capture of THIS tree's kernels
    (profiles/r2_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` run of this command)."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# baselines: the reference's algorithm (oracle port, pinned to the unmodified reference by tests/golden) in eager torch
# ----------------------------------------------------------------------------------------------------------------
def reference_step_fn(device, rays_per_step):
    """The reference training step (render_rays fwd + two-loss backward + torch.optim.Adam, nerf/run.py:684-699) on
    `device` through the oracle port; returns a callable that runs one step and the rays per step."""
    from oracle import nerf_oracle as O
    shapes = O.mlp_param_shapes()
    pc = {k: v.to(device).requires_grad_() for k, v in O.make_params(shapes, 21).items()}
    pf = {k: v.to(device).requires_grad_() for k, v in O.make_params(shapes, 55).items()}
    opt = torch.optim.Adam(list(pc.values()) + list(pf.values()), lr=5e-4, betas=(0.9, 0.999))
    rays = torch.from_numpy(O.blender_rays(rays_per_step, 5)).to(device)
    target = torch.from_numpy(np.random.RandomState(6).uniform(0, 1, (rays_per_step, 3)).astype(np.float32)).to(device)

    def step():
        opt.zero_grad()
        ret = O.render_rays(rays, pc, pf, N_SAMPLES, N_IMPORTANCE, perturb=1.0, white_bkgd=True)
        loss = ((ret["rgb_map"] - target) ** 2).mean() + ((ret["rgb0"] - target) ** 2).mean()
        loss.backward()
        opt.step()
        return loss
    return step


def cpu_reference_run(steps, warmup, rays_per_step, threads, budget_s=240.0, anomaly=False):
    """Host-core run.  Starts at `rays_per_step` (4096 = the full step); if the first step shows that steps+warmup of
    them would not finish inside `budget_s`, the per-step sample is halved until they do (and the line says so).
    anomaly=True times the step the way the reference ships it: utils.py:2 turns autograd anomaly detection ON for
    every runner (a stack capture per op)."""
    torch.set_num_threads(threads)
    if anomaly:
        with torch.autograd.set_detect_anomaly(True):
            return cpu_reference_run(steps, warmup, rays_per_step, threads, budget_s, anomaly=False)
    n = rays_per_step
    while True:
        step = reference_step_fn(torch.device("cpu"), n)
        t0 = time.perf_counter()
        step()
        t1 = time.perf_counter() - t0
        if t1 * (steps + warmup) <= budget_s or n <= 128:
            break
        n //= 2
    for _ in range(max(warmup - 1, 0)):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return n * len(times) / sum(times), sum(times) / len(times), n


def gpu_eager_run(dev, steps=5, warmup=3):
    """The same oracle port, unchanged, on the B200 in eager fp32 (TF32 off: torch's default, and the reference's):
    the honest same-box comparator - no kernel of this repo on its path."""
    step = reference_step_fn(dev, N_RAND)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": N_RAND / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "steps": steps,
            "what": "reference algorithm (oracle port of nerf/run.py render_rays + loss.backward + torch.optim.Adam), eager "
                    "PyTorch fp32 on this GPU, same 4096-ray step, inputs resident", "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}


# ----------------------------------------------------------------------------------------------------------------
# config #5: MultiRes D-NeRF pyramid step, data-parallel
# ----------------------------------------------------------------------------------------------------------------
MULTIRES_LEVELS = [((20, 8, 20), 1024), ((10, 4, 10), 256), ((10, 4, 10), 64), ((-1, -1, -1), 16)]   # multires_dnerf.py:665


def multires_bench(S, dev, rank, world, precision, steps=10, warmup=3, use_graph=True):
    """Four level networks (PE (20,8,20) / (10,4,10) / (10,4,10) / identity; 1024 / 256 / 64 / 16 rays per step), every
    level's rays sharded over the ranks, ONE loss.backward over all levels (multires_dnerf.py:1005), ONE all-reduce of
    the flat gradient buffer that holds all four models, Adam.  Strong scaling by construction (1360 rays per step)."""
    import tempfile
    from argparse import Namespace
    import torch.distributed as dist
    from swnerf_b200 import dnerf, parallel, synth
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "e"), exist_ok=True)
    args = Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64,
                     netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=1 << 30, lrate=5e-4,
                     ft_path=None, basedir=tmp, expname="e", no_reload=True, perturb=1.0, white_bkgd=True,
                     raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False, nerf_type="direct_temporal",
                     use_two_models_for_fine=False, not_zero_canonical=False, do_half_precision=False,
                     swnerf_precision=precision)
    levels, params = [], []
    stdout_fd = os.dup(1)
    for li, (ch, n) in enumerate(MULTIRES_LEVELS):
        kw, _, _, gv, _ = dnerf.create_nerf_multires(args, ch, li, device=dev)
        model = kw["network_fn"]
        model.load_state_dict(synth.scene_params(model, 700 + li)); model.to(dev)
        kw = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}
        lo, hi = parallel.shard_bounds(n, rank, world)
        rays = torch.from_numpy(synth.blender_rays(n, 710 + li, frame_time=0.37)[lo:hi]).to(dev)
        rays._swnerf_frame_time = 0.37
        tgt = torch.from_numpy(np.random.RandomState(720 + li).uniform(0, 1, (n, 3)).astype(np.float32)[lo:hi]).to(dev)
        levels.append((kw, rays, tgt, n))
        params += gv
    os.close(stdout_fd)
    flat = parallel.FlatParams(params)
    opt = parallel.FlatAdam(flat, lr=5e-4)
    fused = [bool(kw["network_query_fn"].uses_tc(kw["network_fn"], True)) for kw, _, _, _ in levels]

    def step():
        flat.zero_()
        loss = None
        for kw, rays, tgt, n in levels:
            if rays.shape[0] == 0:
                continue
            ret = dnerf.render_rays(rays, **kw)
            l = parallel.sharded_mse(ret["rgb_map"], tgt, n)          # F.mse_loss over the level's full patch
            loss = l if loss is None else loss + l
        if loss is not None:
            loss.backward()                                          # one backward over all levels
        flat.all_reduce()
        opt.step()
        return loss
    def fwd_bwd():
        flat.zero_()
        loss = None
        for kw, rays, tgt, n in levels:
            if rays.shape[0] == 0:
                continue
            ret = dnerf.render_rays(rays, **kw)
            l = parallel.sharded_mse(ret["rgb_map"], tgt, n)
            loss = l if loss is None else loss + l
        if loss is not None:
            loss.backward()
        return loss

    def timed_steps(fn):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    ms_eager = timed_steps(step)
    # the step is launch-bound (1360 rays: ~0.7 ms of tensor work behind ~200 launches): replay forward + backward of all
    # four levels from ONE CUDA graph (static ray / target buffers, the per-step weight re-pack captured too), then the
    # all-reduce and Adam - what bench.py does for the vanilla step
    ms_graph, launches_graph = None, None
    if use_graph:
        try:
            from swnerf_b200 import tc as _tc, _lib as _l
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fwd_bwd()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            _tc.GENERATION += 1
            _l.launch_count(reset=True)
            with torch.cuda.graph(g):
                fwd_bwd()
            launches_graph = _l.launch_count()

            def step_graph():
                g.replay()
                flat.all_reduce()
                opt.step()
            ms_graph = timed_steps(step_graph)
        except Exception as e:                      # noqa: BLE001
            sys.stderr.write("multires: cuda graph capture failed, eager number only: %r\n" % (e,))
            torch.cuda.synchronize()
    ms_best = ms_graph if ms_graph is not None else ms_eager
    n_rays = sum(n for _, n in MULTIRES_LEVELS)
    return {"ms_per_step": ms_best, "ms_per_step_eager": ms_eager, "cuda_graph": ms_graph is not None,
            "our_launches_per_step": launches_graph,
            "rays_per_step": n_rays, "value": n_rays / (ms_best * 1e-3),
            "unit": "rays/s", "scaling": "strong", "levels": [{"pe": list(ch), "rays": n, "fused_tcgen05": f}
                                                              for (ch, n), f in zip(MULTIRES_LEVELS, fused)],
            "params_in_one_allreduce": int(flat.flat.numel()), "steps": steps,
            "what": "MultiRes D-NeRF pyramid step (multires_dnerf.py:905-1008): 4 level networks, coarse no-grad + fine "
                    "pass each, one backward, one all-reduce, flat Adam"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="swnerf_b200")
    ap.add_argument("--precision", default=None, help="tc (fused tcgen05, default when built) or fp32 (check mode)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 4096 rays per GPU per step (default); strong: 4096 rays per step over all GPUs")
    ap.add_argument("--cpu-sample-rays", type=int, default=N_RAND)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-render-frame", dest="render_frame", action="store_false")
    ap.add_argument("--no-multires", dest="multires", action="store_false")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="do not capture fwd+bwd in a CUDA graph")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="N > 1: one all-reduce after the whole backward instead of the fine net's slice under the coarse backward")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    # stdout carries exactly ONE line, the JSON record: everything libraries print while the run is in progress
    # (NCCL's version banner, warnings written with print) is routed to stderr at the file-descriptor level
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    rays_per_gpu = N_RAND if args.scaling == "weak" else N_RAND // world
    config = {"workload": "vanilla NeRF lego training step: N_rand=4096 rays/GPU, 64 coarse + 128 fine samples, "
                          "coarse+fine 8x256 MLP (PE L=10/4), fwd+bwd+Adam, synthetic 800x800 Blender-shaped rays",
              "rays_per_gpu": rays_per_gpu, "N_samples": N_SAMPLES, "N_importance": N_IMPORTANCE,
              "parallelism": "ray-sharded dp%d" % world,
              "l2": "every step streams >126 MB of activations (working set larger than L2)"}

    if args.impl == "reference":
        if rank != 0:
            return
        rps, spp, n_used = cpu_reference_run(args.steps, args.warmup, args.cpu_sample_rays, threads)
        sample = ("%d rays per step%s, fwd+bwd+Adam, oracle port of the reference (eager torch CPU fp32, anomaly mode off), "
                  "%d threads" % (n_used, "" if n_used == N_RAND else " (bounded sample of the 4096-ray step: the full step "
                                  "would not fit the time budget)", threads))
        line = {"impl": "reference", "metric": "train_rays_per_s", "value": rps, "unit": "rays/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spp * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch.distributed as dist
    import swnerf_b200 as S
    from swnerf_b200 import _lib, tc, parallel, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.call("swnerf_device_ok")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision or ("tc" if tc.available() else "fp32")

    def build(prec):
        mc = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mc.load_state_dict(synth.scene_params(mc, 21)); mc.to(dev)
        mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(synth.scene_params(mf, 55)); mf.to(dev)
        q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision=prec)
        # loss + optimizer of the step: parameters and gradients live in two flat buffers, the
        # backward kernels accumulate straight into the gradient buffer, ONE Adam kernel
        flat = parallel.FlatParams(list(mc.parameters()) + list(mf.parameters()))
        return mc, mf, q, flat

    # ---- N > 1: numerical check of the data-parallel step on the real kernels, before anything is timed ----------
    dp_check = None
    if world > 1:
        dp_check = {}
        for prec, n_local in (("fp32", 256), (precision, rays_per_gpu)):
            mc_, mf_, q_, flat_ = build(prec)
            kw_ = dict(network_fn=mc_, network_query_fn=q_, N_samples=N_SAMPLES, perturb=0.0, N_importance=N_IMPORTANCE,
                       network_fine=mf_, white_bkgd=True, raw_noise_std=0.0)
            r_loc = torch.from_numpy(synth.blender_rays(n_local, 900 + rank)).to(dev)
            t_loc = torch.from_numpy(np.random.RandomState(950 + rank).uniform(0, 1, (n_local, 3)).astype(np.float32)).to(dev)
            flat_.zero_()
            ret = S.render_rays(r_loc, **kw_)
            parallel.two_loss_mse(ret["rgb_map"], ret["rgb0"], t_loc, n_local * world).backward()
            flat_.all_reduce()
            sharded = flat_.flat.clone()
            r_all = [torch.empty_like(r_loc) for _ in range(world)]; t_all = [torch.empty_like(t_loc) for _ in range(world)]
            dist.all_gather(r_all, r_loc); dist.all_gather(t_all, t_loc)
            flat_.zero_()
            ret = S.render_rays(torch.cat(r_all, 0), **kw_)
            parallel.two_loss_mse(ret["rgb_map"], ret["rgb0"], torch.cat(t_all, 0), n_local * world).backward()
            single = flat_.flat
            rel = float((sharded - single).norm() / single.norm())
            dp_check["grad_rel_l2_%s_%dx%d_rays" % (prec, world, n_local)] = rel
            # fp32 mode is the exactness proof (summation order only).  In tc mode the backward scales d_raw by a power of two
            # chosen per CALL from max|d_raw| (fp16 range), so a shard and the whole batch round their fp16 operands
            # differently: measured 1e-4 (2 ranks) .. 6e-4 (8 ranks), an order below the operand-format floor (3e-3)
            assert rel < (1e-5 if prec == "fp32" else 3e-3), "sharded gradient differs from the single-process one: %g" % rel
            del mc_, mf_, q_, flat_
        torch.cuda.empty_cache()

    torch.manual_seed(1234 + rank)
    mc, mf, q, flat = build(precision)
    opt = parallel.FlatAdam(flat, lr=5e-4, betas=(0.9, 0.999))
    n_global = rays_per_gpu * world

    nbatch = 4
    host_rays = [torch.from_numpy(synth.blender_rays(rays_per_gpu, 100 + rank * 10 + i)).pin_memory() for i in range(nbatch)]
    host_tgt = [torch.from_numpy(np.random.RandomState(200 + rank * 10 + i).uniform(0, 1, (rays_per_gpu, 3))
                                 .astype(np.float32)).pin_memory() for i in range(nbatch)]
    dev_rays = [r.to(dev) for r in host_rays]
    dev_tgt = [t.to(dev) for t in host_tgt]
    kw = dict(network_fn=mc, network_query_fn=q, N_samples=N_SAMPLES, perturb=1.0, N_importance=N_IMPORTANCE,
              network_fine=mf, white_bkgd=True, raw_noise_std=0.0)

    # The step in two phases, so that at N > 1 the fine network's gradient slice can be all-reduced while the coarse
    # network's backward runs: the two losses' graphs are independent (z_fine is detached, nerf/run.py:398).
    #   phase A: forward (coarse + fine), both losses, backward of the fine loss
    #   phase B: backward of the coarse loss
    st = {}

    def phase_a(rays, tgt):
        flat.zero_()
        ret = S.render_rays(rays, **kw)
        loss, da, db = parallel.two_loss_mse_grads(ret["rgb_map"], ret["rgb0"], tgt, n_global)     # nerf/run.py:689-697
        torch.autograd.backward([ret["rgb_map"]], [da])
        st["rgb0"], st["db"] = ret["rgb0"], db
        return loss

    def phase_b():
        torch.autograd.backward([st["rgb0"]], [st["db"]])
        st["rgb0"] = st["db"] = None

    # fwd+bwd is launch-bound between the big kernels (~60 small launches): capture each phase once in a CUDA graph on
    # static input buffers and replay; the all-reduces and Adam follow the replays.
    graph = {"a": None, "b": None, "rays": None, "tgt": None, "loss": None, "launches": 0}
    if args.graph:
        try:
            graph["rays"], graph["tgt"] = dev_rays[0].clone(), dev_tgt[0].clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    phase_a(graph["rays"], graph["tgt"]); phase_b()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            tc.GENERATION += 1          # stale-mark the packed fp16 weight images so the re-pack kernels are captured too
            _lib.launch_count(reset=True)
            pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(ga, pool=pool):
                graph["loss"] = phase_a(graph["rays"], graph["tgt"])
            with torch.cuda.graph(gb, pool=pool):
                phase_b()
            graph["launches"] = _lib.launch_count()      # our kernels inside one replay of both graphs
            graph["a"], graph["b"] = ga, gb
        except Exception as e:                      # noqa: BLE001
            sys.stderr.write("cuda graph capture failed, running eagerly: %r\n" % (e,))
            graph["a"] = graph["b"] = None
            torch.cuda.synchronize()

    lo_f, hi_f = flat.span(list(mf.parameters()))
    lo_c, hi_c = flat.span(list(mc.parameters()))
    comm_stream = torch.cuda.Stream() if world > 1 else None
    overlap = world > 1 and args.overlap

    def step(rays, tgt):
        if graph["a"] is not None:
            graph["rays"].copy_(rays, non_blocking=True)
            graph["tgt"].copy_(tgt, non_blocking=True)
            graph["a"].replay()
            loss = graph["loss"]
        else:
            loss = phase_a(rays, tgt)
        if overlap:
            # the fine slice is final: reduce it on the side stream while phase B runs
            comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm_stream):
                dist.all_reduce(flat.flat[lo_f:hi_f], op=dist.ReduceOp.SUM)
        if graph["b"] is not None:
            graph["b"].replay()
        else:
            phase_b()
        if overlap:
            dist.all_reduce(flat.flat[lo_c:hi_c], op=dist.ReduceOp.SUM)
            torch.cuda.current_stream().wait_stream(comm_stream)
        else:
            flat.all_reduce()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident(i):
        step(dev_rays[i % nbatch], dev_tgt[i % nbatch])

    last = {}
    # the step's loss is read back EVERY step, but the host does not stall on it: the 4-byte copy into pinned memory is
    # queued behind the step and the host reads the PREVIOUS step's value once its event has fired (a training loop that
    # logs the loss one step late), so the next step's launches overlap this step's execution
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e(i):
        r = host_rays[i % nbatch].to(dev, non_blocking=True)
        t = host_tgt[i % nbatch].to(dev, non_blocking=True)
        loss = step(r, t)
        loss_host[i & 1].copy_(loss.reshape(()), non_blocking=True)
        loss_ev[i & 1].record()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            last["loss"] = float(loss_host[(i - 1) & 1])

    for i in range(args.warmup):
        resident(i)
    assert flat.check_views(), "param.grad views were replaced; the flat all-reduce buffer is stale"
    if world > 1:       # parameters bit-identical across ranks after the warm-up Adam steps (>= 3)
        pmax, pmin = flat.flat_p.clone(), flat.flat_p.clone()
        dist.all_reduce(pmax, op=dist.ReduceOp.MAX); dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
        dp_check["params_bit_identical_after_%d_adam_steps" % args.warmup] = bool(torch.equal(pmax, pmin))
        assert dp_check["params_bit_identical_after_%d_adam_steps" % args.warmup], "ranks diverged"
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.launch_count(reset=True)
    ms = timed(resident, args.steps)
    launches = _lib.launch_count()
    for i in range(2):
        e2e(i)
    ms_e2e = timed(e2e, args.steps)           # (timed() ends with a device synchronize: the last read-back is inside)
    last["loss"] = float(loss_host[(args.steps - 1) & 1])
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions

    # per-kernel device times (CUDA events on the launching stream) over a few instrumented steps
    use_graph = graph["a"] is not None
    graph["a"] = graph["b"] = None                  # the instrumented steps below run eagerly
    _lib.TIMING = {}
    bwd_split = {"data": 0.0, "weight": 0.0}
    if precision == "tc":
        import ctypes
        _lib.call("swnerf_tc_set_profiling", 1)
        orig_call = _lib.call

        def call_and_split(name, *a):
            orig_call(name, *a)
            if name == "swnerf_tc_mlp_bwd":
                d, w = ctypes.c_float(), ctypes.c_float()
                orig_call("swnerf_tc_last_bwd_ms", ctypes.byref(d), ctypes.byref(w))
                bwd_split["data"] += d.value / 3.0
                bwd_split["weight"] += w.value / 3.0
        tc.call = call_and_split
    for i in range(3):
        resident(i)
    torch.cuda.synchronize()
    if precision == "tc":
        tc.call = orig_call
        _lib.call("swnerf_tc_set_profiling", 0)
    ktimes = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) * (len(v) / 3.0) for k, v in _lib.TIMING.items()}
    kcalls = {k: len(v) / 3.0 for k, v in _lib.TIMING.items()}
    _lib.TIMING = None

    pk = peaks()
    burst, sust = pk["bf16_tflops"], pk["bf16_tflops_sustained"] or pk["bf16_tflops"]
    evals = rays_per_gpu * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE)
    traffic = measured_traffic()
    if precision == "tc":
        t_fwd = ktimes.get("swnerf_tc_mlp_fwd", 0.0)
        t_bwd = ktimes.get("swnerf_tc_mlp_bwd", 0.0)
        # dominant kernel = fused forward (2 launches/step: coarse 64 + fine 192 samples per ray).  Each call is bracketed
        # by its own CUDA events inside a 3-step (~15 ms) region: a kernel timed alone -> the BURST peak is the roof.
        ach = FLOP_PER_EVAL_FWD * evals / (t_fwd * 1e-3) / 1e12 if t_fwd > 0 else 0.0
        tr_fwd = (traffic or {}).get("fwd_train_bytes_per_step") if rays_per_gpu == N_RAND else None
        roof = {"bound": "tensor", "kernel": "mlp_fwd4_kernel<1> via swnerf_tc_mlp_fwd (fused points + PE + 8x256 MLP, tcgen05 "
                                             "cta_group::2, training variant: saves activations)",
                "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst,
                "peak_sustained": sust, "frac_sustained": ach / sust,
                "achieved_executed": ach * MAC_EXEC_PER_EVAL / MAC_PER_EVAL,
                "frac_executed": ach * MAC_EXEC_PER_EVAL / MAC_PER_EVAL / burst,
                "algorithmic_flop_per_launch_pair": FLOP_PER_EVAL_FWD * evals,
                "executed_over_algorithmic_macs": MAC_EXEC_PER_EVAL / MAC_PER_EVAL,
                "traffic": tr_fwd,
                "traffic_src": (traffic or {}).get("src") if tr_fwd else "no ncu capture of this tree for this shape: null",
                "peak_src": pk["src"] + ": cuBLAS bf16 8192^3, best of 10 (burst) / back to back for 4 s (sustained)",
                "ms_per_step": t_fwd,
                "bwd": {"ms_per_step": t_bwd, "data_kernel_ms": bwd_split["data"], "weight_kernel_ms": bwd_split["weight"],
                        "achieved": 2 * FLOP_PER_EVAL_FWD * evals / (t_bwd * 1e-3) / 1e12 if t_bwd > 0 else 0.0,
                        "frac": (2 * FLOP_PER_EVAL_FWD * evals / (t_bwd * 1e-3) / 1e12 / burst) if t_bwd > 0 else 0.0,
                        "traffic": (traffic or {}).get("bwd_bytes_per_step") if rays_per_gpu == N_RAND else None,
                        "weight_kernel_hbm": {"bound": "hbm", "algorithmic_bytes_per_sample": 9728,
                                              "achieved_gbs": 9728.0 * evals / (bwd_split["weight"] * 1e-3) / 1e9
                                              if bwd_split["weight"] > 0 else 0.0, "peak_gbs": pk["hbm_gbs"]}}}
    else:
        t_mm = ktimes.get("swnerf_sgemm", 0.0)
        ach = 3 * FLOP_PER_EVAL_FWD * evals / (t_mm * 1e-3) / 1e12 if t_mm > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "swnerf_sgemm (fp32 SIMT check path; not the tensor-core kernel)",
                "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst, "traffic": None,
                "peak_src": pk["src"], "ms_per_step": t_mm}

    extra = {}
    if args.render_frame:
        # config #3: full 800x800 frame, test-time kwargs (perturb=0, no noise), chunk 32768, rays sharded over ranks:
        # frames rendered through parallel.render_path (device frame loop, maps gathered, double-buffered async D2H)
        H = W = 800
        focal = 0.5 * W / np.tan(0.5 * 0.6911112)
        Kmat = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
        poses = [torch.from_numpy(synth.pose_spherical(40.0 + 9.0 * i, -30.0, 4.0)[:3, :4]) for i in range(3)]
        kwt = dict(kw); kwt["perturb"] = 0.0
        parallel.render_path(poses[:1], (H, W, focal), Kmat, 1024 * 32, kwt, near=2.0, far=6.0)      # warm-up frame
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        rgbs, disps = parallel.render_path(poses[1:], (H, W, focal), Kmat, 1024 * 32, kwt, near=2.0, far=6.0)
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) / 2
        fm = torch.tensor([e0.elapsed_time(e1) / 2], device=dev)
        if world > 1:
            dist.all_reduce(fm, op=dist.ReduceOp.MAX)
        extra["render_ms_per_frame_800x800"] = float(fm.item())
        extra["render_frame"] = {"ms_per_frame": float(fm.item()), "frames_timed": 2, "H": H, "W": W, "chunk": 32768,
                                 "scaling": "strong (640,000 rays per frame over %d GPUs)" % world,
                                 "wall_ms_per_frame_incl_d2h": wall * 1e3, "d2h_bytes_per_frame": H * W * 4 * 4,
                                 "tflops_algorithmic": FLOP_PER_EVAL_FWD * H * W * 256 / (float(fm.item()) * 1e-3) / 1e12,
                                 "rgb_mean": float(np.mean(rgbs[-1])) if rank == 0 else None,
                                 "what": "render_path over 2 poses after 1 warm-up frame: ray assembly kernel, coarse + "
                                         "resample + fine per 32768-ray chunk (perturb=0), maps all-gathered, frames "
                                         "copied to pinned host memory on a side stream"}
    if args.multires:
        try:
            extra["multires_dp"] = multires_bench(S, dev, rank, world, precision, use_graph=args.graph)
        except Exception as e:                      # noqa: BLE001
            extra["multires_dp"] = {"error": repr(e)}

    if rank == 0:
        cpu = eager = None
        if world == 1 and not args.no_eager_baseline:
            try:
                torch.cuda.reset_peak_memory_stats()
                eager = gpu_eager_run(dev)
            except Exception as e:                  # noqa: BLE001
                eager = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            rps, spp, n_used = cpu_reference_run(3, 1, args.cpu_sample_rays, threads, budget_s=120.0)
            cpu = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                   "sample": "%d rays/step x 3 steps (after 1 warm-up) of the same training step (fwd+bwd+Adam), oracle "
                             "port of the reference, eager torch CPU fp32, anomaly mode off" % n_used}
            try:        # the reference AS SHIPPED runs with autograd anomaly detection on (utils.py:2): context, 2 steps
                rps_a, _, n_a = cpu_reference_run(2, 1, args.cpu_sample_rays, threads, budget_s=60.0, anomaly=True)
                cpu["as_shipped_anomaly_mode_on"] = {"value": rps_a, "unit": "rays/s", "rays_per_step": n_a, "steps": 2}
            except Exception as e:                  # noqa: BLE001
                cpu["as_shipped_anomaly_mode_on"] = {"error": repr(e)}
        rays_total = rays_per_gpu * world * args.steps
        line = {"metric": "train_rays_per_s", "value": rays_total / (ms * 1e-3), "unit": "rays/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate (tcgen05)" if precision == "tc" else "f32",
                "data": "synthetic", "config": config, "precision_mode": precision,
                "e2e": {"value": rays_total / (ms_e2e * 1e-3), "unit": "rays/s",
                        "h2d_bytes_per_step": rays_per_gpu * (11 + 3) * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps, "last_loss": last.get("loss"),
                        "loss_readback": "every step (4 B to pinned host memory), consumed one step late so the host never stalls on it"},
                "gpu_launches": launches + (graph["launches"] * args.steps if use_graph else 0),
                "cuda_graph": use_graph, "allreduce_overlap": overlap, "roofline": roof, "cpu_baseline": cpu,
                "gpu_eager_baseline": eager, "dp_check": dp_check, "clocks": clocks,
                "kernel_ms_per_step": ktimes, "kernel_calls_per_step": kcalls}
        if eager and "value" in eager:
            line["speedup_vs_gpu_eager"] = {"device_timed": line["value"] / eager["value"], "e2e": line["e2e"]["value"] / eager["value"]}
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
