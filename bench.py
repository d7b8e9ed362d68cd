#!/usr/bin/env python
"""Benchmark of the SW-NeRF per-ray rendering hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision tc|fp32]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Step = one vanilla-NeRF lego training step on one batch of synthetic Blender-shaped rays per GPU
(configs[1]: N_rand=4096, 64 coarse + 128 fine samples, coarse+fine 8x256 networks, forward +
backward + gradient all-reduce + Adam), weak scaling (4096 rays per GPU).  Prints ONE JSON line.
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
MAC_PER_EVAL = 593408                      # BASELINE.md section 4 (model.py:22-35)
FLOP_PER_EVAL_FWD = 2 * MAC_PER_EVAL


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "src": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "src": "fallback"}


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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


def cpu_reference_run(steps, warmup, sample_rays, threads):
    """The reference's algorithm on the host cores: the oracle port (oracle/nerf_oracle.py, pinned to
    the unmodified reference by tests/golden) - render_rays forward + two-loss backward + Adam."""
    from oracle import nerf_oracle as O
    torch.set_num_threads(threads)
    shapes = O.mlp_param_shapes()
    pc = {k: v.requires_grad_() for k, v in O.make_params(shapes, 21).items()}
    pf = {k: v.requires_grad_() for k, v in O.make_params(shapes, 55).items()}
    opt = torch.optim.Adam(list(pc.values()) + list(pf.values()), lr=5e-4, betas=(0.9, 0.999))
    rays = torch.from_numpy(O.blender_rays(sample_rays, 5))
    target = torch.from_numpy(np.random.RandomState(6).uniform(0, 1, (sample_rays, 3)).astype(np.float32))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        ret = O.render_rays(rays, pc, pf, N_SAMPLES, N_IMPORTANCE, perturb=1.0, white_bkgd=True)
        loss = ((ret["rgb_map"] - target) ** 2).mean() + ((ret["rgb0"] - target) ** 2).mean()
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sample_rays * len(times) / sum(times), sum(times) / len(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="swnerf_b200")
    ap.add_argument("--precision", default=None, help="tc (fused tcgen05, default when built) or fp32 (check mode)")
    ap.add_argument("--cpu-sample-rays", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="do not capture fwd+bwd in a CUDA graph")
    ap.add_argument("--render-frame", action="store_true", help="also time one 800x800 frame render (config #3)")
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
    config = {"workload": "vanilla NeRF lego training step: N_rand=4096 rays/GPU, 64 coarse + 128 fine samples, "
                          "coarse+fine 8x256 MLP (PE L=10/4), fwd+bwd+Adam, synthetic 800x800 Blender-shaped rays",
              "rays_per_gpu": N_RAND, "N_samples": N_SAMPLES, "N_importance": N_IMPORTANCE,
              "parallelism": "ray-sharded dp%d" % world,
              "l2": "every step streams >126 MB of activations (working set larger than L2)"}

    if args.impl == "reference":
        if rank != 0:
            return
        rps, spp = cpu_reference_run(args.steps, args.warmup, args.cpu_sample_rays, threads)
        line = {"impl": "reference", "metric": "train_rays_per_s", "value": rps, "unit": "rays/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spp * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                                 "sample": "%d rays per step (bounded sample of the 4096-ray step), fwd+bwd+Adam, "
                                           "torch CPU fp32, %d threads" % (args.cpu_sample_rays, threads)},
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

    torch.manual_seed(1234 + rank)
    mc = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mc.load_state_dict(synth.scene_params(mc, 21)); mc.to(dev)
    mf = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True); mf.load_state_dict(synth.scene_params(mf, 55)); mf.to(dev)
    q = S.NetworkQuery(S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 3, 0)[0], precision=precision)
    params = list(mc.parameters()) + list(mf.parameters())
    # loss + optimizer of the step: parameters and gradients live in two flat buffers, the
    # backward kernels accumulate straight into the gradient buffer, ONE all-reduce, ONE Adam kernel
    # (same update as the reference's torch.optim.Adam(lr, betas=(0.9, 0.999)), tests/test_gpu_next_rows.py)
    flat = parallel.FlatParams(params)
    opt = parallel.FlatAdam(flat, lr=5e-4, betas=(0.9, 0.999))
    n_global = N_RAND * world

    nbatch = 4
    host_rays = [torch.from_numpy(synth.blender_rays(N_RAND, 100 + rank * 10 + i)).pin_memory() for i in range(nbatch)]
    host_tgt = [torch.from_numpy(np.random.RandomState(200 + rank * 10 + i).uniform(0, 1, (N_RAND, 3))
                                 .astype(np.float32)).pin_memory() for i in range(nbatch)]
    dev_rays = [r.to(dev) for r in host_rays]
    dev_tgt = [t.to(dev) for t in host_tgt]
    kw = dict(network_fn=mc, network_query_fn=q, N_samples=N_SAMPLES, perturb=1.0, N_importance=N_IMPORTANCE,
              network_fine=mf, white_bkgd=True, raw_noise_std=0.0)

    def fwd_bwd(rays, tgt):
        flat.zero_()
        ret = S.render_rays(rays, **kw)
        loss = parallel.two_loss_mse(ret["rgb_map"], ret["rgb0"], tgt, n_global)     # nerf/run.py:689-697
        loss.backward()
        return loss

    # The forward+backward of one step is launch-bound between the big kernels (~60 small launches): capture it
    # once in a CUDA graph on static input buffers and replay it; the all-reduce and Adam follow the replay.
    graph = {"g": None, "rays": None, "tgt": None, "loss": None}
    if args.graph:
        try:
            graph["rays"], graph["tgt"] = dev_rays[0].clone(), dev_tgt[0].clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fwd_bwd(graph["rays"], graph["tgt"])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            tc.GENERATION += 1          # stale-mark the packed fp16 weight images so the re-pack kernels are captured too
            _lib.launch_count(reset=True)
            with torch.cuda.graph(g_):
                graph["loss"] = fwd_bwd(graph["rays"], graph["tgt"])
            graph["launches"] = _lib.launch_count()      # our kernels inside one replay
            graph["g"] = g_
        except Exception as e:                      # noqa: BLE001
            sys.stderr.write("cuda graph capture failed, running eagerly: %r\n" % (e,))
            graph["g"] = None
            torch.cuda.synchronize()

    def step(rays, tgt):
        if graph["g"] is not None:
            graph["rays"].copy_(rays, non_blocking=True)
            graph["tgt"].copy_(tgt, non_blocking=True)
            graph["g"].replay()
            loss = graph["loss"]
        else:
            loss = fwd_bwd(rays, tgt)
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

    def e2e(i):
        r = host_rays[i % nbatch].to(dev, non_blocking=True)
        t = host_tgt[i % nbatch].to(dev, non_blocking=True)
        last["loss"] = float(step(r, t).item())

    for i in range(args.warmup):
        resident(i)
    assert flat.check_views(), "param.grad views were replaced; the flat all-reduce buffer is stale"
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.launch_count(reset=True)
    ms = timed(resident, args.steps)
    launches = _lib.launch_count()
    for i in range(2):
        e2e(i)
    ms_e2e = timed(e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions

    # per-kernel device times (CUDA events on the launching stream) over a few instrumented steps
    use_graph = graph["g"] is not None
    graph["g"] = None                               # the instrumented steps below run eagerly
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
    evals = N_RAND * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE)
    if precision == "tc":
        t_fwd = ktimes.get("swnerf_tc_mlp_fwd", 0.0)
        t_bwd = ktimes.get("swnerf_tc_mlp_bwd", 0.0)
        # dominant kernel = fused forward (2 launches/step: coarse 64 + fine 192 samples per ray)
        ach = FLOP_PER_EVAL_FWD * evals / (t_fwd * 1e-3) / 1e12 if t_fwd > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "swnerf_tc_mlp_fwd (fused PE + 8x256 MLP, tcgen05)",
                "achieved": ach, "peak": pk["bf16_tflops_sustained"] or pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / (pk["bf16_tflops_sustained"] or pk["bf16_tflops"]),
                # dram read+write of the training forward per step from profiles/ (ncu --set full, r1): 3.85 GB per
                # 786,432-sample launch = 4894 B/sample (saved activation images), x 1,048,576 samples
                "traffic": 4894.0 * evals,
                "peak_src": pk["src"] + " cuBLAS bf16 (sustained: kernel timed inside a long step)",
                "ms_per_step": t_fwd,
                "bwd": {"ms_per_step": t_bwd, "data_kernel_ms": bwd_split["data"], "weight_kernel_ms": bwd_split["weight"],
                        "achieved": 2 * FLOP_PER_EVAL_FWD * evals / (t_bwd * 1e-3) / 1e12 if t_bwd > 0 else 0.0,
                        "weight_kernel_hbm": {"bound": "hbm", "algorithmic_bytes_per_sample": 9728,
                                              "achieved_gbs": 9728.0 * evals / (bwd_split["weight"] * 1e-3) / 1e9
                                              if bwd_split["weight"] > 0 else 0.0, "peak_gbs": pk["hbm_gbs"]}}}
    else:
        t_mm = ktimes.get("swnerf_sgemm", 0.0)
        ach = 3 * FLOP_PER_EVAL_FWD * evals / (t_mm * 1e-3) / 1e12 if t_mm > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "swnerf_sgemm (fp32 SIMT check path; not the tensor-core kernel)",
                "achieved": ach, "peak": pk["bf16_tflops_sustained"] or pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / (pk["bf16_tflops_sustained"] or pk["bf16_tflops"]), "traffic": None,
                "peak_src": pk["src"], "ms_per_step": t_mm}

    extra = {}
    if args.render_frame:
        # config #3: full 800x800 frame, test-time kwargs (perturb=0, no noise), chunk 32768, rays sharded over ranks
        H = W = 800
        focal = 0.5 * W / np.tan(0.5 * 0.6911112)
        Kmat = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
        c2w = torch.from_numpy(synth.pose_spherical(40.0, -30.0, 4.0)[:3, :4])
        kwt = dict(kw); kwt["perturb"] = 0.0
        with torch.no_grad():
            def frame_fn():
                return parallel.render_frame_sharded(H, W, Kmat, c2w, 2.0, 6.0, 1024 * 32, S.render_rays, **kwt)
            frame_fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); frame_fn(); frame_fn(); e1.record()
            barrier()
            fm = torch.tensor([e0.elapsed_time(e1) / 2], device=dev)
            if world > 1:
                dist.all_reduce(fm, op=dist.ReduceOp.MAX)
        extra["render_ms_per_frame_800x800"] = float(fm.item())

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rps, spp = cpu_reference_run(3, 1, args.cpu_sample_rays, threads)
            cpu = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                   "sample": "%d rays/step x 3 steps of the same training step (fwd+bwd+Adam), oracle port, torch "
                             "CPU fp32" % args.cpu_sample_rays}
        rays_total = N_RAND * world * args.steps
        line = {"metric": "train_rays_per_s", "value": rays_total / (ms * 1e-3), "unit": "rays/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate (tcgen05)" if precision == "tc" else "f32",
                "data": "synthetic", "config": config, "precision_mode": precision,
                "e2e": {"value": rays_total / (ms_e2e * 1e-3), "unit": "rays/s",
                        "h2d_bytes_per_step": N_RAND * (11 + 3) * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps, "last_loss": last.get("loss")},
                "gpu_launches": launches + (graph.get("launches", 0) * args.steps if use_graph else 0),
                "cuda_graph": use_graph, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
                "kernel_ms_per_step": ktimes, "kernel_calls_per_step": kcalls}
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
