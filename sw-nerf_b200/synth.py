"""Synthetic workload of the benchmarks: Blender-shaped pinhole rays and a non-degenerate
random scene.  There is no network for datasets or checkpoints, so bench.py and tools/ build their inputs here;
numpy's legacy MT19937 RandomState keeps them bit-stable across machines.  (The CPU checker under tests/ keeps its own
copy of these generators for the golden vectors; tests/test_host_logic.py holds the two copies bit-identical.)
"""
import math

import numpy as np
import torch


def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    """Camera-to-world matrix of dataloader/load_blender.py:9-35: translate along z, rotate by phi about x and by theta
    about y, flip to the Blender axes."""
    t = np.eye(4, dtype=np.float64); t[2, 3] = radius
    ph, th = phi_deg / 180.0 * np.pi, theta_deg / 180.0 * np.pi
    rp = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1.0]])
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1.0]])
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    return (flip @ (rt @ rp @ t)).astype(np.float32)


def blender_rays(n_rays: int, seed: int, H: int = 800, W: int = 800, near: float = 2.0, far: float = 6.0,
                 frame_time=None) -> np.ndarray:
    """Flat ray batch [N, 11] (o, d, near, far, unit viewdir) or [N, 12] (+ frame_time before the viewdir) of random
    pixels of one random camera on the lego orbit: focal from camera_angle_x = 0.6911112 (load_blender.py:133-134),
    c2w = pose_spherical(theta, -30, 4), directions as ray.py:42-72, near 2 / far 6 (nerf/run.py:466-467)."""
    rs = np.random.RandomState(seed)
    focal = 0.5 * W / np.tan(0.5 * 0.6911112)
    c2w = pose_spherical(rs.uniform(-180.0, 180.0), -30.0, 4.0)
    pix = rs.randint(0, H * W, size=n_rays)
    i, j = (pix % W).astype(np.float32), (pix // W).astype(np.float32)
    dirs = np.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -np.ones_like(i)], -1).astype(np.float32)
    rays_d = np.sum(dirs[:, None, :] * c2w[:3, :3], -1).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(np.float32)
    cols = [rays_o, rays_d, np.full((n_rays, 1), near, np.float32), np.full((n_rays, 1), far, np.float32)]
    if frame_time is not None:
        cols.append(np.full((n_rays, 1), frame_time, np.float32))
    cols.append((rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)).astype(np.float32))
    return np.concatenate(cols, -1).astype(np.float32)


def scene_params(module: torch.nn.Module, seed: int) -> dict:
    """state_dict of a random but NON-EMPTY scene for `module`: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's
    default init, with the density head scaled (x24, bias 0.5) and the colour head scaled (x6) - a default-init network
    renders acc = 0 everywhere, which makes every map and every gradient degenerate."""
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    rs = np.random.RandomState(seed)
    out = {}
    for name, sh in shapes.items():
        fan_in = sh[1] if len(sh) == 2 else shapes[name.replace(".bias", ".weight")][1]
        b = 1.0 / math.sqrt(fan_in)
        out[name] = torch.from_numpy(rs.uniform(-b, b, size=sh).astype(np.float32))
    for name in out:
        if name.endswith(("alpha_linear.weight", "density.0.weight")):
            out[name] = out[name] * 24.0
        elif name.endswith(("alpha_linear.bias", "density.0.bias")):
            out[name] = out[name] * 0.0 + 0.5
        elif name.endswith(("rgb_linear.weight", "color.0.weight")):
            out[name] = out[name] * 6.0
        elif name.endswith("_time_out.weight"):
            out[name] = out[name] * 2.0
    return out
