"""swnerf_b200: B200-native implementation of SW-NeRF's per-ray volumetric rendering hot path.

The directory is `sw-nerf_b200/`; it is importable as `swnerf_b200` through the alias module at the
repository root.  Layout mirrors the reference's import surface: embedder / model / ray (the three
files the runners `from X import *`), render (nerf/run.py's five functions), dnerf
(d_nerf/run_dnerf.py's), multires (multires_dnerf.py's create_nerf), tnerf (t_nerf/run_tnerf.py's),
parallel (ray-sharded DP).
"""
from . import _lib  # noqa: F401
from . import ops, embedder, model, ray, render, tc  # noqa: F401
from .embedder import get_embedder, Embedder  # noqa: F401
from .model import vallina_NeRF, NeRFOriginal, DirectTemporalNeRF, NeRF, TNeRF  # noqa: F401
from .ray import sample_pdf, raw2outputs, get_rays, get_rays_np, ndc_rays  # noqa: F401
from .render import batchify, run_network, batchify_rays, render_rays, create_nerf, render, render_path, NetworkQuery  # noqa: F401
from .ops import searchsorted  # noqa: F401

__version__ = "0.1.0"
