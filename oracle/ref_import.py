"""Import harness for the UNMODIFIED reference (test infrastructure only).

Only usable where /root/reference exists (the build container).  It is used by
oracle/make_golden.py to pin the oracle; nothing that runs on the GPU box may
import this module.

The reference runners import I/O-only packages that are not installed here
(imageio, lpips, skimage, matplotlib, configargparse, trimesh); they are
replaced by empty stub modules so that nerf/run.py and d_nerf/run_dnerf.py
import unmodified and their real create_nerf/render/render_rays execute.
"""
import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SWNERF_REFERENCE", "/root/reference")

_STUBS = [
    "imageio", "lpips", "skimage", "skimage.metrics", "matplotlib",
    "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d",
    "configargparse", "trimesh", "mcubes",
]


def available() -> bool:
    return os.path.isdir(REF_ROOT) and os.path.isfile(os.path.join(REF_ROOT, "ray.py"))


def _install_stubs():
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package
        sys.modules[name] = m
    sm = sys.modules["skimage.metrics"]
    for fn in ("structural_similarity", "peak_signal_noise_ratio"):
        if not hasattr(sm, fn):
            setattr(sm, fn, lambda *a, **k: 0.0)
    plt = sys.modules["matplotlib.pyplot"]
    if not hasattr(plt, "rcParams"):
        plt.rcParams = {}
    ax = sys.modules["mpl_toolkits.mplot3d"]
    if not hasattr(ax, "Axes3D"):
        ax.Axes3D = object


def _load(name, relpath):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with the reference modules: embedder, model, ray,
    run (nerf/run.py) and run_dnerf (d_nerf/run_dnerf.py)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    ns.embedder = importlib.import_module("embedder")
    ns.model = importlib.import_module("model")
    ns.ray = importlib.import_module("ray")
    ns.run = _load("ref_nerf_run", "nerf/run.py")
    ns.run_dnerf = _load("ref_dnerf_run", "d_nerf/run_dnerf.py")
    return ns


def load_reference_tnerf():
    """t_nerf/run_tnerf.py, unmodified (it resolves its own `device` at import: CPU here)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return _load("ref_tnerf_run", "t_nerf/run_tnerf.py")


def load_reference_searchsorted_numpy():
    """The reference's own numpy oracle for its torchsearchsorted extension
    (d_nerf/torchsearchsorted/src/torchsearchsorted/utils.py:4-14, UTF-16LE)."""
    p = os.path.join(REF_ROOT, "d_nerf/torchsearchsorted/src/torchsearchsorted/utils.py")
    src = open(p, "rb").read().decode("utf-16")
    src = src.replace("np.long", "np.int64")  # numpy>=1.24 removed the alias
    g = {}
    exec(compile(src, p, "exec"), g)
    return g["numpy_searchsorted"]
