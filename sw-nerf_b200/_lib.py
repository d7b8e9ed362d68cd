"""ctypes binding of libswnerf_b200.so (the C ABI in include/swnerf_b200.h).

There is no CPU fallback: if the shared library is missing, or a tensor is not a contiguous CUDA
tensor of the right dtype, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswnerf_b200.so")
_lib = None

c_f32p = ctypes.c_void_p
_I64, _I32, _VP, _F32 = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_float

# name -> argtypes (restype int unless listed in _RET)
_SIG = {
    "swnerf_version": [],
    "swnerf_device_ok": [],
    "swnerf_launch_count": [_I32],
    "swnerf_searchsorted": [_VP, _VP, _VP, _I64, _I64, _I64, _I64, _I32, _VP],
    "swnerf_stratified_z": [_VP, _I32, _I32, _VP, _VP, _I64, _I32, _I32, _I32, _VP],
    "swnerf_embed_fwd": [_VP, _VP, _I64, _I32, _I32, _VP],
    "swnerf_embed_bwd": [_VP, _VP, _VP, _I64, _I32, _I32, _VP],
    "swnerf_encode_points": [_VP, _I32, _I32, _VP, _VP, _I64, _I32, _I32, _I32, _I32, _VP],
    "swnerf_composite_fwd": [_VP, _I32, _VP, _VP, _I32, _I32, _VP, _I32, _I64, _I32, _VP, _VP, _VP, _VP, _VP, _VP],
    "swnerf_composite_bwd": [_VP, _I32, _VP, _VP, _I32, _I32, _VP, _I32, _I64, _I32, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                             _VP, _VP],
    "swnerf_sample_pdf": [_VP, _VP, _VP, _VP, _I32, _I64, _I32, _I32, _VP, _VP, _VP],
    "swnerf_resample": [_VP, _VP, _VP, _I32, _I64, _I32, _I32, _VP, _VP, _VP, _VP],
    "swnerf_resample_check": [_VP, _VP, _VP, _VP, _I32, _I64, _I32, _I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP],
    "swnerf_sgemm": [_I32, _VP, _I64, _VP, _I64, _VP, _I64, _I64, _I64, _I64, _VP, _I32, _I32, _VP, _I64, _VP],
    "swnerf_colsum": [_VP, _I64, _I64, _I32, _VP, _I32, _VP],
    "swnerf_hgemm_tc_supported": [_I64, _I64],
    "swnerf_hgemm_tc": [_I32, _VP, _I64, _VP, _I64, _VP, _I64, _I64, _I64, _I64, _VP, _I32, _I32, _VP, _I64, _F32, _VP,
                        _VP],
    "swnerf_pow2_scale": [_VP, _I64, _F32, _VP, _VP],
    "swnerf_hgemm_tc_wgrad": [_VP, _I64, _VP, _I64, _VP, _I64, _I64, _I64, _I64, _F32, _VP, _VP],
    "swnerf_act_bwd": [_VP, _I64, _VP, _I64, _I64, _I32, _I32, _VP, _I64, _VP],
    "swnerf_tc_packed_bytes": [],
    "swnerf_tc_packed_t_bytes": [],
    "swnerf_tc_workspace_bytes": [_I64, _I32, _I32],
    "swnerf_tc_pack_weights": [_VP, _I32, _VP, _VP],
    "swnerf_tc_pack_weights_t": [_VP, _I32, _VP, _VP, _VP],
    "swnerf_tc_mlp_fwd": [_VP, _I32, _I32, _VP, _I64, _I32, _VP, _I32, _VP, _VP, _I32, _VP],
    "swnerf_tc_mlp_fwd_points": [_VP, _I32, _I32, _VP, _I64, _I32, _VP, _I32, _VP, _VP, _I32, _VP],
    "swnerf_tc_mlp_bwd_points": [_VP, _I64, _I32, _VP, _VP, _I32, _VP, _VP, _VP, _F32, _VP, _VP, _VP],
    "swnerf_tc_pack_weights_time": [_VP, _VP, _I32, _VP, _VP],
    "swnerf_tc_pack_weights_time_t": [_VP, _I32, _VP, _VP, _VP],
    "swnerf_tc_time_fwd": [_VP, _I32, _I32, _VP, _I64, _I32, _VP, _I32, _VP, _VP, _I32, _VP],
    "swnerf_tc_time_bwd": [_VP, _I64, _I32, _VP, _VP, _I32, _VP, _VP, _VP, _VP, _F32, _VP],
    "swnerf_make_rays": [_I32, _I32, _F32, _F32, _F32, _F32, _VP, _VP, _I64, _F32, _F32, _F32, _I32, _I32, _I32, _F32,
                         ctypes.c_double, _VP, _I32, _VP],
    "swnerf_pick_batch": [_I32, _I32, _F32, _F32, _F32, _F32, _VP, _VP, _I32, _I32, _I32, _I32, ctypes.c_uint64, _I64, _F32,
                          _F32, _F32, _I32, _I32, _I32, _F32, ctypes.c_double, _VP, _I32, _VP, _VP, _VP],
    "swnerf_adam_flat": [_VP, _VP, _VP, _VP, _I64, _F32, _F32, _F32, _F32, _I64, _VP],
    "swnerf_mse2": [_VP, _VP, _VP, _I64, _F32, _VP, _VP, _VP, _VP],
    "swnerf_tc_set_profiling": [_I32],
    "swnerf_tc_last_bwd_ms": [_VP, _VP],
    "swnerf_tc_selftest": [_I32, _VP, _VP, _VP, _I32, _I32, _VP, _VP],
    "swnerf_tc_probe": [_I32, _I32, _I32, _VP, _VP],
    "swnerf_tc_set_fwd_variant": [_I32],
    "swnerf_tc_set_bwd_variant": [_I32],
    "swnerf_set_resample_variant": [_I32],
    "swnerf_resample_fallbacks": [_VP, _I32, _VP],
    "swnerf_tc_selftest_pair": [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _VP, _VP, _VP],
    "swnerf_tc_mlp_bwd": [_VP, _I64, _I32, _VP, _VP, _I32, _VP, _VP, _VP, _F32, _VP],
}
_RET = {
    "swnerf_launch_count": _I64,
    "swnerf_tc_packed_bytes": _I64,
    "swnerf_tc_packed_t_bytes": _I64,
    "swnerf_tc_workspace_bytes": _I64,
}


def declared_symbols():
    return ["swnerf_last_error"] + sorted(_SIG)


def lib():
    """Load the C-ABI library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "swnerf_b200: %s is missing - build it with `python sw-nerf_b200/build.py` "
                "(there is no CPU or PyTorch fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.swnerf_last_error.restype = ctypes.c_char_p
        L.swnerf_last_error.argtypes = []
        for name, args in _SIG.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RET.get(name, _I32)
        _lib = L
    return _lib


def last_error():
    return lib().swnerf_last_error().decode("utf-8", "replace")


# Optional per-entry-point device timing (bench.py): TIMING = {} enables it; every call then records a
# CUDA event pair on the launching stream, appended to TIMING[name].
TIMING = None


def call(name, *args):
    """Call an int-status entry point; raise RuntimeError(message) on failure."""
    if TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        TIMING.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, last_error()))


def launch_count(reset=False):
    return int(lib().swnerf_launch_count(1 if reset else 0))


def resample_fallbacks(reset=True):
    """Rays the eight-lane resample kernel handed to its exact generic routine since the last reset (diagnostic)."""
    n = ctypes.c_ulonglong(0)
    call("swnerf_resample_fallbacks", ctypes.addressof(n), 1 if reset else 0, stream())
    return int(n.value)


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t, dtype=torch.float32, name="tensor", allow_none=False):
    """Device pointer of a contiguous CUDA tensor; rejects anything else (no silent copies)."""
    if t is None:
        if allow_none:
            return None
        raise ValueError("%s must not be None" % name)
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("swnerf_b200: %s is on %s - this library has no CPU path" % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t.data_ptr()


def ptr_array(tensors):
    """Host array of device pointers (const float* const*)."""
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr
