"""ctypes binding of ``libdiga_b200.so`` (the C ABI declared in ``include/diga_b200.h``).

There is no CPU fallback and no alternative backend: if the shared library has not been built
(``python diga_b200/build.py``) importing this module raises, and every wrapper refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import functools
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DIGA_B200_LIB") or os.path.join(_HERE, "libdiga_b200.so")   # override: A/B builds of the same ABI

if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA extension is not built. Run `python diga_b200/build.py` "
        "(nvcc, sm_100a). diga_b200 has no CPU or PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i64 = C.c_int64
_f = C.c_float
_d = C.c_double
_i = C.c_int

# name -> (restype, argtypes); kept in the same order as include/diga_b200.h
SIGNATURES = {
    "diga_version": (_i, []),
    "diga_last_error_string": (C.c_char_p, []),
    "diga_launch_count": (_i64, []),
    "diga_set_tunable": (_i, [C.c_char_p, _i]),
    "diga_kd_workspace_bytes": (C.c_size_t, []),
    "diga_kd_fwd": (_i, [_p, _p, _i64, _i64, _i64, _f, _p, _p, _p]),
    "diga_kd_bwd": (_i, [_p, _p, _i64, _i64, _i64, _f, _p, _p, _p]),
    "diga_kd_fwd_bwd": (_i, [_p, _p, _i64, _i64, _i64, _f, _f, _p, _p, _p, _p]),
    "diga_pseudo_label": (_i, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_class_presence": (_i, [_p, _i64, _i64, _p, _p, _p]),
    "diga_classmix_blend": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_centroid_clsw_bytes": (_i64, [_i64, _i64]),
    "diga_centroid_clsw_build": (_i, [_p, _i64, _i64, _i64, _p, _p]),
    "diga_centroid_assign": (_i, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_centroid_assign_fullres": (_i, [_p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_centroid_accum": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _p]),
    "diga_centroid_means": (_i, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_centroid_update": (_i, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _i, _i, _d, _p]),
    "diga_centroid_finish_supported": (_i, [_i64, _i64]),
    "diga_centroid_finish": (_i, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _i, _i, _d, _p]),
    "diga_centroid_chain_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64]),
    "diga_centroid_chain": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _i, _i, _d, _p]),
    "diga_centroid_chain_sums": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_centroid_chain_reduce": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "diga_centroid_means_scatter": (_i, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _p]),
    "diga_centroid_update_sharded": (_i, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _i, _i, _d, _p]),
    "diga_centroid_update_single": (_i, [_p, _i64, _i64, _i64, _p, _p, _i, _i, _d, _p]),
    "diga_centroid_reduce_images": (_i, [_p, _p, _p, _i64, _i64, _i64, _p, _p]),
    "diga_onehot_labels": (_i, [_p, _i64, _i64, _i64, _p, _p]),
    "diga_proto_workspace_bytes": (C.c_size_t, [_i64, _i64]),
    "diga_proto_distance": (_i, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_proto_prepare": (_i, [_p, _i64, _i64, _p, _p]),
    "diga_proto_distance_prepared": (_i, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_consensus_select": (_i, [_p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "diga_consensus_select_u8": (_i, [_p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "diga_upsample_bilinear": (_i, [_p, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "diga_pseudo_label_upsampled": (_i, [_p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "diga_ce_workspace_bytes": (C.c_size_t, []),
    "diga_cross_entropy2d_fwd": (_i, [_p, _p, _p, _i64, _i64, _i64, _i, _p, _p, _p, _p]),
    "diga_cross_entropy2d_bwd": (_i, [_p, _p, _p, _i64, _i64, _i64, _i, _p, _p, _p, _p]),
    "diga_loss_up_workspace_bytes": (C.c_size_t, [_i64, _i64, _i64, _i64, _i64, _i64]),
    "diga_loss_up_fwd": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f, _i, _p, _p, _p, _p, _p]),
    "diga_loss_up_bwd": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f, _i, _p, _p, _p, _p, _p, _p]),
    "diga_ce_up_fwd_bwd": (_i, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "diga_seg_kd_up_fwd_bwd": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f, _i, _f, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p]),
    "diga_loss_up_scratch_bytes": (C.c_size_t, [_i64, _i64, _i64, _i64, _i64, _i64]),
    "diga_loss_up_gather": (_i, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "diga_kd_up_fwd_bwd": (_i, [_p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _f, _f, _p, _p, _p, _p, _p]),
    "diga_ohem_up_workspace_bytes": (C.c_size_t, []),
    "diga_ohem_up_fwd": (_i, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "diga_ohem_up_bwd": (_i, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "diga_ema_update": (_i, [_p, _p, _p, _i64, _d, _p]),
    "diga_label_resize_remap": (_i, [_p, _i64, _i64, _i64, _p, _p, _i64, _i64, _p, _p, _p]),
    "diga_confusion_matrix": (_i, [_p, _i, _p, _i, _i64, _i64, _p, _p, _p]),
    "diga_png_deflate_capacity": (_i64, [_i64, _i64]),
    "diga_png_deflate_scratch_bytes": (_i64, [_i64, _i64]),
    "diga_png_deflate": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p, _p, _p]),
    "diga_png_write_file": (_i, [C.c_char_p, _p, _i64, _i64, _i64, _p, _i64]),
    "diga_png_crc": (_i, [_p, _i64, _i64, _p, _p, _p]),
    "diga_png_write_file_crc": (_i, [C.c_char_p, _p, _i64, _i64, _i64, _p, _i64, C.c_uint32]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)      # AttributeError here == the .so is stale; rebuild
    _fn.restype = _res
    _fn.argtypes = _args

UPDATE_MEAN, UPDATE_MOVING_AVERAGE = 0, 1


def last_error() -> str:
    return lib.diga_last_error_string().decode()


def check(rc: int) -> None:
    """Non-zero C return code -> RuntimeError carrying diga_last_error_string()."""
    if rc != 0:
        raise RuntimeError(f"diga_b200 error {rc}: {last_error()}")


def launch_count() -> int:
    return int(lib.diga_launch_count())


def set_tunable(name: str, value: int) -> None:
    lib.diga_set_tunable(name.encode(), int(value))


def ptr(t):
    """data_ptr of a tensor, or NULL for None."""
    return None if t is None else t.data_ptr()


def stream() -> int:
    """Raw handle of torch's current stream ON THE CURRENT DEVICE.  Every public wrapper runs under :func:`on_device`, which
    makes the device of its tensor arguments current first, so this is always a stream of the device the pointers live on."""
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors, what="input") -> None:
    """All given tensors (``None`` skipped) must be CUDA tensors on ONE device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"diga_b200: {what} must be a CUDA tensor (there is no CPU fallback); got {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"diga_b200: {what} tensors live on different devices ({dev} and {t.device})")


def _tensor_device(args, kwargs):
    dev = None
    for a in (*args, *kwargs.values()):
        if isinstance(a, (list, tuple)) and a and isinstance(a[0], torch.Tensor):
            a = a[0]
        if isinstance(a, torch.Tensor) and a.is_cuda:
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise RuntimeError(f"diga_b200: arguments live on different CUDA devices ({dev} and {a.device})")
    return dev


def on_device(fn):
    """Device guard of the ctypes layer: the kernels are launched on the CURRENT device's current stream, so a wrapper
    called with tensors of another device (``Class_Features(device='cuda:1')`` while cuda:0 is current) first makes that
    device current for the duration of the call (``torch.cuda.device``); tensors on two different devices raise."""
    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        dev = _tensor_device(args, kwargs)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return guarded


def f32c(t):
    """Contiguous fp32 view of ``t`` (the reference tensors already are; this is a no-op for them)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def i64c(t):
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


_workspaces = {}


def _workspace(kind: str, nbytes: int, device: torch.device) -> torch.Tensor:
    key = (kind, device.index, stream())
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def loss_up_workspace(n, c, h, w, hh, ww, device: torch.device) -> torch.Tensor:
    """Workspace of the fused up-sampling losses (reduction partials + gradient patches), one per geometry/stream."""
    nbytes = int(lib.diga_loss_up_workspace_bytes(n, c, h, w, hh, ww))
    return _workspace(("loss_up", n, c, h, w, hh, ww), nbytes, device)


def ohem_workspace(device: torch.device) -> torch.Tensor:
    """Zero-initialised OHEM selection state (histogram, rank, reduction partials), one per (device, stream)."""
    return _workspace("ohem", int(lib.diga_ohem_up_workspace_bytes()), device)


def kd_workspace(device: torch.device) -> torch.Tensor:
    """Zero-initialised KD reduction workspace, one per (device, stream)."""
    return _workspace("kd", int(lib.diga_kd_workspace_bytes()), device)


def ce_workspace(device: torch.device) -> torch.Tensor:
    """Zero-initialised cross-entropy reduction workspace, one per (device, stream)."""
    return _workspace("ce", int(lib.diga_ce_workspace_bytes()), device)
