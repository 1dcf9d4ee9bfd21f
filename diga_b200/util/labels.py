"""Label maps as ``util/loader/CityLoader.py`` hands them to the training step — the reader half of SURVEY.md §8f row 3.

The reference decodes the label / pseudo-label PNG with PIL, resizes it with ``Image.NEAREST`` (:93-95) and re-assigns ids
with one full-image numpy pass per id (``label_copy[label == k] = v``, :115-132: 34 passes for ground truth, 19 for pseudo-
labels), all inside the DataLoader worker.  Here the decoded uint8 map goes to the GPU as it is (1 B/px over PCIe instead
of an int64 map) and ONE kernel does the resize gather and the id look-up (csrc/labels.cu).  The PNG decode itself stays
PIL; augmentation transforms (random crop / flip, ``self.transform``) are the loader's business and out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L

# util/loader/CityLoader.py:49-56
ID_TO_TRAINID = {7: 0, 8: 1, 11: 2, 12: 3, 13: 4, 17: 5, 19: 6, 20: 7, 21: 8, 22: 9, 23: 10, 24: 11, 25: 12, 26: 13, 27: 14,
                 28: 15, 31: 16, 32: 17, 33: 18}


def pil_nearest_table(n_in: int, n_out: int) -> np.ndarray:
    """Source index of every output index under ``Image.resize(..., Image.NEAREST)``: Pillow's ImagingScaleAffine walks
    ``xo = scale * 0.5; xin = (int)xo; xo += scale`` in double precision (the accumulated sum, not a product, decides
    the exact-boundary cases), ``scale = n_in / n_out``."""
    scale = np.float64(n_in) / np.float64(n_out)
    xo = np.add.accumulate(np.concatenate(([scale * np.float64(0.5)], np.full(n_out - 1, scale, dtype=np.float64))))
    idx = xo.astype(np.int64)                                     # truncation of a non-negative double
    return np.minimum(idx, n_in - 1).astype(np.int32)


def pseudo_label_lut(n_classes: int = 19) -> np.ndarray:
    """CityLoader.py:129-131: ``pseudo_label_copy = 255; pseudo_label_copy[pseudo_label == v] = v for v < n_classes``."""
    lut = np.full(256, 255, dtype=np.uint8)
    lut[:n_classes] = np.arange(n_classes, dtype=np.uint8)
    return lut


def trainid_lut(id_to_trainid=None) -> np.ndarray:
    """CityLoader.py:116-118 / :125-127: ``label_copy = 255; label_copy[label == k] = v``."""
    lut = np.full(256, 255, dtype=np.uint8)
    for k, v in (id_to_trainid or ID_TO_TRAINID).items():
        lut[k] = v
    return lut


_tables = {}


def _device_tables(in_size, out_size, device):
    key = (tuple(in_size), tuple(out_size), device.index)
    t = _tables.get(key)
    if t is None:
        yt = torch.from_numpy(pil_nearest_table(in_size[0], out_size[0])).to(device)
        xt = torch.from_numpy(pil_nearest_table(in_size[1], out_size[1])).to(device)
        t = _tables[key] = (yt, xt)
    return t


@L.on_device
def resize_remap_labels(src_u8: torch.Tensor, size=None, lut: np.ndarray = None) -> torch.Tensor:
    """``src_u8`` ``[N,h0,w0]`` (or ``[h0,w0]``) uint8 CUDA tensor of decoded PNG values -> int64 ``[N,H,W]``:
    PIL-NEAREST resize to ``size`` = (H, W) (``None``: keep) followed by the 256-entry id look-up."""
    L.require_cuda(src_u8, what="resize_remap_labels input")
    if src_u8.dtype != torch.uint8:
        raise ValueError("resize_remap_labels: expected the uint8 values of the decoded PNG")
    squeeze = src_u8.dim() == 2
    src = (src_u8.unsqueeze(0) if squeeze else src_u8).contiguous()
    n, h0, w0 = src.shape
    hh, ww = (h0, w0) if size is None else (int(size[0]), int(size[1]))
    yt, xt = _device_tables((h0, w0), (hh, ww), src.device)
    lut = np.ascontiguousarray(pseudo_label_lut() if lut is None else lut, dtype=np.uint8)
    if lut.shape != (256,):
        raise ValueError("resize_remap_labels: lut must hold 256 uint8 entries")
    out = torch.empty((n, hh, ww), dtype=torch.int64, device=src.device)
    L.check(L.lib.diga_label_resize_remap(src.data_ptr(), n, h0, w0, yt.data_ptr(), xt.data_ptr(), hh, ww, lut.ctypes.data,
                                          out.data_ptr(), L.stream()))
    return out[0] if squeeze else out


def _decode(path_or_image) -> np.ndarray:
    from PIL import Image
    img = Image.open(path_or_image) if not hasattr(path_or_image, "getpixel") else path_or_image
    arr = np.asarray(img)
    if arr.ndim != 2 or arr.dtype != np.uint8:
        raise ValueError("label PNGs must be single-channel 8-bit ('L' or 'P' mode)")
    return arr


def read_pseudo_label(path_or_image, crop_size=None, n_classes: int = 19, device="cuda") -> torch.Tensor:
    """CityLoader.py:62-70 (file), :88 (open), :94-95 (NEAREST resize to ``crop_size`` = (H, W)), :123,:129-131 (ids >=
    ``n_classes`` -> 255): the pseudo-label map of one image as an int64 CUDA tensor ``[H,W]``."""
    src = torch.from_numpy(np.ascontiguousarray(_decode(path_or_image))).to(device, non_blocking=True)
    return resize_remap_labels(src, crop_size, pseudo_label_lut(n_classes))


def read_label(path_or_image, crop_size=None, id_to_trainid=None, device="cuda") -> torch.Tensor:
    """CityLoader.py:86,:93,:113-118: gtFine ``labelIds`` PNG -> trainId map (everything else 255), int64 CUDA ``[H,W]``."""
    src = torch.from_numpy(np.ascontiguousarray(_decode(path_or_image))).to(device, non_blocking=True)
    return resize_remap_labels(src, crop_size, trainid_lut(id_to_trainid))
