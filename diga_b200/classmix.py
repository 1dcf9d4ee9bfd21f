"""Cross-domain ClassMix: class-presence bitmap on the GPU, class choice on the host, LUT blend on the GPU.

The reference has no function for this; it is the statement block
``train_DiGA_gta2city_self_training.py:259-275`` (image only), ``:306-325`` (DACS: image + label),
``train_DiGA_gta2city_warm_up.py:240-259`` and ``calc_centroids.py:47-58``.  ``classmix`` has tensor-in /
tensor-out semantics equal to those blocks; a patched script replaces each block by one call
(INTEGRATION.md).  The host draws ``random.sample(sorted_present_classes, len // 2)`` per image exactly as
the reference does, so a seeded ``random`` produces the same masks.
"""
from __future__ import annotations

import random as _random

import numpy as np
import torch

from . import _lib as L

IGNORE = 255


_pinned = {}
_PINNED_SLOTS = 4          # outstanding asynchronous presence requests per size


def _pinned_i32(n: int) -> torch.Tensor:
    slots = _pinned.setdefault(n, {"next": 0, "bufs": []})
    if len(slots["bufs"]) < _PINNED_SLOTS:
        slots["bufs"].append(torch.empty((n,), dtype=torch.int32).pin_memory())
        return slots["bufs"][-1]
    slots["next"] = (slots["next"] + 1) % _PINNED_SLOTS
    return slots["bufs"][slots["next"]]


class PresentClasses:
    """Handle of an in-flight presence pass: the kernel and the 32 B/image D2H copy are queued on the current stream at
    construction; ``result()`` waits for THAT copy only (a CUDA event), so work queued behind it keeps the GPU busy."""

    def __init__(self, slabel: torch.Tensor):
        L.require_cuda(slabel, what="classmix label")
        lab = L.i64c(slabel)
        self.b = lab.shape[0]
        self._present = [] if self.b == 0 else None
        if self.b == 0:
            return
        hw = lab.shape[1] * lab.shape[2]
        # one device buffer [b*8 bitmap words | 1 flag word] and one pinned host mirror: a single D2H copy
        self._buf = torch.empty((self.b * 8 + 1,), dtype=torch.int32, device=lab.device)
        L.check(L.lib.diga_class_presence(lab.data_ptr(), self.b, hw, self._buf.data_ptr(), self._buf.data_ptr() + self.b * 32,
                                          L.stream()))
        self._host = _pinned_i32(self.b * 8 + 1)
        self._host.copy_(self._buf, non_blocking=True)
        self._done = torch.cuda.Event()
        self._done.record()

    def result(self):
        if self._present is None:
            self._done.synchronize()                                              # the one host sync
            host = self._host.numpy().view(np.uint32)
            if host[-1]:
                raise ValueError("classmix: labels must lie in [0, 255] (trainIds plus the 255 ignore value)")
            bits = np.unpackbits(host[:-1].copy().view(np.uint8).reshape(self.b, 32), axis=1, bitorder="little")
            self._present = [np.nonzero(row)[0].tolist() for row in bits]
            self._buf = None
        return self._present


@L.on_device
def present_classes_async(slabel: torch.Tensor) -> PresentClasses:
    """Start the presence pass for ``slabel`` and return its handle (pass it to ``classmix(..., present=handle)``).  Both
    ClassMix blocks of a self-training step use the same ``slabelv`` (:265, :310), so one pass issued when the batch
    arrives serves both, and its host round trip hides behind whatever is queued after it."""
    return PresentClasses(slabel)


@L.on_device
def present_classes(slabel: torch.Tensor):
    """Per image, the sorted list of label values present — ``torch.unique(slabel[i]).tolist()`` (:265) — from a
    256-bit device bitmap (32 B per image over PCIe instead of a sort + sync per image)."""
    return PresentClasses(slabel).result()


def select_classes(present, rng=_random):
    """``random.sample(label_list, len(label_list) // 2)`` then append 255 if absent (:266-268)."""
    chosen = []
    for label_list in present:
        sel = rng.sample(label_list, len(label_list) // 2)
        if IGNORE not in sel:
            sel.append(IGNORE)
        chosen.append(sel)
    return chosen


@L.on_device
def classmix(slabel, a, b, tlabel=None, rng=_random, classes=None, return_mask=True, assume_labelled=False, present=None):
    """ClassMix mask build + blend.

    ``slabel [B,H,W]`` int64 source labels; ``a``, ``b`` ``[B,CH,H,W]`` fp32; optional ``tlabel [B,H,W]`` int64.
    ``mask = 1`` where ``slabel`` is one of the chosen classes; ``mix = a*(1-mask) + b*mask`` (bit-exact with
    the reference expression); with ``tlabel``: ``mixlabel = where(mask, slabel, tlabel)``.
    Returns ``(mask, mix)`` or ``(mask, mix, mixlabel)``; ``mix`` is ``None`` when every label is 255, the case
    in which the reference never creates the tensor (:271 / :321).

    ``present`` (a ``present_classes_async`` handle or the list ``present_classes`` returns) re-uses a presence pass over
    the same ``slabel``; the class choice is still drawn here, in call order, from ``rng``.
    ``classes`` (per-image lists) skips the presence pass and the host draw; the all-255 test then needs its own
    device reduction + sync unless ``assume_labelled=True`` promises that some label differs from 255.
    """
    L.require_cuda(slabel, a, b, tlabel, what="classmix input")
    lab = L.i64c(slabel)
    bsz = lab.shape[0]
    if isinstance(present, PresentClasses):
        present = present.result()
    if classes is None:
        if present is None:
            present = present_classes(lab)
        classes = select_classes(present, rng)
    lut_host = np.zeros((bsz, 256), dtype=np.uint8)
    for i, sel in enumerate(classes):
        lut_host[i, [c for c in sel if 0 <= c <= 255]] = 1
    hw = lab.shape[1] * lab.shape[2]
    fa, fb = L.f32c(a), L.f32c(b)
    if lab.dim() != 3 or fa.shape != fb.shape or fa.dim() != 4 or fa.shape[0] != bsz or fa.shape[2] * fa.shape[3] != hw:
        raise ValueError("classmix: image / label shapes do not match")
    if present is None:
        all_ignore = False if assume_labelled else bool(torch.all(torch.eq(lab, IGNORE)))
    else:
        all_ignore = all(p == [IGNORE] for p in present)
    mask = torch.empty(lab.shape, dtype=torch.float32, device=lab.device) if return_mask else None
    mix = None if all_ignore else torch.empty_like(fa)
    tl = mixlabel = None
    if tlabel is not None:
        tl = L.i64c(tlabel)
        mixlabel = torch.empty_like(tl)
    L.check(L.lib.diga_classmix_blend(lab.data_ptr(), lut_host.ctypes.data, L.ptr(fa), L.ptr(fb), L.ptr(tl), bsz, fa.shape[1],
                                      hw, L.ptr(mask), L.ptr(mix), L.ptr(mixlabel), L.stream()))
    if tlabel is None:
        return mask, mix
    return mask, mix, mixlabel
