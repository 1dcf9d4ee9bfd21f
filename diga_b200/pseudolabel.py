"""Pseudo-label generation: (two-scale max) -> softmax -> argmax + confidence in one kernel.

The reference has no function for this; it is ``pseudolabel_generator.py:77-85`` (twins in the Synthia and
semi-supervised trees).  ``pseudo_label`` equals lines :80-85 on already up-sampled logits;
``pseudo_label_two_scale`` also folds the two bilinear up-samplings of :77-78 into the kernel.
"""
from __future__ import annotations

import torch

from . import _lib as L


def pseudo_label(logits, logits_ds=None, want_conf=True, want_int64=False):
    """``logits`` (and optional ``logits_ds``): ``[N,C,H,W]`` fp32.  Returns ``(label uint8 [N,H,W], conf fp32
    [N,H,W] or None)`` (plus an int64 copy of the labels when ``want_int64``).  ``label`` is what the reference
    stores after ``np.asarray(label, dtype=np.uint8)`` (:92); ``conf`` is the value it computes and discards."""
    L.require_cuda(logits, logits_ds, what="pseudo_label input")
    z = L.f32c(logits.detach())
    z2 = None
    if logits_ds is not None:
        z2 = L.f32c(logits_ds.detach())
        if z2.shape != z.shape:
            raise ValueError("pseudo_label: both logit maps must have the same (up-sampled) shape")
    n, c, h, w = z.shape
    lab = torch.empty((n, h, w), dtype=torch.uint8, device=z.device)
    conf = torch.empty((n, h, w), dtype=torch.float32, device=z.device) if want_conf else None
    lab64 = torch.empty((n, h, w), dtype=torch.int64, device=z.device) if want_int64 else None
    L.check(L.lib.diga_pseudo_label(z.data_ptr(), L.ptr(z2), n, c, h * w, lab.data_ptr(), L.ptr(lab64), L.ptr(conf),
                                    L.stream()))
    return (lab, conf, lab64) if want_int64 else (lab, conf)


def pseudo_label_two_scale(logits, logits_ds=None, size=(1024, 2048), want_conf=True):
    """``pseudolabel_generator.py:77-85`` from the stride-8 logits: ``logits [N,C,h1,w1]`` (full-resolution pass)
    and ``logits_ds [N,C,h2,w2]`` (half-resolution pass) are bilinearly up-sampled (align_corners) to ``size``
    inside the kernel, max-fused, and arg-maxed.  Nothing of size ``[N,C,H,W]`` is ever written."""
    L.require_cuda(logits, logits_ds, what="pseudo_label input")
    z = L.f32c(logits.detach())
    n, c, h1, w1 = z.shape
    z2, h2, w2 = None, 0, 0
    if logits_ds is not None:
        z2 = L.f32c(logits_ds.detach())
        if z2.shape[:2] != z.shape[:2]:
            raise ValueError("pseudo_label_two_scale: batch / class mismatch")
        h2, w2 = z2.shape[2:]
    hh, ww = int(size[0]), int(size[1])
    lab = torch.empty((n, hh, ww), dtype=torch.uint8, device=z.device)
    conf = torch.empty((n, hh, ww), dtype=torch.float32, device=z.device) if want_conf else None
    L.check(L.lib.diga_pseudo_label_upsampled(z.data_ptr(), h1, w1, L.ptr(z2), h2, w2, n, c, hh, ww, lab.data_ptr(),
                                              None, L.ptr(conf), L.stream()))
    return lab, conf
