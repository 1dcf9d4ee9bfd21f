"""Pseudo-label generation: (two-scale max) -> softmax -> argmax + confidence in one kernel.

The reference has no function for this; it is ``pseudolabel_generator.py:77-85`` (twins in the Synthia and
semi-supervised trees).  ``pseudo_label`` equals lines :80-85 on already up-sampled logits;
``pseudo_label_two_scale`` also folds the two bilinear up-samplings of :77-78 into the kernel.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib as L

# Cityscapes trainId palette the reference writes into its 'P'-mode PNGs (pseudolabel_generator.py:38-43), zero-padded
# to 256 entries.
CITYSCAPES_PALETTE = [128, 64, 128, 244, 35, 232, 70, 70, 70, 102, 102, 156, 190, 153, 153, 153, 153, 153, 250, 170, 30,
                      220, 220, 0, 107, 142, 35, 152, 251, 152, 70, 130, 180, 220, 20, 60, 255, 0, 0, 0, 0, 142, 0, 0, 70,
                      0, 60, 100, 0, 80, 100, 0, 0, 230, 119, 11, 32]
CITYSCAPES_PALETTE = CITYSCAPES_PALETTE + [0] * (256 * 3 - len(CITYSCAPES_PALETTE))


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


def colorize_mask(mask):
    """``pseudolabel_generator.py:45-49``: uint8 label map -> PIL 'P' image whose palette index is the trainId."""
    from PIL import Image
    img = Image.fromarray(np.asarray(mask).astype(np.uint8)).convert('P')
    img.putpalette(CITYSCAPES_PALETTE)
    return img


class PseudoLabelWriter:
    """Streams uint8 label maps from the GPU to palette PNGs (next row f3; replaces pseudolabel_generator.py:66,89-105).

    The reference keeps all 2975 label maps in one float64 host array (50 GB) after pulling the 159 MB softmax tensor of
    every image over PCIe, and encodes the PNGs in a second loop.  Here the kernel's uint8 map (2 MB per 2048x1024 image)
    is copied into a pinned staging buffer on a side stream and encoded by a small thread pool while the GPU continues;
    the files are identical in format ('P' mode, palette index = trainId, file name = basename of the image name).
    """

    def __init__(self, output_dir, workers=4, slots=4):
        self.output_dir = output_dir
        os.makedirs(output_dir, exist_ok=True)
        self._pool = ThreadPoolExecutor(max_workers=workers)
        self._copy_stream = torch.cuda.Stream()
        self._slots = [None] * slots        # (pinned buffer, cuda event, pending futures)
        self._next = 0
        self.written = 0

    def _encode(self, arr, name):
        colorize_mask(arr).save(os.path.join(self.output_dir, name.split('/')[-1]))

    def submit(self, label_u8, names):
        """``label_u8 [N,H,W]`` uint8 CUDA tensor, ``names``: N file names (``name.split('/')[-1]`` is used, :102)."""
        L.require_cuda(label_u8, what="pseudo-label map")
        if label_u8.dtype != torch.uint8 or label_u8.dim() != 3 or label_u8.shape[0] != len(names):
            raise ValueError("PseudoLabelWriter.submit: expected a uint8 [N,H,W] tensor and N names")
        i = self._next
        self._next = (self._next + 1) % len(self._slots)
        slot = self._slots[i]
        if slot is not None:
            for f in slot[2]:
                f.result()                                  # the staging buffer is free again
        buf = slot[0] if slot is not None and slot[0].shape == label_u8.shape else torch.empty(
            label_u8.shape, dtype=torch.uint8).pin_memory()
        ev = torch.cuda.Event()
        self._copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            buf.copy_(label_u8, non_blocking=True)
            ev.record(self._copy_stream)
        label_u8.record_stream(self._copy_stream)

        def job(k, name):
            ev.synchronize()
            self._encode(buf[k].numpy(), name)

        futures = [self._pool.submit(job, k, nm) for k, nm in enumerate(names)]
        self._slots[i] = (buf, ev, futures)
        self.written += len(names)

    def close(self):
        for slot in self._slots:
            if slot is not None:
                for f in slot[2]:
                    f.result()
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def generate_pseudo_labels(student, loader, output_dir, size=(1024, 2048), workers=4):
    """The loop of ``pseudolabel_generator.py:69-105`` with the per-pixel math and the output path replaced:
    two forward passes (full and half resolution, :73-76), fused up-sampling + max + argmax on the GPU, PNGs streamed
    out by :class:`PseudoLabelWriter`.  ``student(x)`` returns ``(_, _, logits, _)`` like the reference ``SegModel``;
    ``loader`` yields ``(image, _, name)`` batches."""
    import torch.nn.functional as F
    with PseudoLabelWriter(output_dir, workers=workers) as writer, torch.no_grad():
        for index, batch in enumerate(loader):
            image, _, name = batch
            image = image.cuda(non_blocking=True)
            image_ds = F.interpolate(image, (size[0] // 2, size[1] // 2), mode='bilinear', align_corners=True)   # :73
            _, _, output_ds, _ = student(image_ds)                                                                 # :75
            _, _, output, _ = student(image)                                                                       # :76
            label, _ = pseudo_label_two_scale(output, output_ds, size, want_conf=False)                            # :77-85
            writer.submit(label, list(name))
    return writer.written
