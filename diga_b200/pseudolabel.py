"""Pseudo-label generation: (two-scale max) -> softmax -> argmax + confidence in one kernel.

The reference has no function for this; it is ``pseudolabel_generator.py:77-85`` (twins in the Synthia and
semi-supervised trees).  ``pseudo_label`` equals lines :80-85 on already up-sampled logits;
``pseudo_label_two_scale`` also folds the two bilinear up-samplings of :77-78 into the kernel.
"""
from __future__ import annotations

import ctypes
import os
import struct
import zlib
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


@L.on_device
def pseudo_label(logits, logits_ds=None, want_conf=True, want_int64=False):
    """``logits`` (and optional ``logits_ds``): ``[N,C,H,W]`` fp32.  Returns ``(label uint8 [N,H,W], conf fp32
    [N,H,W] or None)`` (plus an int64 copy of the labels when ``want_int64``).  ``label`` is what the reference
    stores after ``np.asarray(label, dtype=np.uint8)`` (:92); ``conf`` is the value it computes and discards.
    Outputs of ``diga_b200.nn.Upsample`` (``upsample_1024(output)``, :77-78) are consumed at their stride-8 resolution."""
    from .nn import LazyUpsampled
    if isinstance(logits, LazyUpsampled) and (logits_ds is None or isinstance(logits_ds, LazyUpsampled)) and not want_int64 and \
            (logits_ds is None or logits_ds.out_size == logits.out_size):
        return pseudo_label_two_scale(logits.low, None if logits_ds is None else logits_ds.low, logits.out_size, want_conf)
    if isinstance(logits, LazyUpsampled):
        logits = logits.materialize()
    if isinstance(logits_ds, LazyUpsampled):
        logits_ds = logits_ds.materialize()
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


@L.on_device
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


_PALETTE_BYTES = (ctypes.c_ubyte * len(CITYSCAPES_PALETTE))(*CITYSCAPES_PALETTE)
_PNG_SIGNATURE = b"\x89PNG\r\n\x1a\n"


def _png_chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def frame_png(idat_payload, height, width, palette=None):
    """Wrap a finished IDAT payload (the zlib stream :func:`png_deflate` makes on the GPU) into a 'P'-mode PNG file:
    signature, IHDR (8-bit, colour type 3), PLTE (the trainId palette of ``pseudolabel_generator.py:38-43``), IDAT, IEND.
    Pure framing — no pixel ever passes through the host."""
    pal = bytes(CITYSCAPES_PALETTE if palette is None else palette)
    ihdr = struct.pack(">IIBBBBB", int(width), int(height), 8, 3, 0, 0, 0)
    return b"".join((_PNG_SIGNATURE, _png_chunk(b"IHDR", ihdr), _png_chunk(b"PLTE", pal), _png_chunk(b"IDAT", bytes(idat_payload)),
                     _png_chunk(b"IEND", b"")))


@L.on_device
def png_deflate(label_u8):
    """``label_u8 [N,H,W]`` uint8 CUDA tensor -> ``(payload uint8 [N, capacity], lengths int64 [N])``, both on the device:
    ``payload[i, :lengths[i]]`` is image i's complete zlib stream (Up-filtered scanlines, one deflate block with a static Huffman table,
    Adler-32), ready for :func:`frame_png`.  Three kernel launches for the whole batch, no host sync."""
    L.require_cuda(label_u8, what="pseudo-label map")
    if label_u8.dtype != torch.uint8 or label_u8.dim() != 3:
        raise ValueError("png_deflate: expected a uint8 [N,H,W] tensor")
    lab = label_u8.contiguous()
    n, h, w = lab.shape
    cap = int(L.lib.diga_png_deflate_capacity(h, w))
    if cap <= 0:
        raise ValueError("png_deflate: empty image")
    payload = torch.empty((n, cap), dtype=torch.uint8, device=lab.device)
    lengths = torch.empty((n,), dtype=torch.int64, device=lab.device)
    scratch = torch.empty((max(int(L.lib.diga_png_deflate_scratch_bytes(n, h)), 16),), dtype=torch.uint8, device=lab.device)
    L.check(L.lib.diga_png_deflate(lab.data_ptr(), n, h, w, payload.data_ptr(), cap, scratch.data_ptr(), lengths.data_ptr(),
                                   L.stream()))
    return payload, lengths


@L.on_device
def png_crc(payload, lengths):
    """CRC-32 of every image's IDAT chunk (``b"IDAT" + payload[i, :lengths[i]]``) computed on the GPU: ``uint32 [N]`` (stored in
    an int64 tensor so that it copies out with the lengths).  ``payload`` / ``lengths`` as :func:`png_deflate` returns them."""
    L.require_cuda(payload, lengths, what="png_crc input")
    if payload.dtype != torch.uint8 or payload.dim() != 2 or lengths.dtype != torch.int64 or lengths.shape != (payload.shape[0],):
        raise ValueError("png_crc: expected the (payload uint8 [N,capacity], lengths int64 [N]) pair of png_deflate")
    n, cap = payload.shape
    tmp = torch.empty((n,), dtype=torch.int32, device=payload.device)
    L.check(L.lib.diga_png_crc(payload.data_ptr(), n, cap, lengths.data_ptr(), tmp.data_ptr(), L.stream()))
    return tmp.to(torch.int64) & 0xFFFFFFFF


class PseudoLabelWriter:
    """Streams label maps from the GPU to palette PNGs (next row f3; replaces pseudolabel_generator.py:66,89-105).

    The reference keeps all 2975 label maps in one float64 host array (50 GB) after pulling the 159 MB softmax tensor of
    every image over PCIe, and encodes the PNGs in a second loop (Pillow + zlib: 17-43 ms per 2048x1024 map and core).
    ``encoder='gpu'`` (default): the zlib stream of every map is produced on the GPU (:func:`png_deflate`), its first
    ``prefix`` bytes and its length are copied to pinned memory on a side stream, and a small thread pool only frames and
    writes the files (``diga_png_write_file``: CRC-32 + framing in the library, outside the interpreter lock); a stream longer than ``prefix`` (noise-like maps) is fetched with a second copy.
    ``coalesce=k`` gathers k maps per encoder call (the per-image call sequence of config 5 then pays the three launches
    once per k images).  ``encoder='pil'``: the uint8 map itself (2 MB per image) is copied out and encoded by Pillow like the reference does.
    Either way the files are 'P' mode, palette index = trainId, file name = basename of the image name, and decode to the
    same pixels as the reference's.
    """

    def __init__(self, output_dir, workers=4, slots=4, encoder="gpu", prefix=256 * 1024, coalesce=1):
        if encoder not in ("gpu", "pil"):
            raise ValueError("PseudoLabelWriter: encoder must be 'gpu' or 'pil'")
        self.output_dir = output_dir
        os.makedirs(output_dir, exist_ok=True)
        self.encoder = encoder
        self.prefix = int(prefix)
        self._pool = ThreadPoolExecutor(max_workers=workers)
        self._copy_stream = torch.cuda.Stream()
        self._slots = [None] * slots        # (pinned buffers, device tensors kept alive, pending futures)
        self._next = 0
        self.written = 0
        self.bytes_d2h = 0
        self.coalesce = int(coalesce)       # gpu encoder: gather this many maps before encoding (one batch of launches)
        self._pending, self._pending_n = [], 0

    def _path(self, name):
        return os.path.join(self.output_dir, name.split('/')[-1])

    def _encode(self, arr, name):
        colorize_mask(arr).save(self._path(name))

    def _wait_slot(self, i):
        slot = self._slots[i]
        if slot is not None:
            for f in slot[2]:
                f.result()                                  # the staging buffers are free again
        return slot

    @L.on_device
    def submit(self, label_u8, names):
        """``label_u8 [N,H,W]`` uint8 CUDA tensor, ``names``: N file names (``name.split('/')[-1]`` is used, :102)."""
        L.require_cuda(label_u8, what="pseudo-label map")
        if label_u8.dtype != torch.uint8 or label_u8.dim() != 3 or label_u8.shape[0] != len(names):
            raise ValueError("PseudoLabelWriter.submit: expected a uint8 [N,H,W] tensor and N names")
        if self.coalesce > 1 and self.encoder == "gpu":
            if self._pending and self._pending[0][0].shape[1:] != label_u8.shape[1:]:
                self.flush()
            self._pending.append((label_u8, list(names)))
            self._pending_n += len(names)
            if self._pending_n >= self.coalesce:
                self.flush()
            return
        self._submit_now(label_u8, names)

    def flush(self):
        """Encode and hand over the maps gathered so far (``coalesce > 1``)."""
        if not self._pending:
            return
        labs, names = [p[0] for p in self._pending], [nm for p in self._pending for nm in p[1]]
        self._pending, self._pending_n = [], 0
        self._submit_now(labs[0] if len(labs) == 1 else torch.cat(labs), names)

    def _submit_now(self, label_u8, names):
        i = self._next
        self._next = (self._next + 1) % len(self._slots)
        slot = self._wait_slot(i)
        ev = torch.cuda.Event()
        if self.encoder == "pil":
            buf = slot[0][0] if slot is not None and slot[0][0].shape == label_u8.shape else torch.empty(
                label_u8.shape, dtype=torch.uint8).pin_memory()
            self._copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._copy_stream):
                buf.copy_(label_u8, non_blocking=True)
                ev.record(self._copy_stream)
            self.bytes_d2h += label_u8.numel()

            def job(k, name):
                ev.synchronize()
                self._encode(buf[k].numpy(), name)

            keep, pinned = (label_u8,), (buf,)
        else:
            n, h, w = label_u8.shape
            payload, lengths = png_deflate(label_u8)
            meta = torch.stack((lengths, png_crc(payload, lengths)))        # [2, n]: stream lengths and IDAT CRC-32s (both made on the GPU)
            pre = min(self.prefix, payload.shape[1])
            if slot is not None and slot[0][0].shape == (n, pre):
                buf, lens = slot[0]
            else:
                buf, lens = torch.empty((n, pre), dtype=torch.uint8).pin_memory(), torch.empty((2, n), dtype=torch.int64).pin_memory()
            self._copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._copy_stream):
                lens.copy_(meta, non_blocking=True)
                buf.copy_(payload[:, :pre], non_blocking=True)
                ev.record(self._copy_stream)
            self.bytes_d2h += n * (pre + 16)

            base = buf.data_ptr()

            def job(k, name):
                ev.synchronize()
                m, crc = int(lens[0, k]), int(lens[1, k])
                if m <= pre:
                    ptr, host = base + k * pre, None
                else:                                        # a noise-like map: fetch the whole stream
                    with torch.cuda.stream(self._copy_stream):
                        host = payload[k, :m].cpu()
                    ptr = host.data_ptr()
                # framing and the file write happen inside the library, i.e. without the interpreter lock; the IDAT CRC-32 came with
                # the stream, so the host thread touches the payload only to hand it to write()
                L.check(L.lib.diga_png_write_file_crc(os.fsencode(self._path(name)), ptr, m, h, w, _PALETTE_BYTES, len(_PALETTE_BYTES), crc))

            keep, pinned = (payload, lengths, label_u8), (buf, lens)
        futures = [self._pool.submit(job, k, nm) for k, nm in enumerate(names)]
        self._slots[i] = (pinned, keep, futures)
        self.written += len(names)

    def close(self):
        self.flush()
        for i in range(len(self._slots)):
            self._wait_slot(i)
            self._slots[i] = None
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def generate_pseudo_labels(student, loader, output_dir, size=(1024, 2048), workers=4, encoder="gpu", coalesce=8, preprocess=None):
    """The loop of ``pseudolabel_generator.py:69-105`` with the per-pixel math and the output path replaced:
    two forward passes (full and half resolution, :73-76), fused up-sampling + max + argmax on the GPU, PNGs streamed
    out by :class:`PseudoLabelWriter`.  ``student(x)`` returns ``(_, _, logits, _)`` like the reference ``SegModel``;
    ``loader`` yields ``(image, _, name)`` batches (batch size 1 in the reference; ``coalesce`` maps are encoded per call).

    Channel order: this mirrors the GTA5 / Synthia trees, which feed the loader's image to the model as is.  The
    semi-supervised tree flips BGR -> RGB at the call site (``student(image[:, [2, 1, 0], :, :])``,
    semi-supervised_segmentation/pseudolabel_generator.py:74-75): pass ``preprocess=lambda x: x[:, [2, 1, 0]]`` there
    (applied to both resolutions before the model)."""
    import torch.nn.functional as F
    with PseudoLabelWriter(output_dir, workers=workers, encoder=encoder, coalesce=coalesce) as writer, torch.no_grad():
        for index, batch in enumerate(loader):
            image, _, name = batch
            image = image.cuda(non_blocking=True)
            image_ds = F.interpolate(image, (size[0] // 2, size[1] // 2), mode='bilinear', align_corners=True)   # :73
            if preprocess is not None:
                image, image_ds = preprocess(image), preprocess(image_ds)
            _, _, output_ds, _ = student(image_ds)                                                                 # :75
            _, _, output, _ = student(image)                                                                       # :76
            label, _ = pseudo_label_two_scale(output, output_ds, size, want_conf=False)                            # :77-85
            writer.submit(label, list(name))
    return writer.written
