"""Bilateral-consensus ("threshold-free dynamic") pseudo-label selection.

The reference has no function for this; it is ``train_DiGA_gta2city_self_training.py:298-304``:
``feat_weights = upsample_tgt(get_centroid_weight(feat)); feat_pseudo = argmax; pseudo[pseudo != feat_pseudo] = 255``.
"""
from __future__ import annotations

import torch

from . import _lib as L


@L.on_device
def consensus_select(pseudo_prob, feat_weights_lowres, out_size=None, want_feat_pseudo=True):
    """``pseudo_prob [B,H,W]`` int64 stored pseudo-labels, ``feat_weights_lowres [B,C,h,w]`` fp32 prototype weights
    (``Class_Features.get_centroid_weight``).  Returns ``(tlabelv_pseudo, feat_pseudo)``, both ``[B,H,W]`` in the dtype of ``pseudo_prob`` (int64, or uint8 for the offline
    pseudo-label path):
    the kept labels (255 where the two views disagree) and the arg-max of the bilinearly up-sampled weights.
    The up-sampled ``[B,C,H,W]`` tensor is never materialised."""
    L.require_cuda(pseudo_prob, feat_weights_lowres, what="consensus_select input")
    u8 = pseudo_prob.dtype == torch.uint8          # offline path: uint8 label maps in, uint8 out (no int64 round trip)
    pp = pseudo_prob.contiguous() if u8 else L.i64c(pseudo_prob)
    wl = L.f32c(feat_weights_lowres.detach())
    b, hh, ww = pp.shape
    if out_size is not None and (int(out_size[0]), int(out_size[1])) != (hh, ww):
        raise ValueError("consensus_select: out_size must equal the pseudo-label resolution")
    if wl.shape[0] != b:
        raise ValueError("consensus_select: batch mismatch")
    _, c, h, w = wl.shape
    kept = torch.empty_like(pp)
    fp = torch.empty_like(pp) if want_feat_pseudo else None
    fn = L.lib.diga_consensus_select_u8 if u8 else L.lib.diga_consensus_select
    L.check(fn(wl.data_ptr(), pp.data_ptr(), b, c, h, w, hh, ww, kept.data_ptr(), L.ptr(fp), L.stream()))
    return kept, fp


@L.on_device
def upsample_bilinear(x, size):
    """``nn.Upsample(size, mode='bilinear', align_corners=True)`` with the interpolation routine the fused kernels
    use (bit-identical to torch's CUDA kernel); exists so that the routine can be tested on its own."""
    L.require_cuda(x, what="upsample input")
    xi = L.f32c(x.detach())
    n, c, h, w = xi.shape
    out = torch.empty((n, c, int(size[0]), int(size[1])), dtype=torch.float32, device=xi.device)
    L.check(L.lib.diga_upsample_bilinear(xi.data_ptr(), n * c, h, w, int(size[0]), int(size[1]), out.data_ptr(),
                                         L.stream()))
    return out
