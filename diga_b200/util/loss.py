"""Drop-in for ``util/loss.py:125-143`` of the reference (``distillation_loss``).

Same name, arguments, return value and autograd behaviour: a 0-dim fp32 tensor connected to
``student_out`` only (the teacher soft targets are detached, loss.py:128).  The arithmetic is one fused
sm_100a kernel per direction instead of ~45 elementwise passes (diga_b200/csrc/kd.cu).
"""
from __future__ import annotations

import torch

from .. import _lib as L
from ..nn import LazyUpsampled


def _materialized(x):
    return x.materialize() if isinstance(x, LazyUpsampled) else x


def _check(teacher_out: torch.Tensor, student_out: torch.Tensor):
    L.require_cuda(teacher_out, student_out, what="distillation_loss input")
    if teacher_out.shape != student_out.shape or teacher_out.dim() != 4:
        raise ValueError(f"distillation_loss: expected two [2B,C,H,W] tensors, got {tuple(teacher_out.shape)} "
                         f"and {tuple(student_out.shape)}")
    if teacher_out.shape[0] % 2:
        # the reference's chunk(2) on an odd batch pairs views of different size and fails to broadcast
        raise ValueError("distillation_loss: batch must hold two equally sized views (even size)")


class _DistillationLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher_out, student_out, scale):
        t, s = L.f32c(teacher_out.detach()), L.f32c(student_out.detach())
        n2, c, h, w = s.shape
        loss = torch.empty((), dtype=torch.float32, device=s.device)
        L.check(L.lib.diga_kd_fwd(t.data_ptr(), s.data_ptr(), n2, c, h * w, float(scale), loss.data_ptr(),
                                  L.kd_workspace(s.device).data_ptr(), L.stream()))
        ctx.save_for_backward(t, s)
        ctx.scale = float(scale)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        t, s = ctx.saved_tensors
        n2, c, h, w = s.shape
        g = grad_out.to(dtype=torch.float32, device=s.device).contiguous()   # 0-dim device scalar: read on the GPU
        ds = torch.empty_like(s)
        L.check(L.lib.diga_kd_bwd(t.data_ptr(), s.data_ptr(), n2, c, h * w, ctx.scale, g.data_ptr(), ds.data_ptr(),
                                  L.stream()))
        return None, ds, None


@L.on_device
def distillation_loss(teacher_out, student_out, scale=0.5):
    """``util.loss.distillation_loss`` (G/util/loss.py:125; default ``scale`` is 0.25 in the Synthia tree).

    ``teacher_out``, ``student_out``: ``[2B,C,H,W]`` fp32 logits of two stacked views.  Returns
    ``mean CE(softmax(t_view0), s_view1) + scale * mean CE(softmax(t_view1), s_view0)``.
    Outputs of ``diga_b200.nn.Upsample`` (:class:`~diga_b200.nn.LazyUpsampled`) are consumed at their stride-8 resolution
    by the fused up-sampling kernels — same call site as the reference's :289,:351-352.
    """
    if isinstance(teacher_out, LazyUpsampled) and isinstance(student_out, LazyUpsampled) and \
            teacher_out.out_size == student_out.out_size:
        return distillation_loss_upsampled(teacher_out.low, student_out.low, student_out.out_size, scale)
    teacher_out, student_out = _materialized(teacher_out), _materialized(student_out)
    _check(teacher_out, student_out)
    return _DistillationLoss.apply(teacher_out, student_out, scale)


@L.on_device
def distillation_loss_and_grad(teacher_out, student_out, scale=0.5, grad_scale=1.0):
    """Single-pass variant (not in the reference): returns ``(loss, grad_scale * dloss/dstudent)``.

    228 B/px of HBM traffic instead of the 152 + 228 B/px of the autograd pair, for call sites that know
    the upstream factor (``lambda_distil``) when the loss is computed.
    """
    _check(teacher_out, student_out)
    t, s = L.f32c(teacher_out.detach()), L.f32c(student_out.detach())
    n2, c, h, w = s.shape
    loss = torch.empty((), dtype=torch.float32, device=s.device)
    ds = torch.empty_like(s)
    L.check(L.lib.diga_kd_fwd_bwd(t.data_ptr(), s.data_ptr(), n2, c, h * w, float(scale), float(grad_scale),
                                  loss.data_ptr(), ds.data_ptr(), L.kd_workspace(s.device).data_ptr(), L.stream()))
    return loss, ds


class _CrossEntropy2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, target, weight, size_average):
        x = L.f32c(input.detach())
        t = L.i64c(target)
        wt = None if weight is None else L.f32c(weight.detach()).to(x.device)
        n, c, h, w = x.shape
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        denom = torch.empty((), dtype=torch.float32, device=x.device)
        L.check(L.lib.diga_cross_entropy2d_fwd(x.data_ptr(), t.data_ptr(), L.ptr(wt), n, c, h * w, int(bool(size_average)),
                                               loss.data_ptr(), denom.data_ptr(), L.ce_workspace(x.device).data_ptr(),
                                               L.stream()))
        ctx.save_for_backward(x, t, denom)
        ctx.weight = wt
        ctx.size_average = int(bool(size_average))
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        x, t, denom = ctx.saved_tensors
        n, c, h, w = x.shape
        g = grad_out.to(dtype=torch.float32, device=x.device).contiguous()
        dx = torch.empty_like(x)
        L.check(L.lib.diga_cross_entropy2d_bwd(x.data_ptr(), t.data_ptr(), L.ptr(ctx.weight), n, c, h * w, ctx.size_average,
                                               g.data_ptr(), denom.data_ptr(), dx.data_ptr(), L.stream()))
        return dx, None, None, None


@L.on_device
def cross_entropy2d(input, target, weight=None, size_average=True):
    """``util.loss.cross_entropy2d`` (G/util/loss.py:48-62): pixel-wise cross entropy with ``ignore_index=255``;
    pixels with a negative target are dropped; ``size_average`` divides by the number of pixels with target >= 0
    (ignore-255 pixels included, exactly like the reference).  ``input [N,C,H,W]`` fp32, ``target [N,H,W]`` int64.
    An output of ``diga_b200.nn.Upsample`` is consumed at its stride-8 resolution (fused up-sampling, :344,:348-349,:355).
    Targets in ``[C, 255)`` are outside the reference's domain (``F.nll_loss`` raises a device-side assert there): the kernels
    treat them like 255 — ignored in the loss, counted in the ``size_average`` denominator — instead of aborting the step."""
    if isinstance(input, LazyUpsampled):
        if tuple(target.shape[1:]) == input.out_size:
            return cross_entropy2d_upsampled(input.low, target, weight, size_average)
        input = input.materialize()
    L.require_cuda(input, target, weight, what="cross_entropy2d input")
    if input.dim() != 4 or target.dim() != 3 or input.shape[0] != target.shape[0] or input.shape[2:] != target.shape[1:]:
        raise ValueError(f"cross_entropy2d: expected [N,C,H,W] logits and [N,H,W] targets, got {tuple(input.shape)} and "
                         f"{tuple(target.shape)}")
    return _CrossEntropy2d.apply(input, target, weight, size_average)


# ----------------------------------------------------------------------------------------------------------------------
# f1 for the loss consumers: the same losses evaluated straight from the stride-8 logits (csrc/loss_up.cu).  The
# reference up-samples first (train_DiGA_gta2city_self_training.py:289,:344,:348,:351: nn.Upsample(bilinear,
# align_corners=True)) and so materialises, per loss, the up-sampled logits and their gradient (76 B/px each).
# ----------------------------------------------------------------------------------------------------------------------
def _size2(size):
    hh, ww = (int(size[0]), int(size[1]))
    return hh, ww


class _LossesUpsampled(torch.autograd.Function):
    """(loss_ce, loss_kd) of the first ``n_ce`` / all images of ``student_low``; either part may be switched off.

    With a single loss and a student that requires grad, the forward launches the loss+gradient kernel with a unit upstream
    (one pass instead of two) and leaves the per-CTA gradient patches in a scratch buffer of its own; the backward is the
    gather over those patches, multiplied by the upstream scalar on the device (``diga_loss_up_gather``: one launch, where
    gather + scale were two).  With both losses the backward is its own pass, because the two upstream scalars are unknown
    until then."""

    @staticmethod
    def forward(ctx, teacher_low, student_low, target, weight, size, scale, size_average):
        s = L.f32c(student_low.detach())
        t = None if teacher_low is None else L.f32c(teacher_low.detach())
        tg = None if target is None else L.i64c(target)
        wt = None if weight is None else L.f32c(weight.detach()).to(s.device)
        n, c, h, w = s.shape
        hh, ww = size
        n_ce = 0 if tg is None else tg.shape[0]
        dev = s.device
        ctx.set_materialize_grads(False)               # an unused loss arrives as None in backward, not as a zeros() launch
        new = lambda: torch.empty((), dtype=torch.float32, device=dev)
        loss_kd = None if t is None else new()         # the switched-off loss is None (no fill kernel for a value nobody reads)
        loss_ce = None if tg is None else new()
        denom = None if tg is None else new()          # written by the kernel whenever CE is on
        ws = L.loss_up_workspace(n, c, h, w, hh, ww, dev)
        size_average = int(bool(size_average))
        unit = None
        if ctx.needs_input_grad[1] and (t is None) != (tg is None):
            unit = _new_scratch(n, c, h, w, hh, ww, dev)          # the unit gradient as patches: gathered (and scaled) in backward
            if tg is None:
                L.check(L.lib.diga_kd_up_fwd_bwd(t.data_ptr(), s.data_ptr(), n, c, h, w, hh, ww, float(scale), 1.0,
                                                 loss_kd.data_ptr(), None, unit.data_ptr(), ws.data_ptr(), L.stream()))
            else:
                L.check(L.lib.diga_ce_up_fwd_bwd(s.data_ptr(), tg.data_ptr(), L.ptr(wt), n, c, h, w, hh, ww, size_average,
                                                 loss_ce.data_ptr(), denom.data_ptr(), None, unit.data_ptr(), ws.data_ptr(),
                                                 L.stream()))
            ctx.save_for_backward(unit, denom)
            ctx.low_shape = (n, c, h, w)
        else:
            L.check(L.lib.diga_loss_up_fwd(L.ptr(t), s.data_ptr(), L.ptr(tg), L.ptr(wt), n, n_ce, c, h, w, hh, ww,
                                           float(scale), size_average, L.ptr(loss_kd), L.ptr(loss_ce),
                                           L.ptr(denom), ws.data_ptr(), L.stream()))
            ctx.save_for_backward(s, denom)
        ctx.aux = (t, tg, wt, (hh, ww), float(scale), size_average, n_ce, unit is not None)
        return loss_ce, loss_kd

    @staticmethod
    def backward(ctx, g_ce, g_kd):
        s, denom = ctx.saved_tensors
        t, tg, wt, (hh, ww), scale, size_average, n_ce, have_unit = ctx.aux
        n, c, h, w = ctx.low_shape if have_unit else s.shape
        dev = s.device
        as_scalar = lambda g: None if g is None else g.to(dtype=torch.float32, device=dev).contiguous()
        g_ce, g_kd = as_scalar(g_ce), as_scalar(g_kd)                 # 0-dim device scalars, read on the GPU
        if have_unit:                                                 # `s` holds the unit gradient's patches saved by the forward
            g = g_kd if tg is None else g_ce
            if g is None:
                return None, None, None, None, None, None, None
            ds = _gathered(s, g, denom if (tg is not None and size_average) else None, (n, c, h, w), (hh, ww))
            return None, ds, None, None, None, None, None
        if g_ce is None and g_kd is None:
            return None, None, None, None, None, None, None
        zero = None
        if (t is not None and g_kd is None) or (tg is not None and g_ce is None):   # one of two live losses left out of the graph
            zero = torch.zeros((), dtype=torch.float32, device=dev)
        ds = torch.empty_like(s)
        ws = L.loss_up_workspace(n, c, h, w, hh, ww, dev)
        L.check(L.lib.diga_loss_up_bwd(L.ptr(t), s.data_ptr(), L.ptr(tg), L.ptr(wt), n, n_ce, c, h, w, hh, ww, scale,
                                       size_average, L.ptr(zero if (t is not None and g_kd is None) else g_kd),
                                       L.ptr(zero if (tg is not None and g_ce is None) else g_ce), L.ptr(denom), ds.data_ptr(),
                                       ws.data_ptr(), L.stream()))
        return None, ds, None, None, None, None, None


def _new_scratch(n, c, h, w, hh, ww, device):
    """Uninitialised buffer for the per-CTA gradient patches of one loss+gradient launch (diga_loss_up_scratch_bytes)."""
    return torch.empty((int(L.lib.diga_loss_up_scratch_bytes(n, c, h, w, hh, ww)) // 4,), dtype=torch.float32, device=device)


def _gathered(scratch, num, den, low_shape, size):
    """``dlow = (sum of the patches in scratch) * (num / den)`` (``den`` None: ``* num``), 0-dim device scalars: one launch."""
    n, c, h, w = low_shape
    dlow = torch.empty(low_shape, dtype=torch.float32, device=scratch.device)
    L.check(L.lib.diga_loss_up_gather(scratch.data_ptr(), num.data_ptr(), L.ptr(den), n, c, h, w, size[0], size[1], dlow.data_ptr(),
                                      L.stream()))
    return dlow


def _check_low(student_low, size, what):
    if student_low.dim() != 4:
        raise ValueError(f"{what}: expected [N,C,h,w] logits, got {tuple(student_low.shape)}")
    hh, ww = _size2(size)
    if hh < student_low.shape[2] or ww < student_low.shape[3]:
        raise ValueError(f"{what}: output size {(hh, ww)} must not be smaller than the logits {tuple(student_low.shape[2:])} "
                         "(the reference only up-samples)")
    return hh, ww


@L.on_device
def distillation_loss_upsampled(teacher_low, student_low, size, scale=0.5):
    """``distillation_loss(upsample(teacher_low), upsample(student_low), scale)`` with ``upsample =
    nn.Upsample(size, mode='bilinear', align_corners=True)`` (self_training.py:289,:351-352), neither up-sampled tensor
    nor its gradient materialised.  Autograd-connected to ``student_low`` ([2B,C,h,w]) only."""
    L.require_cuda(teacher_low, student_low, what="distillation_loss_upsampled input")
    if teacher_low.shape != student_low.shape or student_low.dim() != 4 or student_low.shape[0] % 2:
        raise ValueError(f"distillation_loss_upsampled: expected two [2B,C,h,w] tensors, got {tuple(teacher_low.shape)} and "
                         f"{tuple(student_low.shape)}")
    size = _check_low(student_low, size, "distillation_loss_upsampled")
    return _LossesUpsampled.apply(teacher_low, student_low, None, None, size, scale, True)[1]


@L.on_device
def cross_entropy2d_upsampled(input_low, target, weight=None, size_average=True):
    """``cross_entropy2d(upsample(input_low), target, weight, size_average)`` with the up-sampling to ``target``'s
    resolution fused (self_training.py:344,:348-349,:355).  ``input_low [N,C,h,w]`` fp32, ``target [N,H,W]`` int64."""
    L.require_cuda(input_low, target, weight, what="cross_entropy2d_upsampled input")
    if target.dim() != 3 or input_low.dim() != 4 or target.shape[0] != input_low.shape[0]:
        raise ValueError(f"cross_entropy2d_upsampled: expected [N,C,h,w] logits and [N,H,W] targets, got "
                         f"{tuple(input_low.shape)} and {tuple(target.shape)}")
    size = _check_low(input_low, target.shape[1:], "cross_entropy2d_upsampled")
    return _LossesUpsampled.apply(None, input_low, target, weight, size, 0.0, size_average)[0]


@L.on_device
def seg_distillation_losses_upsampled(teacher_low, student_low, target, scale=0.5, weight=None, size_average=True):
    """The two losses that share ``s_pred_cat_stu`` in the self-training step (self_training.py:348-352) from ONE pass over
    the stride-8 logits: returns ``(cross_entropy2d(upsample(student_low[:B]), target), distillation_loss(upsample(
    teacher_low), upsample(student_low), scale))``; the backward adds both gradients in one kernel."""
    L.require_cuda(teacher_low, student_low, target, weight, what="seg_distillation_losses_upsampled input")
    if teacher_low.shape != student_low.shape or student_low.dim() != 4 or student_low.shape[0] % 2:
        raise ValueError("seg_distillation_losses_upsampled: expected two [2B,C,h,w] logit tensors")
    if target.dim() != 3 or not (1 <= target.shape[0] <= student_low.shape[0]):
        raise ValueError("seg_distillation_losses_upsampled: target must be [n_ce,H,W] with 1 <= n_ce <= 2B")
    size = _check_low(student_low, target.shape[1:], "seg_distillation_losses_upsampled")
    return _LossesUpsampled.apply(teacher_low, student_low, target, weight, size, scale, size_average)


@L.on_device
def distillation_loss_upsampled_and_grad(teacher_low, student_low, size, scale=0.5, grad_scale=1.0):
    """Single-pass variant: ``(loss, grad_scale * dloss/dstudent_low)`` for call sites that know ``lambda_distil``."""
    L.require_cuda(teacher_low, student_low, what="distillation_loss_upsampled_and_grad input")
    if teacher_low.shape != student_low.shape or student_low.dim() != 4 or student_low.shape[0] % 2:
        raise ValueError("distillation_loss_upsampled_and_grad: expected two [2B,C,h,w] tensors")
    hh, ww = _check_low(student_low, size, "distillation_loss_upsampled_and_grad")
    t, s = L.f32c(teacher_low.detach()), L.f32c(student_low.detach())
    n, c, h, w = s.shape
    loss = torch.empty((), dtype=torch.float32, device=s.device)
    ds = torch.empty_like(s)
    ws = L.loss_up_workspace(n, c, h, w, hh, ww, s.device)
    L.check(L.lib.diga_kd_up_fwd_bwd(t.data_ptr(), s.data_ptr(), n, c, h, w, hh, ww, float(scale), float(grad_scale),
                                     loss.data_ptr(), ds.data_ptr(), None, ws.data_ptr(), L.stream()))
    return loss, ds


class _SegDistillationTotal(torch.autograd.Function):
    """``lambda_seg * CE + lambda_distil * KD`` from the stride-8 logits with the weights known up front: the forward
    launches the loss+gradient kernel once (the weighted gradient comes out of the same pass), the backward scales it."""

    @staticmethod
    def forward(ctx, teacher_low, student_low, target, weight, size, scale, size_average, lambda_seg, lambda_distil,
                targets_nonnegative):
        s, t, tg = L.f32c(student_low.detach()), L.f32c(teacher_low.detach()), L.i64c(target)
        wt = None if weight is None else L.f32c(weight.detach()).to(s.device)
        n, c, h, w = s.shape
        hh, ww = size
        dev = s.device
        ctx.set_materialize_grads(False)
        loss_kd, loss_ce, denom = (torch.empty((), dtype=torch.float32, device=dev) for _ in range(3))
        ws = L.loss_up_workspace(n, c, h, w, hh, ww, dev)
        if ctx.needs_input_grad[1]:
            ds = _new_scratch(n, c, h, w, hh, ww, dev)                   # gradient patches: gathered and scaled in backward
            total = torch.empty((), dtype=torch.float32, device=dev)     # written by the kernel's last CTA: no scalar launches
            L.check(L.lib.diga_seg_kd_up_fwd_bwd(t.data_ptr(), s.data_ptr(), tg.data_ptr(), L.ptr(wt), n, tg.shape[0], c, h, w,
                                                 hh, ww, float(scale), int(bool(size_average)), float(lambda_seg),
                                                 float(lambda_distil), float(tg.numel()) if targets_nonnegative else 0.0,
                                                 loss_kd.data_ptr(), loss_ce.data_ptr(), denom.data_ptr(), total.data_ptr(),
                                                 None, ds.data_ptr(), ws.data_ptr(), L.stream()))
            ctx.save_for_backward(ds)
            ctx.geom = ((n, c, h, w), (hh, ww))
        else:
            L.check(L.lib.diga_loss_up_fwd(t.data_ptr(), s.data_ptr(), tg.data_ptr(), L.ptr(wt), n, tg.shape[0], c, h, w, hh, ww,
                                           float(scale), int(bool(size_average)), loss_kd.data_ptr(), loss_ce.data_ptr(),
                                           denom.data_ptr(), ws.data_ptr(), L.stream()))
            total = lambda_seg * loss_ce + lambda_distil * loss_kd
        ctx.mark_non_differentiable(loss_ce, loss_kd)
        return total, loss_ce, loss_kd

    @staticmethod
    def backward(ctx, g_total, _g_ce, _g_kd):
        (ds,) = ctx.saved_tensors
        if g_total is None:
            return (None,) * 10
        return (None, _gathered(ds, g_total.to(dtype=torch.float32, device=ds.device).contiguous(), None, *ctx.geom)) + (None,) * 8


@L.on_device
def seg_distillation_total_upsampled(teacher_low, student_low, target, lambda_seg=1.0, lambda_distil=0.25, scale=0.5,
                                     weight=None, size_average=True, targets_nonnegative=False):
    """``total = lambda_seg * seg_loss(upsample(student_low[:B]), target) + lambda_distil * distillation_loss(upsample(
    teacher_low), upsample(student_low), scale)`` — self_training.py:348-352 and the source-image part of :382 — with the
    loss weights known when the losses are computed, so loss and gradient cost ONE pass over the stride-8 logits.
    Returns ``(total, loss_seg, loss_distil)``; only ``total`` carries a gradient (the two parts are for logging).
    ``targets_nonnegative=True`` is the caller's promise that no target is negative (the loaders deliver trainIds or 255),
    which makes the ``size_average`` denominator ``#(target >= 0)`` (util/loss.py:56,:60) the number of target pixels and saves
    the counting pass in front of the loss kernel; the kernel still counts, and ``total`` / ``loss_seg`` are NaN if the promise
    was wrong."""
    L.require_cuda(teacher_low, student_low, target, weight, what="seg_distillation_total_upsampled input")
    if teacher_low.shape != student_low.shape or student_low.dim() != 4 or student_low.shape[0] % 2:
        raise ValueError("seg_distillation_total_upsampled: expected two [2B,C,h,w] logit tensors")
    if target.dim() != 3 or not (1 <= target.shape[0] <= student_low.shape[0]):
        raise ValueError("seg_distillation_total_upsampled: target must be [n_ce,H,W] with 1 <= n_ce <= 2B")
    size = _check_low(student_low, target.shape[1:], "seg_distillation_total_upsampled")
    return _SegDistillationTotal.apply(teacher_low, student_low, target, weight, size, scale, size_average, float(lambda_seg),
                                       float(lambda_distil), bool(targets_nonnegative))


# ----------------------------------------------------------------------------------------------------------------------
# OhemCrossEntropy (G/util/loss.py:65-122; the seg loss of the Synthia tree, S/train_DiGA_syn2city_self_training.py:184)
# ----------------------------------------------------------------------------------------------------------------------
class _OhemUpsampled(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, target, weight, ignore_label, thresh, min_kept):
        s, tg = L.f32c(score.detach()), L.i64c(target)
        wt = None if weight is None else L.f32c(weight.detach()).to(s.device)
        n, c, h, w = s.shape
        hh, ww = tg.shape[1], tg.shape[2]
        dev = s.device
        pred = torch.empty((n, hh, ww), dtype=torch.float32, device=dev)
        losspx = torch.empty((n, hh, ww), dtype=torch.float32, device=dev)
        loss, count, thr = (torch.empty((), dtype=torch.float32, device=dev) for _ in range(3))
        L.check(L.lib.diga_ohem_up_fwd(s.data_ptr(), tg.data_ptr(), L.ptr(wt), n, c, h, w, hh, ww, int(ignore_label), float(thresh),
                                       int(min_kept), pred.data_ptr(), losspx.data_ptr(), loss.data_ptr(), count.data_ptr(),
                                       thr.data_ptr(), L.ohem_workspace(dev).data_ptr(), L.stream()))
        ctx.save_for_backward(s, tg, pred, thr, count)
        ctx.aux = (wt, int(ignore_label))
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        s, tg, pred, thr, count = ctx.saved_tensors
        wt, ignore_label = ctx.aux
        n, c, h, w = s.shape
        hh, ww = tg.shape[1], tg.shape[2]
        g = grad_out.to(dtype=torch.float32, device=s.device).contiguous()
        ds = torch.empty_like(s)
        L.check(L.lib.diga_ohem_up_bwd(s.data_ptr(), tg.data_ptr(), L.ptr(wt), n, c, h, w, hh, ww, ignore_label, pred.data_ptr(),
                                       thr.data_ptr(), count.data_ptr(), g.data_ptr(), ds.data_ptr(),
                                       L.loss_up_workspace(n, c, h, w, hh, ww, s.device).data_ptr(), L.stream()))
        return ds, None, None, None, None, None


class OhemCrossEntropy(torch.nn.Module):
    """``util.loss.OhemCrossEntropy`` (G/util/loss.py:65-122): online hard example mining on the target-class probability —
    keep the pixels whose probability is below ``max(thres, the min_kept-th smallest probability)``, mean CE over them.
    Same constructor and ``forward(score, target)``; as in the reference, a ``score`` smaller than ``target`` is up-sampled
    bilinearly (``align_corners=True``) first — here without materialising it, forward or backward."""

    def __init__(self, ignore_label=255, thres=0.7, min_kept=100000, weight=None):
        super().__init__()
        self.thresh = thres
        self.min_kept = max(1, min_kept)
        self.ignore_label = ignore_label
        self.weight = weight

    @L.on_device
    def forward(self, score, target):
        if isinstance(score, LazyUpsampled):       # _ohem_forward up-samples the score itself (:91-95): hand it the stride-8 map
            score = score.low if tuple(target.shape[1:]) == score.out_size else score.materialize()
        L.require_cuda(score, target, self.weight, what="OhemCrossEntropy input")
        if score.dim() != 4 or target.dim() != 3 or score.shape[0] != target.shape[0]:
            raise ValueError(f"OhemCrossEntropy: expected [N,C,h,w] scores and [N,H,W] targets, got {tuple(score.shape)} and "
                             f"{tuple(target.shape)}")
        _check_low(score, target.shape[1:], "OhemCrossEntropy")
        return _OhemUpsampled.apply(score, target, self.weight, self.ignore_label, self.thresh, self.min_kept)
