"""Drop-in for ``util/loss.py:125-143`` of the reference (``distillation_loss``).

Same name, arguments, return value and autograd behaviour: a 0-dim fp32 tensor connected to
``student_out`` only (the teacher soft targets are detached, loss.py:128).  The arithmetic is one fused
sm_100a kernel per direction instead of ~45 elementwise passes (diga_b200/csrc/kd.cu).
"""
from __future__ import annotations

import torch

from .. import _lib as L


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


def distillation_loss(teacher_out, student_out, scale=0.5):
    """``util.loss.distillation_loss`` (G/util/loss.py:125; default ``scale`` is 0.25 in the Synthia tree).

    ``teacher_out``, ``student_out``: ``[2B,C,H,W]`` fp32 logits of two stacked views.  Returns
    ``mean CE(softmax(t_view0), s_view1) + scale * mean CE(softmax(t_view1), s_view0)``.
    """
    _check(teacher_out, student_out)
    return _DistillationLoss.apply(teacher_out, student_out, scale)


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


def cross_entropy2d(input, target, weight=None, size_average=True):
    """``util.loss.cross_entropy2d`` (G/util/loss.py:48-62): pixel-wise cross entropy with ``ignore_index=255``;
    pixels with a negative target are dropped; ``size_average`` divides by the number of pixels with target >= 0
    (ignore-255 pixels included, exactly like the reference).  ``input [N,C,H,W]`` fp32, ``target [N,H,W]`` int64."""
    L.require_cuda(input, target, weight, what="cross_entropy2d input")
    if input.dim() != 4 or target.dim() != 3 or input.shape[0] != target.shape[0] or input.shape[2:] != target.shape[1:]:
        raise ValueError(f"cross_entropy2d: expected [N,C,H,W] logits and [N,H,W] targets, got {tuple(input.shape)} and "
                         f"{tuple(target.shape)}")
    return _CrossEntropy2d.apply(input, target, weight, size_average)
