"""Drop-in for the hot-path helpers of ``util/utils.py`` of the reference: ``process_label`` (:158-163) and the EMA
teacher update ``update_teacher_params`` (:103-116)."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib as L


@L.on_device
def process_label(label, class_numbers=19):
    """``[B,1,H,W]`` float labels -> ``[B,class_numbers+1,H,W]`` fp32 one-hot; ids >= class_numbers go to the
    last channel (G/util/utils.py:158-163; the Synthia tree's default is 16)."""
    L.require_cuda(label, what="process_label input")
    batch, channel, w, h = label.size()
    if channel != 1:
        raise ValueError("process_label: expected a [B,1,H,W] label tensor")
    lab = L.f32c(label)
    out = torch.empty((batch, class_numbers + 1, w, h), dtype=torch.float32, device=label.device)
    L.check(L.lib.diga_onehot_labels(lab.data_ptr(), batch, class_numbers, w * h, out.data_ptr(), L.stream()))
    return out


@L.on_device
def ema_update_tensors(teacher_tensors, student_tensors, alpha):
    """``t = alpha * t + (1 - alpha) * s`` for every pair, in place, up to 512 tensors per kernel launch."""
    ts, ss = list(teacher_tensors), list(student_tensors)
    if len(ts) != len(ss):
        raise ValueError("ema_update_tensors: teacher and student lists differ in length")
    for t, s in zip(ts, ss):
        L.require_cuda(t, s, what="EMA parameter")
        if t.dtype != torch.float32 or s.dtype != torch.float32 or not t.is_contiguous() or not s.is_contiguous():
            raise TypeError("ema_update_tensors: parameters must be contiguous fp32 tensors")
        if t.numel() != s.numel():
            raise ValueError("ema_update_tensors: parameter shapes differ")
    n = len(ts)
    tp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
    sp = (ctypes.c_void_p * n)(*[s.data_ptr() for s in ss])
    ne = (ctypes.c_int64 * n)(*[t.numel() for t in ts])
    L.check(L.lib.diga_ema_update(tp, sp, ne, n, float(alpha), L.stream()))


def update_teacher_params(teacher, student, iteration, stage0=True, mean=False, replace=False):
    """``util.utils.update_teacher_params`` (G/util/utils.py:103-116): EMA of the student's parameters into the teacher
    with ``alpha = min(1 - 1/(iteration+1), 0.999)`` in stage 0, one multi-tensor launch per 512 parameters instead of
    three launches per parameter.  Returns the teacher like the reference (``teacher.cuda()``)."""
    if stage0 == True:          # noqa: E712  (mirrors the reference's flag tests)
        alpha_teacher = min(1 - 1 / (iteration + 1), 0.999)
    elif mean == True:          # noqa: E712
        alpha_teacher = 0.9
    elif replace == True:       # noqa: E712
        alpha_teacher = 0.0
    else:
        alpha_teacher = 0.999
    ema_update_tensors([p.data for p in teacher.parameters()], [p.data for p in student.parameters()], alpha_teacher)
    return teacher.cuda()
